// The producer / consumer hand-off of the weight ring in isolation: 8 consumer warps + 1 producer thread, NS-stage ring of
// full / empty mbarriers, no data movement and no work.  Cycles per stage for several wait / arrive styles.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1;} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void arrive_relaxed(uint64_t* bar) { asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool try_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// raw costs: 64 arrives (release / relaxed) on a barrier that never completes, one lane
__global__ void arrive_cost_kernel(long long* out) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000;" ::"r"(smem_u32(&bar)));
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) arrive(&bar);
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) arrive_relaxed(&bar);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t1;
  }
}
__device__ __forceinline__ void wait_all(uint64_t* bar, uint32_t parity) { while (!try_wait(bar, parity)) {} }
__device__ __forceinline__ void wait_elect(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) while (!try_wait(bar, parity)) {}
  __syncwarp();
}
// mode 0: all lanes wait, 8 arrivals per empty barrier (the kernel today)
// mode 1: elected lane waits
// mode 2: all lanes wait, consumers bar.sync then ONE arrival per empty barrier
// mode 3: elected lane waits + one arrival via bar.sync
// mode 4: mode 0 with relaxed arrives (both sides); mode 5: mode 4 + relaxed try_wait
__global__ void __launch_bounds__(384) ring_kernel(long long* out, int mode, int NS, int n_stages, int work_cycles) {
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ecount = (mode == 2 || mode == 3) ? 1 : 8;
  if (threadIdx.x == 0) for (int i = 0; i < NS; ++i) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[i])), "r"(ecount));
  }
  __syncthreads();
  if (warp < 4) {
    if (warp == 0 && lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < n_stages; ++i) {
        if (mode == 5) { while (!try_wait_relaxed(&empty[st], ph ^ 1)) {} } else wait_all(&empty[st], ph ^ 1);
        if (mode >= 4) arrive_relaxed(&full[st]); else arrive(&full[st]);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int cw = warp - 4;
    int st = 0; uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < n_stages; ++i) {
      if (mode == 1 || mode == 3) wait_elect(&full[st], ph);
      else if (mode == 5) { while (!try_wait_relaxed(&full[st], ph)) {} }
      else wait_all(&full[st], ph);
      if (work_cycles) { long long t = clock64(); while (clock64() - t < work_cycles) {} }
      if (mode == 2 || mode == 3) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 128) arrive(&empty[st]);
      } else {
        __syncwarp();
        if (lane == 0) { if (mode >= 4) arrive_relaxed(&empty[st]); else arrive(&empty[st]); }
      }
      if (++st == NS) { st = 0; ph ^= 1; }
    }
    if (cw == 0 && lane == 0) out[0] = clock64() - t0;
  }
}
int main() {
  long long* out; CK(cudaMalloc(&out, 8));
  const int n = 4096;
  { arrive_cost_kernel<<<1, 32>>>(out); CK(cudaDeviceSynchronize()); long long h2[2]; long long* o2; CK(cudaMalloc(&o2, 16));
    arrive_cost_kernel<<<1, 32>>>(o2); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h2, o2, 16, cudaMemcpyDeviceToHost));
    printf("mbarrier.arrive x64: release %.1f cycles each, relaxed %.1f cycles each\n", h2[0] / 64.0, h2[1] / 64.0); }
  for (int work : {0, 200})
    for (int NS : {2, 4, 8})
      for (int mode = 0; mode < 6; ++mode) {
        ring_kernel<<<1, 384>>>(out, mode, NS, n, work); CK(cudaDeviceSynchronize());
        ring_kernel<<<1, 384>>>(out, mode, NS, n, work); CK(cudaDeviceSynchronize());
        long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
        printf("work %3d cyc, ring %d stages, mode %d: %.1f cycles per stage\n", work, NS, mode, (double)h / n);
      }
  return 0;
}
