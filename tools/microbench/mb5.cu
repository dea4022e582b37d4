// Which side of the full / empty mbarrier hand-off is slow?  Each side alone against barriers that are always ready,
// then the named-barrier alternative for the "stage is free" signal.
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
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }

// mode 0: consumer side alone (8 warps): wait on an always-complete barrier, syncwarp, lane-0 arrive on a never-completing one
// mode 1: producer side alone (1 thread): wait on an always-complete barrier, arrive on a never-completing one
// mode 2: full hand-off, empty signal = mbarrier (count 8)              [reference]
// mode 3: full hand-off, empty signal = named barrier: consumers bar.arrive, producer warp bar.sync
// mode 4: mode 2 with test_wait polling on both sides
__global__ void __launch_bounds__(384) k(long long* out, int mode, int n) {
  __shared__ uint64_t full[4], empty[4], done, never;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&empty[i])));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&never)));
    arrive(&done);                                   // phase 0 of `done` is complete for ever
  }
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {
    if (warp >= 4) {
      for (int i = 0; i < n; ++i) { while (!try_wait(&done, 0)) {} __syncwarp(); if (lane == 0) arrive(&never); }
      if (threadIdx.x == 128) out[0] = clock64() - t0;
    }
  } else if (mode == 1) {
    if (threadIdx.x == 0) {
      for (int i = 0; i < n; ++i) { while (!try_wait(&done, 0)) {} arrive(&never); }
      out[0] = clock64() - t0;
    }
  } else if (mode == 2 || mode == 4) {
    if (warp == 0 && lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < n; ++i) {
        if (mode == 4) { while (!test_wait(&empty[st], ph ^ 1)) {} } else { while (!try_wait(&empty[st], ph ^ 1)) {} }
        arrive(&full[st]);
        if (++st == 4) { st = 0; ph ^= 1; }
      }
    } else if (warp >= 4) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < n; ++i) {
        if (mode == 4) { while (!test_wait(&full[st], ph)) {} } else { while (!try_wait(&full[st], ph)) {} }
        __syncwarp(); if (lane == 0) arrive(&empty[st]);
        if (++st == 4) { st = 0; ph ^= 1; }
      }
      if (threadIdx.x == 128) out[0] = clock64() - t0;
    }
  } else if (mode == 3) {
    if (warp == 0) {
      int st = 0;
      for (int i = 0; i < n; ++i) {
        if (i >= 4) asm volatile("bar.sync %0, 288;" ::"r"(8 + st) : "memory");     // stage st released by all consumers
        if (lane == 0) arrive(&full[st]);
        if (++st == 4) st = 0;
      }
    } else if (warp >= 4) {
      int st = 0; uint32_t ph = 0;
      for (int i = 0; i < n; ++i) {
        while (!try_wait(&full[st], ph)) {}
        if (i + 4 < n) asm volatile("bar.arrive %0, 288;" ::"r"(8 + st) : "memory");
        if (++st == 4) { st = 0; ph ^= 1; }
      }
      if (threadIdx.x == 128) out[0] = clock64() - t0;
    }
  }
}
int main() {
  long long* out; CK(cudaMalloc(&out, 8));
  const int n = 4096;
  const char* names[] = {"consumer side alone", "producer side alone", "hand-off, mbarrier empty (try_wait)", "hand-off, named-barrier empty", "hand-off, mbarrier empty (test_wait)"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 384>>>(out, mode, n); CK(cudaDeviceSynchronize()); }
    long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("%-40s: %.1f cycles per stage\n", names[mode], (double)h / n);
  }
  return 0;
}
