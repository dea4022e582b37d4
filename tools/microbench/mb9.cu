// Which load flavour should the exchange words use?  Latency of one warp reading 16 KB (32 x 16 B per lane) that sits in L2,
// and of a 2-SM ping-pong, for ld.volatile (= .sys strong), ld.relaxed.gpu, ld.global.cg; 1 and 8 warps per SM on all SMs.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1;} } while (0)
template <int MODE> __device__ __forceinline__ uint4 ld16(const void* p) {
  uint4 r;
  if (MODE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  if (MODE == 3) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
template <int MODE, int NLD>
__global__ void __launch_bounds__(256, 1) rd(const uint8_t* buf, int iters, int warps, long long* out, unsigned* sink) {
  extern __shared__ uint8_t pad[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  const uint8_t* base = buf + ((size_t)(blockIdx.x * 8 + warp) % 64) * 65536;   // 4 MB window: L2 resident
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint8_t* p = base + ((it * 7 + (acc & 1)) % 3) * 16384 + lane * 16;     // depends on the previous batch
    uint4 v[NLD];
#pragma unroll
    for (int j = 0; j < NLD; ++j) v[j] = ld16<MODE>(p + j * 512);
#pragma unroll
    for (int j = 0; j < NLD; ++j) acc += v[j].x ^ v[j].w;
    acc = __shfl_xor_sync(0xffffffffu, acc, 1) + acc;
  }
  const long long t1 = clock64();
  if (lane == 0) { sink[blockIdx.x * 8 + warp] = acc; if (blockIdx.x == 3 && warp == 0) out[0] = t1 - t0; }
}
template <int MODE>
__global__ void pingpong(uint64_t* buf, int n, long long* out) {
  if (threadIdx.x != 0 || blockIdx.x > 1) return;
  const int me = blockIdx.x;
  const long long t0 = clock64();
  for (int i = 1; i <= n; ++i) {
    uint64_t* mine = buf + (me == 0 ? 0 : 16), *theirs = buf + (me == 0 ? 16 : 0);
    if (me == 0) {
      if (MODE == 0) asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(mine), "r"(i), "r"(i) : "memory");
      else asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(mine), "r"(i), "r"(i) : "memory");
    }
    unsigned d, s;
    do {
      if (MODE == 0) asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(d), "=r"(s) : "l"(theirs) : "memory");
      else if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(d), "=r"(s) : "l"(theirs) : "memory");
      else asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(d), "=r"(s) : "l"(theirs) : "memory");
    } while (s != (unsigned)i);
    if (me == 1) {
      if (MODE == 0) asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(mine), "r"(i), "r"(i) : "memory");
      else asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" ::"l"(mine), "r"(i), "r"(i) : "memory");
    }
  }
  if (me == 0) out[0] = clock64() - t0;
}
int main() {
  uint8_t* buf; CK(cudaMalloc(&buf, 8 << 20)); CK(cudaMemset(buf, 1, 8 << 20));
  long long* out; CK(cudaMalloc(&out, 16)); unsigned* sink; CK(cudaMalloc(&sink, 148 * 8 * 4));
  const char* names[4] = {"ld.volatile", "ld.relaxed.gpu", "ld.global.cg", "ld.global (L1)"};
  auto report = [&](const char* what, int warps, int nld, int iters) { long long c; cudaDeviceSynchronize(); cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
    printf("%-18s %d warps/SM x %2d loads of 16 B per lane: %7.1f cycles per batch\n", what, warps, nld, (double)c / iters); };
  const int iters = 2000;
  for (int warps : {1, 8}) {
    CK(cudaFuncSetAttribute(rd<0, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(rd<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(rd<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(rd<3, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    rd<0, 32><<<148, 256, 200 * 1024>>>(buf, iters, warps, out, sink); report(names[0], warps, 32, iters);
    rd<1, 32><<<148, 256, 200 * 1024>>>(buf, iters, warps, out, sink); report(names[1], warps, 32, iters);
    rd<2, 32><<<148, 256, 200 * 1024>>>(buf, iters, warps, out, sink); report(names[2], warps, 32, iters);
    rd<3, 32><<<148, 256, 200 * 1024>>>(buf, iters, warps, out, sink); report(names[3], warps, 32, iters);
    rd<0, 8><<<148, 256, 0>>>(buf, iters, warps, out, sink); report(names[0], warps, 8, iters);
    rd<1, 8><<<148, 256, 0>>>(buf, iters, warps, out, sink); report(names[1], warps, 8, iters);
    rd<2, 8><<<148, 256, 0>>>(buf, iters, warps, out, sink); report(names[2], warps, 8, iters);
    rd<0, 1><<<148, 256, 0>>>(buf, iters, warps, out, sink); report(names[0], warps, 1, iters);
    rd<1, 1><<<148, 256, 0>>>(buf, iters, warps, out, sink); report(names[1], warps, 1, iters);
    rd<2, 1><<<148, 256, 0>>>(buf, iters, warps, out, sink); report(names[2], warps, 1, iters);
  }
  uint64_t* pp; CK(cudaMalloc(&pp, 4096));
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(pp, 0, 4096));
    if (mode == 0) pingpong<0><<<148, 32>>>(pp, 20000, out);
    if (mode == 1) pingpong<1><<<148, 32>>>(pp, 20000, out);
    if (mode == 2) pingpong<2><<<148, 32>>>(pp, 20000, out);
    long long c; CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost));
    printf("ping-pong %-16s: %.1f cycles one way\n", names[mode], c / 40000.0);
  }
  return 0;
}
