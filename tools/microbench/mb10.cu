// Why does the 16 KB exchange read of decode_flow.cu take 1.2-2 us in the kernel but 0.4 us in mb9?  Same read, the real
// conditions one at a time: number of SMs reading the SAME buffer at the same moment (L2 slice hot-spotting), row pattern
// (8 rows x 64 B per instruction) vs contiguous 512 B, buffer freshly rewritten by other SMs with 8-byte stores or not.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1;} } while (0)
__device__ __forceinline__ void ld2(const uint64_t* p, uint32_t& d0, uint32_t& s0, uint32_t& d1, uint32_t& s1) {
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d0), "=r"(s0), "=r"(d1), "=r"(s1) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld1(const uint64_t* p, uint32_t& d, uint32_t& s) {
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(d), "=r"(s) : "l"(p) : "memory");
}
__device__ __forceinline__ void st1(uint64_t* p, uint32_t d, uint32_t s) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(d), "r"(s) : "memory");
}
// SMs [0, nprod): producers (warp 0): rewrite their 1/nprod of the 2048-word buffer with stamp it+1 once all consumers have
// acknowledged it; SMs [nprod, nprod + ncons): consumers (warp 0): sentinel poll, then the 32-load batch, timed.
__global__ void __launch_bounds__(256, 1) k(uint64_t* buf, uint32_t* ack, int nprod, int ncons, int iters, int pattern, int rewrite,
                                           long long* out, unsigned* sink, int npoll, int poll_sleep) {
  extern __shared__ uint8_t pad[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sm = blockIdx.x;
  if (warp != 0) {
    // optional interference: the other warps of a consumer SM poll a word that never changes, like waiting units do
    if (npoll > 0 && warp <= npoll && sm >= nprod && sm < nprod + ncons) {
      uint32_t d, s2; unsigned n = 0;
      while (*(volatile uint32_t*)ack < (uint32_t)(iters * ncons)) { ld1(buf + 8192 + warp * 16, d, s2); n += d; if (poll_sleep) __nanosleep(poll_sleep); }
      if (lane == 0) sink[sm] += n;
    }
    return;
  }
  const int qd = lane >> 2, tq = lane & 3;
  if (sm < nprod) {
    for (int it = 0; it < iters; ++it) {
      // wait for the consumers of the previous round
      if (it > 0) { while (*(volatile uint32_t*)ack < (uint32_t)(it * ncons)) {} }
      if (rewrite || it == 0) {
        const int per = 2048 / nprod;
        for (int w = sm * per + lane; w < (sm + 1) * per; w += 32) st1(buf + w, w, rewrite ? it + 1 : 1);
      }
      if (!rewrite && it > 0 && sm == 0 && lane == 0) st1(buf + 4096, 0, it + 1);   // only a go word changes
    }
  } else if (sm < nprod + ncons) {
    long long tot = 0; unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t want = rewrite ? it + 1 : 1;
      uint32_t d, s;
      if (rewrite) { do { ld1(buf, d, s); } while (s != want); }
      else { do { ld1(buf + (it ? 4096 : 0), d, s); } while (s != (uint32_t)(it ? it + 1 : 1)); }
      const long long t0 = clock64();
      uint32_t bad;
      do {
        bad = 0;
        const uint64_t* row = pattern == 0 ? buf + qd * 256 + 2 * tq : buf + lane * 2;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          uint32_t d0, s0, d1, s1, d2, s2, d3, s3;
          if (pattern == 0) { ld2(row + ks * 16, d0, s0, d1, s1); ld2(row + ks * 16 + 8, d2, s2, d3, s3); }
          else { ld2(row + (2 * ks) * 64, d0, s0, d1, s1); ld2(row + (2 * ks + 1) * 64, d2, s2, d3, s3); }
          acc += d0 + d1 + d2 + d3;
          bad |= (s0 ^ want) | (s1 ^ want) | (s2 ^ want) | (s3 ^ want);
        }
      } while (__any_sync(0xffffffffu, bad != 0));
      tot += clock64() - t0;
      __syncwarp();
      if (lane == 0) atomicAdd(ack, 1u);
    }
    if (lane == 0) { sink[sm] = acc; if (sm == nprod) out[0] = tot; }
  }
}
int main() {
  uint64_t* buf; CK(cudaMalloc(&buf, 1 << 20)); uint32_t* ack; CK(cudaMalloc(&ack, 4));
  long long* out; CK(cudaMalloc(&out, 16)); unsigned* sink; CK(cudaMalloc(&sink, 148 * 4));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int iters = 500;
  for (int psleep : {0, 100, 400})
  for (int npoll : {0, 1, 3, 7})
  for (int rewrite : {1})
    for (int pattern : {0})
      for (int ncons : {64}) {
        if (npoll == 0 && psleep) continue;
        CK(cudaMemset(buf, 0, 1 << 20)); CK(cudaMemset(ack, 0, 4));
        k<<<148, 256, 200 * 1024>>>(buf, ack, 16, ncons, iters, pattern, rewrite, out, sink, npoll, psleep);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost));
        printf("%d polling warps per SM (nanosleep %d), %s, %s, %3d SMs reading the same 16 KB: %7.1f cycles per batch (after the sentinel)\n",                npoll, psleep, "rewritten", pattern == 0 ? "8 rows x 64 B per load" : "512 B contiguous per load", ncons, (double)c / iters);
      }
  return 0;
}
