// Instruction-level latencies that bound the decode kernel's short phases: HMMA (mma.sync m16n8k16) dependent-chain latency
// and issue rate, ldmatrix, mbarrier try_wait on a completed phase, named barrier of 256 threads, warp shuffle reduction,
// st.async -> remote mbarrier round trip inside a cluster of 4.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1;} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm(uint32_t addr, uint32_t (&a)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
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

__global__ void __launch_bounds__(288) lat_kernel(long long* out, float* sink) {
  __shared__ __align__(1024) uint8_t tile[32768];
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory"); }
  __syncthreads();
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
  float d[8][4] = {};
  long long t0, t1;
  // (a) dependent chain of 32 HMMAs, one warp
  if (warp == 0) {
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) mma(d[0], a, a[0], a[1]);
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0);
  }
  __syncthreads();
  // (b) 8 independent chains x 8 = 64 HMMAs: one warp alone
  if (warp == 0) {
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) mma(d[c], a, a[0], a[1]);
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0);
  }
  __syncthreads();
  // (c) same, all 8 warps at once (2 per SMSP)
  if (warp < 8) {
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int c = 0; c < 8; ++c) mma(d[c], a, a[0], a[1]);
    t1 = clock64();
    if (lane == 0 && warp == 0) out[2] = (t1 - t0);
  }
  __syncthreads();
  // (d) ldmatrix dependent chain (address depends on the previous result), 16 deep
  if (warp == 0) {
    uint32_t addr = smem_u32(tile) + lane * 16;
    uint32_t r[4];
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) { ldsm(addr, r); addr += (r[0] & 0x10u); }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0);
    d[0][0] += (float)r[1];
  }
  __syncthreads();
  // (e) ldmatrix x8 independent then 8 HMMA (the per-stage body), one warp
  if (warp == 0) {
    const uint32_t base = smem_u32(tile) + lane * 16;
    uint32_t r[8][4];
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) ldsm(base + i * 512, r[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) mma(d[i & 3], r[i], a[0], a[1]);
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0);
  }
  __syncthreads();
  // (f) try_wait / test_wait on a completed phase, 16 in a row
  if (warp == 0) {
    int acc = 0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += try_wait(&bar, 0);
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0);
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += test_wait(&bar, 0);
    t1 = clock64();
    if (lane == 0) out[6] = (t1 - t0);
    d[0][1] += acc;
  }
  __syncthreads();
  // (g) named barrier, 256 threads, 16 in a row
  if (warp < 8) {
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
    t1 = clock64();
    if (threadIdx.x == 0) out[7] = (t1 - t0);
  }
  __syncthreads();
  // (h) 5-step shuffle reduction x 8 dependent
  if (warp == 0) {
    float v = lane;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    t1 = clock64();
    if (lane == 0) out[8] = (t1 - t0);
    d[0][2] += v;
  }
  __syncthreads();
  // (i) LDS.32 dependent chain 16
  if (warp == 0) {
    uint32_t off = lane * 4, v = 0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; ++i) { v = *reinterpret_cast<volatile uint32_t*>(tile + off); off = (off + (v & 4)) & 32767; }
    t1 = clock64();
    if (lane == 0) out[9] = (t1 - t0);
    d[0][3] += v;
  }
  __syncthreads();
  // (j) erff x 8 dependent, expf x8, exp2f x 8
  if (warp == 0) {
    float v = 0.3f + lane * 0.01f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) v = erff(v);
    t1 = clock64();
    if (lane == 0) out[10] = (t1 - t0);
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 8; ++i) v = exp2f(v) - 1.0f;
    t1 = clock64();
    if (lane == 0) out[11] = (t1 - t0);
    d[1][0] += v;
  }
  float s = 0;
  for (int c = 0; c < 8; ++c) for (int e = 0; e < 4; ++e) s += d[c][e];
  if (s == 1234.5f) sink[0] = s;
}


// the GEMM stage body of the decode kernel with all 8 warps: 8 LDSM.x4 + 8 HMMA per warp per 32 KB stage, 16 stages
__global__ void __launch_bounds__(288) stage_kernel(long long* out, float* sink, int mode) {
  extern __shared__ __align__(1024) uint8_t ring[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * 8192; i += blockDim.x) reinterpret_cast<uint32_t*>(ring)[i] = 0x3f803f80u;
  __syncthreads();
  if (warp >= 8) return;
  const int lrow = ((lane >> 3) & 1) * 8 + (lane & 7), lchunk = lane >> 4;
  uint32_t foff[2][4];
  for (int t = 0; t < 2; ++t) {
    const int i = warp * 32 + t * 16 + lrow;
    for (int ks = 0; ks < 4; ++ks) foff[t][ks] = (i >> 7) * 16384 + (i & 127) * 128 + (((ks * 2 + lchunk) ^ (i & 7)) << 4);
  }
  float acc[2][4] = {}, acc2[2][4] = {};
  const uint32_t base = smem_u32(ring);
  uint32_t b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 4
    for (int st = 0; st < 16; ++st) {
      const uint32_t sb = base + (st & 3) * 32768;
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a[4];
          ldsm(sb + foff[t][ks], a);
          if (ks & 1) mma(acc2[t], a, b0, b1); else mma(acc[t], a, b0, b1);
        }
    }
  } else if (mode == 3 || mode == 4) {   // mode 0 + the kernel's per-stage synchronisation (try_wait on a completed phase, syncwarp, arrive)
    __shared__ uint64_t fullb[4], emptyb[4];
    if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fullb[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&emptyb[i])));
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullb[i])) : "memory");
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    t0 = clock64();
    const int per = mode == 3 ? 1 : 4;          // stages per synchronisation
#pragma unroll 1
    for (int st0 = 0; st0 < 16; st0 += per) {
      for (int q = 0; q < per; ++q) while (!try_wait(&fullb[(st0 + q) & 3], 0)) {}
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q >= per) break;
        const uint32_t sb = base + ((st0 + q) & 3) * 32768;
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t a[4];
            ldsm(sb + foff[t][ks], a);
            if (ks & 1) mma(acc2[t], a, b0, b1); else mma(acc[t], a, b0, b1);
          }
      }
      __syncwarp();
      if (lane == 0) for (int q = 0; q < per; ++q) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&emptyb[(st0 + q) & 3])) : "memory");
    }
  } else if (mode == 5) {   // explicit software pipeline: ldmatrix of stage i + 1 before the MMAs of stage i, per-stage sync kept
    __shared__ uint64_t fullb[4], emptyb[4];
    if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&fullb[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(smem_u32(&emptyb[i])));
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullb[i])) : "memory");
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    t0 = clock64();
    uint32_t fa[2][2][4][4];
    while (!try_wait(&fullb[0], 0)) {}
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm(base + foff[t][ks], fa[0][t][ks]);
#pragma unroll
    for (int st = 0; st < 16; ++st) {
      if (st + 1 < 16) {
        while (!try_wait(&fullb[(st + 1) & 3], 0)) {}
        const uint32_t sb = base + ((st + 1) & 3) * 32768;
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) ldsm(sb + foff[t][ks], fa[(st + 1) & 1][t][ks]);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int t = 0; t < 2; ++t) { if (ks & 1) mma(acc2[t], fa[st & 1][t][ks], b0, b1); else mma(acc[t], fa[st & 1][t][ks], b0, b1); }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&emptyb[st & 3])) : "memory");
    }
  } else if (mode == 1) {   // LDSM only
    uint32_t x = 0;
#pragma unroll 4
    for (int st = 0; st < 16; ++st) {
      const uint32_t sb = base + (st & 3) * 32768;
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { uint32_t a[4]; ldsm(sb + foff[t][ks], a); x ^= a[0] ^ a[3]; }
    }
    acc[0][0] += x;
  } else {                  // HMMA only
    uint32_t a[4] = {b0, b0, b0, b0};
#pragma unroll 4
    for (int st = 0; st < 16; ++st)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { if (ks & 1) mma(acc2[t], a, b0, b1); else mma(acc[t], a, b0, b1); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[mode] = t1 - t0;
  float s = 0;
  for (int t = 0; t < 2; ++t) for (int e = 0; e < 4; ++e) s += acc[t][e] + acc2[t][e];
  if (s == 1234.5f) sink[0] = s;
}

// st.async ping-pong between CTA 0 and CTA 1 of a cluster: round-trip latency
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(32) pingpong_kernel(long long* out, int rounds) {
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0 && rank < 2) {
    const uint32_t peer = rank ^ 1;
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(&slot)), "r"(peer));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(&bar)), "r"(peer));
    long long t0 = clock64();
    for (int i = 0; i < rounds; ++i) {
      if (rank == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(i), "r"(rb) : "memory");
        while (!try_wait(&bar, i & 1)) {}
      } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(smem_u32(&bar)) : "memory");
        while (!try_wait(&bar, i & 1)) {}
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(i), "r"(rb) : "memory");
      }
    }
    if (rank == 0) out[0] = (clock64() - t0) / rounds;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  long long* out; float* sink; CK(cudaMalloc(&out, 64 * 8)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(out, 0, 64 * 8));
  lat_kernel<<<1, 288>>>(out, sink); CK(cudaDeviceSynchronize());
  lat_kernel<<<1, 288>>>(out, sink); CK(cudaDeviceSynchronize());
  long long h[16]; CK(cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost));
  printf("HMMA dependent chain x32: %lld cyc -> %.1f per HMMA\n", h[0], h[0] / 32.0);
  printf("HMMA 8 chains x8 (64), 1 warp: %lld cyc -> %.1f per HMMA\n", h[1], h[1] / 64.0);
  printf("HMMA 8 chains x8 (64) per warp, 8 warps: %lld cyc -> %.1f per HMMA per warp\n", h[2], h[2] / 64.0);
  printf("LDSM.x4 dependent x16: %lld -> %.1f each\n", h[3], h[3] / 16.0);
  printf("stage body (8 LDSM + 8 HMMA, 4 chains): %lld cyc\n", h[4]);
  printf("mbarrier try_wait (done) x16: %lld -> %.1f each; test_wait x16: %lld -> %.1f each\n", h[5], h[5] / 16.0, h[6], h[6] / 16.0);
  printf("bar.sync 256 thr x16: %lld -> %.1f each\n", h[7], h[7] / 16.0);
  printf("warp shuffle allreduce (5 steps) x8: %lld -> %.1f each\n", h[8], h[8] / 8.0);
  printf("LDS dependent x16: %lld -> %.1f each\n", h[9], h[9] / 16.0);
  printf("erff x8: %lld -> %.1f each; exp2f+sub x8: %lld -> %.1f each\n", h[10], h[10] / 8.0, h[11], h[11] / 8.0);
  CK(cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768));
  for (int rep = 0; rep < 2; ++rep) for (int mode = 0; mode < 6; ++mode) { stage_kernel<<<1, 288, 4 * 32768>>>(out, sink, mode); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, out, 6 * 8, cudaMemcpyDeviceToHost));
  printf("GEMM stage body, 8 warps, 16 stages: LDSM+HMMA %.1f cyc/stage, LDSM only %.1f, HMMA only %.1f, with per-stage mbarrier sync %.1f, sync per 4 stages %.1f, explicit pipeline + per-stage sync %.1f\n", h[0] / 16.0, h[1] / 16.0, h[2] / 16.0, h[3] / 16.0, h[4] / 16.0, h[5] / 16.0);
  CK(cudaMemset(out, 0, 64 * 8));
  pingpong_kernel<<<4, 32>>>(out, 1000); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
  printf("st.async ping-pong round trip (2 hops): %lld cyc -> %.1f per hop\n", h[0], h[0] / 2.0);
  return 0;
}
