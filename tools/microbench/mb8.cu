// LL-style hops ("stamp in the data"): every 8-byte word of an exchange buffer = (32-bit payload, 32-bit step stamp), written
// with ONE 64-bit store and read with 64-bit loads, so a consumer needs no flag, no fence and no atomic: it polls one sentinel
// word, then reads its whole input and re-reads any word whose stamp is still old.  Same phase structure as mb6.cu.
//   mode 1: LL hops;  mode 2: LL hops + consumer-side fence.acq_rel.gpu per hop;  mode 3: LL + producer __threadfence after the stores
// plus the K/V bulk-copy stream of mb6 in warps 8..15.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1;} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// two (payload, stamp) words per 16-byte volatile load (each 8-byte half is single-copy atomic)
__device__ __forceinline__ void ld_ll2(const void* p, uint32_t& d0, uint32_t& s0, uint32_t& d1, uint32_t& s1) {
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d0), "=r"(s0), "=r"(d1), "=r"(s1) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_ll2(void* p, uint32_t d0, uint32_t s0, uint32_t d1, uint32_t s1) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(d0), "r"(s0), "r"(d1), "r"(s1) : "memory");
}
__device__ __forceinline__ uint2 ld_ll1(const void* p) {
  uint2 r; asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory"); return r;
}

constexpr int NPH = 22;
__constant__ int c_nprod[NPH];
__constant__ int c_off[NPH];
__constant__ int c_words[NPH];               // 8-byte words of phase p's output buffer (multiple of 512)
constexpr int BUF_WORDS = 4096;              // 32 KB per (group, phase, parity)

struct Params {
  uint64_t* act;          // [groups][NPH][2 parity][BUF_WORDS]
  const uint8_t* kv;
  size_t kv_per_warp;
  int sentinel;
  int groups, steps, nsm, mode, stream_warps, stage_bytes, nstage, stream_reps;
  int* err;
  long long* out;
  unsigned long long* sink;
};

__global__ void __launch_bounds__(512, 1) k(Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sm = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) mbar_init(&bars[i][j], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 8) {
    if (p.mode == 0 || warp >= p.groups) return;
    const int g = warp;
    uint64_t* act = p.act + (size_t)g * NPH * 2 * BUF_WORDS;
    unsigned long long acc = 0;
    for (int step = 0; step < p.steps; ++step) {
      for (int ph = 0; ph < NPH; ++ph) {
        const int rel = (sm - c_off[ph] + 2 * p.nsm) % p.nsm;
        if (rel >= c_nprod[ph]) continue;
        const int prev = (ph + NPH - 1) % NPH;
        const int pstep = ph == 0 ? step - 1 : step;          // step in which the input was produced
        uint32_t r = 0;
        if (pstep >= 0) {
          const uint32_t want = (uint32_t)pstep + 1u;
          const uint64_t* src = act + ((size_t)prev * 2 + (pstep & 1)) * BUF_WORDS;
          uint32_t tries = 0;
          if (p.sentinel) {
            while (ld_ll1(src).y != want) {
              if (++tries > (1u << 22)) { *p.err = 1 + ph; return; }
              if ((tries & 1023) == 0 && *(volatile int*)p.err) return;
            }
          }
          const int nb = c_words[prev] / 1024;                // batches of 1024 words (8 KB): 16 loads of 16 B per lane
          for (int b = 0; b < nb; ++b) {
            uint32_t d[32], s[32];
            bool ok;
            do {
#pragma unroll
              for (int j = 0; j < 16; ++j) ld_ll2(src + (size_t)b * 1024 + j * 64 + lane * 2, d[2 * j], s[2 * j], d[2 * j + 1], s[2 * j + 1]);
              ok = true;
#pragma unroll
              for (int j = 0; j < 32; ++j) ok &= (s[j] == want);
              if (++tries > (1u << 22)) { *p.err = 50 + ph; return; }
            } while (!__all_sync(0xffffffffu, ok));
#pragma unroll
            for (int j = 0; j < 32; ++j) r += d[j];
          }
          if (p.mode == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        for (int o = 16; o; o >>= 1) r ^= __shfl_xor_sync(0xffffffffu, r, o);
        acc += r;
        // this unit's slice of the output buffer
        const int nw = c_words[ph], np = c_nprod[ph];
        const int lo = (int)((long long)nw * rel / np) & ~1, hi = rel + 1 == np ? nw : ((int)((long long)nw * (rel + 1) / np) & ~1);
        uint64_t* dst = act + ((size_t)ph * 2 + (step & 1)) * BUF_WORDS;
        const uint32_t stamp = (uint32_t)step + 1u;
        for (int w = lo + lane * 2; w < hi; w += 64) st_ll2(dst + w, r + w, stamp, r ^ w, stamp);
        if (p.mode == 3) __threadfence();
      }
    }
    if (lane == 0) p.sink[sm * 16 + warp] = acc;
    if (sm == 0 && g == 0 && lane == 0) p.out[0] = clock64() - t0;
  } else {
    const int sw = warp - 8;
    if (sw >= p.stream_warps) return;
    uint8_t* ring = smem + (size_t)sw * p.nstage * p.stage_bytes;
    const uint8_t* src = p.kv + ((size_t)sm * p.stream_warps + sw) * p.kv_per_warp;
    const int nreg = (int)(p.kv_per_warp / p.stage_bytes);
    const int ntile = nreg * p.stream_reps;
    const int half = p.stage_bytes / 2;
    if (lane == 0)
      for (int i = 0; i < p.nstage && i < ntile; ++i) {
        expect_tx(&bars[sw][i], p.stage_bytes);
        bulk_load(ring + (size_t)i * p.stage_bytes, src + (size_t)(i % nreg) * p.stage_bytes, half, &bars[sw][i]);
        bulk_load(ring + (size_t)i * p.stage_bytes + half, src + (size_t)(i % nreg) * p.stage_bytes + half, half, &bars[sw][i]);
      }
    uint32_t acc = 0;
    int st = 0; uint32_t phs = 0;
    for (int t = 0; t < ntile; ++t) {
      uint32_t tries = 0;
      while (!try_wait(&bars[sw][st], phs)) { if (++tries > (1u << 22)) { *p.err = 100; return; } }
      const uint8_t* tile = ring + (size_t)st * p.stage_bytes;
      for (int i = lane * 16; i < p.stage_bytes; i += 512) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
      }
      __syncwarp();
      if (lane == 0 && t + p.nstage < ntile) {
        const size_t o = (size_t)((t + p.nstage) % nreg) * p.stage_bytes;
        expect_tx(&bars[sw][st], p.stage_bytes);
        bulk_load(ring + (size_t)st * p.stage_bytes, src + o, half, &bars[sw][st]);
        bulk_load(ring + (size_t)st * p.stage_bytes + half, src + o + half, half, &bars[sw][st]);
      }
      if (++st == p.nstage) { st = 0; phs ^= 1; }
    }
    if (lane == 0) p.sink[sm * 16 + warp] = acc;
    if (lane == 0 && sw == 0 && sm == 0) p.out[1] = clock64() - t0;
  }
}

__global__ void pingpong(uint64_t* buf, int n, long long* out) {
  // CTA 0 writes word 0 and waits for word 16 (another 128-byte line); CTA 1 the opposite
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x;
  if (me > 1) return;
  const long long t0 = clock64();
  for (int i = 1; i <= n; ++i) {
    if (me == 0) {
      asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(buf), "r"(i), "r"(i) : "memory");
      while (ld_ll1(buf + 16).y != (uint32_t)i) {}
    } else {
      while (ld_ll1(buf).y != (uint32_t)i) {}
      asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(buf + 16), "r"(i), "r"(i) : "memory");
    }
  }
  if (me == 0) out[0] = clock64() - t0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("SMs %d, clock %d kHz\n", nsm, clk_khz);
  int nprod[NPH], off[NPH], words[NPH];
  int o = 0;
  for (int l = 0; l < 4; ++l) {
    const int np[5] = {48, 128, 16, 64, 16};
    const int wd[5] = {1024, 4096, 2048, 4096, 2048};     // q+k+v new rows / attention partials (S = 2) / x1 fp32 / hidden bf16x2 / x fp32
    for (int j = 0; j < 5; ++j) { nprod[l * 5 + j] = np[j]; off[l * 5 + j] = o % nsm; words[l * 5 + j] = wd[j]; o += np[j]; }
  }
  nprod[20] = nsm; off[20] = 0; words[20] = 1024;          // head -> tile maxima
  nprod[21] = 8; off[21] = 17; words[21] = 2048;           // sampler -> next step's embedded x
  CK(cudaMemcpyToSymbol(c_nprod, nprod, sizeof(nprod)));
  CK(cudaMemcpyToSymbol(c_off, off, sizeof(off)));
  CK(cudaMemcpyToSymbol(c_words, words, sizeof(words)));
  Params p{};
  const int G = 8;
  const size_t act_bytes = (size_t)G * NPH * 2 * BUF_WORDS * 8;
  CK(cudaMalloc(&p.act, act_bytes));
  const size_t kv_total = (size_t)3 << 30;
  uint8_t* kv; CK(cudaMalloc(&kv, kv_total)); CK(cudaMemset(kv, 3, kv_total));
  p.kv = kv;
  CK(cudaMalloc(&p.err, 4)); CK(cudaMalloc(&p.out, 16)); CK(cudaMalloc(&p.sink, 160 * 16 * 8));
  p.nsm = nsm;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, int groups, int steps, int mode, int stream_warps, int stage_bytes, int nstage, size_t per_warp, int sreps, int sentinel = 1) -> int {
    p.sentinel = sentinel;
    p.groups = groups; p.steps = steps; p.mode = mode; p.stream_warps = stream_warps; p.stage_bytes = stage_bytes; p.nstage = nstage; p.stream_reps = sreps;
    p.kv_per_warp = per_warp / stage_bytes * stage_bytes;
    float best = 1e30f; long long out[2] = {0, 0};
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(p.act, 0, act_bytes)); CK(cudaMemset(p.err, 0, 4)); CK(cudaMemset(p.out, 0, 16));
      CK(cudaEventRecord(e0));
      k<<<nsm, 512, 200 * 1024>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      int err; CK(cudaMemcpy(&err, p.err, 4, cudaMemcpyDeviceToHost));
      if (err) { printf("%s: TIMEOUT code %d\n", name, err); return 0; }
      if (ms < best) { best = ms; CK(cudaMemcpy(out, p.out, 16, cudaMemcpyDeviceToHost)); }
    }
    const double hop_us = mode ? out[0] / (clk_khz * 1e-3) / ((double)steps * NPH) : 0.0;
    const double bytes = (double)nsm * stream_warps * p.kv_per_warp * sreps;
    printf("%-48s kernel %8.3f ms | hop %6.3f us (chain of %d: %6.2f us/step) | stream %7.1f GB/s\n", name, best, hop_us, NPH,
           hop_us * NPH, stream_warps ? bytes / (out[1] / (clk_khz * 1e3)) * 1e-9 : 0.0);
    return 0;
  };
  const size_t MB = 1 << 20;
  {
    CK(cudaMemset(p.act, 0, 4096)); CK(cudaMemset(p.out, 0, 16));
    pingpong<<<148, 32>>>(p.act, 20000, p.out);
    CK(cudaDeviceSynchronize());
    long long o2[2]; CK(cudaMemcpy(o2, p.out, 16, cudaMemcpyDeviceToHost));
    printf("ping-pong (1 word LL, SM 0 <-> SM 1): %.1f cycles = %.3f us per ONE-WAY hop\n", o2[0] / 40000.0, o2[0] / 40000.0 / (clk_khz * 1e-3));
  }
  run("LL hops, 1 group", 1, 400, 1, 0, 4096, 2, 0, 1);
  run("LL hops, 1 group, no sentinel", 1, 400, 1, 0, 4096, 2, 0, 1, 0);
  run("LL hops, 8 groups, no sentinel", 8, 400, 1, 0, 4096, 2, 0, 1, 0);
  run("LL hops, 8 groups", 8, 400, 1, 0, 4096, 2, 0, 1);
  run("LL hops + consumer fence, 8 groups", 8, 400, 2, 0, 4096, 2, 0, 1);
  run("LL hops + producer threadfence, 8 groups", 8, 400, 3, 0, 4096, 2, 0, 1);
  run("LL hops 8 groups + stream 8 x 3 x 8 KB", 8, 400, 1, 8, 8192, 3, 2 * MB, 16);
  run("LL hops 8 groups no sentinel + stream 8 x 3 x 8 KB", 8, 400, 1, 8, 8192, 3, 2 * MB, 16, 0);
  run("LL hops+cons fence 8 groups + stream 8 x 3 x 8 KB", 8, 400, 2, 8, 8192, 3, 2 * MB, 16);
  printf("done\n");
  return 0;
}
