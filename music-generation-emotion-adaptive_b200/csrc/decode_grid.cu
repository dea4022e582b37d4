// Grid-synchronous persistent decode kernel for sm_100a: the whole generation loop of reference api_cache.py:159-184
// (embedding -> L x [LN1, in_proj, attention over the cache, out_proj, +x, LN2, mlp.0, GELU, mlp.2, +x] -> head -> /T ->
// top-k -> softmax -> multinomial -> EOS) in ONE cooperative launch over all SMs, for up to 64 sequences that move through
// the step in lock-step.
//
// Decomposition (the opposite of decode_mega.cu, where a 4-CTA cluster owns a few sequences and streams all weights):
//   * the batch is the N dimension of every contraction (8 sequences = one n8 tile of mma.sync.m16n8k16 = one warp), a
//     work item is 16 rows of one weight matrix, so every weight byte is read ONCE per step by ONE SM -- straight from
//     L1 / L2 into A fragments (fragment-major packing, ld.global.nc: the per-SM weight set of a step is ~100 KB for the
//     train_large geometry and stays L1-resident from step to step);
//   * a step is 5 L + 2 phases separated by a grid barrier (one counter, release by one thread per CTA behind a
//     __syncthreads, polled by one lane per warp); activations between phases live in L2 and are read with L1-bypassing
//     loads; LayerNorm is applied by the consumer while it builds its B fragments (no LayerNorm phase), the residual adds sit
//     in the out_proj / mlp.2 epilogues, the embedding is recomputed where it is needed;
//   * attention: the (sequence, head, key range) units of a step are dealt evenly over ALL warps of ALL SMs by total K/V
//     bytes (balanced partition recomputed every step from the cache lengths), each unit is the warp-level tensor-core
//     flash-decoding of attn_tc.cuh; partial results are merged by the out_proj consumers while they load their operand;
//   * sampling: one CTA per sequence, logits in registers, threshold from the per-thread maxima, exact ranking of the
//     (small) candidate superset, Philox draw; general path = sampler.cuh on a global scratch row.
// Any d_model in {256, 512} with d_ff = 4 d_model and head_dim 32 / 64 (the paper's train_large2 geometry included), any
// cache length (config 4 uses all SMs), B <= 64.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "attn_tc.cuh"
#include "decode_grid.cuh"
#include "mg_engine.h"
#include "ptx.cuh"
#include "sampler.cuh"

namespace mg {
namespace grid {

namespace {

using namespace attn;

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kCandCap = 1024;         // candidate superset capacity of the sampler's fast path
constexpr unsigned long long kWatchdogNs = 2000000000ull;   // a barrier that is not reached within 2 s aborts the launch

// ---- L1-bypassing loads of data other SMs rewrite during the launch (ld.volatile: the fastest flavour measured, mb9) ----
__device__ __forceinline__ uint4 ldv4u(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float4 ldv4f(const float* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float2 ldv2f(const float* p) {
  float2 r;
  asm volatile("ld.volatile.global.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ldvf(const float* p) {
  float r;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ int ldvi(const int32_t* p) {
  int r;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ unsigned ldvu(const unsigned* p) {
  unsigned r;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ int ldvb(const uint8_t* p) {
  unsigned r;
  asm volatile("ld.volatile.global.u8 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return static_cast<int>(r);
}
// weights: immutable for the whole launch, read-only path, allocate in L1
__device__ __forceinline__ uint4 ldw16(const uint8_t* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// cache append: L2-only store (no copy of the line is left in the appending SM's L1)
__device__ __forceinline__ void st_kv(bf16* p, bf16 v) {
  asm volatile("st.global.cg.u16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<const unsigned short*>(&v)) : "memory");
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

struct GridSmem {
  alignas(16) float4 red[8][32];                   // k-split partial accumulators [warp][lane]
  float qs[8][64];                     // per-warp attention scratch: query (fp32, log2-scaled), new K / V rows, partial result
  bf16 kn[8][64];
  bf16 vn[8][64];
  float part[8][68];
  int nws[kMaxSeqs];                   // attention workers per (sequence, head) in this step (0: finished)
  int istart[kMaxSeqs];                // first attention unit of the sequence
  int len[kMaxSeqs];
  int fin[kMaxSeqs];
  int total_units;
  int upc;                             // attention units per CTA in this step (consecutive units = consecutive warps of one CTA)
  // sampler
  alignas(16) float mx[kThreads];
  alignas(16) uint2 cand[kCandCap];
  uint2 sorted[kThreads];
  int wtot[8];
  float tau;
  int flag;
  int result;
  int next_tok;                        // sampler -> embedding producer of the same CTA (-1: the sequence has just finished)
  SampleSmem ss;
  GridItem items[kMaxItems];
};

// Tile layout (grid_pack_kernel): a tile = 16 rows x K of a matrix; per pair of k-steps j (32 columns) two 512-byte blocks, block
// e = the A fragments of MMA 2 j + e in lane order: lane (g, t) holds {W[g][c], W[g + 8][c], W[g][c + 2], W[g + 8][c + 2]} (bf16
// pairs), c = 32 j + 8 t + 4 e.  The B fragments use the same permutation of the contraction index: lane (g, t) of sequence g
// holds columns 32 j + 8 t .. + 7 as ONE 16-byte load, (x, y) feed MMA 2 j, (z, w) MMA 2 j + 1.
template <int NP>
__device__ __forceinline__ void mma_tile(const uint8_t* __restrict__ tile, int lane, const uint4 (&bq)[NP], int np, float (&acc)[4]) {
  // groups of G k-step pairs; the loads of group g + 1 are issued before the MMAs of group g (asm volatile keeps this order: left to
  // itself the compiler sinks every load next to its MMA to save registers, and the tile costs one L1 / L2 round trip per pair)
  constexpr int G = NP >= 4 ? 4 : NP, NG = NP / G;
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  const uint8_t* tp = tile + lane * 16;
  uint4 wa[2][2 * G];
#pragma unroll
  for (int i = 0; i < G; ++i) { wa[0][2 * i] = ldw16(tp + i * 1024); wa[0][2 * i + 1] = ldw16(tp + i * 1024 + 512); }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g + 1 < NG) {
#pragma unroll
      for (int i = 0; i < G; ++i) {
        wa[(g + 1) & 1][2 * i] = ldw16(tp + ((g + 1) * G + i) * 1024);
        wa[(g + 1) & 1][2 * i + 1] = ldw16(tp + ((g + 1) * G + i) * 1024 + 512);
      }
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
      mma16816(a0, wa[g & 1][2 * i], bq[g * G + i].x, bq[g * G + i].y);
      mma16816(a1, wa[g & 1][2 * i + 1], bq[g * G + i].z, bq[g * G + i].w);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] = a0[e] + a1[e];
}

__device__ __forceinline__ uint32_t pk(float a, float b) { return pack_bf16(a, b); }

template <int DM, int HD>
__global__ void __launch_bounds__(kThreads, 1) decode_grid_kernel(const GridParams p) {
  constexpr int NPD = DM / 32;           // k-step pairs across d_model
  constexpr int DFF = 4 * DM;
  constexpr int PS = HD + 4;             // floats per attention partial
  constexpr int NT16 = DM / 16;          // 16-row tiles across d_model = LayerNorm partial statistics per sequence
  __shared__ GridSmem sm;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, qd = lane >> 2, tq = lane & 3;
  const int cta = blockIdx.x, n_cta = p.n_cta, B = p.B, H = p.H, L = p.L;
  const int n_items = p.n_items[cta];
  for (int i = tid; i < n_items; i += kThreads) sm.items[i] = p.items[static_cast<size_t>(cta) * kMaxItems + i];
  const SampleParams sp = *p.sp;
  const float scale_log2 = kLog2e / sqrtf(static_cast<float>(HD));
  if (tid == 0) sm.flag = 1;
  const int n_active0 = __syncthreads_count(tid < B && p.st.finished[tid] == 0);
  unsigned* const bar = p.ctrl;
  unsigned epoch = 0;                    // grid barriers passed so far
  bool alive = true;
  unsigned long long* tr = nullptr;     // fine trace (clock64) of one dense phase of CTA 0: p.prof[96..]
  auto trace = [&]() { if (tr) *tr++ = clock64(); };

  // ---- grid barrier: arrive (release by one thread behind the CTA barrier) / wait (one polling lane per warp) ----
  // Grid barrier: one release-only RED per CTA on a monotonic counter (behind the CTA barrier), ONE polling thread per CTA.
  // Measured alternatives (profiles/r2n_grid_barrier_ab.txt): a poller per warp 108 -> 132 us per step (the arrivals queue behind the
  // polls in the counter's L2 slice); last arriver publishes the epoch in 16 flag lines that the CTAs poll instead 108 -> 114 (one more
  // L2 hop); __threadfence() + atomicAdd invalidates this SM's L1 (MEMBAR.SC + CCTL.IVALL), i.e. the weight tiles it keeps there.
  auto arrive = [&]() {
    trace();
    __syncthreads();
    trace();
    ++epoch;
    if (tid == 0) {
      if (p.fence_mode == 1) { __threadfence(); atomicAdd(bar, 1u); }
      else asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    }
    trace();
  };
  auto wait = [&]() {
    if (tid == 0) {
      const unsigned target = epoch * static_cast<unsigned>(n_cta);
      unsigned spins = 0;
      unsigned long long t0 = 0;
      int ok = 1;
      while (ldvu(bar) < target) {
        if ((++spins & 0xfffu) == 0) {
          const unsigned long long now = ptx::global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > kWatchdogNs || ldvu(p.ctrl + 2) != 0) { ok = 0; atomicExch(p.ctrl + 2, 1u); break; }
        }
      }
      if (!ok) sm.flag = 0;
      trace();
    }
    __syncthreads();
    if (!sm.flag) alive = false;
  };
  const bool prof_on = p.prof != nullptr && cta == 0 && tid == 0;
  int prof_i = 0;

  const int s_own = 8 * warp + qd;                        // sequence of this lane's B fragments when a warp owns n-tile `warp`
  const int sa = 8 * warp + 2 * tq, sb = sa + 1;          // sequences of this lane's accumulators (same n-tile)

  // embedding x = tok_emb[tok] + pos_emb[0] (api_cache.py:99 with T == 1) of sequence b by its owner CTA: fp32 + bf16 copies and the
  // per-16-feature LayerNorm statistics, exactly what an mlp.2 epilogue leaves for the next block
  auto embed = [&](int b, int tok) {
    const bf16* te = p.tok_emb + static_cast<size_t>(tok) * DM;
    for (int f = tid; f < DM; f += kThreads) {
      const float v = __bfloat162float(te[f]) + __bfloat162float(p.pos_emb[f]);
      p.x[static_cast<size_t>(b) * DM + f] = v;
      p.xb[static_cast<size_t>(b) * DM + f] = __float2bfloat16_rn(v);
      float a = v, q = v * v;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if ((lane & 15) == 0) *reinterpret_cast<float2*>(p.sx + (static_cast<size_t>(b) * NT16 + (f >> 4)) * 2) = make_float2(a, q);
    }
  };
  if (cta < B && p.st.finished[cta] == 0) embed(cta, min(max(p.st.cur_tok[cta], 0), p.V - 1));
  arrive(); wait();

  for (int step = 0; step < p.n_steps && alive; ++step) {
    auto stamp = [&]() { if (prof_on && step == p.prof_step) p.prof[prof_i++] = ptx::global_timer_ns(); };
    // ---------------- per-step state + attention partition ----------------
    if (tid < kMaxSeqs) {
      const bool in = tid < B;
      sm.len[tid] = in ? ldvi(p.st.lens + tid) : 0;
      sm.fin[tid] = in ? ldvb(p.st.finished + tid) : 1;
    }
    __syncthreads();
    if (warp == 0) {
      const int nb0 = sm.fin[lane] ? 0 : (sm.len[lane] + 31) >> 5, nb1 = sm.fin[lane + 32] ? 0 : (sm.len[lane + 32] + 31) >> 5;
      const bool act0 = !sm.fin[lane], act1 = !sm.fin[lane + 32];
      int tb = nb0 + nb1, mxb = max(nb0, nb1);
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        tb += __shfl_xor_sync(0xffffffffu, tb, o);
        mxb = max(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
      }
      const int NW = n_cta * 8;
      int c = max(1, max((tb * H + NW - 1) / NW, (mxb + kMaxSplits - 1) / kMaxSplits));
      int w0 = 0, w1 = 0, tot = 0;
      for (;; ++c) {
        w0 = act0 ? max(1, (nb0 + c - 1) / c) : 0;
        w1 = act1 ? max(1, (nb1 + c - 1) / c) : 0;
        tot = w0 + w1;
#pragma unroll
        for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (tot * H <= NW || c > 4096) break;
      }
      int i0 = w0 * H, i1 = w1 * H;                       // inclusive scans over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
        if (lane >= o) { i0 += t0; i1 += t1; }
      }
      const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
      sm.nws[lane] = w0; sm.nws[lane + 32] = w1;
      sm.istart[lane] = i0 - w0 * H; sm.istart[lane + 32] = tot0 + i1 - w1 * H;
      if (lane == 0) { sm.total_units = tot * H; sm.upc = max(1, (tot * H + n_cta - 1) / n_cta); }
    }
    __syncthreads();
    // this warp's attention unit of the step (the same for every layer).  Units are numbered sequence-major, head, key range
    // fastest, and dealt in runs of `upc` consecutive units per CTA: the key ranges of one (sequence, head) sit in neighbouring warps
    // of one CTA (two at a run boundary) and are merged in shared memory before anything is written
    int at_b = -1, at_h = 0, at_wi = 0, at_nws = 1, at_len = 0, at_run = 0, at_slot = 0;
    const int upc = sm.upc;
    {
      const int gw = cta * upc + warp;
      const bool mine = warp < upc && gw < sm.total_units;
      const bool in0 = mine && sm.nws[lane] > 0 && sm.istart[lane] <= gw && gw < sm.istart[lane] + sm.nws[lane] * H;
      const bool in1 = mine && sm.nws[lane + 32] > 0 && sm.istart[lane + 32] <= gw && gw < sm.istart[lane + 32] + sm.nws[lane + 32] * H;
      const unsigned m0 = __ballot_sync(0xffffffffu, in0), m1 = __ballot_sync(0xffffffffu, in1);
      at_b = m0 ? __ffs(m0) - 1 : (m1 ? 32 + __ffs(m1) - 1 : -1);
      if (at_b >= 0) {
        const int r = gw - sm.istart[at_b];
        at_nws = sm.nws[at_b];
        at_h = r / at_nws;
        at_wi = r - at_h * at_nws;
        at_len = sm.len[at_b];
        // the first warp of a run of key ranges of the same (sequence, head) inside this CTA merges the run
        if (warp == 0 || at_wi == 0) at_run = min(at_nws - at_wi, min(upc - warp, sm.total_units - gw));
        at_slot = cta - (gw - at_wi) / upc;              // partial index = CTAs since the one that holds key range 0
      }
    }
    stamp();
    int it = 0;
    const int dslot = !p.dbg_logits ? -1 : (p.dbg_slot ? p.dbg_slot[step] : step);

    // ---------------- dense phases whose operand is a whole (LayerNorm-ed) residual row: in_proj, mlp.0, head ----------------
    auto phase_rows = [&](int ph, auto kind_c, int l) {
      constexpr int kind = decltype(kind_c)::value;
      if (!(it < n_items && sm.items[it].phase == ph)) return;
      const bool wactive = 8 * warp < B;
      uint4 bq[NPD];
      float mean_a = 0.f, rstd_a = 1.f, mean_b = 0.f, rstd_b = 1.f;
      if (prof_on && step == p.prof_step && kind == K_MLP1 && l == 1) tr = p.prof + 96;
      trace();
      if (wactive) {
        // operand = the bf16 copy of the residual stream its producer left next to the fp32 one, used AS IS: LayerNorm is folded
        // into the packed weights (GridLayer) and finished in the epilogue from the row statistics
        const int sq = min(s_own, B - 1);
        const bf16* row = (kind == K_MLP1 ? p.x1b : p.xb) + static_cast<size_t>(sq) * DM + 8 * tq;
#pragma unroll
        for (int j = 0; j < NPD; ++j) bq[j] = ldv4u(row + 32 * j);
        if constexpr (kind != K_HEAD) {
          // row statistics of this lane's two accumulator sequences from the producer's per-tile (sum, sum of squares): the eight
          // lanes of a column group read different tiles
          const float* stp = (kind == K_MLP1 ? p.sx1 : p.sx);
          const float* st_a = stp + static_cast<size_t>(min(sa, B - 1)) * (2 * NT16) + 2 * qd;
          const float* st_b = stp + static_cast<size_t>(min(sb, B - 1)) * (2 * NT16) + 2 * qd;
          float sum_a = 0.f, ssq_a = 0.f, sum_b = 0.f, ssq_b = 0.f;
#pragma unroll
          for (int i = 0; i < NT16 / 8; ++i) {
            const float2 ta = ldv2f(st_a + 16 * i), tb = ldv2f(st_b + 16 * i);
            sum_a += ta.x; ssq_a += ta.y; sum_b += tb.x; ssq_b += tb.y;
          }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            sum_a += __shfl_xor_sync(0xffffffffu, sum_a, o); ssq_a += __shfl_xor_sync(0xffffffffu, ssq_a, o);
            sum_b += __shfl_xor_sync(0xffffffffu, sum_b, o); ssq_b += __shfl_xor_sync(0xffffffffu, ssq_b, o);
          }
          mean_a = sum_a * (1.0f / DM); mean_b = sum_b * (1.0f / DM);
          rstd_a = rsqrtf(fmaxf(ssq_a * (1.0f / DM) - mean_a * mean_a, 0.f) + 1e-5f);
          rstd_b = rsqrtf(fmaxf(ssq_b * (1.0f / DM) - mean_b * mean_b, 0.f) + 1e-5f);
          trace();
        }
      }
      trace();
      while (it < n_items && sm.items[it].phase == ph) {
        const int rt = sm.items[it].row_tile;
        ++it;
        if (!wactive) continue;
        const int ra = 16 * rt + qd, rb = ra + 8;
        float acc[4];
        if constexpr (kind == K_QKV) {
          const GridLayer& lw = p.layers[l];
          const float ca = __ldg(lw.c_in + ra), cb = __ldg(lw.c_in + rb), da = __ldg(lw.d_in + ra), db = __ldg(lw.d_in + rb);
          mma_tile<NPD>(p.packed + lw.w_in + static_cast<size_t>(rt) * (16 * DM * 2), lane, bq, NPD, acc);
          const int part = ra / DM, fa = ra - part * DM, fb = fa + 8;        // a tile never straddles q | k | v
          const float va[2] = {fmaf(rstd_a, acc[0] - mean_a * ca, da), fmaf(rstd_b, acc[1] - mean_b * ca, da)};
          const float vb[2] = {fmaf(rstd_a, acc[2] - mean_a * cb, db), fmaf(rstd_b, acc[3] - mean_b * cb, db)};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int s = sa + e;
            if (s >= B) continue;
            if (part == 0) {
              p.q[static_cast<size_t>(s) * DM + fa] = va[e] * scale_log2;
              p.q[static_cast<size_t>(s) * DM + fb] = vb[e] * scale_log2;
            } else {
              const bf16 xa = __float2bfloat16_rn(va[e]), xb = __float2bfloat16_rn(vb[e]);
              bf16* nw = part == 1 ? p.knew : p.vnew;
              nw[static_cast<size_t>(s) * DM + fa] = xa;
              nw[static_cast<size_t>(s) * DM + fb] = xb;
              if (!sm.fin[s]) {                                         // append to the cache (api_cache.py:66-67)
                const int key = sm.len[s], hh = fa / HD, d0 = fa - hh * HD, d1 = d0 + 8;   // both rows lie in the same head
                const size_t hb = (static_cast<size_t>(s) * H + hh) * p.Tvt * HD;
                if (part == 1) {
                  st_kv(lw.kh + hb + static_cast<size_t>(key) * HD + d0, xa);
                  st_kv(lw.kh + hb + static_cast<size_t>(key) * HD + d1, xb);
                } else {
                  const int ki = key & 31, pos = 8 * ((ki >> 1) & 3) + 2 * (ki >> 3) + (ki & 1);      // see attn_tc
                  bf16* vb2 = lw.vt + hb + static_cast<size_t>(key >> 5) * (HD * 32) + pos;
                  st_kv(vb2 + d0 * 32, xa);
                  st_kv(vb2 + d1 * 32, xb);
                }
              }
            }
          }
        } else if constexpr (kind == K_MLP1) {
          const GridLayer& lw = p.layers[l];
          const float ca = __ldg(lw.c_1 + ra), cb = __ldg(lw.c_1 + rb), da = __ldg(lw.d_1 + ra), db = __ldg(lw.d_1 + rb);
          mma_tile<NPD>(p.packed + lw.w1 + static_cast<size_t>(rt) * (16 * DM * 2), lane, bq, NPD, acc);
          if (tr) { asm volatile("" ::"f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3])); trace(); }
          if (sa < B) {
            p.h[static_cast<size_t>(sa) * DFF + ra] = __float2bfloat16_rn(gelu_erf_f(fmaf(rstd_a, acc[0] - mean_a * ca, da)));
            p.h[static_cast<size_t>(sa) * DFF + rb] = __float2bfloat16_rn(gelu_erf_f(fmaf(rstd_a, acc[2] - mean_a * cb, db)));
          }
          if (sb < B) {
            p.h[static_cast<size_t>(sb) * DFF + ra] = __float2bfloat16_rn(gelu_erf_f(fmaf(rstd_b, acc[1] - mean_b * ca, da)));
            p.h[static_cast<size_t>(sb) * DFF + rb] = __float2bfloat16_rn(gelu_erf_f(fmaf(rstd_b, acc[3] - mean_b * cb, db)));
          }
        } else {
          const float ba = ra < p.V ? __ldg(p.head_b + ra) : 0.f, bb = rb < p.V ? __ldg(p.head_b + rb) : 0.f;
          mma_tile<NPD>(p.packed + p.w_head + static_cast<size_t>(rt) * (16 * DM * 2), lane, bq, NPD, acc);
          const float va[2] = {acc[0] + ba, acc[1] + ba}, vb[2] = {acc[2] + bb, acc[3] + bb};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int s = sa + e;
            if (s >= B) continue;
            if (ra < p.V) p.logits[static_cast<size_t>(s) * p.ldl + ra] = va[e];
            if (rb < p.V) p.logits[static_cast<size_t>(s) * p.ldl + rb] = vb[e];
            if (dslot >= 0) {
              float* dl = p.dbg_logits + (static_cast<size_t>(dslot) * B + s) * p.V;
              if (ra < p.V) dl[ra] = va[e];
              if (rb < p.V) dl[rb] = vb[e];
            }
          }
        }
      }
    };

    // ---------------- dense phases with k-splits across the warps of a CTA: out_proj (operand = merged attention partials) and
    //                  mlp.2 (operand = bf16 hidden activations); both add the residual stream in their epilogue ----------------
    auto phase_split = [&](int ph, auto kind_c, int l) {
      constexpr int kind = decltype(kind_c)::value;
      const int TN = p.tn[kind], KS = p.ks[kind];
      const int ntl = warp % TN, ksp = warp / TN;
      const GridLayer& lw = p.layers[l];
      while (it < n_items && sm.items[it].phase == ph) {
        const int rt = sm.items[it].row_tile, ntile = sm.items[it].group * TN + ntl;
        ++it;
        const bool wactive = 8 * ntile < B;
        const int sq = min(8 * ntile + qd, B - 1);
        const int s0 = 8 * ntile + 2 * tq, s1 = s0 + 1;
        const int ra = 16 * rt + qd, rb = ra + 8;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float res[4] = {0.f, 0.f, 0.f, 0.f}, ba = 0.f, bb = 0.f;
        if (wactive && ksp == 0) {
          // residual + bias of this lane's four outputs, requested ahead of the operand
          const float* bias = kind == K_OUT ? lw.b_out : lw.b2;
          ba = __ldg(bias + ra); bb = __ldg(bias + rb);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int s = min(s0 + e, B - 1);
            const float* xr = (kind == K_OUT ? p.x : p.x1) + static_cast<size_t>(s) * DM;
            res[e] = ldvf(xr + ra);
            res[2 + e] = ldvf(xr + rb);
          }
        }
        auto contract = [&](auto npw_c) {
          constexpr int npw = decltype(npw_c)::value;          // k-step pairs of this warp: compile-time, so that all loads are batched
          uint4 bq[npw];
          if constexpr (kind == K_MLP2) {
            const bf16* hrow = p.h + static_cast<size_t>(sq) * DFF + ksp * npw * 32 + 8 * tq;
#pragma unroll
            for (int j = 0; j < npw; ++j) bq[j] = ldv4u(hrow + 32 * j);
            mma_tile<npw>(p.packed + lw.w2 + static_cast<size_t>(rt) * (16 * DFF * 2) + static_cast<size_t>(ksp) * npw * 1024, lane, bq, npw, acc);
          } else {
            const int nws_q = sm.nws[sq];
            // partial results of the attention phase: (sequence, head) -> nsp partials (one per CTA its key ranges ran on)
            int nsp_j[npw];
            const float* pb_j[npw];
            int fo_j[npw];
            int nsp_max = 0;
#pragma unroll
            for (int j = 0; j < npw; ++j) {
              const int fb0 = 32 * (ksp * npw + j), hh = fb0 / HD;
              fo_j[j] = fb0 - hh * HD + 8 * tq;
              pb_j[j] = p.part + (static_cast<size_t>(sq) * H + hh) * kMaxSplits * PS;
              const int g0 = sm.istart[sq] + hh * nws_q;
              nsp_j[j] = nws_q == 0 ? 0 : (g0 + nws_q - 1) / upc - g0 / upc + 1;
              nsp_max = max(nsp_max, nsp_j[j]);
            }
            auto merge_pack = [&](const float2 (&ml)[4], const float4 (&oa)[4], const float4 (&ob)[4], float (&o)[8], float& M, float& Ls) {
              float Mn = M;
#pragma unroll
              for (int u = 0; u < 4; ++u) Mn = fmaxf(Mn, ml[u].x);
              const float corr = fast_exp2(M - Mn);                    // M == -inf: exp2(-inf) = 0 (Mn is finite: worker 0 folds the new token)
              Ls *= corr;
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] *= corr;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float w = fast_exp2(ml[u].x - Mn);
                Ls = fmaf(ml[u].y, w, Ls);
                o[0] = fmaf(oa[u].x, w, o[0]); o[1] = fmaf(oa[u].y, w, o[1]); o[2] = fmaf(oa[u].z, w, o[2]); o[3] = fmaf(oa[u].w, w, o[3]);
                o[4] = fmaf(ob[u].x, w, o[4]); o[5] = fmaf(ob[u].y, w, o[5]); o[6] = fmaf(ob[u].z, w, o[6]); o[7] = fmaf(ob[u].w, w, o[7]);
              }
              M = Mn;
            };
            if (npw <= 4 && __all_sync(0xffffffffu, nsp_max <= 2)) {
              // the usual case (a run of key ranges is cut by at most one CTA boundary): ALL loads of the warp's operand first, one L2
              // round trip instead of one per k-step pair
              float2 ml[npw][2];
              float4 oa[npw][2], ob[npw][2];
#pragma unroll
              for (int j = 0; j < npw; ++j)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  ml[j][u] = make_float2(-INFINITY, 0.f);
                  oa[j][u] = ob[j][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (u < nsp_j[j]) {
                    const float* pp = pb_j[j] + static_cast<size_t>(u) * PS;
                    ml[j][u] = ldv2f(pp + HD);
                    oa[j][u] = ldv4f(pp + fo_j[j]);
                    ob[j][u] = ldv4f(pp + fo_j[j] + 4);
                  }
                }
#pragma unroll
              for (int j = 0; j < npw; ++j) {
                const float M = fmaxf(ml[j][0].x, ml[j][1].x);
                const float w0 = fast_exp2(ml[j][0].x - M), w1 = fast_exp2(ml[j][1].x - M);
                const float Ls = ml[j][0].y * w0 + ml[j][1].y * w1;
                const float inv = Ls > 0.f ? __fdividef(1.0f, Ls) : 0.f;
                const float4 a = oa[j][0], c = oa[j][1], e = ob[j][0], g = ob[j][1];
                bq[j] = make_uint4(pk((a.x * w0 + c.x * w1) * inv, (a.y * w0 + c.y * w1) * inv), pk((a.z * w0 + c.z * w1) * inv, (a.w * w0 + c.w * w1) * inv),
                                   pk((e.x * w0 + g.x * w1) * inv, (e.y * w0 + g.y * w1) * inv), pk((e.z * w0 + g.z * w1) * inv, (e.w * w0 + g.w * w1) * inv));
              }
            } else {
#pragma unroll
              for (int j = 0; j < npw; ++j) {
                float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, M = -INFINITY, Ls = 0.f;
                for (int s4 = 0; s4 < nsp_j[j]; s4 += 4) {
                  float2 ml[4];
                  float4 oa[4], ob[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float* pp = pb_j[j] + static_cast<size_t>(s4 + u) * PS;
                    ml[u] = make_float2(-INFINITY, 0.f);
                    oa[u] = ob[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s4 + u < nsp_j[j]) {
                      ml[u] = ldv2f(pp + HD);
                      oa[u] = ldv4f(pp + fo_j[j]);
                      ob[u] = ldv4f(pp + fo_j[j] + 4);
                    }
                  }
                  merge_pack(ml, oa, ob, o, M, Ls);
                }
                const float inv = Ls > 0.f ? __fdividef(1.0f, Ls) : 0.f;
                bq[j] = make_uint4(pk(o[0] * inv, o[1] * inv), pk(o[2] * inv, o[3] * inv), pk(o[4] * inv, o[5] * inv), pk(o[6] * inv, o[7] * inv));
              }
            }
            mma_tile<npw>(p.packed + lw.w_out + static_cast<size_t>(rt) * (16 * DM * 2) + static_cast<size_t>(ksp) * npw * 1024, lane, bq, npw, acc);
          }
        };
        if (wactive) {
          const int npw_rt = (kind == K_MLP2 ? DFF / 32 : NPD) / KS;
          switch (npw_rt) {
            case 1: contract(std::integral_constant<int, 1>{}); break;
            case 2: contract(std::integral_constant<int, 2>{}); break;
            case 4: contract(std::integral_constant<int, 4>{}); break;
            case 8: contract(std::integral_constant<int, 8>{}); break;
            default: contract(std::integral_constant<int, 16>{}); break;
          }
        }
        if (KS > 1) {
          sm.red[warp][lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          __syncthreads();
          if (ksp == 0) {
            for (int s = 1; s < KS; ++s) {
              const float4 r = sm.red[s * TN + ntl][lane];
              acc[0] += r.x; acc[1] += r.y; acc[2] += r.z; acc[3] += r.w;
            }
          }
        }
        if (wactive && ksp == 0) {
          float* dst = kind == K_OUT ? p.x1 : p.x;
          bf16* dstb = kind == K_OUT ? p.x1b : p.xb;
          float* dsts = kind == K_OUT ? p.sx1 : p.sx;
          const float y0 = res[0] + (acc[0] + ba), y1 = res[1] + (acc[1] + ba), y2 = res[2] + (acc[2] + bb), y3 = res[3] + (acc[3] + bb);
          if (s0 < B) {
            dst[static_cast<size_t>(s0) * DM + ra] = y0; dstb[static_cast<size_t>(s0) * DM + ra] = __float2bfloat16_rn(y0);
            dst[static_cast<size_t>(s0) * DM + rb] = y2; dstb[static_cast<size_t>(s0) * DM + rb] = __float2bfloat16_rn(y2);
          }
          if (s1 < B) {
            dst[static_cast<size_t>(s1) * DM + ra] = y1; dstb[static_cast<size_t>(s1) * DM + ra] = __float2bfloat16_rn(y1);
            dst[static_cast<size_t>(s1) * DM + rb] = y3; dstb[static_cast<size_t>(s1) * DM + rb] = __float2bfloat16_rn(y3);
          }
          // (sum, sum of squares) of this tile's 16 rows per sequence: the next LayerNorm's statistics are assembled by its consumer
          float a0 = y0 + y2, q0 = fmaf(y0, y0, y2 * y2), a1 = y1 + y3, q1 = fmaf(y1, y1, y3 * y3);
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
          }
          if (qd == 0) {
            if (s0 < B) *reinterpret_cast<float2*>(dsts + (static_cast<size_t>(s0) * NT16 + rt) * 2) = make_float2(a0, q0);
            if (s1 < B) *reinterpret_cast<float2*>(dsts + (static_cast<size_t>(s1) * NT16 + rt) * 2) = make_float2(a1, q1);
          }
        }
        if (KS > 1 && it < n_items && sm.items[it].phase == ph) __syncthreads();     // `red` is reused by the next item
      }
    };

    // Between the arrival at a grid barrier and the wait: pull the weight tile of this CTA's first item of the NEXT phase into L1
    // (immutable data, independent of the barrier), so that the phase starts with ONE L2 round trip (its operand) instead of two
    auto prefetch_next = [&](int ph_next) {
      const int kind = ph_next >= 5 * L ? K_HEAD : ph_next % 5, l = ph_next >= 5 * L ? 0 : ph_next / 5;
      const GridLayer& lw = p.layers[l];
      const size_t mat = kind == K_QKV ? lw.w_in : (kind == K_OUT ? lw.w_out : (kind == K_MLP1 ? lw.w1 : (kind == K_MLP2 ? lw.w2 : p.w_head)));
      const int bytes = 32 * (kind == K_MLP2 ? DFF : DM);
      for (int k = it; k < n_items && sm.items[k].phase == ph_next; ++k) {        // every item of the phase (the head: 3-4 tiles per CTA)
        const uint8_t* base = p.packed + mat + static_cast<size_t>(sm.items[k].row_tile) * bytes;
        for (int i = tid; i < bytes / 128; i += kThreads) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + i * 128));
      }
    };
    auto sync_phase = [&](int ph_next) { stamp(); arrive(); prefetch_next(ph_next); wait(); stamp(); };

    for (int l = 0; l < L && alive; ++l) {
      const GridLayer& lw = p.layers[l];
      phase_rows(5 * l + K_QKV, std::integral_constant<int, K_QKV>{}, l);
      stamp(); arrive();
      // the first K/V block of this warp's attention unit holds older tokens only (the row this step appends is folded from the
      // in_proj output, not read from the cache): requested behind the barrier arrival, its HBM round trip overlaps the barrier
      uint4 kq0[4][HD / 32], vq0[HD / 8];
      const size_t hb = at_b >= 0 ? (static_cast<size_t>(at_b) * H + at_h) * p.Tvt * HD : 0;
      constexpr bool kAttnPre = HD == 32;          // head_dim 64: 64 more live registers across the barrier spill (240 -> 263 us per step)
      if (kAttnPre && at_b >= 0 && (at_wi << 5) < at_len) attn_load_block<HD, true>(lw.kh + hb, lw.vt + hb, at_len, at_wi, lane, kq0, vq0);
      wait(); stamp();
      if (!alive) break;
      // ---------------- attention: this warp's (sequence, head, key range) unit ----------------
      if (at_b >= 0) {
        const size_t so = static_cast<size_t>(at_b) * DM + at_h * HD;
        if (lane < HD / 4) *reinterpret_cast<float4*>(&sm.qs[warp][lane * 4]) = ldv4f(p.q + so + lane * 4);
        if (lane < HD / 8) {
          *reinterpret_cast<uint4*>(&sm.kn[warp][lane * 8]) = ldv4u(p.knew + so + lane * 8);
          *reinterpret_cast<uint4*>(&sm.vn[warp][lane * 8]) = ldv4u(p.vnew + so + lane * 8);
        }
        __syncwarp();
        attn_tc<HD, kAttnPre, true>(lw.kh + hb, lw.vt + hb, at_len, at_wi, at_nws, lane, sm.qs[warp], sm.kn[warp], sm.vn[warp], at_wi == 0,
                                sm.part[warp], kq0, vq0);
        __syncwarp();
      }
      __syncthreads();
      if (at_run > 0) {
        // merge the run's partial results (each: numerators | m | l, log2 domain) and write ONE partial for this CTA
        float M = -INFINITY;
        for (int i = 0; i < at_run; ++i) M = fmaxf(M, sm.part[warp + i][64]);
        float o0 = 0.f, o1 = 0.f, Ls = 0.f;
        for (int i = 0; i < at_run; ++i) {
          const float w = fast_exp2(sm.part[warp + i][64] - M);
          Ls = fmaf(sm.part[warp + i][65], w, Ls);
          o0 = fmaf(sm.part[warp + i][lane], w, o0);
          if (HD == 64) o1 = fmaf(sm.part[warp + i][32 + lane], w, o1);
        }
        float* dst = p.part + ((static_cast<size_t>(at_b) * H + at_h) * kMaxSplits + at_slot) * PS;
        dst[lane] = o0;
        if (HD == 64) dst[32 + lane] = o1;
        if (lane == 0) *reinterpret_cast<float2*>(dst + HD) = make_float2(M, Ls);
      }
      sync_phase(5 * l + K_OUT);
      if (!alive) break;
      phase_split(5 * l + K_OUT, std::integral_constant<int, K_OUT>{}, l);
      sync_phase(5 * l + K_MLP1);
      if (!alive) break;
      phase_rows(5 * l + K_MLP1, std::integral_constant<int, K_MLP1>{}, l);
      sync_phase(5 * l + K_MLP2);
      tr = nullptr;
      if (!alive) break;
      phase_split(5 * l + K_MLP2, std::integral_constant<int, K_MLP2>{}, l);
      sync_phase(5 * l + 5);
    }
    if (!alive) break;
    phase_rows(5 * L + K_QKV, std::integral_constant<int, K_HEAD>{}, 0);
    sync_phase(5 * L + 1);
    if (!alive) break;

    // ---------------- sampler (api_cache.py:169-181): CTA b owns sequence b ----------------
    if (cta < B && !sm.fin[cta]) {
      const int b = cta;
      const float* row = p.logits + static_cast<size_t>(b) * p.ldl;
      const int V = p.V, k = sp.top_k;
      const uint64_t seq = sp.seq_base + static_cast<uint64_t>(p.st.seq_idx ? p.st.seq_idx[b] : b);
      // the sequence's state words are requested first: their L2 round trips overlap the logits loads
      const uint32_t nn = static_cast<uint32_t>(ldvi(p.st.n_new + b));
      const int out_pos = ldvi(p.st.out_len + b), budget = ldvi(p.st.max_new + b);
      int tok = -1;
      const bool fast = V <= kThreads * kSampMaxPer && k >= 1 && k <= kThreads && k < V;
      bool done = false;
      if (p.forced) {
        tok = p.forced[static_cast<size_t>(b) * p.forced_stride + step];
        done = true;
      } else if (fast) {
        // logits / temperature (api_cache.py:169) in registers: element 4 (tid + 256 j) + e
        float z[kSampMaxPer];
        const float inv_temp = 1.0f / sp.temperature;
#pragma unroll
        for (int j = 0; j < kSampMaxPer / 4; ++j) {
          const int i0 = 4 * (tid + kThreads * j);
          float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i0 < V) q4 = ldv4f(row + i0);                                   // rows are padded to a multiple of 16 floats
          z[4 * j] = i0 < V ? q4.x * inv_temp : -INFINITY;
          z[4 * j + 1] = i0 + 1 < V ? q4.y * inv_temp : -INFINITY;
          z[4 * j + 2] = i0 + 2 < V ? q4.z * inv_temp : -INFINITY;
          z[4 * j + 3] = i0 + 3 < V ? q4.w * inv_temp : -INFINITY;
        }
        float tmax = -INFINITY;
        int timax = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < kSampMaxPer; ++j)
          if (z[j] > tmax) { tmax = z[j]; timax = 4 * (tid + kThreads * (j >> 2)) + (j & 3); }   // ascending index: the lowest index wins ties
        if (k == 1) {
          // greedy: arg-max, lowest index on ties (torch.topk / multinomial over a one-hot distribution)
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, tmax, o);
            const int oi = __shfl_xor_sync(0xffffffffu, timax, o);
            if (ov > tmax || (ov == tmax && oi < timax)) { tmax = ov; timax = oi; }
          }
          if (lane == 0) { sm.mx[warp] = tmax; sm.wtot[warp] = timax; }
          __syncthreads();
          float bv = sm.mx[0];
          int bi = sm.wtot[0];
#pragma unroll
          for (int w = 1; w < 8; ++w)
            if (sm.mx[w] > bv || (sm.mx[w] == bv && sm.wtot[w] < bi)) { bv = sm.mx[w]; bi = sm.wtot[w]; }
          tok = bi;
          done = true;
        } else {
          // threshold = the k-th largest of G group maxima (G = 64 quads of threads for k <= 64, else the 256 threads): a lower
          // bound of the k-th largest logit, so {z >= tau} is a (small) superset of the top-k
          const int G = k <= 64 ? 64 : kThreads;
          float gm = tmax;
          if (G == 64) {
            gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 1));
            gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, 2));
            if (tq == 0) sm.mx[tid >> 2] = gm;
          } else {
            sm.mx[tid] = gm;
          }
          __syncthreads();
          if (tid < G) {
            const float mine = sm.mx[tid];
            int rank = 0;
            for (int u4 = 0; u4 < G / 4; ++u4) {
              const float4 m4 = *reinterpret_cast<const float4*>(&sm.mx[4 * u4]);
              const int u = 4 * u4;
              rank += (m4.x > mine || (m4.x == mine && u < tid)) + (m4.y > mine || (m4.y == mine && u + 1 < tid)) +
                      (m4.z > mine || (m4.z == mine && u + 2 < tid)) + (m4.w > mine || (m4.w == mine && u + 3 < tid));
            }
            if (rank == k - 1) sm.tau = mine;
          }
          __syncthreads();
          const float tau = sm.tau;
          int cnt = 0;
#pragma unroll
          for (int j = 0; j < kSampMaxPer; ++j) cnt += z[j] >= tau && 4 * (tid + kThreads * (j >> 2)) + (j & 3) < V;
          int inc = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
          }
          if (lane == 31) sm.wtot[warp] = inc;
          __syncthreads();
          int base = 0, total = 0;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const int t = sm.wtot[w];
            if (w < warp) base += t;
            total += t;
          }
          if (total <= kCandCap) {
            int pos = base + inc - cnt;
#pragma unroll
            for (int j = 0; j < kSampMaxPer; ++j)
              {
                const int idx = 4 * (tid + kThreads * (j >> 2)) + (j & 3);
                if (z[j] >= tau && idx < V) sm.cand[pos++] = make_uint2(__float_as_uint(z[j]), static_cast<uint32_t>(idx));
              }
            __syncthreads();
            // exact rank of every candidate: value descending, index ascending
            for (int c = tid; c < total; c += kThreads) {
              const uint2 me = sm.cand[c];
              const float mv = __uint_as_float(me.x);
              int r = 0;
              for (int u = 0; u < total; ++u) {
                const uint2 o = sm.cand[u];
                const float ov = __uint_as_float(o.x);
                r += (ov > mv) || (ov == mv && o.y < me.y);
              }
              if (r < k) sm.sorted[r] = me;
            }
            __syncthreads();
            if (warp == 0) {
              const float z0 = __uint_as_float(sm.sorted[0].x);
              float w[8], run = 0.f;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int r = lane * 8 + e;
                w[e] = r < k ? expf(__uint_as_float(sm.sorted[r].x) - z0) : 0.f;
                run += w[e];
              }
              float incw = run;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, incw, o);
                if (lane >= o) incw += t;
              }
              const float totalw = __shfl_sync(0xffffffffu, incw, 31);
              const float target = philox_uniform(sp.seed, seq, nn) * totalw;
              float cum = incw - run;
              int pick = -1;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                cum += w[e];
                if (pick < 0 && lane * 8 + e < k && cum > target) pick = lane * 8 + e;
              }
              const unsigned hit = __ballot_sync(0xffffffffu, pick >= 0);
              const int src = hit ? __ffs(hit) - 1 : 0;
              int pk2 = __shfl_sync(0xffffffffu, pick, src);
              if (!hit) pk2 = k - 1;                                     // rounding corner: the last kept candidate
              if (lane == 0) sm.result = static_cast<int>(sm.sorted[pk2].y);
            }
            __syncthreads();
            tok = sm.result;
            done = true;
          }
        }
      }
      if (!done) {
        // general path (top_k = 0 / >= V / > 256, huge vocabularies, more than kCandCap ties at the threshold)
        // the row is staged with L1-bypassing loads into a scratch row only this CTA ever touches (sample_row uses plain loads)
        float* vr = p.vals + static_cast<size_t>(b) * p.ldl;
        for (int i = tid; i < V; i += kThreads) vr[i] = ldvf(row + i);
        __syncthreads();
        tok = sample_row(vr, V, sp.temperature, k, sp.seed, seq, nn, vr, sm.ss);
      }
      if (tid == 0) {
        tok = min(max(tok, 0), p.V - 1);                       // never index the embedding table out of range
        const int pos = out_pos;
        p.st.out_ids[static_cast<size_t>(b) * p.st.out_stride + pos] = tok;      // api_cache.py:179
        if (p.st.step_ns && b == 0) p.st.step_ns[step] = ptx::global_timer_ns();   // per-token latency read-out
        p.st.out_len[b] = pos + 1;
        p.st.cur_tok[b] = tok;
        p.st.lens[b] = sm.len[b] + 1;
        const int n = static_cast<int>(nn) + 1;
        p.st.n_new[b] = n;
        const bool fin_now = tok == sp.eos_id || n >= budget;               // api_cache.py:181
        if (fin_now) {
          p.st.finished[b] = 1;
          atomicAdd(p.ctrl + 1, 1u);
        }
        sm.next_tok = fin_now ? -1 : tok;
      }
      __syncthreads();
      if (sm.next_tok >= 0) embed(b, sm.next_tok);                  // the next step's input row
    }
    sync_phase(0);
    if (!alive) break;
    if (p.early_exit && static_cast<int>(ldvu(p.ctrl + 1)) >= n_active0) break;
  }
}

// ---- weight packing: bf16 [rows, K] row-major -> tiles (see mma_tile) ----
__device__ __forceinline__ uint32_t scale2(uint32_t w2, const float* __restrict__ g, int col) {
  if (!g) return w2;
  const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w2);
  return pack_bf16(__bfloat162float(v.x) * g[col], __bfloat162float(v.y) * g[col + 1]);
}

// c[r] = sum_k bf16(W[r][k] gamma[k]) (exactly the values the MMAs see), d[r] = sum_k W[r][k] beta[k] + bias[r]; one warp per row
__global__ void grid_fold_kernel(const bf16* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ bias, int rows, int K, float* __restrict__ c, float* __restrict__ d) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float cs = 0.f, ds = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = __bfloat162float(w[static_cast<size_t>(r) * K + k]);
    cs += __bfloat162float(__float2bfloat16_rn(wv * gamma[k]));
    ds = fmaf(wv, beta[k], ds);
  }
  cs = warp_sum(cs);
  ds = warp_sum(ds);
  if (lane == 0) { c[r] = cs; d[r] = ds + bias[r]; }
}

__global__ void grid_pack_kernel(const bf16* __restrict__ w, int rows, int K, uint4* __restrict__ dst, size_t n_chunks,
                                 const float* __restrict__ gamma) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n_chunks; idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int per_tile = K * 2;                               // 16-byte chunks per tile: 16 rows x K x 2 B / 16
    const int tile = static_cast<int>(idx / per_tile), c = static_cast<int>(idx % per_tile);
    const int j = c >> 6, e = (c >> 5) & 1, lane = c & 31, g = lane >> 2, t = lane & 3;
    const int col = 32 * j + 8 * t + 4 * e, r0 = 16 * tile + g, r1 = r0 + 8;
    uint32_t v[4] = {0u, 0u, 0u, 0u};
    if (r0 < rows) {
      v[0] = scale2(*reinterpret_cast<const uint32_t*>(w + static_cast<size_t>(r0) * K + col), gamma, col);
      v[2] = scale2(*reinterpret_cast<const uint32_t*>(w + static_cast<size_t>(r0) * K + col + 2), gamma, col + 2);
    }
    if (r1 < rows) {
      v[1] = scale2(*reinterpret_cast<const uint32_t*>(w + static_cast<size_t>(r1) * K + col), gamma, col);
      v[3] = scale2(*reinterpret_cast<const uint32_t*>(w + static_cast<size_t>(r1) * K + col + 2), gamma, col + 2);
    }
    dst[idx] = make_uint4(v[0], v[1], v[2], v[3]);
  }
}

// prefill caches [B][d / 64][Tmax][64] -> kh [B][H][Tvt][hd], vt [B][H][Tvt / 32][hd][32] (attn_tc.cuh)
__global__ void grid_relayout_kv_kernel(const bf16* __restrict__ ksrc0, const bf16* __restrict__ vsrc0, bf16* __restrict__ kdst0,
                                        bf16* __restrict__ vdst0, const int32_t* __restrict__ lens, const int32_t* __restrict__ slots, int ns,
                                        int Tmax, int Tvt, int hd) {
  // (sequence, slice); `slots` (optional) = the sequences to convert (continuous batching: the newly admitted ones)
  const int bs = slots ? slots[blockIdx.x / ns] * ns + blockIdx.x % ns : blockIdx.x, b = bs / ns;
  const int len = lens[b];
  const bf16* ksrc = ksrc0 + static_cast<size_t>(bs) * Tmax * 64;
  const bf16* vsrc = vsrc0 + static_cast<size_t>(bs) * Tmax * 64;
  bf16* kdst = kdst0 + static_cast<size_t>(bs) * Tvt * 64;
  bf16* vdst = vdst0 + static_cast<size_t>(bs) * Tvt * 64;
  for (int i = threadIdx.x; i < len * 64; i += blockDim.x) {
    const int t = i >> 6, f = i & 63, h = f / hd, d = f - h * hd, ki = t & 31;
    const int pos = 8 * ((ki >> 1) & 3) + 2 * (ki >> 3) + (ki & 1);
    kdst[(static_cast<size_t>(h) * Tvt + t) * hd + d] = ksrc[i];
    vdst[(static_cast<size_t>(h) * (Tvt >> 5) + (t >> 5)) * (hd * 32) + d * 32 + pos] = vsrc[i];
  }
}

template <typename F>
auto grid_dispatch(int d_model, int hd, F&& f) {
  if (d_model == 256) return hd == 32 ? f(decode_grid_kernel<256, 32>) : f(decode_grid_kernel<256, 64>);
  return hd == 32 ? f(decode_grid_kernel<512, 32>) : f(decode_grid_kernel<512, 64>);
}

}  // namespace

bool grid_eligible(int d_model, int d_ff, int n_head, int n_layer, int V) {
  if (d_model != 256 && d_model != 512) return false;
  if (d_ff != 4 * d_model || n_head <= 0 || d_model % n_head) return false;
  const int hd = d_model / n_head;
  if (hd != 32 && hd != 64) return false;
  return n_layer >= 1 && n_layer <= kMaxLayers && V >= 2;
}

static size_t tiles_bytes(int rows, int K) { return static_cast<size_t>((rows + 15) / 16) * 16 * K * 2; }

size_t grid_packed_bytes(int d_model, int d_ff, int n_layer, int V) {
  return n_layer * (tiles_bytes(3 * d_model, d_model) + tiles_bytes(d_model, d_model) + tiles_bytes(d_ff, d_model) + tiles_bytes(d_model, d_ff)) +
         tiles_bytes(V, d_model);
}

int grid_pack_weights(cudaStream_t s, const GridPackSrc* src, const bf16* head, int n_layer, int d_model, int d_ff, int V, uint8_t* packed,
                      float* fold, GridLayer* layers, size_t* w_head) {
  size_t off = 0;
  auto pack = [&](const bf16* w, int rows, int K, const float* gamma, size_t* at) -> int {
    *at = off;
    const size_t bytes = tiles_bytes(rows, K);
    grid_pack_kernel<<<296, 256, 0, s>>>(w, rows, K, reinterpret_cast<uint4*>(packed + off), bytes / 16, gamma);
    MG_LAUNCH_CHECK();
    off += bytes;
    return MG_OK;
  };
  auto foldv = [&](const bf16* w, const float* gamma, const float* beta, const float* bias, int rows, int K, float* c, float* d) -> int {
    grid_fold_kernel<<<(rows + 7) / 8, 256, 0, s>>>(w, gamma, beta, bias, rows, K, c, d);
    MG_LAUNCH_CHECK();
    return MG_OK;
  };
  const int per_layer = 2 * (3 * d_model + d_ff);
  for (int l = 0; l < n_layer; ++l) {
    const GridPackSrc& q = src[l];
    float* f = fold + static_cast<size_t>(l) * per_layer;
    float *c_in = f, *d_in = f + 3 * d_model, *c_1 = f + 6 * d_model, *d_1 = f + 6 * d_model + d_ff;
    layers[l] = GridLayer{c_in, d_in, c_1, d_1, q.b_out, q.b2, q.kh, q.vt, 0, 0, 0, 0};
    MG_TRY(pack(q.w_in, 3 * d_model, d_model, q.ln1w, &layers[l].w_in));
    MG_TRY(pack(q.w_out, d_model, d_model, nullptr, &layers[l].w_out));
    MG_TRY(pack(q.w1, d_ff, d_model, q.ln2w, &layers[l].w1));
    MG_TRY(pack(q.w2, d_model, d_ff, nullptr, &layers[l].w2));
    MG_TRY(foldv(q.w_in, q.ln1w, q.ln1b, q.b_in, 3 * d_model, d_model, c_in, d_in));
    MG_TRY(foldv(q.w1, q.ln2w, q.ln2b, q.b1, d_ff, d_model, c_1, d_1));
  }
  MG_TRY(pack(head, V, d_model, nullptr, w_head));
  return MG_OK;
}

int grid_plan(int d_model, int d_ff, int n_layer, int V, int B, int n_cta, int* tn, int* ks, GridItem* items, int32_t* n_items) {
  const int NT = (B + 7) / 8;
  for (int k = 0; k < 8; ++k) { tn[k] = 8; ks[k] = 1; }
  // k-split phases: n-tiles per item x k-splits = 8 warps; the fewer sequences, the more warps split K
  auto shape = [&](int kind, int K, int want_tn) {
    int t = std::min(want_tn, 8);
    while (t > 1 && t / 2 >= NT) t /= 2;                  // no point in more n-tiles per item than the batch has
    int s = 8 / t;
    while ((K / 32) % s != 0 && s > 1) { s /= 2; t = 8 / s; }
    while ((K / 32) / s > 16 && s < 8) { s *= 2; t = 8 / s; }   // <= 16 k-step pairs per warp (registers)
    tn[kind] = t; ks[kind] = s;
  };
  // n-tiles per item of the k-split phases, measured (profiles/r2n_grid_tn_sweep.txt): one n-tile x 8 k-splits for d_model 256
  // (config 3 105.3 -> 102.6, config 4 94.5 -> 92.3 us per step), two n-tiles x 4 k-splits for d_model 512 (240 vs 252)
  int want_out = d_model <= 256 ? 1 : 2, want_mlp2 = d_model <= 256 ? 1 : 2;
  if (const char* e = std::getenv("MG_GRID_TN_OUT")) want_out = std::max(1, std::atoi(e));
  if (const char* e = std::getenv("MG_GRID_TN_MLP2")) want_mlp2 = std::max(1, std::atoi(e));
  shape(K_OUT, d_model, want_out);
  shape(K_MLP2, d_ff, want_mlp2);
  if ((d_model / 32) / ks[K_OUT] > 16 || (d_ff / 32) / ks[K_MLP2] > 16) return MG_E_SHAPE;
  for (int c = 0; c < n_cta; ++c) n_items[c] = 0;
  int cursor = 0;
  auto add = [&](int phase, int tiles, int groups) -> int {
    for (int t = 0; t < tiles; ++t)
      for (int g = 0; g < groups; ++g) {
        const int c = cursor % n_cta;
        ++cursor;
        if (n_items[c] >= kMaxItems) return MG_E_SHAPE;
        items[static_cast<size_t>(c) * kMaxItems + n_items[c]++] = GridItem{static_cast<int16_t>(phase), static_cast<int16_t>(t), static_cast<int16_t>(g), 0};
      }
    return MG_OK;
  };
  for (int l = 0; l < n_layer; ++l) {
    MG_TRY(add(5 * l + K_QKV, 3 * d_model / 16, 1));
    MG_TRY(add(5 * l + K_OUT, d_model / 16, (NT + tn[K_OUT] - 1) / tn[K_OUT]));
    MG_TRY(add(5 * l + K_MLP1, d_ff / 16, 1));
    MG_TRY(add(5 * l + K_MLP2, d_model / 16, (NT + tn[K_MLP2] - 1) / tn[K_MLP2]));
  }
  MG_TRY(add(5 * n_layer, (V + 15) / 16, 1));
  return MG_OK;
}

int grid_init() {
  for (int d : {256, 512})
    for (int hd : {32, 64}) {
      const cudaError_t e = grid_dispatch(d, hd, [&](auto* k) {
        return cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 20);
      });
      MG_CUDA_OK(e);
    }
  return MG_OK;
}

int grid_max_ctas(int d_model, int hd) {
  int per_sm = 0, dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  const cudaError_t e = grid_dispatch(d_model, hd, [&](auto* k) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0); });
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return per_sm > 0 ? prop.multiProcessorCount : 0;          // one CTA per SM
}

int grid_relayout_kv(cudaStream_t s, const bf16* const* kc, const bf16* const* vc, bf16* const* kh, bf16* const* vt, const int32_t* lens,
                     const int32_t* slots, int B, int n_layer, int d_model, int hd, int Tmax, int Tvt) {
  const int ns = d_model / 64;
  for (int l = 0; l < n_layer; ++l) {
    grid_relayout_kv_kernel<<<B * ns, 256, 0, s>>>(kc[l], vc[l], kh[l], vt[l], lens, slots, ns, Tmax, Tvt, hd);
    MG_LAUNCH_CHECK();
  }
  return MG_OK;
}

int launch_decode_grid(cudaStream_t s, const GridParams& p, int d_model, int hd) {
  if ((d_model != 256 && d_model != 512) || (hd != 32 && hd != 64)) return MG_E_SHAPE;
  GridParams pc = p;
  void* args[] = {&pc};
  const cudaError_t e = grid_dispatch(d_model, hd, [&](auto* k) {
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k), dim3(p.n_cta), dim3(kThreads), args, 0, s);
  });
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(MG_E_CUDA, std::string("decode_grid_kernel launch: ") + cudaGetErrorString(e));
  }
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return MG_OK;
}

}  // namespace grid
}  // namespace mg
