// Weight-stationary persistent decode kernel: the WHOLE decode loop of reference api_cache.py:166-182 in ONE launch.
//
// Round-1's cluster kernel (decode_mega.cu) gave every 4-CTA cluster its own sequences and streamed all 10.6 MB of weights
// through every cluster every step: 339 MB/step out of L2, every byte twice through shared memory, and a 48 us chain of
// dependent phases per step that the K/V stream (22 us) could not overlap.  This kernel turns the decomposition around:
//
//   * WEIGHTS NEVER MOVE.  The GEMMs of a step are cut into 16-row tiles (m16 of mma.sync m16n8k16) and every tile lives in
//     the shared memory of ONE SM for the whole generation (10.3 MB over 148 SMs = 70 KB each, host-side plan: flow_plan).
//     An SM owns at most one tile per layer (in_proj 48 tiles, out_proj 16, mlp.0 64, mlp.2 16 = 144 <= 148 SMs) plus 3-4
//     vocabulary tiles of the head.  LayerNorm scale/shift and the attention scale are folded into the packed weights.
//   * SEQUENCES FLOW.  The batch is cut into <= 8 groups of <= 8 sequences (the N of the MMA).  Warp g of EVERY SM works for
//     group g and nothing else, so groups never synchronise with each other and 8 of them are in flight on every SM: while
//     one group's attention units stream K/V, the other groups' dense phases run on the same SMs (the overlap the cluster
//     kernel could not have).
//   * PHASES HAND OVER THROUGH L2 WITH NO FENCE, FLAG OR ATOMIC.  Every exchange buffer is made of 8-byte words
//     (32-bit payload, 32-bit stamp = step * 32 + layer), written with one 64-bit store and read with volatile loads; a consumer
//     polls one sentinel word and then re-reads any word whose stamp is still old.  Measured (profiles/r2a_*): 0.45 us one way
//     between two SMs against 1.9 us for store + fence + atomic flag + acquire.
//   * ATTENTION = split flash-decoding over ALL SMs: unit = (sequence, head, key range); K/V tiles (64 keys) come through a
//     per-warp 2-stage shared-memory ring filled by cp.async.bulk (mbarrier complete_tx, L2 evict-first), the next unit's first
//     tiles are requested as soon as the previous unit ends (before its query exists); S = q K^T and O = P V are
//     mma.sync m16n8k16 with ldmatrix / ldmatrix.trans on a 16-byte-chunk swizzled cache layout (conflict free).  Partial
//     (max, sum, normalised output) go to the out_proj units, which merge them.
//   * SAMPLING: head units publish logits and per-tile maxima; one warp per sequence finds the top-k superset from the tile
//     maxima (16-bit bisection), gathers the candidate tiles, selects exactly k, draws with Philox and publishes the token
//     together with the next step's embedded row.
//
// One CTA per SM (cooperative launch: all CTAs must be co-resident because they wait on each other), 8 warps, no
// __syncthreads after set-up.  Every wait is bounded: a lost hand-over raises status != 0 and the kernel drains instead
// of hanging.  Eligibility: bf16, d_model 256, d_ff 1024, head_dim 32 / 64, <= 64 sequences, 1 <= top_k <= 64.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "decode_flow.cuh"
#include "mg_engine.h"
#include "ptx.cuh"

namespace mg {
namespace flow {

namespace {

constexpr int D = 256;
constexpr int DFF = 1024;
constexpr uint32_t kFull = 0xffffffffu;
constexpr float kLog2e = 1.4426950408889634f;
#ifndef MG_FLOW_MAX_TRIES
#define MG_FLOW_MAX_TRIES (1u << 22)
#endif
constexpr uint32_t kMaxTries = MG_FLOW_MAX_TRIES;      // ~1-2 s of polling: a lost hand-over, fail instead of hanging
constexpr int kCandCap = 128;                          // sampler candidates (value, index) per sequence

enum FlowStatus { FS_OK = 0, FS_TIMEOUT_LL = 1, FS_TIMEOUT_BAR = 2, FS_CAND_OVERFLOW = 3 };

// ---- cache layout ---------------------------------------------------------------------------------
// One (sequence, head) region = Tcap rows of head_dim bf16; the 16-byte chunks of a row are XOR-swizzled with the row index so
// that the 8 row addresses of an ldmatrix 8x8 matrix (8 consecutive keys, same chunk) fall into 8 different bank groups.
__host__ __device__ __forceinline__ int swz_chunk(int hd, int key, int c) { return hd == 32 ? (c ^ ((key >> 1) & 3)) : (c ^ (key & 7)); }
__host__ __device__ __forceinline__ size_t kv_elem_offset(int hd, int key, int dim) {            // in bf16 elements
  return static_cast<size_t>(key) * hd + swz_chunk(hd, key, dim >> 3) * 8 + (dim & 7);
}

// ---- LL words -------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_ld2(const uint64_t* p, uint32_t& d0, uint32_t& s0, uint32_t& d1, uint32_t& s1) {
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d0), "=r"(s0), "=r"(d1), "=r"(s1) : "l"(p) : "memory");
}
__device__ __forceinline__ void ll_ld1(const uint64_t* p, uint32_t& d, uint32_t& s) {
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(d), "=r"(s) : "l"(p) : "memory");
}
__device__ __forceinline__ void ll_st2(uint64_t* p, uint32_t d0, uint32_t d1, uint32_t stamp) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(d0), "r"(stamp), "r"(d1), "r"(stamp) : "memory");
}
__device__ __forceinline__ void ll_st1(uint64_t* p, uint32_t d, uint32_t stamp) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(d), "r"(stamp) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t seq, uint32_t step) {      // same stream as kernels.cu
  uint32_t c[4] = {static_cast<uint32_t>(seq), static_cast<uint32_t>(seq >> 32), step, 0u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ uint64_t make_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}

// ---- per-warp state -------------------------------------------------------------------------------
struct Warp {
  int g, lane, sm, nseq;               // group, lane, SM, sequences of this group
  int qd, tq;                          // lane >> 2 (MMA row / sequence column owner), lane & 3
  uint64_t* xg;                        // exchange buffers of this group
  int32_t* status;
  // per-sequence state, sequence i in lane i (lanes >= 8 mirror lane & 7)
  int len;                             // cached positions before the current step
  int fin;                             // finished (EOS or budget) -- or no such sequence
  int nnew;                            // tokens generated so far
  int maxT;                            // max len over the running sequences of the group
  // K/V stages: a pool of kPoolStages stages shared by the 8 warps of the SM (a warp that streams takes up to kMaxInFlight of
  // them, a warp that waits for its query holds the 2 it requested ahead); `q` = FIFO of the stages this warp has in flight
  uint32_t ring, bars;                 // shared-memory addresses: stage 0, barrier 0
  uint32_t pool;                       // shared-memory address of {free mask, parity mask}
  int step;
  int out0, maxnew;                    // out_len at kernel start, token budget (sequence i in lane i)
  uint32_t q;                          // 4 bits per entry, oldest in the low nibble
  int qn;
  const bf16* pf_k;                    // first K tile of the unit whose tiles were requested ahead (pf_n of them)
  int pf_n;
  uint64_t pol;
  bool dead;
  unsigned long long* prof;            // timeline of this warp's units in the profiled step (null otherwise)
  unsigned long long t_ready;
};
#define FLOW_READY(w) do { if ((w).prof) (w).t_ready = ptx::global_timer_ns(); } while (0)

__device__ __noinline__ void flow_report(int32_t* status, int code, int sm, int g, int detail) {
  if (atomicCAS(status, 0, code) == 0) { status[1] = sm; status[2] = g; status[3] = detail; }
}
__device__ __forceinline__ void flow_fail(Warp& w, int code, int detail) {      // (Warp stays in registers: the call takes values)
  flow_report(w.status, code, w.sm, w.g, detail);
  w.dead = true;
}
// true = keep polling
// what this warp was waiting for when the kernel was aborted: status[8 + 2 (sm * 8 + group)] = {detail, step + 1}
__device__ __noinline__ void flow_waitlog(int32_t* status, int sm, int g, int detail, int step) {
  if ((threadIdx.x & 31) == 0) { status[8 + 2 * (sm * kMaxGroups + g)] = detail; status[9 + 2 * (sm * kMaxGroups + g)] = step + 1; }
}
__device__ __forceinline__ bool poll_ok(Warp& w, uint32_t& tries, int detail) {
  if (++tries > kMaxTries) { flow_fail(w, FS_TIMEOUT_LL, detail); flow_waitlog(w.status, w.sm, w.g, detail, w.step); return false; }
  if ((tries & 255u) == 0 && *reinterpret_cast<volatile int32_t*>(w.status) != 0) { w.dead = true; flow_waitlog(w.status, w.sm, w.g, detail, w.step); return false; }
  return true;
}

__device__ __forceinline__ uint32_t stamp_of(int step, int layer) { return static_cast<uint32_t>(step + 1) * 32u + static_cast<uint32_t>(layer); }

// Wait until ONE word carries `want` (cheap: one sector per iteration for the whole warp).
__device__ __forceinline__ bool ll_wait_word(Warp& w, const uint64_t* p, uint32_t want, int detail) {
  uint32_t d, s, tries = 0;
  while (true) {
    ll_ld1(p, d, s);
    if (s == want) return true;
    if (!poll_ok(w, tries, detail)) return false;
  }
}

// ---- B operand from a bf16 row buffer [8][128 pair words] (payload = two bf16, stored at word 8 ks + 2 tq + j so that one
// 16-byte load brings the two pairs of a k-step): lane (qd, tq) gets, for its sequence qd and every k-step, features
// 16 ks + {2 tq, 2 tq + 1} and {2 tq + 8, 2 tq + 9}; optional LayerNorm (scale / shift are folded into the weights).  The
// residual stream itself stays fp32 (separate words, read by the units that add to it); only this GEMM operand is bf16.
__device__ __forceinline__ int pair_pos(int pair) { return (pair & ~7) | (2 * (pair & 3)) | ((pair >> 2) & 1); }
__device__ __forceinline__ bool load_rows_bf16(Warp& w, const uint64_t* buf, uint32_t want, bool ln, uint32_t (&b)[16][2], int detail) {
  const bool act = w.qd < w.nseq;
  if (!ll_wait_word(w, buf, want, detail)) return false;
  if (w.prof && w.lane == 0) w.prof[3] = ptx::global_timer_ns();
  // All loads are issued back to back (no branch, no short-circuit between them: a conditional around a volatile load makes
  // ptxas wait for each load before the next -- measured 7-8 us for a batch instead of 0.5); lanes without a sequence read
  // row 0 and ignore what they get.
  const uint64_t* row = buf + (act ? w.qd : 0) * (D / 2) + 2 * w.tq;
  uint32_t tries = 0;
  while (true) {
    uint32_t bad = 0;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      uint32_t s0, s1;
      ll_ld2(row + ks * 8, b[ks][0], s0, b[ks][1], s1);
      bad |= (s0 ^ want) | (s1 ^ want);
    }
    if (__all_sync(kFull, bad == 0 || !act)) break;
    if (!poll_ok(w, tries, detail + 1)) return false;
  }
  if (w.prof && w.lane == 0) { w.prof[4] = tries; w.prof[5] = ptx::global_timer_ns(); }
  if (ln) {
    float s = 0.f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) s += (bf_lo(b[ks][0]) + bf_hi(b[ks][0])) + (bf_lo(b[ks][1]) + bf_hi(b[ks][1]));
    s += __shfl_xor_sync(kFull, s, 1);
    s += __shfl_xor_sync(kFull, s, 2);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const float c0 = bf_lo(b[ks][0]) - mean, c1 = bf_hi(b[ks][0]) - mean, c2 = bf_lo(b[ks][1]) - mean, c3 = bf_hi(b[ks][1]) - mean;
      q = fmaf(c0, c0, q); q = fmaf(c1, c1, q); q = fmaf(c2, c2, q); q = fmaf(c3, c3, q);
    }
    q += __shfl_xor_sync(kFull, q, 1);
    q += __shfl_xor_sync(kFull, q, 2);
    const float rstd = rsqrtf(q * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      b[ks][0] = pack_bf16((bf_lo(b[ks][0]) - mean) * rstd, (bf_hi(b[ks][0]) - mean) * rstd);
      b[ks][1] = pack_bf16((bf_lo(b[ks][1]) - mean) * rstd, (bf_hi(b[ks][1]) - mean) * rstd);
    }
  }
  if (!act) {
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) { b[ks][0] = 0u; b[ks][1] = 0u; }
  }
  return true;
}

// One 16-row tile x K = 16 * NKS against the B fragments: weights fragment-major in shared memory (512 bytes per k-step).
template <int NKS>
__device__ __forceinline__ void tile_mma(uint32_t wt, int lane, const uint32_t (*b)[2], float (&acc)[4]) {
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ks = 0; ks < NKS; ks += 2) {
    const uint4 f0 = lds128(wt + ks * 512 + lane * 16);
    const uint4 f1 = lds128(wt + (ks + 1) * 512 + lane * 16);
    mma_bf16_16816(a0, f0.x, f0.y, f0.z, f0.w, b[ks][0], b[ks][1]);
    mma_bf16_16816(a1, f1.x, f1.y, f1.z, f1.w, b[ks + 1][0], b[ks + 1][1]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] = a0[e] + a1[e];
}
__device__ __forceinline__ float tile_bias(uint32_t wt, int nks, int row) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(wt + nks * 512 + row * 4));
  return v;
}


// ---- dense units ----------------------------------------------------------------------------------
// Row order of the tiles whose output is published as bf16 pairs (in_proj, mlp.0): MMA row r < 8 holds weight row 2 r, MMA row
// r + 8 holds weight row 2 r + 1 (flow_pack_kernel), so lane (qd, tq) owns the PAIR of consecutive outputs 2 qd, 2 qd + 1 for
// sequences 2 tq (c0, c2) and 2 tq + 1 (c1, c3): one 32-bit payload each, no shuffle.

// in_proj tile (16 rows of q | k | v): LN1(x) -> pairs -> qkv words; k / v pairs are also appended to the cache.
template <int HD>
__device__ __forceinline__ bool unit_qkv(Warp& w, const FlowParams& p, int layer, int tile, uint32_t wt) {
  const uint32_t st = stamp_of(w.step, layer);
  uint32_t b[16][2];
  if (!load_rows_bf16(w, w.xg + p.xc.off_xb, st, true, b, 100 + layer * 10)) return false;
  FLOW_READY(w);
  float acc[4];
  tile_mma<16>(wt, w.lane, b, acc);
  const float b_lo = tile_bias(wt, 16, w.qd), b_hi = tile_bias(wt, 16, w.qd + 8);
  const uint32_t w0 = pack_bf16(acc[0] + b_lo, acc[2] + b_hi);       // sequence 2 tq
  const uint32_t w1 = pack_bf16(acc[1] + b_lo, acc[3] + b_hi);       // sequence 2 tq + 1
  const int which = tile >> 4, pair = which * 128 + (tile & 15) * 8 + w.qd;
  uint64_t* q = w.xg + p.xc.off_qkv;
  const int s0 = 2 * w.tq, s1 = s0 + 1;
  if (which > 0) {                                                   // cache append at row len (running sequences only)
    const int f = (tile & 15) * 16 + 2 * w.qd, head = f / HD, dim = f % HD;
    bf16* cache = which == 1 ? p.layers[layer].kc : p.layers[layer].vc;
    const int len0 = __shfl_sync(kFull, w.len, s0), len1 = __shfl_sync(kFull, w.len, s1);
    const int fin0 = __shfl_sync(kFull, w.fin, s0), fin1 = __shfl_sync(kFull, w.fin, s1);
    if (s0 < w.nseq && !fin0) {
      const size_t reg = (static_cast<size_t>(s0 * p.n_groups + w.g) * p.n_head + head) * p.Tcap * HD;
      *reinterpret_cast<uint32_t*>(cache + reg + kv_elem_offset(HD, len0, dim)) = w0;
    }
    if (s1 < w.nseq && !fin1) {
      const size_t reg = (static_cast<size_t>(s1 * p.n_groups + w.g) * p.n_head + head) * p.Tcap * HD;
      *reinterpret_cast<uint32_t*>(cache + reg + kv_elem_offset(HD, len1, dim)) = w1;
    }
  }
  ll_st1(q + s0 * 384 + pair, w0, st);
  ll_st1(q + s1 * 384 + pair, w1, st);
  return true;
}

// Residual slice of this tile's 16 features for the 4 accumulators of a lane.  Every tile but the head's has the paired row order
// (MMA row r < 8 = feature 2 r, row r + 8 = feature 2 r + 1), so the lane owns features f0 + 2 qd, + 1 of sequences 2 tq (values 0, 2)
// and 2 tq + 1 (values 1, 3): one 16-byte word pair per sequence in the fp32 buffer.
__device__ __forceinline__ bool load_slice(Warp& w, const uint64_t* buf, uint32_t want, int f0, float (&r)[4], int detail) {
  const int s0 = 2 * w.tq;
  const bool a0_ok = s0 < w.nseq, a1_ok = s0 + 1 < w.nseq;
  const uint64_t* a0 = buf + (a0_ok ? s0 : 0) * D + f0 + 2 * w.qd;
  const uint64_t* a1 = buf + (a1_ok ? s0 + 1 : 0) * D + f0 + 2 * w.qd;
  uint32_t tries = 0;
  while (true) {
    uint32_t d0, d1, d2, d3, t0, t1, t2, t3;
    ll_ld2(a0, d0, t0, d2, t2); ll_ld2(a1, d1, t1, d3, t3);
    const uint32_t bad = (a0_ok ? (t0 ^ want) | (t2 ^ want) : 0u) | (a1_ok ? (t1 ^ want) | (t3 ^ want) : 0u);
    if (__all_sync(kFull, bad == 0)) {
      r[0] = a0_ok ? __uint_as_float(d0) : 0.f; r[1] = a1_ok ? __uint_as_float(d1) : 0.f;
      r[2] = a0_ok ? __uint_as_float(d2) : 0.f; r[3] = a1_ok ? __uint_as_float(d3) : 0.f;
      return true;
    }
    if (!poll_ok(w, tries, detail)) return false;
  }
}
// fp32 words (for the units that add to the residual stream) + bf16 pair words (the next GEMM's operand)
__device__ __forceinline__ void store_slice(Warp& w, uint64_t* buf, uint64_t* bbuf, uint32_t st, int tile, const float (&r)[4]) {
  const int s0 = 2 * w.tq, f = tile * 16 + 2 * w.qd, pos = tile * 8 + 2 * (w.qd & 3) + (w.qd >> 2);
  ll_st1(bbuf + s0 * (D / 2) + pos, pack_bf16(r[0], r[2]), st);
  ll_st1(bbuf + (s0 + 1) * (D / 2) + pos, pack_bf16(r[1], r[3]), st);
  ll_st2(buf + s0 * D + f, __float_as_uint(r[0]), __float_as_uint(r[2]), st);
  ll_st2(buf + (s0 + 1) * D + f, __float_as_uint(r[1]), __float_as_uint(r[3]), st);
}

// Attention partials of a group: per (sequence, key split) one block of 128 output words (bf16 pairs of o / l, head-major,
// inside a head in the order the k-steps of the out_proj MMA read them: word 8 ks + 2 tq + j) and one block of (max, sum) pairs
// per head.  kMaxSplits = 4.
constexpr int kMaxSplits = 4;
__device__ __forceinline__ int part_o_off(int seq, int split) { return (seq * kMaxSplits + split) * (D / 2); }
__device__ __forceinline__ int part_ml_off(int seq, int split, int head) { return kGroupSeqs * kMaxSplits * (D / 2) + ((seq * kMaxSplits + split) * 8 + head) * 2; }

// out_proj tile: the attention outputs of every (sequence, head) -- merged over the S key splits when S > 1 -- are the bf16 B
// operand; x1 = x + W_out att + b_out for this tile's 16 features.
template <int HD>
__device__ __forceinline__ bool unit_out(Warp& w, const FlowParams& p, int layer, int tile, uint32_t wt, int S) {
  constexpr int H = D / HD, KS = HD / 16, HPL = H / 4;               // heads whose (max, sum) this lane fetches
  const uint32_t st = stamp_of(w.step, layer);
  const bool act = w.qd < w.nseq;
  const uint64_t* part = w.xg + p.xc.off_part;
  uint32_t b[16][2];
  float xs[4];                                       // the residual slice is old news (published before the attention): fetch it first,
  if (!load_slice(w, w.xg + p.xc.off_xin, st, tile * 16, xs, 202 + layer * 10)) return false;   // hidden behind the wait for the partials
  if (!ll_wait_word(w, part + part_o_off(0, 0), st, 200 + layer * 10)) return false;
  if (S == 1) {
    // one split: the payloads ARE the B registers (o is published normalised)
    const uint64_t* o = part + part_o_off(act ? w.qd : 0, 0) + 2 * w.tq;
    uint32_t tries = 0;
    while (true) {
      uint32_t bad = 0;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) { uint32_t s0, s1; ll_ld2(o + ks * 8, b[ks][0], s0, b[ks][1], s1); bad |= (s0 ^ st) | (s1 ^ st); }
      if (__all_sync(kFull, bad == 0 || !act)) break;
      if (!poll_ok(w, tries, 201 + layer * 10)) return false;
    }
    if (!act) {
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) { b[ks][0] = 0u; b[ks][1] = 0u; }
    }
  } else {
    float M[H], L[H], acc[16][4];
#pragma unroll
    for (int h = 0; h < H; ++h) { M[h] = -1e30f; L[h] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) { acc[ks][0] = acc[ks][1] = acc[ks][2] = acc[ks][3] = 0.f; }
    for (int s = 0; s < S; ++s) {
      const int sq = act ? w.qd : 0;
      const uint64_t* o = part + part_o_off(sq, s) + 2 * w.tq;
      uint32_t d[16][2], ml[HPL][2];
      uint32_t tries = 0;
      while (true) {
        uint32_t bad = 0;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) { uint32_t s0, s1; ll_ld2(o + ks * 8, d[ks][0], s0, d[ks][1], s1); bad |= (s0 ^ st) | (s1 ^ st); }
#pragma unroll
        for (int j = 0; j < HPL; ++j) {
          uint32_t s0, s1;
          ll_ld2(part + part_ml_off(sq, s, w.tq * HPL + j), ml[j][0], s0, ml[j][1], s1);
          bad |= (s0 ^ st) | (s1 ^ st);
        }
        if (__all_sync(kFull, bad == 0 || !act)) break;
        if (!poll_ok(w, tries, 201 + layer * 10)) return false;
      }
      if (!act) {
#pragma unroll
        for (int j = 0; j < HPL; ++j) { ml[j][0] = __float_as_uint(-1e30f); ml[j][1] = 0u; }
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) { d[ks][0] = d[ks][1] = 0u; }
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        // (max, sum) of head h sits in lane tq = h / HPL of this quad
        const int src = (w.lane & ~3) | (h / HPL);
        const float m = __uint_as_float(__shfl_sync(kFull, ml[h % HPL][0], src)), l = __uint_as_float(__shfl_sync(kFull, ml[h % HPL][1], src));
        const float Mn = fmaxf(M[h], m);
        const float ca = fast_exp2(M[h] - Mn), cb = l * fast_exp2(m - Mn);
        L[h] = L[h] * ca + cb;
        M[h] = Mn;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          const int ks = h * KS + kk;
          acc[ks][0] = acc[ks][0] * ca + cb * bf_lo(d[ks][0]); acc[ks][1] = acc[ks][1] * ca + cb * bf_hi(d[ks][0]);
          acc[ks][2] = acc[ks][2] * ca + cb * bf_lo(d[ks][1]); acc[ks][3] = acc[ks][3] * ca + cb * bf_hi(d[ks][1]);
        }
      }
    }
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const float inv = L[ks / KS] > 0.f ? 1.0f / L[ks / KS] : 0.f;
      b[ks][0] = act ? pack_bf16(acc[ks][0] * inv, acc[ks][1] * inv) : 0u;
      b[ks][1] = act ? pack_bf16(acc[ks][2] * inv, acc[ks][3] * inv) : 0u;
    }
  }
  FLOW_READY(w);
  float acc4[4];
  tile_mma<16>(wt, w.lane, b, acc4);
  const float b_lo = tile_bias(wt, 16, w.qd), b_hi = tile_bias(wt, 16, w.qd + 8);
  const float r[4] = {xs[0] + acc4[0] + b_lo, xs[1] + acc4[1] + b_lo, xs[2] + acc4[2] + b_hi, xs[3] + acc4[3] + b_hi};
  store_slice(w, w.xg + p.xc.off_x1, w.xg + p.xc.off_x1b, st, tile, r);
  return true;
}

// mlp.0 tile: LN2(x1) -> GELU(W1 . + b1) pairs -> h words (laid out so that the mlp.2 units read 16 bytes per k-step)
__device__ __forceinline__ bool unit_mlp1(Warp& w, const FlowParams& p, int layer, int tile, uint32_t wt) {
  const uint32_t st = stamp_of(w.step, layer);
  uint32_t b[16][2];
  if (!load_rows_bf16(w, w.xg + p.xc.off_x1b, st, true, b, 300 + layer * 10)) return false;
  FLOW_READY(w);
  float acc[4];
  tile_mma<16>(wt, w.lane, b, acc);
  const float b_lo = tile_bias(wt, 16, w.qd), b_hi = tile_bias(wt, 16, w.qd + 8);
  const uint32_t w0 = pack_bf16(gelu_erf_f(acc[0] + b_lo), gelu_erf_f(acc[2] + b_hi));
  const uint32_t w1 = pack_bf16(gelu_erf_f(acc[1] + b_lo), gelu_erf_f(acc[3] + b_hi));
  const int pos = tile * 8 + 2 * (w.qd & 3) + (w.qd >> 2);
  uint64_t* h = w.xg + p.xc.off_h;
  ll_st1(h + (2 * w.tq) * 512 + pos, w0, st);
  ll_st1(h + (2 * w.tq + 1) * 512 + pos, w1, st);
  return true;
}

// mlp.2 tile: K = 1024 from the h words (payloads ARE the B registers) -> x_next = x1 + W2 h + b2 for 16 features
__device__ __forceinline__ bool unit_mlp2(Warp& w, const FlowParams& p, int layer, int tile, uint32_t wt) {
  const uint32_t st = stamp_of(w.step, layer);
  const bool act = w.qd < w.nseq;
  const uint64_t* h = w.xg + p.xc.off_h + (act ? w.qd : 0) * 512 + 2 * w.tq;
  float xs[4];                                       // residual slice first (published two phases ago), hidden behind the wait for h
  if (!load_slice(w, w.xg + p.xc.off_x1, st, tile * 16, xs, 402 + layer * 10)) return false;
  if (!ll_wait_word(w, w.xg + p.xc.off_h, st, 400 + layer * 10)) return false;
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    uint32_t b[32][2];
    uint32_t tries = 0;
    while (true) {
      uint32_t bad = 0;
#pragma unroll
      for (int ks = 0; ks < 32; ++ks) {
        uint32_t s0, s1;
        ll_ld2(h + (half * 32 + ks) * 8, b[ks][0], s0, b[ks][1], s1);
        bad |= (s0 ^ st) | (s1 ^ st);
      }
      if (__all_sync(kFull, bad == 0 || !act)) break;
      if (!poll_ok(w, tries, 401 + layer * 10)) return false;
    }
    if (!act) {
#pragma unroll
      for (int ks = 0; ks < 32; ++ks) { b[ks][0] = 0u; b[ks][1] = 0u; }
    }
    const uint32_t wh = wt + half * 32 * 512;
#pragma unroll
    for (int ks = 0; ks < 32; ks += 2) {
      const uint4 f0 = lds128(wh + ks * 512 + w.lane * 16);
      const uint4 f1 = lds128(wh + (ks + 1) * 512 + w.lane * 16);
      mma_bf16_16816(a0, f0.x, f0.y, f0.z, f0.w, b[ks][0], b[ks][1]);
      mma_bf16_16816(a1, f1.x, f1.y, f1.z, f1.w, b[ks + 1][0], b[ks + 1][1]);
    }
  }
  FLOW_READY(w);
  const float b_lo = tile_bias(wt, 64, w.qd), b_hi = tile_bias(wt, 64, w.qd + 8);
  const float r[4] = {xs[0] + a0[0] + a1[0] + b_lo, xs[1] + a0[1] + a1[1] + b_lo, xs[2] + a0[2] + a1[2] + b_hi,
                      xs[3] + a0[3] + a1[3] + b_hi};
  store_slice(w, w.xg + p.xc.off_xin, w.xg + p.xc.off_xb, stamp_of(w.step, layer + 1), tile, r);
  return true;
}

// head: this SM's vocabulary tiles against the final residual stream (no final LayerNorm, api_cache.py:105); logits words,
// per-tile maxima for the sampler, raw logits for the parity path.
__device__ __forceinline__ bool unit_head(Warp& w, const FlowParams& p, const SmProgram& prog, uint32_t blob) {
  const uint32_t st = stamp_of(w.step, p.n_layer);
  uint32_t b[16][2];
  if (!load_rows_bf16(w, w.xg + p.xc.off_xb, st, false, b, 500)) return false;
  FLOW_READY(w);
  const int dslot = !p.dbg_logits ? -1 : (p.dbg_slot ? p.dbg_slot[w.step] : w.step);
  const int s0 = 2 * w.tq, s1 = s0 + 1;
  const int ldl = p.xc.nt * 16;
  for (int i = 0; i < prog.n_head; ++i) {
    const int vt = prog.head_tile[i];
    const uint32_t wt = blob + prog.head_off[i];
    float acc[4];
    tile_mma<16>(wt, w.lane, b, acc);
    const int r_lo = vt * 16 + w.qd, r_hi = r_lo + 8;
    const float b_lo = tile_bias(wt, 16, w.qd), b_hi = tile_bias(wt, 16, w.qd + 8);
    const float v00 = r_lo < p.V ? acc[0] + b_lo : -INFINITY, v01 = r_lo < p.V ? acc[1] + b_lo : -INFINITY;
    const float v10 = r_hi < p.V ? acc[2] + b_hi : -INFINITY, v11 = r_hi < p.V ? acc[3] + b_hi : -INFINITY;
    uint64_t* lg = w.xg + p.xc.off_logits;
    ll_st1(lg + s0 * ldl + r_lo, __float_as_uint(v00), st);
    ll_st1(lg + s1 * ldl + r_lo, __float_as_uint(v01), st);
    ll_st1(lg + s0 * ldl + r_hi, __float_as_uint(v10), st);
    ll_st1(lg + s1 * ldl + r_hi, __float_as_uint(v11), st);
    float m0 = fmaxf(v00, v10), m1 = fmaxf(v01, v11);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      m0 = fmaxf(m0, __shfl_xor_sync(kFull, m0, o));
      m1 = fmaxf(m1, __shfl_xor_sync(kFull, m1, o));
    }
    if (w.qd == 0) {
      uint64_t* tm = w.xg + p.xc.off_tmax;
      ll_st1(tm + s0 * p.xc.nt_pad + vt, __float_as_uint(m0), st);
      ll_st1(tm + s1 * p.xc.nt_pad + vt, __float_as_uint(m1), st);
    }
    if (dslot >= 0) {
      float* dl = p.dbg_logits + static_cast<size_t>(dslot) * p.B * p.V;
      if (s0 < w.nseq) {
        float* row = dl + static_cast<size_t>(s0 * p.n_groups + w.g) * p.V;
        if (r_lo < p.V) row[r_lo] = v00;
        if (r_hi < p.V) row[r_hi] = v10;
      }
      if (s1 < w.nseq) {
        float* row = dl + static_cast<size_t>(s1 * p.n_groups + w.g) * p.V;
        if (r_lo < p.V) row[r_lo] = v01;
        if (r_hi < p.V) row[r_hi] = v11;
      }
    }
  }
  return true;
}

// ---- attention ------------------------------------------------------------------------------------
template <int HD> struct AttnCfg {
  static constexpr int TILE = HD == 32 ? 64 : 32;      // keys per ring stage
  static constexpr int ROWB = HD * 2;                  // bytes per cached row
  static constexpr int HALF = TILE * ROWB;             // bytes of the K (or V) part of a stage = 4096
  static constexpr int KS = HD / 16;                   // k-steps of the score MMAs
  static constexpr int NT = HD / 8;                    // n-tiles of the P V MMAs
  static constexpr int PW = HD / 2 + 2;                // words of one partial
};

__device__ __forceinline__ int split_count(int nseq, int n_head, int maxT, int n_sm) {
  int s = (maxT + 511) / 512;                          // ~512 keys per unit
  s = min(s, n_sm / (nseq * n_head));
  return max(1, min(s, kMaxSplits));
}
// Which attention unit (if any) SM `sm` runs for (group, layer, step): units are dealt round-robin from a rotating start.
__device__ __forceinline__ int attn_unit_of(int sm, int n_sm, int g, int layer, int step, int n_units) {
  const int off = (g * 37 + layer * 53 + step * 29) % n_sm;
  const int u = (sm - off + n_sm) % n_sm;
  return u < n_units ? u : -1;
}

constexpr int kPoolStages = kMaxGroups * kStages;        // 16 stages of 8 KB per SM
constexpr int kMaxInFlight = 6;                          // per streaming warp (48 KB: ~45 GB/s at 1.1 us of HBM latency)
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void atoms_or(uint32_t addr, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_xor(uint32_t addr, uint32_t v) { asm volatile("red.shared.xor.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Take a free stage and start the two bulk copies of one tile into it; false = no stage free right now.
template <int HD>
__device__ __forceinline__ bool ring_issue(Warp& w, const bf16* ksrc, const bf16* vsrc) {
  using C = AttnCfg<HD>;
  int stage = -1;
  if (w.lane == 0) {
    uint32_t m = lds32_volatile(w.pool);
    while (m) {
      const int bit = __ffs(m) - 1;
      const uint32_t old = atoms_cas(w.pool, m, m & ~(1u << bit));
      if (old == m) { stage = bit; break; }
      m = old;
    }
    if (stage >= 0) {
      const uint32_t bar = w.bars + stage * 8, dst = w.ring + stage * kStageBytes;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * C::HALF) : "memory");
      bulk_load_hint(dst, ksrc, C::HALF, bar, w.pol);
      bulk_load_hint(dst + C::HALF, vsrc, C::HALF, bar, w.pol);
    }
  }
  stage = __shfl_sync(kFull, stage, 0);
  if (stage < 0) return false;
  w.q |= static_cast<uint32_t>(stage) << (4 * w.qn);
  w.qn += 1;
  return true;
}
// Wait for the oldest stage in flight; returns its index (-1: watchdog).
__device__ __forceinline__ int ring_wait(Warp& w) {
  // test_wait polling, not try_wait: a thread suspended inside try_wait is woken thousands of cycles after the phase completes
  // (round 1 measured 3100 cycles of skew), and this wait sits on the critical chain of every K/V tile
  const int stage = static_cast<int>(w.q & 15u);
  const uint32_t bar = w.bars + stage * 8, parity = (lds32_volatile(w.pool + 4) >> stage) & 1u;
  uint32_t tries = 0;
  while (true) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return stage;
    if (++tries > 16 * kMaxTries) { flow_fail(w, FS_TIMEOUT_BAR, stage); flow_waitlog(w.status, w.sm, w.g, 900 + stage, w.step); return -1; }
    if ((tries & 4095u) == 0 && *reinterpret_cast<volatile int32_t*>(w.status) != 0) { w.dead = true; flow_waitlog(w.status, w.sm, w.g, 900 + stage, w.step); return -1; }
  }
}
// Give the oldest stage back (every lane has finished reading it).
__device__ __forceinline__ void ring_pop(Warp& w) {
  __syncwarp();
  if (w.lane == 0) {
    const uint32_t bit = 1u << (w.q & 15u);
    atoms_xor(w.pool + 4, bit);                          // next fill of this stage completes the other phase
    atoms_or(w.pool, bit);
  }
  w.q >>= 4;
  w.qn -= 1;
}
// tiles that were requested ahead for a unit that does not run after all
__device__ __forceinline__ bool ring_drain(Warp& w) {
  while (w.qn > 0) { if (ring_wait(w) < 0) return false; ring_pop(w); }
  w.pf_n = 0;
  w.pf_k = nullptr;
  return true;
}

struct UnitGeom { const bf16 *k, *v; int n_tiles, key0; };     // first tile of the unit's key range
template <int HD>
__device__ __forceinline__ UnitGeom unit_geom(const FlowParams& p, int layer, int b, int head, int split, int S, int T) {
  using C = AttnCfg<HD>;
  const int ntile = (T + C::TILE - 1) / C::TILE, tps = (ntile + S - 1) / S;
  const int t0 = min(split * tps, ntile), t1 = min(t0 + tps, ntile);
  const size_t reg = (static_cast<size_t>(b) * p.n_head + head) * p.Tcap * HD + static_cast<size_t>(t0) * C::TILE * HD;
  return UnitGeom{p.layers[layer].kc + reg, p.layers[layer].vc + reg, t1 - t0, t0 * C::TILE};
}

// Request the first tiles of the unit this warp will run next (next layer, or layer 0 of the next step), if it has one.
template <int HD>
__device__ __forceinline__ void attn_prefetch_next(Warp& w, const FlowParams& p, int layer) {
  using C = AttnCfg<HD>;
  int nl = layer + 1, nstep = w.step, grow = 0;
  if (nl == p.n_layer) { nl = 0; nstep = w.step + 1; grow = 1; }
  if (nstep >= p.n_steps || w.pf_n != 0 || w.qn != 0) return;
  const int S = split_count(w.nseq, p.n_head, w.maxT + grow, p.n_sm);
  const int u = attn_unit_of(w.sm, p.n_sm, w.g, nl, nstep, w.nseq * p.n_head * S);
  if (u < 0) return;
  const int seq = u / (p.n_head * S), head = (u / S) % p.n_head, split = u % S;
  const int fin = __shfl_sync(kFull, w.fin, seq), T = __shfl_sync(kFull, w.len, seq) + grow;
  if (fin) return;
  const UnitGeom ug = unit_geom<HD>(p, nl, seq * p.n_groups + w.g, head, split, S, T);
  const int n = min(ug.n_tiles, kStages);
  if (n <= 0) return;
  w.pf_k = ug.k;
  int got = 0;
  for (; got < n; ++got)
    if (!ring_issue<HD>(w, ug.k + static_cast<size_t>(got) * C::TILE * HD, ug.v + static_cast<size_t>(got) * C::TILE * HD)) break;
  w.pf_n = got;
  if (got == 0) w.pf_k = nullptr;
}

// One (sequence, head, key range) unit: flash-decoding on the CUDA cores.  A decode query is ONE row: in mma.sync m16n8k16 form it
// fills 1 of 16 rows, and the 32 HMMAs of a 64-key tile measured 0.85 us for the single warp that owns the unit (legacy HMMA
// latency, nothing to overlap it with) -- 8 KB / 0.85 us = 10 GB/s per unit.  Here: scores = every lane owns the keys
// lane, lane + 32, ... of the tile (16-byte shared-memory loads, conflict free thanks to the chunk swizzle, fp32 FMAs against the
// query held in registers); O += P V = every lane owns two output dims and a 1 / G share of the keys (G = 64 / head_dim lane
// groups), probabilities passed through 256 bytes of per-warp shared memory.
template <int HD>
__device__ __forceinline__ bool unit_attn(Warp& w, const FlowParams& p, int layer, int u, int S, uint32_t pbuf) {
  using C = AttnCfg<HD>;
  constexpr int KPL = C::TILE / 32;                                  // keys per lane in the score pass
  constexpr int LPR = HD / 2;                                        // lanes per key row in the P V pass (two dims each)
  constexpr int G = 32 / LPR;                                        // key groups in the P V pass
  const uint32_t st = stamp_of(w.step, layer);
  const int seq = u / (p.n_head * S), head = (u / S) % p.n_head, split = u % S;
  const int fin = __shfl_sync(kFull, w.fin, seq), T = __shfl_sync(kFull, w.len, seq);
  uint64_t* part_o = w.xg + p.xc.off_part + part_o_off(seq, split) + head * (HD / 2);
  uint64_t* part_ml = w.xg + p.xc.off_part + part_ml_off(seq, split, head);
  const UnitGeom ug = fin ? UnitGeom{nullptr, nullptr, 0, 0} : unit_geom<HD>(p, layer, seq * p.n_groups + w.g, head, split, S, T);
  // ---- reconcile with what was requested ahead ----
  int issued = 0;
  if (w.pf_n > 0 || w.qn > 0) {
    if (w.pf_k == ug.k && ug.n_tiles >= w.pf_n && w.qn == w.pf_n) { issued = w.pf_n; w.pf_n = 0; w.pf_k = nullptr; }
    else if (!ring_drain(w)) return false;
  }
  const bool neutral = fin || (ug.n_tiles == 0 && split != 0);       // nothing to attend to: a neutral partial (after the query wait)
  // keep up to `cap` tiles in flight (as many as the pool gives)
  auto top_up = [&](int cap) {
    while (issued < ug.n_tiles && w.qn < cap) {
      if (!ring_issue<HD>(w, ug.k + static_cast<size_t>(issued) * C::TILE * HD, ug.v + static_cast<size_t>(issued) * C::TILE * HD)) break;
      ++issued;
    }
  };
  top_up(3);                                                        // the query is not here yet: do not hog the pool
  // ---- query: HD / 2 bf16 pairs, already scaled by log2(e) / sqrt(hd); every lane holds the whole vector in fp32 ----
  const uint64_t* qrow = w.xg + p.xc.off_qkv + seq * 384 + head * (HD / 2);
  float q[HD];
  {
    uint32_t tries = 0;
    while (true) {
      uint32_t bad = 0;
#pragma unroll
      for (int j = 0; j < HD / 4; ++j) {
        uint32_t d0, s0, d1, s1;
        ll_ld2(qrow + 2 * j, d0, s0, d1, s1);
        q[4 * j] = bf_lo(d0); q[4 * j + 1] = bf_hi(d0); q[4 * j + 2] = bf_lo(d1); q[4 * j + 3] = bf_hi(d1);
        bad |= (s0 ^ st) | (s1 ^ st);
      }
      if (__all_sync(kFull, bad == 0)) break;
      if (!poll_ok(w, tries, 600 + layer * 10)) return false;
    }
  }
  FLOW_READY(w);
  if (neutral) {
    // Published only AFTER this layer's query has arrived, like every other partial: the query is what orders this unit behind the
    // out_proj units of the previous layer, which may still be reading the words this unit overwrites (a unit that publishes
    // without consuming its input runs ahead of the chain -- that hung ragged / EOS batches)
    if (w.lane < HD / 4) ll_st2(part_o + 2 * w.lane, 0u, 0u, st);
    if (w.lane == HD / 4) ll_st2(part_ml, __float_as_uint(-1e30f), 0u, st);
    return true;
  }
  top_up(kMaxInFlight);
  const int grp = w.lane / LPR, jd = w.lane % LPR;                  // P V pass: key group, dim pair (dims 2 jd, 2 jd + 1)
  float m = -1e30f, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int it = 0; it < ug.n_tiles; ++it) {
    if (w.qn == 0) {                                                  // the pool was empty when this tile was due: insist
      uint32_t tries = 0;
      while (true) {
        top_up(kMaxInFlight);
        if (w.qn > 0) break;
        if (!poll_ok(w, tries, 602 + layer * 10)) return false;
      }
    }
    const unsigned long long tw0 = w.prof ? ptx::global_timer_ns() : 0ull;
    const int stg = ring_wait(w);
    if (stg < 0) return false;
    if (w.prof && w.lane == 0) { w.prof[3] += ptx::global_timer_ns() - tw0; w.prof[4] += 1; }
    const uint32_t kt = w.ring + stg * kStageBytes, vt = kt + C::HALF;
    const int key_base = ug.key0 + it * C::TILE;
    const long long tc0 = w.prof ? clock64() : 0;
    // ---- scores of this lane's keys (log2 domain) ----
    float sc[KPL];
#pragma unroll
    for (int a = 0; a < KPL; ++a) {
      const int key = w.lane + 32 * a;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int c = 0; c < C::ROWB / 16; ++c) {
        const uint4 kv = lds128(kt + key * C::ROWB + (swz_chunk(HD, key, c) << 4));
        acc0 = fmaf(q[8 * c + 0], bf_lo(kv.x), acc0); acc1 = fmaf(q[8 * c + 1], bf_hi(kv.x), acc1);
        acc0 = fmaf(q[8 * c + 2], bf_lo(kv.y), acc0); acc1 = fmaf(q[8 * c + 3], bf_hi(kv.y), acc1);
        acc0 = fmaf(q[8 * c + 4], bf_lo(kv.z), acc0); acc1 = fmaf(q[8 * c + 5], bf_hi(kv.z), acc1);
        acc0 = fmaf(q[8 * c + 6], bf_lo(kv.w), acc0); acc1 = fmaf(q[8 * c + 7], bf_hi(kv.w), acc1);
      }
      sc[a] = key_base + key < T ? acc0 + acc1 : -1e30f;
    }
    float tm = sc[0];
#pragma unroll
    for (int a = 1; a < KPL; ++a) tm = fmaxf(tm, sc[a]);
    tm = warp_max(tm);
    const float mn = fmaxf(m, tm), corr = fast_exp2(m - mn);
    m = mn;
    l *= corr; o0 *= corr; o1 *= corr;
    __syncwarp();                                                    // the previous tile's P V pass has read the probabilities
#pragma unroll
    for (int a = 0; a < KPL; ++a) {
      const float pr = sc[a] > -1e29f ? fast_exp2(sc[a] - mn) : 0.f;
      l += pr;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(pbuf + (w.lane + 32 * a) * 4), "f"(pr) : "memory");
    }
    __syncwarp();
    const long long tc1 = w.prof ? clock64() : 0;
    // ---- O += P V: this lane's two dims over the keys G i + grp ----
#pragma unroll 8
    for (int i = 0; i < C::TILE / G; ++i) {
      const int key = G * i + grp;
      uint32_t vw; float pr;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(vw) : "r"(vt + key * C::ROWB + (swz_chunk(HD, key, jd >> 2) << 4) + (jd & 3) * 4));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pr) : "r"(pbuf + key * 4));
      o0 = fmaf(pr, bf_lo(vw), o0);
      o1 = fmaf(pr, bf_hi(vw), o1);
    }
    if (w.prof && w.lane == 0) { const long long tc2 = clock64(); w.prof[6] += tc1 - tc0; w.prof[7] += tc2 - tc1; }
    ring_pop(w);
    top_up(kMaxInFlight);
  }
  if (w.prof && w.lane == 0) w.prof[5] = ptx::global_timer_ns();
  attn_prefetch_next<HD>(w, p, layer);
  if (G == 2) { o0 += __shfl_xor_sync(kFull, o0, 16); o1 += __shfl_xor_sync(kFull, o1, 16); }
  l = warp_sum(l);
  // ---- the new token's own key / value (split 0): from the qkv words, not from the cache ----
  if (split == 0) {
    const uint64_t* krow = qrow + 128, *vrow = qrow + 256;
    uint32_t kn, vn;
    uint32_t tries = 0;
    while (true) {
      uint32_t s0, s1;
      ll_ld1(krow + jd, kn, s0);
      ll_ld1(vrow + jd, vn, s1);
      if (__all_sync(kFull, ((s0 ^ st) | (s1 ^ st)) == 0)) break;
      if (!poll_ok(w, tries, 601 + layer * 10)) return false;
    }
    // lane jd of the first group multiplies its pair of dims (q is indexed statically through a select chain)
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int j = 0; j < LPR; ++j) { qa = jd == j ? q[2 * j] : qa; qb = jd == j ? q[2 * j + 1] : qb; }
    float s = w.lane < LPR ? fmaf(qa, bf_lo(kn), qb * bf_hi(kn)) : 0.f;
    s = warp_sum(s);
    const float mn = fmaxf(m, s), corr = fast_exp2(m - mn), pn = fast_exp2(s - mn);
    m = mn;
    l = l * corr + pn;
    o0 = o0 * corr + pn * bf_lo(vn);
    o1 = o1 * corr + pn * bf_hi(vn);
  }
  // ---- publish: normalised output as bf16 pairs in the order the out_proj units read them, then (max, sum) ----
  if (w.lane < LPR) {
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    ll_st1(part_o + pair_pos(jd), pack_bf16(o0 * inv, o1 * inv), st);
    if (w.lane == 0) ll_st2(part_ml, __float_as_uint(m), __float_as_uint(l), st);
  }
  return true;
}

// ---- sampler --------------------------------------------------------------------------------------
// Token of sequence `seq` (+ finished flag) and the embedded row of the NEXT step: tok word, x row, decode state.
__device__ __forceinline__ void publish_token(Warp& w, const FlowParams& p, int seq, int tok, int fin, int step_done) {
  // step_done = the step that produced the token (-1 = the prompt's last token at kernel start)
  uint64_t* tw = w.xg + p.xc.off_tok + ((step_done + 1) & 1) * 8 + seq;
  const uint32_t stx = stamp_of(step_done + 1, 0);
  uint64_t* xr = w.xg + p.xc.off_xin + seq * D + w.lane * 8;
  float x[8];
  if (!fin) {
    const uint4 te = *reinterpret_cast<const uint4*>(p.tok_emb + static_cast<size_t>(tok) * D + w.lane * 8);
    const uint4 pe = *reinterpret_cast<const uint4*>(p.pos_emb + w.lane * 8);        // pos_emb[0] on every decode step (api_cache.py:99)
    const uint32_t tv[4] = {te.x, te.y, te.z, te.w}, pv[4] = {pe.x, pe.y, pe.z, pe.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { x[2 * e] = bf_lo(tv[e]) + bf_lo(pv[e]); x[2 * e + 1] = bf_hi(tv[e]) + bf_hi(pv[e]); }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
  }
  uint64_t* xb = w.xg + p.xc.off_xb + seq * (D / 2);
#pragma unroll
  for (int e = 0; e < 4; ++e) ll_st1(xb + pair_pos(w.lane * 4 + e), pack_bf16(x[2 * e], x[2 * e + 1]), stx);
#pragma unroll
  for (int e = 0; e < 4; ++e) ll_st2(xr + 2 * e, __float_as_uint(x[2 * e]), __float_as_uint(x[2 * e + 1]), stx);
  if (w.lane == 0) ll_st1(tw, static_cast<uint32_t>(tok) | (fin ? 0x80000000u : 0u), static_cast<uint32_t>(step_done + 2));
}

// k-th largest of the keys held as `nv` values per lane (0 = absent): bisection over the top `bits` bits (a lower bound of
// the true k-th largest key when bits < 32).
template <int NV>
__device__ __forceinline__ uint32_t kth_largest_key(const uint32_t (&key)[NV], int k, int bits) {
  uint32_t prefix = 0;
  for (int bit = 31; bit >= 32 - bits; --bit) {
    const uint32_t cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < NV; ++j) c += key[j] >= cand;
    c = __reduce_add_sync(kFull, c);
    if (c >= k) prefix = cand;
  }
  return prefix;
}

constexpr int kTmaxLoads = 9;            // 16-byte loads per lane over the tile maxima: nt_pad <= 576 (V <= 9216)

__device__ __forceinline__ bool unit_sample(Warp& w, const FlowParams& p, int seq, uint8_t* scratch) {
  const SampleParams sp = *p.sp;
  const int b = seq * p.n_groups + w.g;
  const int fin_before = __shfl_sync(kFull, w.fin, seq);
  const int nnew = __shfl_sync(kFull, w.nnew, seq), maxnew = __shfl_sync(kFull, w.maxnew, seq);
  const int out0 = __shfl_sync(kFull, w.out0, seq), len = __shfl_sync(kFull, w.len, seq);
  const uint32_t st = stamp_of(w.step, p.n_layer);
  const int nt = p.xc.nt, ldl = nt * 16;
  const uint64_t* tm = w.xg + p.xc.off_tmax + seq * p.xc.nt_pad;
  const uint64_t* lg = w.xg + p.xc.off_logits + seq * ldl;
  int tok = 0;
  // ---- tile maxima: tile 64 j + 2 lane + e.  EVERY path waits for them, also a finished or teacher-forced sequence that does
  // not look at the logits: publishing the next step's row overwrites this sequence's x words, which the head units of other SMs
  // may still be reading -- all tile maxima of this step present = every head unit has consumed the row. ----
  uint32_t key[2 * kTmaxLoads];
  {
    const int nld = p.xc.nt_pad / 64;
    if (!ll_wait_word(w, tm, st, 700)) return false;
    uint32_t tries = 0;
    while (true) {
      uint32_t bad = 0;
#pragma unroll
      for (int j = 0; j < kTmaxLoads; ++j) {
        // tiles past the end are not produced: read tile pair 0 instead (always valid) and drop the result -- no branch around a load
        const int t0 = 64 * j + 2 * w.lane;
        const bool in0 = j < nld && t0 < nt, in1 = j < nld && t0 + 1 < nt;
        uint32_t d0, s0, d1, s1;
        ll_ld2(tm + (in0 ? t0 : 0), d0, s0, d1, s1);
        bad |= (in0 ? s0 ^ st : 0u) | (in1 ? s1 ^ st : 0u);
        key[2 * j] = in0 ? float_key(__uint_as_float(d0)) : 0u;
        key[2 * j + 1] = in1 ? float_key(__uint_as_float(d1)) : 0u;
      }
      if (__all_sync(kFull, bad == 0)) break;
      if (!poll_ok(w, tries, 701)) return false;
    }
    FLOW_READY(w);
  }
  if (fin_before) { publish_token(w, p, seq, 0, 1, w.step); return true; }
  if (p.forced) {
    tok = w.step + 1 < p.n_steps ? p.forced[static_cast<size_t>(b) * p.forced_stride + w.step] : 0;
  } else {
    const int top_k = sp.top_k;
    if (top_k == 1) {
      // greedy: the best tile (lowest index on ties), then the best row of it (lowest index on ties)
      uint32_t bk = 0; int bt = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < 2 * kTmaxLoads; ++j) {
        const int t = 64 * (j >> 1) + 2 * w.lane + (j & 1);
        if (key[j] > bk || (key[j] == bk && key[j] != 0u && t < bt)) { bk = key[j]; bt = t; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ok2 = __shfl_xor_sync(kFull, bk, o); const int ot = __shfl_xor_sync(kFull, bt, o);
        if (ok2 > bk || (ok2 == bk && ot < bt)) { bk = ok2; bt = ot; }
      }
      uint32_t d = 0, s = st, tr2 = 0;
      while (true) {
        ll_ld1(lg + bt * 16 + (w.lane & 15), d, s);
        if (__all_sync(kFull, s == st)) break;
        if (!poll_ok(w, tr2, 702)) return false;
      }
      uint32_t vk = w.lane < 16 ? float_key(__uint_as_float(d)) : 0u; int vi = bt * 16 + w.lane;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const uint32_t ok2 = __shfl_xor_sync(kFull, vk, o); const int oi = __shfl_xor_sync(kFull, vi, o);
        if (ok2 > vk || (ok2 == vk && oi < vi)) { vk = ok2; vi = oi; }
      }
      tok = __shfl_sync(kFull, vi, 0);
    } else {
      // ---- superset threshold: a lower bound of the k-th largest tile maximum (16-bit bisection); every one of the k tiles
      // with the largest maxima holds at least one logit >= it, so the k largest logits are all >= it ----
      uint32_t thr = kth_largest_key(key, top_k, 16);
      float* cval = reinterpret_cast<float*>(scratch);                 // [kCandCap]
      int* cidx = reinterpret_cast<int*>(scratch + kCandCap * 4);      // [kCandCap]
      uint16_t* tlist = reinterpret_cast<uint16_t*>(scratch + kCandCap * 8);   // [256] candidate tiles
      int n_c = 0;
      for (int attempt = 0; attempt < 2; ++attempt) {
        // candidate tiles, compacted lane-major
        int mine = 0;
#pragma unroll
        for (int j = 0; j < 2 * kTmaxLoads; ++j) mine += key[j] >= thr && key[j] != 0u;
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, inc, o); if (w.lane >= o) inc += t; }
        const int n_t = __shfl_sync(kFull, inc, 31);
        int pos = inc - mine;
#pragma unroll
        for (int j = 0; j < 2 * kTmaxLoads; ++j)
          if (key[j] >= thr && key[j] != 0u) { if (pos < 256) tlist[pos] = static_cast<uint16_t>(64 * (j >> 1) + 2 * w.lane + (j & 1)); ++pos; }
        __syncwarp();
        // gather the logits of those tiles (8 row pairs per tile, 16 bytes per load), keep the ones >= thr
        n_c = 0;
        const int items = min(n_t, 256) * 8;
        bool overflow = n_t > 256;
        for (int base = 0; base < items && !overflow; base += 512) {
          uint32_t d[16][2];
          uint32_t tr3 = 0;
          while (true) {
            uint32_t bad = 0;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int itc = min(base + e * 32 + w.lane, items - 1);             // past the end: re-read the last pair, drop it
              uint32_t s0, s1;
              ll_ld2(lg + tlist[itc >> 3] * 16 + (itc & 7) * 2, d[e][0], s0, d[e][1], s1);
              bad |= (s0 ^ st) | (s1 ^ st);
            }
            if (__all_sync(kFull, bad == 0)) break;
            if (!poll_ok(w, tr3, 703)) return false;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int it = base + e * 32 + w.lane;
            if (base + e * 32 >= items) break;                                    // warp-uniform
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const bool keep = it < items && float_key(__uint_as_float(d[e][hh])) >= thr;
              const uint32_t bal = __ballot_sync(kFull, keep);
              const int at = n_c + __popc(bal & ((1u << w.lane) - 1u));
              if (keep && at < kCandCap) { cval[at] = __uint_as_float(d[e][hh]); cidx[at] = tlist[it >> 3] * 16 + (it & 7) * 2 + hh; }
              n_c += __popc(bal);
            }
          }
          if (n_c > kCandCap) overflow = true;
        }
        if (!overflow) break;
        if (attempt == 0) thr = kth_largest_key(key, top_k, 32);       // exact k-th largest tile maximum, try again
        else n_c = min(n_c, kCandCap);                                 // massive ties at the threshold: any of them are a valid top-k
      }
      __syncwarp();
      if (w.prof && w.lane == 0) w.prof[3] = ptx::global_timer_ns();
      // ---- exactly top_k of the candidates: values above the k-th largest candidate, plus the first ties in list order ----
      uint32_t ck[kCandCap / 32]; float cv[kCandCap / 32];
#pragma unroll
      for (int e = 0; e < kCandCap / 32; ++e) {
        const int j = e * 32 + w.lane;
        cv[e] = j < n_c ? cval[j] : -INFINITY;
        ck[e] = j < n_c ? float_key(cv[e]) : 0u;
      }
      const int k_eff = min(top_k, n_c);
      uint32_t kth;
      if (n_c - k_eff <= 8) {
        // the usual case (a handful of extra candidates): walk the distinct values upwards until the n_c - k_eff keys to drop
        // are used up; the value the cut falls into is the k-th largest
        const int drop = n_c - k_eff;
        uint32_t cur = 0;
        int removed = 0;
        kth = 0;
        for (int r = 0; r <= drop; ++r) {
          uint32_t mn = 0xffffffffu;
#pragma unroll
          for (int e = 0; e < kCandCap / 32; ++e) if (ck[e] > cur && ck[e] < mn) mn = ck[e];     // absent entries are key 0
          mn = __reduce_min_sync(kFull, mn);
          int cnt = 0;
#pragma unroll
          for (int e = 0; e < kCandCap / 32; ++e) cnt += ck[e] == mn;
          cnt = __reduce_add_sync(kFull, cnt);
          kth = mn;
          if (removed + cnt > drop) break;
          removed += cnt;
          cur = mn;
        }
      } else {
        kth = kth_largest_key(ck, k_eff, 32);
      }
      int above = 0;
#pragma unroll
      for (int e = 0; e < kCandCap / 32; ++e) above += ck[e] > kth;
      above = __reduce_add_sync(kFull, above);
      int ties_left = k_eff - above;
      float zmax = -INFINITY;
#pragma unroll
      for (int e = 0; e < kCandCap / 32; ++e) zmax = fmaxf(zmax, cv[e]);
      zmax = warp_max(zmax);
      const float scale = kLog2e / sp.temperature;
      float wgt[kCandCap / 32], run = 0.f, cum[kCandCap / 32];
#pragma unroll
      for (int e = 0; e < kCandCap / 32; ++e) {
        const bool tie = ck[e] == kth && ck[e] != 0u;
        const uint32_t tb = __ballot_sync(kFull, tie);
        const bool keep = ck[e] > kth || (tie && __popc(tb & ((1u << w.lane) - 1u)) < ties_left);
        ties_left -= min(ties_left, __popc(tb));
        wgt[e] = keep ? fast_exp2((cv[e] - zmax) * scale) : 0.f;
        float inc = wgt[e];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(kFull, inc, o); if (w.lane >= o) inc += t; }
        cum[e] = run + inc;                                            // inclusive prefix in list order
        run += __shfl_sync(kFull, inc, 31);
      }
      const float target = philox_uniform(sp.seed, sp.seq_base + static_cast<uint64_t>(b), static_cast<uint32_t>(nnew)) * run;
      int pick = 0x7fffffff, last = -1;                                // first kept entry whose inclusive prefix exceeds the target
#pragma unroll
      for (int e = 0; e < kCandCap / 32; ++e) {
        const int j = e * 32 + w.lane;
        if (wgt[e] > 0.f) { last = j; if (cum[e] > target && j < pick) pick = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { pick = min(pick, __shfl_xor_sync(kFull, pick, o)); last = max(last, __shfl_xor_sync(kFull, last, o)); }
      if (pick == 0x7fffffff) pick = max(last, 0);                     // rounding corner: the last kept entry
      tok = cidx[pick];
      if (w.prof && w.lane == 0) w.prof[4] = ptx::global_timer_ns();
    }
  }
  const int n = nnew + 1;
  const int fin_now = (!p.forced && tok == sp.eos_id) || n >= maxnew;
  publish_token(w, p, seq, tok, fin_now, w.step);
  if (w.prof && w.lane == 0) w.prof[5] = ptx::global_timer_ns();
  if (w.lane == 0) {
    p.st.out_ids[static_cast<size_t>(b) * p.st.out_stride + out0 + nnew] = tok;       // api_cache.py:179
    if (fin_now || w.step + 1 >= p.n_steps) {
      p.st.out_len[b] = out0 + n; p.st.lens[b] = len + 1; p.st.n_new[b] = n; p.st.cur_tok[b] = tok;
      p.st.finished[b] = fin_now ? 1 : 0;                                             // api_cache.py:181
    }
  }
  return true;
}

// ---- the kernel -----------------------------------------------------------------------------------
__device__ __forceinline__ int sampler_sm(int n_sm, int g, int seq, int step) { return ((g * 41 + seq * 17 + (step + 1) * 13 + 7) % n_sm + n_sm) % n_sm; }

// ---- unit entry points ----------------------------------------------------------------------------
// Every unit is its own (non-inlined) function: with everything inlined into the step loop, ptxas hoisted the address
// arithmetic of ALL units out of the loop and spilled it (1 KB of local memory per thread next to 225 KB of shared memory,
// i.e. an L1 of ~30 KB: every spill was an L2 round trip).  The per-warp state a unit needs travels by value.
struct RingState { uint32_t q; int qn; int pf_n; const bf16* pf_k; };            // stages one warp has in flight (shared memory)

struct UnitCtx {                        // what the step loop hands to a unit
  int g, nseq, step, sm;
  int len, fin, nnew, maxnew, out0, maxT;
  unsigned long long* prof;            // slot of this unit in the timeline (null: not profiled)
};

__device__ __forceinline__ Warp make_warp(const FlowParams& p, const UnitCtx& c, uint32_t ring, uint32_t bars, const RingState* rs, uint32_t pool = 0) {
  Warp w;
  w.g = c.g; w.lane = threadIdx.x & 31; w.sm = c.sm; w.nseq = c.nseq; w.qd = w.lane >> 2; w.tq = w.lane & 3;
  w.xg = p.xc.base + static_cast<size_t>(c.g) * p.xc.group_words;
  w.status = p.status;
  w.len = c.len; w.fin = c.fin; w.nnew = c.nnew; w.maxnew = c.maxnew; w.out0 = c.out0; w.maxT = c.maxT;
  w.step = c.step;
  w.ring = ring; w.bars = bars; w.pool = pool;
  if (rs) { w.q = rs->q; w.qn = rs->qn; w.pf_n = rs->pf_n; w.pf_k = rs->pf_k; }
  else { w.q = 0; w.qn = 0; w.pf_n = 0; w.pf_k = nullptr; }
  w.pol = 0;
  w.dead = false;
  w.prof = c.prof; w.t_ready = 0;
  return w;
}
__device__ __forceinline__ bool unit_done(Warp& w, bool ok, unsigned long long t0) {
  if (w.prof && w.lane == 0) { w.prof[0] = t0; w.prof[1] = w.t_ready; w.prof[2] = ptx::global_timer_ns(); }
  return ok && !w.dead;
}
#define FLOW_T0(c) ((c).prof ? ptx::global_timer_ns() : 0ull)

template <int HD>
__device__ __noinline__ bool run_qkv(const FlowParams& p, UnitCtx c, int layer, int tile, uint32_t wt) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, 0, 0, nullptr);
  return unit_done(w, unit_qkv<HD>(w, p, layer, tile, wt), t0);
}
template <int HD>
__device__ __noinline__ bool run_out(const FlowParams& p, UnitCtx c, int layer, int tile, uint32_t wt, int S) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, 0, 0, nullptr);
  return unit_done(w, unit_out<HD>(w, p, layer, tile, wt, S), t0);
}
__device__ __noinline__ bool run_mlp1(const FlowParams& p, UnitCtx c, int layer, int tile, uint32_t wt) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, 0, 0, nullptr);
  return unit_done(w, unit_mlp1(w, p, layer, tile, wt), t0);
}
__device__ __noinline__ bool run_mlp2(const FlowParams& p, UnitCtx c, int layer, int tile, uint32_t wt) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, 0, 0, nullptr);
  return unit_done(w, unit_mlp2(w, p, layer, tile, wt), t0);
}
__device__ __noinline__ bool run_head(const FlowParams& p, UnitCtx c, const SmProgram* prog, uint32_t blob) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, 0, 0, nullptr);
  return unit_done(w, unit_head(w, p, *prog, blob), t0);
}
__device__ __noinline__ bool run_sample(const FlowParams& p, UnitCtx c, int seq, uint8_t* scratch) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, 0, 0, nullptr);
  return unit_done(w, unit_sample(w, p, seq, scratch), t0);
}
template <int HD>
__device__ __noinline__ bool run_attn(const FlowParams& p, UnitCtx c, int layer, int u, int S, uint32_t ring, uint32_t bars, RingState* rs, uint32_t pool, uint32_t pbuf) {
  const unsigned long long t0 = FLOW_T0(c);
  Warp w = make_warp(p, c, ring, bars, rs, pool);
  w.pol = make_evict_first_policy();
  const bool ok = unit_attn<HD>(w, p, layer, u, S, pbuf);
  __syncwarp();
  if (w.lane == 0) { rs->q = w.q; rs->qn = w.qn; rs->pf_n = w.pf_n; rs->pf_k = w.pf_k; }
  __syncwarp();
  return unit_done(w, ok, t0);
}
__device__ __noinline__ void run_publish(const FlowParams& p, UnitCtx c, int seq, int tok, int fin) {
  Warp w = make_warp(p, c, 0, 0, nullptr);
  publish_token(w, p, seq, tok, fin, c.step);
}
__device__ __noinline__ void run_drain(const FlowParams& p, UnitCtx c, uint32_t ring, uint32_t bars, RingState* rs, uint32_t pool) {
  Warp w = make_warp(p, c, ring, bars, rs, pool);
  ring_drain(w);
}

// ---- the kernel -----------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kThreads, 1) decode_flow_kernel(const __grid_constant__ FlowParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ SmProgram prog;
  __shared__ RingState rstate[kMaxGroups];
  __shared__ uint32_t pool_words[2];                                 // {free stage mask, parity mask}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sm = blockIdx.x;
  // shared memory: [rings 8 x 2 x 8 KB][scratch 8 x 1.5 KB][barriers][weight blob]
  uint8_t* rings = smem;
  uint8_t* scratch = rings + kMaxGroups * kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + kMaxGroups * kScratchBytes);
  uint8_t* blob = reinterpret_cast<uint8_t*>(bars) + 128;
  {
    if (threadIdx.x < sizeof(SmProgram) / 4) reinterpret_cast<int32_t*>(&prog)[threadIdx.x] = reinterpret_cast<const int32_t*>(&p.prog[sm])[threadIdx.x];
    if (threadIdx.x < kMaxGroups) rstate[threadIdx.x] = RingState{0u, 0, 0, nullptr};
    if (threadIdx.x == 0) { pool_words[0] = (1u << kPoolStages) - 1u; pool_words[1] = 0u; }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(p.packed + prog.blob_off);
    uint4* dst = reinterpret_cast<uint4*>(blob);
    for (int i = threadIdx.x; i < prog.blob_bytes / 16; i += kThreads) dst[i] = src[i];
    if (threadIdx.x < kMaxGroups * kStages) ptx::mbar_init(&bars[threadIdx.x], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (warp >= p.n_groups) return;

  UnitCtx c;
  c.g = warp; c.sm = sm; c.step = -1; c.prof = nullptr;
  c.nseq = (p.B - warp + p.n_groups - 1) / p.n_groups;               // sequences b = i * n_groups + g
  const uint32_t ring = ptx::smem_u32(rings), rbars = ptx::smem_u32(&bars[0]), pool = ptx::smem_u32(&pool_words[0]);
  const uint32_t blob_addr = ptx::smem_u32(blob);
  uint64_t* const xg = p.xc.base + static_cast<size_t>(warp) * p.xc.group_words;
  {
    const int i = lane & 7, b = i * p.n_groups + warp;
    const bool have = i < c.nseq;
    c.len = have ? p.st.lens[b] : 0;
    c.nnew = have ? p.st.n_new[b] : 0;
    c.maxnew = have ? p.st.max_new[b] : 0;
    c.out0 = have ? p.st.out_len[b] - c.nnew : 0;
    c.fin = have ? (p.st.finished[b] != 0 || c.nnew >= c.maxnew) : 1;
    c.maxT = 0;
  }
  // step -1: whoever owns (group, sequence) publishes the last prompt token (fed again, api_cache.py:167-168) and its row
  for (int i = 0; i < c.nseq; ++i)
    if (sampler_sm(p.n_sm, warp, i, -1) == sm) {
      const int fin = __shfl_sync(kFull, c.fin, i);
      run_publish(p, c, i, fin ? 0 : p.st.cur_tok[i * p.n_groups + warp], fin);
    }

  // Groups start `stagger_ns` apart: all groups in the same phase at the same moment means 8 warps of an SM pulling their 16-32 KB
  // of exchange words through the SM's one L2 port together (measured: 2 us per read instead of 0.5) and K/V streams that
  // all run, then all pause.
  if (p.stagger_ns > 0 && warp > 0) {
    const unsigned long long t_go = ptx::global_timer_ns() + static_cast<unsigned long long>(warp) * p.stagger_ns;
    while (ptx::global_timer_ns() < t_go) __nanosleep(200);
  }
  bool alive = true;
#pragma unroll 1
  for (int step = 0; step < p.n_steps && alive; ++step) {
    c.step = step;
    // ---- header: the tokens of the previous step -> per-sequence state ----
    {
      const uint64_t* tw = xg + p.xc.off_tok + (step & 1) * 8 + (lane & 7);
      uint32_t d = 0, s = static_cast<uint32_t>(step + 1), tries = 0;
      while (true) {
        if ((lane & 7) < c.nseq) ll_ld1(tw, d, s);
        if (__all_sync(kFull, s == static_cast<uint32_t>(step + 1))) break;
        if (++tries > kMaxTries) { flow_report(p.status, FS_TIMEOUT_LL, sm, warp, 800); flow_waitlog(p.status, sm, warp, 800, step); alive = false; break; }
        if ((tries & 255u) == 0 && *reinterpret_cast<volatile int32_t*>(p.status) != 0) { flow_waitlog(p.status, sm, warp, 800, step); alive = false; break; }
      }
      if (!alive) break;
      if (step > 0 && !c.fin) { c.nnew += 1; c.len += 1; }
      if ((lane & 7) < c.nseq && (d & 0x80000000u)) c.fin = 1;
      int mt = c.fin ? 0 : c.len;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) mt = max(mt, __shfl_xor_sync(kFull, mt, o));
      c.maxT = mt;
      if (__all_sync(kFull, c.fin != 0)) break;                         // every sequence of the group has finished
    }
    const int S = split_count(c.nseq, p.n_head, c.maxT, p.n_sm);
    // optional timeline of group 0 in ONE step: per SM and unit slot (layer * 5 + kind | 5 L head | 5 L + 1 sampler): entry, inputs
    // complete, done (globaltimer ns)
    unsigned long long* const prof = (p.prof && warp == 0 && step == p.prof_steps) ? p.prof + static_cast<size_t>(sm) * 48 * 8 : nullptr;
#define FLOW_SLOT(slot) (c.prof = prof ? prof + (slot) * 8 : nullptr)
    for (int layer = 0; layer < p.n_layer && alive; ++layer) {
      const int ut = prog.unit_type[layer], tile = prog.unit_tile[layer];
      const uint32_t wt = blob_addr + prog.unit_off[layer];
      if (ut == U_QKV) { FLOW_SLOT(layer * 5 + 0); alive = run_qkv<HD>(p, c, layer, tile, wt); }
      if (!alive) break;
      const int u = attn_unit_of(sm, p.n_sm, warp, layer, step, c.nseq * p.n_head * S);
      if (u >= 0) { FLOW_SLOT(layer * 5 + 1); alive = run_attn<HD>(p, c, layer, u, S, ring, rbars, &rstate[warp], pool, ptx::smem_u32(scratch + warp * kScratchBytes)); }
      if (!alive) break;
      if (ut == U_OUT) { FLOW_SLOT(layer * 5 + 2); alive = run_out<HD>(p, c, layer, tile, wt, S); }
      else if (ut == U_MLP1) { FLOW_SLOT(layer * 5 + 3); alive = run_mlp1(p, c, layer, tile, wt); }
      else if (ut == U_MLP2) { FLOW_SLOT(layer * 5 + 4); alive = run_mlp2(p, c, layer, tile, wt); }
    }
    if (!alive) break;
    if (prog.n_head > 0) { FLOW_SLOT(p.n_layer * 5); alive = run_head(p, c, &prog, blob_addr); }
    for (int i = 0; i < c.nseq && alive; ++i)
      if (sampler_sm(p.n_sm, warp, i, step) == sm) { FLOW_SLOT(p.n_layer * 5 + 1); alive = run_sample(p, c, i, scratch + warp * kScratchBytes); }
  }
  // tiles still in flight must land before the CTA may exit
  c.prof = nullptr;
  if (rstate[warp].qn > 0) run_drain(p, c, ring, rbars, &rstate[warp], pool);
}

// ---- weight packing -------------------------------------------------------------------------------
struct PackTile {
  int32_t type;            // 0 in_proj, 1 out_proj, 2 mlp.0, 3 mlp.2, 4 head
  int32_t layer, tile;
  int64_t dst;             // byte offset in the packed buffer
};
struct PackSrc {
  FlowWeightSrc layers[kMaxLayers];
  const float* head_w;
  const float* head_b;
  int V, n_head;
};

// One block per tile: fragment-major A operand of mma.sync m16n8k16 (k-step ks, lane l -> 16 bytes = a0..a3) + 16 fp32 biases.
// LayerNorm scale is folded into the columns and LayerNorm shift into the bias of in_proj / mlp.0; the attention scale
// log2(e) / sqrt(hd) into the q rows (scores come out in the log2 domain).
__global__ void flow_pack_kernel(const PackTile* tiles, PackSrc src, uint8_t* packed) {
  const PackTile t = tiles[blockIdx.x];
  const bool perm = t.type != 4;                                     // every tile but the head's: outputs published as pairs
  const int K = t.type == 3 ? DFF : D;
  const float* W; const float* bias; const float* cs = nullptr; const float* cb = nullptr;
  int rows;
  float rscale = 1.0f;
  if (t.type == 4) { W = src.head_w; bias = src.head_b; rows = src.V; }
  else {
    const FlowWeightSrc& L = src.layers[t.layer];
    if (t.type == 0) { W = L.w_in; bias = L.b_in; cs = L.ln1w; cb = L.ln1b; rows = 3 * D; if (t.tile < 16) rscale = kLog2e * rsqrtf(static_cast<float>(D / src.n_head)); }
    else if (t.type == 1) { W = L.w_out; bias = L.b_out; rows = D; }
    else if (t.type == 2) { W = L.w1; bias = L.b1; cs = L.ln2w; cb = L.ln2b; rows = DFF; }
    else { W = L.w2; bias = L.b2; rows = D; }
  }
  uint8_t* dst = packed + t.dst;
  const int nks = K / 16;
  for (int i = threadIdx.x; i < nks * 32; i += blockDim.x) {
    const int ks = i >> 5, lane = i & 31, qd = lane >> 2, tq = lane & 3;
    uint32_t regs[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int mrow = qd + (r & 1) * 8;                              // a0 / a2: row qd, a1 / a3: row qd + 8
      const int wrow = t.tile * 16 + (perm ? (mrow < 8 ? 2 * mrow : 2 * (mrow - 8) + 1) : mrow);
      const int col = ks * 16 + 2 * tq + (r >> 1) * 8;
      float v0 = 0.f, v1 = 0.f;
      if (wrow < rows) {
        v0 = W[static_cast<size_t>(wrow) * K + col] * rscale; v1 = W[static_cast<size_t>(wrow) * K + col + 1] * rscale;
        if (cs) { v0 *= cs[col]; v1 *= cs[col + 1]; }
      }
      regs[r] = pack_bf16(v0, v1);
    }
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(i) * 16) = make_uint4(regs[0], regs[1], regs[2], regs[3]);
  }
  for (int mrow = threadIdx.x; mrow < 16; mrow += blockDim.x) {
    const int wrow = t.tile * 16 + (perm ? (mrow < 8 ? 2 * mrow : 2 * (mrow - 8) + 1) : mrow);
    float bv = 0.f;
    if (wrow < rows) {
      bv = bias[wrow];
      if (cb) for (int c = 0; c < K; ++c) bv = fmaf(W[static_cast<size_t>(wrow) * K + c], cb[c], bv);
      bv *= rscale;
    }
    reinterpret_cast<float*>(dst + static_cast<size_t>(nks) * 512)[mrow] = bv;
  }
}

// prefill caches [B][d/64][Tmax][64] (kernels.cu kv_append) -> flow caches [B][H][Tcap][hd] swizzled, rows < lens[b]
__global__ void flow_relayout_kernel(const bf16* __restrict__ kc, const bf16* __restrict__ vc, bf16* __restrict__ fk, bf16* __restrict__ fv,
                                     const int32_t* __restrict__ lens, int n_head, int hd, int Tmax, int Tcap) {
  const int b = blockIdx.y, len = lens[b];
  // one thread per 16-byte chunk (8 features) of one cached row
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len * (D / 8); i += gridDim.x * blockDim.x) {
    const int t = i / (D / 8), c = i % (D / 8), f = c * 8, slice = f / 64, head = f / hd, dim = f % hd;
    const size_t srco = ((static_cast<size_t>(b) * (D / 64) + slice) * Tmax + t) * 64 + (f % 64);
    const size_t dsto = (static_cast<size_t>(b) * n_head + head) * Tcap * hd + kv_elem_offset(hd, t, dim);
    *reinterpret_cast<uint4*>(fk + dsto) = *reinterpret_cast<const uint4*>(kc + srco);
    *reinterpret_cast<uint4*>(fv + dsto) = *reinterpret_cast<const uint4*>(vc + srco);
  }
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------
int flow_init() {
  // 227 KB per CTA in total: the static part (SM program, ring states) comes off the dynamic maximum
  cudaFuncAttributes fa{};
  MG_CUDA_OK(cudaFuncGetAttributes(&fa, decode_flow_kernel<32>));
  MG_CUDA_OK(cudaFuncSetAttribute(decode_flow_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - static_cast<int>(fa.sharedSizeBytes)));
  MG_CUDA_OK(cudaFuncGetAttributes(&fa, decode_flow_kernel<64>));
  MG_CUDA_OK(cudaFuncSetAttribute(decode_flow_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - static_cast<int>(fa.sharedSizeBytes)));
  return MG_OK;
}

bool flow_eligible(int d_model, int d_ff, int n_head, int n_layer, int V, int n_sm) {
  if (d_model != D || d_ff != DFF || n_head <= 0 || d_model % n_head) return false;
  const int hd = d_model / n_head;
  if (hd != 32 && hd != 64) return false;
  if (n_layer < 1 || n_layer > kMaxLayers) return false;
  if (n_sm < 144 || n_sm > kMaxSM) return false;                      // one dense unit per layer and SM
  const int nt = (V + 15) / 16;
  if (nt < kMaxTopK || (nt + 63) / 64 > kTmaxLoads) return false;
  return (nt + n_sm - 1) / n_sm <= kMaxHeadTiles;
}

int flow_tcap(int max_seq, int head_dim) {
  const int tile = head_dim == 32 ? 64 : 32;
  return (max_seq + tile - 1) / tile * tile + tile;                    // + one tile: whole tiles are read
}

FlowExchange flow_exchange_layout(int V, int n_head, int head_dim, int n_sm) {
  FlowExchange x{};
  x.nt = (V + 15) / 16;
  x.nt_pad = (x.nt + 63) / 64 * 64;
  int o = 0;
  x.off_xin = o; o += kGroupSeqs * D;
  x.off_xb = o; o += kGroupSeqs * (D / 2);
  x.off_qkv = o; o += kGroupSeqs * 384;
  x.off_part = o; o += kGroupSeqs * 4 * (D / 2) + kGroupSeqs * 4 * 8 * 2;     // outputs + (max, sum) pairs, kMaxSplits = 4
  x.off_x1 = o; o += kGroupSeqs * D;
  x.off_x1b = o; o += kGroupSeqs * (D / 2);
  x.off_h = o; o += kGroupSeqs * 512;
  x.off_tok = o; o += 16;
  x.off_tmax = o; o += kGroupSeqs * x.nt_pad;
  x.off_logits = o; o += kGroupSeqs * x.nt * 16;
  x.group_words = (o + 15) / 16 * 16;
  (void)n_head; (void)n_sm;
  return x;
}

// Unit -> SM assignment.  Layer l deals its 144 units (in_proj 48, out_proj 16, mlp.0 64, mlp.2 16) to SMs (u + 37 l) mod n_sm,
// so that no SM gets the 32 KB mlp.2 tile of more than one layer (the mlp.2 range is 16 wide, the shift 37); vocabulary tiles
// then go, one at a time, to the SM with the smallest blob.
int flow_plan(int n_layer, int V, int n_sm, FlowPlan* plan) {
  FlowPlan& pl = *plan;
  pl = FlowPlan{};
  pl.n_sm = n_sm;
  pl.nt = (V + 15) / 16;
  const int tile_k256 = 16 * 512 + kTileBiasBytes, tile_k1024 = 64 * 512 + kTileBiasBytes;
  std::vector<int> bytes(n_sm, 0);
  for (int s = 0; s < n_sm; ++s) {
    for (int l = 0; l < kMaxLayers; ++l) { pl.prog[s].unit_type[l] = U_NONE; pl.prog[s].unit_tile[l] = 0; pl.prog[s].unit_off[l] = 0; }
    pl.prog[s].n_head = 0;
  }
  for (int l = 0; l < n_layer; ++l)
    for (int u = 0; u < 144; ++u) {
      const int s = (u + 37 * l) % n_sm;
      int type, tile;
      if (u < 48) { type = U_QKV; tile = u; }
      else if (u < 64) { type = U_OUT; tile = u - 48; }
      else if (u < 128) { type = U_MLP1; tile = u - 64; }
      else { type = U_MLP2; tile = u - 128; }
      if (pl.prog[s].unit_type[l] != U_NONE) return fail(MG_E_ARG, "flow_plan: two dense units of one layer on one SM");
      pl.prog[s].unit_type[l] = type; pl.prog[s].unit_tile[l] = tile; pl.prog[s].unit_off[l] = bytes[s];
      bytes[s] += type == U_MLP2 ? tile_k1024 : tile_k256;
    }
  for (int t = 0; t < pl.nt; ++t) {
    int best = 0;
    for (int s = 1; s < n_sm; ++s)
      if (bytes[s] < bytes[best]) best = s;
    SmProgram& pr = pl.prog[best];
    if (pr.n_head >= kMaxHeadTiles) return fail(MG_E_OOM, "flow_plan: too many vocabulary tiles per SM");
    pr.head_tile[pr.n_head] = t; pr.head_off[pr.n_head] = bytes[best]; pr.n_head += 1;
    bytes[best] += tile_k256;
  }
  size_t off = 0;
  for (int s = 0; s < n_sm; ++s) {
    pl.prog[s].blob_off = static_cast<int32_t>(off);
    pl.prog[s].blob_bytes = bytes[s];
    off += (bytes[s] + 127) / 128 * 128;
    pl.max_blob = std::max<size_t>(pl.max_blob, bytes[s]);
  }
  pl.packed_bytes = off;
  pl.smem_bytes = static_cast<size_t>(kMaxGroups) * kStages * kStageBytes + kMaxGroups * kScratchBytes + 128 + (pl.max_blob + 127) / 128 * 128;
  if (pl.smem_bytes > 227 * 1024 - 2048) return fail(MG_E_OOM, "flow_plan: weight tiles do not fit the shared memory of the GPU");
  return MG_OK;
}

int flow_pack_weights(cudaStream_t s, const FlowPlan& plan, const FlowWeightSrc* layers, int n_layer, int n_head, const float* head_w,
                      const float* head_b, int V, uint8_t* packed) {
  std::vector<PackTile> tiles;
  for (int sm = 0; sm < plan.n_sm; ++sm) {
    const SmProgram& pr = plan.prog[sm];
    for (int l = 0; l < n_layer; ++l)
      if (pr.unit_type[l] != U_NONE) tiles.push_back(PackTile{pr.unit_type[l], l, pr.unit_tile[l], static_cast<int64_t>(pr.blob_off) + pr.unit_off[l]});
    for (int i = 0; i < pr.n_head; ++i) tiles.push_back(PackTile{4, 0, pr.head_tile[i], static_cast<int64_t>(pr.blob_off) + pr.head_off[i]});
  }
  PackSrc src{};
  for (int l = 0; l < n_layer; ++l) src.layers[l] = layers[l];
  src.head_w = head_w; src.head_b = head_b; src.V = V; src.n_head = n_head;
  PackTile* d_tiles = nullptr;
  MG_CUDA_OK(cudaMalloc(&d_tiles, tiles.size() * sizeof(PackTile)));
  cudaError_t e = cudaMemcpyAsync(d_tiles, tiles.data(), tiles.size() * sizeof(PackTile), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(packed, 0, plan.packed_bytes, s);
  if (e == cudaSuccess) {
    flow_pack_kernel<<<static_cast<unsigned>(tiles.size()), 256, 0, s>>>(d_tiles, src, packed);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);                 // `tiles` is pageable host memory
  cudaFree(d_tiles);
  if (e != cudaSuccess) return fail(MG_E_CUDA, std::string("flow_pack_weights: ") + cudaGetErrorString(e));
  return MG_OK;
}

int flow_relayout_kv(cudaStream_t s, const bf16* kc, const bf16* vc, bf16* fk, bf16* fv, const int32_t* lens, int B, int n_head,
                     int head_dim, int Tmax, int Tcap) {
  flow_relayout_kernel<<<dim3(8, B), 256, 0, s>>>(kc, vc, fk, fv, lens, n_head, head_dim, Tmax, Tcap);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int launch_decode_flow(cudaStream_t s, const FlowParams& p, size_t smem_bytes) {
  FlowParams pp = p;
  void* args[] = {&pp};
  const void* fn = p.head_dim == 32 ? reinterpret_cast<const void*>(decode_flow_kernel<32>) : reinterpret_cast<const void*>(decode_flow_kernel<64>);
  // cooperative launch: every CTA waits on words written by the others, so all of them must be co-resident
  MG_CUDA_OK(cudaLaunchCooperativeKernel(fn, dim3(p.n_sm), dim3(kThreads), args, smem_bytes, s));
  MG_LAUNCH_CHECK();
  return MG_OK;
}

}  // namespace flow
}  // namespace mg
