// Persistent cluster decode kernel: the WHOLE decode loop of reference api_cache.py:166-182 for a
// group of sequences in ONE launch -- embedding, L pre-LN blocks (api_cache.py:51-74), head and the
// top-k Philox sampler -- with no host round trip and no grid-wide synchronisation.
//
// Decomposition (bf16, d_model = 256, d_ff = 1024):
//   * a cluster of CL = 4 CTAs owns S <= SMAX sequences for the whole generation; clusters never talk
//     to each other (sequences are independent: reference sample_kvcache is per prompt);
//   * CTA r of a cluster owns the feature slice [64r, 64r+64) of every layer: its q/k/v rows of
//     in_proj, its K/V cache slice, the matching 64 input columns of out_proj, hidden units
//     [256r, 256r+256) of the MLP and vocabulary rows [r*VS, (r+1)*VS) of the head (Megatron-style
//     column-parallel -> row-parallel pairs, so a layer needs two exchanges only);
//   * the weights were re-laid out at load time (mega_pack_weights) as the exact shared-memory image of
//     every 32 KB stage [256 weight rows x 64 K, ldmatrix-fragment-major] in consumption order, so the producer warp
//     streams them L2 -> shared memory with ONE cp.async.bulk per stage into a 4-stage ring ("full" = mbarrier
//     with complete_tx, "free" = named barrier: bar.arrive by the compute warps, bar.sync by the producer warp);
//   * the contractions are warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 -> fp32) issued by all eight
//     compute warps straight from the ring with ldmatrix (weights = the M = 16 operand, the cluster's <= 8
//     sequences = the N = 8 operand), software-pipelined by hand (ldmatrix of stage i + 1 before the MMAs of
//     stage i).  A tcgen05/TMEM formulation of the same step (swap-AB, M = 128,
//     N = 16, single issuing thread) was built and measured first: at 2-4 sequences per cluster it is bound
//     by instruction issue (~85 cycles per tcgen05.mma, 704 per step = 31 us) and by the commit -> mbarrier
//     -> tcgen05.ld hand-offs (profiles/r1_mega_tcgen05_timeline.txt); tcgen05 stays where the contraction
//     is dense (prefill, batched multi-kernel decode, the classifier: gemm_tc.cu);
//   * row-parallel partial sums are all-gathered across the cluster with st.async (distributed shared
//     memory writes that complete_tx on the receiver's mbarrier), summed in a fixed order, and the residual
//     add + the NEXT LayerNorm are done by the same warp in the same phase;
//   * attention is tensor-core flash-decoding straight from global memory over re-laid-out caches (K head-major,
//     V transposed per 32-key block; see attn_tc): one warp per (sequence, head, key range), 16-byte loads,
//     no shuffles between the score and the output MMAs, two register sets of loads in flight;
//   * sampling: every CTA selects the top-k of its vocabulary slice (threshold from the per-thread maxima,
//     atomic-free gather, radix-select fallback; arg-max fast path for top_k = 1), the candidates go to the
//     sequence's owner CTA, which ranks them, draws with Philox and broadcasts the token.
//
// Warp roles (384 threads): a producer warpgroup (warp 0 streams the weights; setmaxnreg.dec) and two compute
// warpgroups (LayerNorm, MMA, attention, epilogues, exchange, sampler; setmaxnreg.inc to 232 registers).
// Where the time of a step goes, and what was measured and rejected: profiles/r1e_mega_timeline.txt.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>

#include "decode_mega.cuh"
#include "mg_engine.h"
#include "ptx.cuh"
#include "attn_tc.cuh"

// -DMG_MEGA_TRACE: clock64 trace of one layer of the profiled step (costs registers; off in the product build)
#ifdef MG_MEGA_TRACE
#define MG_TR(tr) do { if (tr) *tr++ = clock64(); } while (0)
#else
#define MG_TR(tr) do { } while (0)
#endif

namespace mg {
namespace mega {

namespace {

using namespace attn;

constexpr int CL = kMegaCluster;       // CTAs per cluster
constexpr int FS = 64;                 // features per CTA slice
constexpr int D = 256;                 // d_model
constexpr int HS = 256;                // hidden units per CTA (d_ff / CL)
constexpr int NCW = 8;                 // compute warps (all of them stream K/V in the attention phase; 12 measured no faster)
constexpr int GW = 8;                  // ... of which the first 8 run the GEMMs, epilogues and the sampler
constexpr int NCT = NCW * 32;          // compute threads
constexpr int NPW = 4;                 // producer warpgroup: warp 0 lane 0 streams the weights, warps 1..3 idle (setmaxnreg is per warpgroup)
constexpr int NTHREADS = (NPW + NCW) * 32;
constexpr int XP = 264;                // activation row pitch (bf16 elements): K = 256 + 8 pad, bank-conflict free
constexpr int AP = 72;                 // attention-output row pitch: K = 64 + 8 pad
constexpr int STAGE_BYTES = kMegaStageBytes;   // one stage: two weight tiles [128 rows x 64 K] bf16 (rows 0..255)
constexpr int KMAX = kMegaMaxTopK;     // top_k limit of the in-kernel sampler
constexpr int kCandCap = 128;          // per-sequence superset capacity of the sampler's fast path
constexpr float kLog2e = 1.4426950408889634f;

// named barriers: 1 = the compute warps, 3 + sequence = a sampler group, BAR_STAGE_FREE + ring stage = "every compute warp has read it"
enum { BAR_COMPUTE = 1, BAR_STAGE_FREE = 8 };
// per-layer parameter block kept in shared memory for the whole generation (floats):
//   ln1w 256 | ln1b 256 | ln2w 256 | ln2b 256 | b_q 64 | b_k 64 | b_v 64 | b_out 256 | b1 slice 256 | b2 256
constexpr int kLayerParamFloats = 4 * 256 + 3 * 64 + 3 * 256;
enum { P_LN1W = 0, P_LN1B = 256, P_LN2W = 512, P_LN2B = 768, P_BQKV = 1024, P_BOUT = 1216, P_B1 = 1472, P_B2 = 1728 };

template <int SMAX, int NSTAGE>
struct Smem {
  static constexpr int kRing = 0;
  static constexpr int kBx = kRing + NSTAGE * STAGE_BYTES;           // [8][XP] bf16: LN output / head input
  static constexpr int kBh = kBx + 8 * XP * 2;                       // [8][XP] bf16: GELU(mlp.0) of this CTA's hidden slice
  static constexpr int kBatt = kBh + 8 * XP * 2;                     // [8][AP] bf16: attention output of this CTA's slice
  static constexpr int kSlots = kBatt + 8 * AP * 2;                  // [2][CL][256][SMAX] fp32 exchange slots
  static constexpr int kLogits = kSlots + 2 * CL * D * SMAX * 4;     // [SMAX][NL] fp32 (NL <= kMegaMaxNL)
  static constexpr int kX = kLogits + SMAX * kMegaMaxNL * 4;         // [SMAX][256] fp32 residual stream
  static constexpr int kQ = kX + SMAX * D * 4;                       // [SMAX][64] fp32 (pre-scaled, log2 domain)
  static constexpr int kKnew = kQ + SMAX * FS * 4;                   // [SMAX][64] bf16
  static constexpr int kVnew = kKnew + SMAX * FS * 2;
  static constexpr int kPart = kVnew + SMAX * FS * 2;                // [SMAX][NCW][68] fp32 attention partials
  static constexpr int kCand = kPart + SMAX * NCW * 68 * 4;          // [CL][KMAX] (value, index) at the owner
  static constexpr int kLocal = kCand + CL * KMAX * 8;               // [SMAX][KMAX] local candidates + [KMAX] sorted list
  static constexpr int kHist = kLocal + (SMAX + 1) * KMAX * 8;       // [SMAX][256] u32 radix histograms
  static constexpr int kParams = kHist + SMAX * 1024;                // per-layer LN / bias slices (fp32)
  static constexpr int kMisc = kParams + kMegaMaxLayersSmem * kLayerParamFloats * 4;                  // small scalars
  static constexpr int kBars = kMisc + 512;
  // K/V staging slots (MG_MEGA_KVSLOT): one 4 KB slot per compute warp = one 32-key block of one head (K 2 KB | V^T 2 KB), filled by
  // cp.async.bulk and tracked by an mbarrier -- a third block in flight per warp next to the two register sets
  static constexpr int kKvSlot = kBars + 512;
  static constexpr bool kHasKvSlot = MG_MEGA_KVSLOT && SMAX <= 2;    // no room next to the wider buffers of 3..4 sequences
  static constexpr int kTotal = kKvSlot + (kHasKvSlot ? NCW * 4096 : 0);
};

struct MiscSmem {
  uint32_t prefix;
  int remaining;
  int count;
  int result;
  float total;
  int tok[8];
  int len[8];
  int nnew[8];
  int fin[8];
  int maxnew[8];
  int go;
  int sel[4][4];                      // per sequence: remaining, gathered count, radix prefix, exact flag
  bf16* kvp[kMegaMaxLayersSmem][2];   // K / V cache base of every layer (the MegaLayer table lives in global memory)
  int wtot[4][8];                     // sampler: per-warp candidate counts of a sequence's group
  int outlen[8];                      // tokens written so far per sequence (owner CTA; global copy updated at the end)
};

static_assert(sizeof(MiscSmem) <= 512, "MiscSmem outgrew its shared-memory slot");

struct Bars {
  uint64_t full[16];
  uint64_t empty[16];
  uint64_t xchg[2];
  uint64_t cand;
  uint64_t tok;
  uint64_t step_go;                   // compute warps -> producer: the next step will run (early-exit mode)
  uint64_t kvslot[8];                 // per compute warp: "the block in my staging slot has landed" (MG_MEGA_KVSLOT)
};

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
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t seq, uint32_t step) {
  uint32_t c[4] = {static_cast<uint32_t>(seq), static_cast<uint32_t>(seq >> 32), step, 0u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---- warp-level tensor-core helpers ---------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&a)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}

// ---- ring bookkeeping shared by the producer and the compute warps -------------------------------
struct RingPos {
  int stage = 0;
  uint32_t phase = 0;
  template <int NSTAGE> __device__ __forceinline__ void advance() {
    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
  }
};

// A-operand (weight) fragments of one stage for this warp: 2 m16 tiles x 4 k-steps, double-buffered across stages.
struct WFrag {
  uint32_t a[2][2][4][4];
  uint32_t off;                   // this lane's offset inside a stage: warp * 4 KB + lane * 16 (fragment-major layout)
  RingPos ld;                     // next stage to load (runs ahead of the release position)
  int dbg;                        // timing experiments: 1 = skip ldmatrix, 2 = skip the MMAs (results are garbage)
};
// Stage layout (mega_pack_kernel): fragment-major.  The 32 KB of a stage are 64 blocks of 512 bytes, block (w, t, ks) =
// the ldmatrix.x4 fragment of warp w for its m16 tile t and k-step ks, stored in ldmatrix lane order (lane l supplies the
// 16-byte row l of the block: matrix l / 8 = rows (m & 1) * 8.. of the tile, k chunk m >> 1).  Every ldmatrix of a warp is
// one contiguous, conflict-free 512-byte read at an IMMEDIATE offset from one per-lane register.
__device__ __forceinline__ void wfrag_init(WFrag& wf, int cw, int lane) { wf.off = cw * 4096 + lane * 16; }
template <int NSTAGE>
__device__ __forceinline__ void wfrag_load(WFrag& wf, int buf, uint32_t ring_addr, uint64_t* full, bool active) {
  ptx::mbar_wait(&full[wf.ld.stage], wf.ld.phase);
  if (active && !(wf.dbg & 1)) {
    const uint32_t sbase = ring_addr + wf.ld.stage * STAGE_BYTES;
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(sbase + wf.off + (t * 4 + ks) * 512, wf.a[buf][t][ks]);
  }
  wf.ld.template advance<NSTAGE>();
}

// in_proj of this CTA's slice: 192 rows (q | k | v, 64 each) x K = 256 = 12 m16 tiles x 16 k-steps = 192 fragments = exactly
// THREE stages with every warp busy (the row-major [256 x 64] stage format needed four, a quarter of them padding):
//   warp w owns the k / v tile 4 + w over all 16 k-steps (stages 0 and 1: fragment f = k-step 8 s + f) and HALF of the
//   k-steps of q tile w / 2 (stage 2: fragment f = k-step 8 (w % 2) + f); the two warps of a q tile add their halves into the
//   fp32 q vector in shared memory (two commutative additions onto zero: deterministic).
//   acc_kv: rows frow / frow + 8 of tile 4 + w, acc_q: this warp's half of tile w / 2; sequences fs, fs + 1.
template <int NSTAGE>
__device__ __forceinline__ void gemm_qkv(uint8_t* ring, uint64_t* full, RingPos& rp, WFrag& wf, const bf16* act, int pitch, int cw,
                                         int lane, float (&acc_kv)[4], float (&acc_q)[4], unsigned long long*& tr) {
  const uint32_t ring_addr = ptx::smem_u32(ring);
  const bf16* bp = act + (lane >> 2) * pitch + (lane & 3) * 2;
  uint32_t bq[2][8][2];
  auto load_b = [&](int buf, int ks0) {
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      bq[buf][f][0] = *reinterpret_cast<const uint32_t*>(bp + (ks0 + f) * 16);
      bq[buf][f][1] = *reinterpret_cast<const uint32_t*>(bp + (ks0 + f) * 16 + 8);
    }
  };
  float acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) { acc_kv[e] = 0.f; acc_q[e] = 0.f; }
  load_b(0, 0);
  wf.ld = rp;
  wfrag_load<NSTAGE>(wf, 0, ring_addr, full, true);
  MG_TR(tr);
#pragma unroll
  for (int st = 0; st < 3; ++st) {
    if (st + 1 < 3) {
      load_b((st + 1) & 1, st == 0 ? 8 : 8 * (cw & 1));
      wfrag_load<NSTAGE>(wf, (st + 1) & 1, ring_addr, full, true);
    }
    if (!(wf.dbg & 2)) {
      if (st == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { acc_kv[e] += acc2[e]; acc2[e] = 0.f; }
      }
#pragma unroll
      for (int f = 0; f < 8; ++f) {
        float (&dst)[4] = (f & 1) ? acc2 : (st == 2 ? acc_q : acc_kv);
        mma_bf16_16816(dst, wf.a[st & 1][f >> 2][f & 3], bq[st & 1][f][0], bq[st & 1][f][1]);
      }
    }
    ptx::named_bar_arrive(BAR_STAGE_FREE + rp.stage, NCT + 32);
    rp.template advance<NSTAGE>();
    MG_TR(tr);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) acc_q[e] += acc2[e];
}

// ONE stage holding 64 weight rows x K = 256 (4 m16 tiles x 16 k-steps): warp w owns tile w / 2 and the k-half w % 2
// (fragment f = k-step 8 (w % 2) + f); the two warps of a tile add their halves in shared memory.  The tail of the head
// (the last <= 64 vocabulary rows of the slice) goes through this instead of a padded 256-row tile pair.
// PRE: buffer 0 was loaded by the previous GEMM (preload_next of a pair).
template <int NSTAGE, bool PRE>
__device__ __forceinline__ void gemm_khalf(uint8_t* ring, uint64_t* full, RingPos& rp, WFrag& wf, const bf16* act, int pitch, int cw,
                                           int lane, float (&acc)[4]) {
  const bf16* bp = act + (lane >> 2) * pitch + (lane & 3) * 2 + 8 * (cw & 1) * 16;
  uint32_t bq[8][2];
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    bq[f][0] = *reinterpret_cast<const uint32_t*>(bp + f * 16);
    bq[f][1] = *reinterpret_cast<const uint32_t*>(bp + f * 16 + 8);
  }
  if (!PRE) { wf.ld = rp; wfrag_load<NSTAGE>(wf, 0, ptx::smem_u32(ring), full, true); }
  float acc2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] = 0.f;
  if (!(wf.dbg & 2)) {
#pragma unroll
    for (int f = 0; f < 8; ++f) mma_bf16_16816((f & 1) ? acc2 : acc, wf.a[0][f >> 2][f & 3], bq[f][0], bq[f][1]);
  }
  ptx::named_bar_arrive(BAR_STAGE_FREE + rp.stage, NCT + 32);
  rp.template advance<NSTAGE>();
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] += acc2[e];
}

// One GEMM "pair": NKB stages of [256 weight rows x 64 K]; this warp owns weight rows [32 cw, 32 cw + 32)
// (two m16 tiles) and accumulates D[16 rows x 8 sequences] per tile over the stages.
//   act   : activations bf16 [8][pitch] (row = sequence), K contiguous
//   acc   : [2][4] fp32, thread holds rows (lane/4, lane/4 + 8) x sequences ((lane%4)*2, +1) of each tile
// Explicit software pipeline: the eight ldmatrix of stage kb + 1 are issued BEFORE the eight MMAs of stage kb, so the
// shared-memory pipe (256 cycles per 32 KB stage for the whole CTA) and the tensor pipe overlap even though the mbarrier
// waits / arrives pin the instruction order (measured: 275 instead of 390-510 cycles per stage, tools/microbench/mb3.cu).
// PRE: buffer 0 already holds the first stage (loaded by the previous call); preload_next: the first stage of the NEXT GEMM
// is loaded into buffer 0 during the last stage (NKB even) -- the head chains its tile pairs this way.
template <int NKB, int NSTAGE, bool PRE = false>
__device__ __forceinline__ void gemm_pair(uint8_t* ring, uint64_t* full, uint64_t* empty, RingPos& rp, WFrag& wf, const bf16* act,
                                          int pitch, int lane, bool active, bool preload_next, bool next_active,
                                          float (&acc)[2][4], unsigned long long*& tr) {
  const uint32_t ring_addr = ptx::smem_u32(ring);
  const bf16* bp = act + (lane >> 2) * pitch + (lane & 3) * 2;   // B fragments: sequence lane / 4, k = (lane % 4) * 2 + {0, 1, 8, 9}
  uint32_t bq[2][4][2];
  auto load_b = [&](int buf, int kb) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      bq[buf][ks][0] = *reinterpret_cast<const uint32_t*>(bp + (kb * 4 + ks) * 16);
      bq[buf][ks][1] = *reinterpret_cast<const uint32_t*>(bp + (kb * 4 + ks) * 16 + 8);
    }
  };
  float acc2[2][4];                                        // odd k-steps: halves the dependent MMA chains
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[t][e] = 0.f; acc2[t][e] = 0.f; }
  load_b(0, 0);
  if (!PRE) { wf.ld = rp; wfrag_load<NSTAGE>(wf, 0, ring_addr, full, active); }
  MG_TR(tr);
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    if (kb + 1 < NKB) {
      load_b((kb + 1) & 1, kb + 1);
      wfrag_load<NSTAGE>(wf, (kb + 1) & 1, ring_addr, full, active);
    } else if (preload_next) {
      wfrag_load<NSTAGE>(wf, (kb + 1) & 1, ring_addr, full, next_active);
    }
    if (active && !(wf.dbg & 2)) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (ks & 1) mma_bf16_16816(acc2[t], wf.a[kb & 1][t][ks], bq[kb & 1][ks][0], bq[kb & 1][ks][1]);
          else mma_bf16_16816(acc[t], wf.a[kb & 1][t][ks], bq[kb & 1][ks][0], bq[kb & 1][ks][1]);
        }
    }
    // stage kb is free: its ldmatrix have returned (the MMAs above were issued).  A named barrier is the cheap way to tell the
    // producer warp (92 vs 235 cycles per hand-off for an 8-arrival mbarrier, tools/microbench/mb5.cu)
#ifndef MG_MEGA_MBAR_RELEASE
    ptx::named_bar_arrive(BAR_STAGE_FREE + rp.stage, NCT + 32);
#else
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[rp.stage]);
#endif
    rp.template advance<NSTAGE>();
    MG_TR(tr);
  }
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[t][e] += acc2[t][e];
}


template <int SMAX, int NSTAGE, int HD>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NTHREADS, 1)
decode_mega_kernel(const MegaParams p) {
  using L = Smem<SMAX, NSTAGE>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array itself: a round trip through uintptr_t makes the
  // compiler lose the shared address space and turn every LDS / STS into a generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(smem + L::kBars);
  MiscSmem& misc = *reinterpret_cast<MiscSmem*>(smem + L::kMisc);
  float* xs = reinterpret_cast<float*>(smem + L::kX);
  float* qs = reinterpret_cast<float*>(smem + L::kQ);
  bf16* knew = reinterpret_cast<bf16*>(smem + L::kKnew);
  bf16* vnew = reinterpret_cast<bf16*>(smem + L::kVnew);
  bf16* xb = reinterpret_cast<bf16*>(smem + L::kBx);
  bf16* hb = reinterpret_cast<bf16*>(smem + L::kBh);
  bf16* attb = reinterpret_cast<bf16*>(smem + L::kBatt);
  float* part = reinterpret_cast<float*>(smem + L::kPart);
  float* logits = reinterpret_cast<float*>(smem + L::kLogits);
  float* slots = reinterpret_cast<float*>(smem + L::kSlots);
  uint2* cand = reinterpret_cast<uint2*>(smem + L::kCand);
  uint2* local_list = reinterpret_cast<uint2*>(smem + L::kLocal);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L::kHist);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster = static_cast<int>(ptx::cluster_id_x());
  const int S = min(p.S, p.B - cluster * p.S);             // sequences of this cluster (may be <= 0)
  const int b0 = cluster * p.S;                            // first global sequence index
  const int n_layer = p.n_layer, NL = 256 * p.NP + 64 * p.head_tail;   // head rows per CTA, padded to tile pairs (+ tail tiles)
  constexpr int hd = HD;

  // ---------------- one-time setup ----------------
  for (int i = threadIdx.x; i < (L::kSlots - L::kBx) / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + L::kBx)[i] = 0u;
  if (p.dbg_skip_loads)
    for (int i = threadIdx.x; i < L::kBx / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  {
    float* prm = reinterpret_cast<float*>(smem + L::kParams);
    const int r0 = static_cast<int>(rank);
    for (int l = 0; l < n_layer; ++l) {
      const MegaLayer& lw = p.layers[l];
      float* pl = prm + l * kLayerParamFloats;
      for (int i = threadIdx.x; i < 256; i += NTHREADS) {
        pl[P_LN1W + i] = lw.ln1w[i]; pl[P_LN1B + i] = lw.ln1b[i];
        pl[P_LN2W + i] = lw.ln2w[i]; pl[P_LN2B + i] = lw.ln2b[i];
        pl[P_BOUT + i] = lw.b_out[i]; pl[P_B2 + i] = lw.b2[i];
        pl[P_B1 + i] = lw.b1[r0 * HS + i];
        if (i < 192) pl[P_BQKV + i] = lw.b_in[(i >> 6) * D + r0 * FS + (i & 63)];
      }
    }
  }
  for (int i = threadIdx.x; i < CL * KMAX; i += NTHREADS) cand[i] = make_uint2(__float_as_uint(-INFINITY), 0xffffffffu);
  for (int i = threadIdx.x; i < SMAX * FS; i += NTHREADS) qs[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&bars.full[s], 1); ptx::mbar_init(&bars.empty[s], GW); }
    ptx::mbar_init(&bars.step_go, 1);
    ptx::mbar_init(&bars.xchg[0], 1);
    ptx::mbar_init(&bars.xchg[1], 1);
    ptx::mbar_init(&bars.cand, 1);
    ptx::mbar_init(&bars.tok, 1);
    for (int w = 0; w < 8; ++w) ptx::mbar_init(&bars.kvslot[w], 1);
    ptx::fence_mbar_init();
    for (int s = 0; s < 8; ++s) {
      const bool live = s < S;
      misc.tok[s] = live ? p.st.cur_tok[b0 + s] : 0;
      misc.len[s] = live ? p.st.lens[b0 + s] : 0;
      misc.nnew[s] = live ? p.st.n_new[b0 + s] : 0;
      misc.maxnew[s] = live ? p.st.max_new[b0 + s] : 0;
      misc.fin[s] = live ? static_cast<int>(p.st.finished[b0 + s]) : 1;
      misc.outlen[s] = live ? p.st.out_len[b0 + s] : 0;
    }
    for (int l = 0; l < n_layer; ++l) { misc.kvp[l][0] = p.layers[l].kh; misc.kvp[l][1] = p.layers[l].vt; }
  }
  __syncthreads();
  // every CTA of the cluster must have initialised its barriers before any remote st.async lands
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const int n_steps = p.n_steps;
  // optional start offset per cluster group: de-synchronises the clusters' phases (weight streaming from L2 vs K/V streaming
  // from HBM are different shared resources; clusters that run in lock-step hit each of them all at once)
  if (p.stagger_groups > 1 && p.stagger_ns > 0) {
    const uint64_t wait_ns = static_cast<uint64_t>(cluster % p.stagger_groups) * static_cast<uint64_t>(p.stagger_ns);
    const uint64_t t_start = ptx::global_timer_ns();
    while (ptx::global_timer_ns() - t_start < wait_ns) {}
  }

  if (S > 0) {
    if (warp < NPW) {
      // =========================== bulk-copy producer warpgroup ===========================
      ptx::setmaxnreg_dec<40>();
      if (warp == 0) {
        // warp 0: all lanes take part in the named-barrier waits, lane 0 issues the copies
        RingPos rp;
        const int stages_per_step = n_layer * kMegaStagesPerLayer + 4 * p.NP + p.head_tail;
        const uint8_t* src0 = p.packed + static_cast<size_t>(rank) * stages_per_step * STAGE_BYTES;
        int issued = 0;
        for (int step = 0; step < n_steps; ++step) {
          if (p.early_exit && step > 0) {                     // do not stream weights for a step that will not run
            ptx::mbar_wait(&bars.step_go, (step - 1) & 1);
            if (*reinterpret_cast<volatile int*>(&misc.go) == 0) break;
          }
          const uint8_t* src = src0;
          for (int i = 0; i < stages_per_step; ++i, src += STAGE_BYTES, ++issued) {
#ifndef MG_MEGA_MBAR_RELEASE
            if (issued >= NSTAGE) ptx::named_bar_sync(BAR_STAGE_FREE + rp.stage, NCT + 32);
#else
            ptx::mbar_wait(&bars.empty[rp.stage], rp.phase ^ 1);
#endif
            if (lane == 0) {
              if (p.dbg_skip_loads) {
                ptx::mbar_arrive(&bars.full[rp.stage]);
              } else {
                ptx::mbar_arrive_expect_tx(&bars.full[rp.stage], STAGE_BYTES);
                ptx::bulk_load_1d(smem + L::kRing + rp.stage * STAGE_BYTES, src, STAGE_BYTES, &bars.full[rp.stage]);
              }
            }
            rp.template advance<NSTAGE>();
          }
        }
        // match the compute warps' arrivals for the last stages (no refill follows)
#ifdef MG_MEGA_MBAR_RELEASE
        issued = 0;
#endif
        for (int i = 0; i < NSTAGE && i < issued; ++i) {
          ptx::named_bar_sync(BAR_STAGE_FREE + rp.stage, NCT + 32);
          rp.template advance<NSTAGE>();
        }
      }
      __syncwarp();
    } else {
      // =========================== compute warps (two warpgroups) ===========================
      ptx::setmaxnreg_inc<232>();
      const int cw = warp - NPW;                             // 0..7
      const int ct = threadIdx.x - NPW * 32;                 // 0..255
      uint32_t xuse = 0;
      RingPos rp;
      WFrag wf;
      wfrag_init(wf, cw, lane);
      wf.dbg = p.dbg_gemm;
      if (wf.dbg) for (int i = 0; i < 64; ++i) (&wf.a[0][0][0][0])[i] = 0u;
      const SampleParams sp = *p.sp;
      const float scale_log2 = kLog2e / sqrtf(static_cast<float>(hd));
      const float inv_temp = 1.0f / sp.temperature;          // logits / temperature (api_cache.py:169) as one multiply
      const int cph = hd / 8;                                // 16-byte chunks per head (4 or 8)
      constexpr int hd_shift = HD == 64 ? 6 : 5, nh_shift = 6 - hd_shift;   // heads per 64-wide slice: 1 << nh_shift
      // attention work assignment: (sequence, head) pairs dealt round-robin to the warps
      const int n_pairs = S << nh_shift;
      int att_pair = cw, att_wi = 0;
      while (att_pair >= n_pairs) { att_pair -= n_pairs; ++att_wi; }
      const int att_nws = (NCW - att_pair + n_pairs - 1) / n_pairs;
      // work assignment without run-time integer divisions (S <= 4): warp cw serves sequence cw % S as its (cw / S)-th
      // worker; a sequence has NCW / S workers (3, 3, 2 when S == 3)
      auto seq_of = [&](int w) { return S == 3 ? w % 3 : (w & (S - 1)); };
      auto worker_of = [&](int w) { return S == 3 ? w / 3 : (S == 1 ? w : (S == 2 ? w >> 1 : w >> 2)); };
      auto workers = [&](int sq) { return S == 1 ? 8 : (S == 2 ? 4 : (S == 4 ? 2 : (sq < 2 ? 3 : 2))); };
      const int r = static_cast<int>(rank);
      uint8_t* ring = smem + L::kRing;
      // accumulator fragment coordinates: rows frow / frow + 8 of tile t (weight row 32 cw + 16 t + ...), sequences fs, fs + 1
      const int frow = lane >> 2, fs = (lane & 3) * 2;

      auto bar_compute = [&]() { ptx::named_bar_sync(BAR_COMPUTE, NCT); };
      int stamp_id = 0;
      const bool prof_on = p.prof != nullptr && cluster == 0 && rank == 0 && ct == p.prof_thread;
      auto stamp = [&](int step_now) {
        if (prof_on && step_now == p.prof_step && stamp_id < 64) p.prof[stamp_id++] = ptx::global_timer_ns();
      };
      // fine-grained trace (clock64) of layer 1 of the profiled step: p.prof[64..127]
      unsigned long long* tr = nullptr;
      auto fst = [&]() { MG_TR(tr); };
      // Row-wise epilogue of warp s < S over v[8] = features {64 j + 2 lane, +1}: keep the fp32 residual row, LayerNorm it
      // (one-pass mean / E[x^2], fp32; plain cast when w == nullptr) and stage the bf16 GEMM operand.
      auto ln_store = [&](int s, float (&v)[8], const float* __restrict__ w, const float* __restrict__ b) {
        float wv[8], bv[8];                                    // every load ahead of the first store
        if (w) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float2 w2 = *reinterpret_cast<const float2*>(w + 64 * j4 + 2 * lane);
            const float2 b2 = *reinterpret_cast<const float2*>(b + 64 * j4 + 2 * lane);
            wv[2 * j4] = w2.x; wv[2 * j4 + 1] = w2.y; bv[2 * j4] = b2.x; bv[2 * j4 + 1] = b2.y;
          }
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          *reinterpret_cast<float2*>(xs + s * D + 64 * j4 + 2 * lane) = make_float2(v[2 * j4], v[2 * j4 + 1]);
        if (w) {
          float sum = 0.f, sq = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) { sum += v[e]; sq = fmaf(v[e], v[e], sq); }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          const float mean = sum * (1.0f / D);
          const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd * wv[e] + bv[e];
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          *reinterpret_cast<uint32_t*>(xb + s * XP + 64 * j4 + 2 * lane) = pack_bf16(v[2 * j4], v[2 * j4 + 1]);
      };
      // embedding x = tok_emb[tok] + pos_emb[0] (api_cache.py:99 with T == 1) fused with the first LayerNorm
      auto embed_ln = [&](const float* __restrict__ w, const float* __restrict__ b) {
        if (cw < S) {
          const int s = cw;
          const bf16* te = p.tok_emb + static_cast<size_t>(misc.tok[s]) * D;
          float v[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const __nv_bfloat162 t2 = *reinterpret_cast<const __nv_bfloat162*>(te + 64 * j4 + 2 * lane);
            const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(p.pos_emb + 64 * j4 + 2 * lane);
            v[2 * j4] = __bfloat162float(t2.x) + __bfloat162float(p2.x);
            v[2 * j4 + 1] = __bfloat162float(t2.y) + __bfloat162float(p2.y);
          }
          ln_store(s, v, w, b);
        }
        if (p.head_tail && ct >= NCT - 64) {
          // the head's tail tiles ADD their two k-halves into the logits: real rows start from 0, padding rows stay -inf
          const int lr = 256 * p.NP + (ct - (NCT - 64));
          const float init = (lr < p.VS && r * p.VS + lr < p.V) ? 0.f : -INFINITY;
          for (int s = 0; s < S; ++s) logits[s * NL + lr] = init;
        }
        bar_compute();
      };
      // all-gather of the row-parallel partial sums: this warp's 32 output features x S sequences
      auto exchange_send = [&](const float (&acc)[2][4]) {
        const int buf = xuse & 1;
        const uint32_t bar_addr = ptx::smem_u32(&bars.xchg[buf]);
        if (fs < SMAX) {
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              const int f = cw * 32 + t * 16 + frow + h8 * 8;
              const float v0 = acc[t][h8 * 2], v1 = acc[t][h8 * 2 + 1];
              float* my = slots + ((buf * CL + r) * D + f) * SMAX + fs;
              my[0] = v0;
              my[1] = v1;
              const uint32_t my_addr = ptx::smem_u32(my);
#pragma unroll
              for (int q = 1; q < CL; ++q) {
                const uint32_t peer = (rank + q) % CL;
                ptx::st_async_v2b32(ptx::map_to_cta(my_addr, peer), __float_as_uint(v0), __float_as_uint(v1),
                                    ptx::map_to_cta(bar_addr, peer));
              }
            }
        }
      };
      // receive the peers' partial sums, x += sum + bias (fixed order: identical in every CTA of the cluster), then the
      // NEXT LayerNorm (or the plain cast in front of the head) by the same warp: one pass, two barriers
      auto exchange_finish_ln = [&](const float* __restrict__ bias, const float* __restrict__ w, const float* __restrict__ b) {
        const int buf = xuse & 1;
        if (cw == 0) ptx::mbar_wait_spin(&bars.xchg[buf], (xuse >> 1) & 1);   // one warp polls (see mbar_wait_spin) ...
        fst();                                                // peers' partial sums have landed
        bar_compute();                                        // ... and releases the others; local slot writes visible
        if (ct == 0) ptx::mbar_arrive_expect_tx(&bars.xchg[buf], (CL - 1) * D * SMAX * 4);   // re-arm for use + 2
        if (cw < S) {
          const int s = cw;
          float v[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int f0 = 64 * j4 + 2 * lane;
            const float2 xv = *reinterpret_cast<const float2*>(xs + s * D + f0);
            const float2 bv = *reinterpret_cast<const float2*>(bias + f0);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int src = 0; src < CL; ++src) {
              const float* sl = slots + ((buf * CL + src) * D + f0) * SMAX;
              a0 += sl[s];
              a1 += sl[SMAX + s];
            }
            v[2 * j4] = xv.x + (a0 + bv.x);
            v[2 * j4 + 1] = xv.y + (a1 + bv.y);
          }
          ln_store(s, v, w, b);
        }
        ++xuse;
        bar_compute();
      };

      if (ct == 0) {                                          // arm the receive barriers once
        ptx::mbar_arrive_expect_tx(&bars.xchg[0], (CL - 1) * D * SMAX * 4);
        ptx::mbar_arrive_expect_tx(&bars.xchg[1], (CL - 1) * D * SMAX * 4);
      }
      uint32_t cand_use = 0, tok_use = 0;
      uint8_t* const kvs_slot = smem + L::kKvSlot + cw * 4096;     // this warp's K/V staging slot (MG_MEGA_KVSLOT)
      uint32_t kvs_phase = 0;

      for (int step = 0; step < n_steps; ++step) {
        // ---- embedding + LayerNorm 1 of the first block ----
        embed_ln(reinterpret_cast<const float*>(smem + L::kParams) + P_LN1W, reinterpret_cast<const float*>(smem + L::kParams) + P_LN1B);
        stamp(step);                                                        // 0: embedding done
        for (int l = 0; l < n_layer; ++l) {
          const float* pl = reinterpret_cast<const float*>(smem + L::kParams) + l * kLayerParamFloats;
#ifdef MG_MEGA_TRACE
          tr = (prof_on && step == p.prof_step && l == 1) ? p.prof + 64 : nullptr;
#endif
          fst();                                                            // t0: layer start
          // ---- attention prefetch: this warp's first 32-key block does not depend on this step's q / k / v ----
#ifdef MG_MEGA_NO_ATTN_PRE
          constexpr bool kAttnPre = false;
#else
          constexpr bool kAttnPre = HD == 32;
#endif
          uint4 kq0[4][HD / 32], vq0[HD / 8];
          const size_t kv_seq = static_cast<size_t>(CL) * p.Tvt * FS;        // K cache elements per sequence
          const size_t vt_seq = static_cast<size_t>(CL) * p.Tvt * FS;        // V cache elements per sequence
          bf16* const kbase = misc.kvp[l][0] + (static_cast<size_t>(b0) * CL + r) * p.Tvt * FS;    // [head][Tvt][hd]
          bf16* const vbase = misc.kvp[l][1] + (static_cast<size_t>(b0) * CL + r) * p.Tvt * FS;    // [head][T / 32][hd][32]
          const int at_s = att_pair >> nh_shift, at_h = att_pair & ((1 << nh_shift) - 1);
          const int at_len = misc.fin[at_s] ? 0 : misc.len[at_s];
          const bf16* at_kh = kbase + at_s * kv_seq + static_cast<size_t>(at_h) * p.Tvt * hd;
          const bf16* at_vt = vbase + at_s * vt_seq + static_cast<size_t>(at_h) * p.Tvt * hd;
          if (kAttnPre && (att_wi << 5) < at_len) attn_load_block<HD>(at_kh, at_vt, at_len, att_wi, lane, kq0, vq0);
          // third block of this warp: into its staging slot by bulk copy, also ahead of the QKV GEMM
          bool kvs_issued = false;
          if (L::kHasKvSlot && HD == 32 && kAttnPre && ((att_wi + 2 * att_nws) << 5) < at_len) {
            kvslot_issue<HD>(at_kh, at_vt, att_wi + 2 * att_nws, lane, kvs_slot, &bars.kvslot[cw]);
            kvs_issued = true;
          }
          // ---- QKV (LN1 was applied by the previous epilogue): stage rows 0..63 = q, 64..127 = k, 128..191 = v slice ----
          if (cw < GW) {
            // biases ahead of the MMAs (no load behind a store in the epilogue): k / v tile 4 + cw, q tile cw / 2
            const int f_kv = (cw & 3) * 16 + frow, f_q = (cw >> 1) * 16 + frow;   // features inside the 64-wide slice (+ 8 h8)
            const int part_kv = 1 + (cw >> 2);                                 // 1 = k (warps 0..3), 2 = v (warps 4..7)
            float bkv[2], bqq[2];
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              bkv[h8] = pl[P_BQKV + part_kv * FS + f_kv + h8 * 8];
              bqq[h8] = (cw & 1) ? 0.f : pl[P_BQKV + f_q + h8 * 8];            // the bias goes in with the first k-half
            }
            float acc_kv[4], acc_q[4];
            gemm_qkv<NSTAGE>(ring, bars.full, rp, wf, xb, XP, cw, lane, acc_kv, acc_q, tr);
            if (fs < S) {
              bf16* kvdst = part_kv == 1 ? knew : vnew;
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int s = fs + e;
                  if (s < S) {
                    kvdst[s * FS + f_kv + h8 * 8] = __float2bfloat16_rn(acc_kv[h8 * 2 + e] + bkv[h8]);
                    atomicAdd(qs + s * FS + f_q + h8 * 8, (acc_q[h8 * 2 + e] + bqq[h8]) * scale_log2);
                  }
                }
            }
          }
          fst();                                                            // QKV epilogue
          bar_compute();
          stamp(step);                                                      // +1: QKV done
          // ---- append the new K row / V^T column (api_cache.py:66-67) + flash-decoding over this CTA's slice ----
          if (cw < S && lane < 8 && !misc.fin[cw]) {
            const int s = cw, c = lane, h = (8 * c) >> hd_shift, d0 = (8 * c) & (hd - 1);
            bf16* dst = kbase + s * kv_seq + (static_cast<size_t>(h) * p.Tvt + misc.len[s]) * hd + d0;
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(knew + s * FS + c * 8);
          }
          if (ct < S * FS) {
            const int s = ct >> 6, f = ct & 63, h = f >> hd_shift, d = f & (hd - 1);
            if (!misc.fin[s]) {
              const int key = misc.len[s], ki = key & 31;
              const int pos = 8 * ((ki >> 1) & 3) + 2 * (ki >> 3) + (ki & 1);          // see attn_tc
              vbase[s * vt_seq + (static_cast<size_t>(h) * (p.Tvt >> 5) + (key >> 5)) * (hd * 32) + d * 32 + pos] = vnew[s * FS + f];
            }
          }
          fst();                                                            // append + prefetch issued
          {
            // warp cw is the att_wi-th worker of (sequence, head) pair att_pair
            const int s = at_s, h = at_h;
            attn_tc<HD, kAttnPre>(at_kh, at_vt, at_len, att_wi, att_nws, lane, qs + s * FS + h * HD, knew + s * FS + h * HD,
                                  vnew + s * FS + h * HD, att_wi == 0, part + cw * 68, kq0, vq0,
                                  L::kHasKvSlot ? kvs_slot : nullptr, &bars.kvslot[cw], &kvs_phase, kvs_issued);
          }
          fst();                                                            // attention stream done (this warp)
          bar_compute();
          fst();
          // final merge over the workers of the (sequence, head) pair (worker 0 always holds the new token: finite maximum)
          if (ct < S * FS) {
            const int mg_s = ct >> 6, mg_f = ct & 63, mg_h = mg_f >> hd_shift, mg_d = mg_f & (hd - 1);
            const int pair = (mg_s << nh_shift) + mg_h;
            float mw[NCW], lw_[NCW], ow[NCW];
#pragma unroll
            for (int k = 0; k < NCW; ++k) {                                   // all loads first, then the arithmetic
              const int w = pair + k * n_pairs;
              const bool on = w < NCW;
              const float* pp = part + (on ? w : 0) * 68;
              mw[k] = on ? pp[64] : -INFINITY;
              lw_[k] = on ? pp[65] : 0.f;
              ow[k] = on ? pp[mg_d] : 0.f;
            }
            float M = mw[0];
#pragma unroll
            for (int k = 1; k < NCW; ++k) M = fmaxf(M, mw[k]);
            float Lsum = 0.f, o = 0.f;
#pragma unroll
            for (int k = 0; k < NCW; ++k) {
              const float fw = fast_exp2(mw[k] - M);                        // exp2(-inf) = 0 for an idle worker
              Lsum = fmaf(lw_[k], fw, Lsum);
              o = fmaf(ow[k], fw, o);
            }
            attb[mg_s * AP + mg_f] = __float2bfloat16_rn(__fdividef(o, Lsum));
            qs[mg_s * FS + mg_f] = 0.f;                                     // the next in_proj ADDS its two k-halves into q
          }
          fst();                                                            // merge done
          bar_compute();
          stamp(step);                                                      // +2: attention done
          // ---- out_proj (row-parallel over this CTA's 64 attention features) -> exchange -> x += attn ----
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<1, NSTAGE>(ring, bars.full, bars.empty, rp, wf, attb, AP, lane, true, false, false, acc, tr);
            exchange_send(acc);
          }
          fst();                                                            // out_proj sent
          exchange_finish_ln(pl + P_BOUT, pl + P_LN2W, pl + P_LN2B);
          fst();
          stamp(step);                                                      // +3: out_proj + exchange + LN2 done
          // ---- MLP1 (+GELU, this CTA's 256 hidden units) -> MLP2 (row-parallel) -> exchange ----
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<4, NSTAGE>(ring, bars.full, bars.empty, rp, wf, xb, XP, lane, true, false, false, acc, tr);
            // bias + exact-erf GELU: the valid accumulators sit in quad lanes 0 .. SMAX/2-1 (two sequences each); spread
            // them over the four lanes of the quad so that every lane evaluates erff for 2 (SMAX = 2) or 4 values
            {
              constexpr int NSRC = SMAX / 2, PER = 2 * NSRC;
              const int ql = lane & 3, qbase = lane & ~3;
              float mine[PER], b1v[PER];
#pragma unroll
              for (int idx = 0; idx < PER; ++idx) {
                const int i = (ql * PER + idx) & 7;
                b1v[idx] = pl[P_B1 + cw * 32 + (i >> 2) * 16 + frow + ((i >> 1) & 1) * 8];
              }
#pragma unroll
              for (int src = 0; src < NSRC; ++src)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float got = __shfl_sync(0xffffffffu, acc[i >> 2][i & 3], qbase + src);
                  const int v = src * 8 + i;                               // value id inside the quad: owner lane v / PER
                  if (v / PER == ql) mine[v % PER] = got;
                }
#pragma unroll
              for (int idx = 0; idx < PER; ++idx) {
                const int v = ql * PER + idx, src = v >> 3, i = v & 7;
                const int t = i >> 2, h8 = (i >> 1) & 1, sq = 2 * src + (i & 1);
                const int j = cw * 32 + t * 16 + frow + h8 * 8;            // hidden unit inside this CTA's slice
                if (sq < S) hb[sq * XP + j] = __float2bfloat16_rn(gelu_erf_f(mine[idx] + b1v[idx]));
              }
            }
          }
          fst();                                                            // GELU epilogue
          bar_compute();
          stamp(step);                                                      // +4: MLP1 done
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<4, NSTAGE>(ring, bars.full, bars.empty, rp, wf, hb, XP, lane, true, false, false, acc, tr);
            exchange_send(acc);
          }
          fst();                                                            // mlp.2 sent
          {
            // ... + LayerNorm 1 of the next block, or the plain cast in front of the head (no final LayerNorm, api_cache.py:105)
            const float* pn = pl + kLayerParamFloats;
            const bool more = l + 1 < n_layer;
            exchange_finish_ln(pl + P_B2, more ? pn + P_LN1W : nullptr, more ? pn + P_LN1B : nullptr);
          }
          fst();
          stamp(step);                                                      // +5: MLP2 + exchange done
        }
        // ---- head: logits of this CTA's vocabulary slice ----
        {
          const int v_lo = r * p.VS, v_hi = min(p.V, (r + 1) * p.VS);
          auto head_pair = [&](int pr, bool first) {
            float hbv[2][2];                                   // head bias of this thread's 4 rows, loaded ahead of the MMAs
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                const int vr = v_lo + pr * 256 + cw * 32 + t * 16 + frow + h8 * 8;
                hbv[t][h8] = vr < v_hi ? __ldg(p.head_b + vr) : 0.f;
              }
            float acc[2][4];
            const bool more = pr + 1 < p.NP + p.head_tail;         // another pair, or the single-stage tail, follows
            if (first) gemm_pair<4, NSTAGE, false>(ring, bars.full, bars.empty, rp, wf, xb, XP, lane, true, more, true, acc, tr);
            else gemm_pair<4, NSTAGE, true>(ring, bars.full, bars.empty, rp, wf, xb, XP, lane, true, more, true, acc, tr);
            if (fs < S) {
              const int lr0 = pr * 256 + cw * 32 + frow;                     // row inside the slice: lr0 + 16 t + 8 h8
#pragma unroll
              for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  const int lr = lr0 + t * 16 + h8 * 8;
                  const bool ok = v_lo + lr < v_hi;
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (fs + e < S) logits[(fs + e) * NL + lr] = ok ? (acc[t][h8 * 2 + e] + hbv[t][h8]) * inv_temp : -INFINITY;
                }
              const int dslot = !p.dbg_logits ? -1 : (p.dbg_slot ? p.dbg_slot[step] : step);   // mg_step_logits_at: selected steps only
              if (dslot >= 0) {                                              // parity / debug path only (mg_step_logits)
                float* dl = p.dbg_logits + (static_cast<size_t>(dslot) * p.B + b0 + fs) * p.V + v_lo + lr0;
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                  for (int h8 = 0; h8 < 2; ++h8)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                      if (fs + e < S && v_lo + lr0 + t * 16 + h8 * 8 < v_hi)
                        dl[static_cast<size_t>(e) * p.V + t * 16 + h8 * 8] = acc[t][h8 * 2 + e] + hbv[t][h8];
              }
            }
          };
          if (cw < GW) {
            if (p.NP > 0) head_pair(0, true);
            for (int pr = 1; pr < p.NP; ++pr) head_pair(pr, false);
            if (p.head_tail) {
              // tail: tile cw / 2 of the last 64 rows, k-half cw % 2; the bias goes in with the first half
              const int lr0 = 256 * p.NP + (cw >> 1) * 16 + frow;
              float tb[2];
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                const int vr = v_lo + lr0 + h8 * 8;
                tb[h8] = (!(cw & 1) && vr < v_hi) ? __ldg(p.head_b + vr) : 0.f;
              }
              float acc[4];
              if (p.NP > 0) gemm_khalf<NSTAGE, true>(ring, bars.full, rp, wf, xb, XP, cw, lane, acc);
              else gemm_khalf<NSTAGE, false>(ring, bars.full, rp, wf, xb, XP, cw, lane, acc);
              if (fs < S) {
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8)
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (fs + e < S && v_lo + lr0 + h8 * 8 < v_hi)
                      atomicAdd(logits + (fs + e) * NL + lr0 + h8 * 8, (acc[h8 * 2 + e] + tb[h8]) * inv_temp);
              }
            }
          }
        }
        bar_compute();
        const int dslot_t = !p.dbg_logits ? -1 : (p.dbg_slot ? p.dbg_slot[step] : step);
        if (dslot_t >= 0 && p.head_tail && ct < 64) {         // parity / debug path: raw logits of the tail rows
          const int lr = 256 * p.NP + ct, vr = r * p.VS + lr;
          if (lr < p.VS && vr < p.V)
            for (int s = 0; s < S; ++s)
              p.dbg_logits[(static_cast<size_t>(dslot_t) * p.B + b0 + s) * p.V + vr] = logits[s * NL + lr] * sp.temperature;
        }
        stamp(step);                                                        // head done

        // ---- sampler (api_cache.py:169-181): local top-k -> owner CTA ranks, draws, broadcasts ----
        // The compute warps are dealt to the sequences (warp cw -> sequence cw % S); each group radix-selects
        // the k largest logits of this CTA's vocabulary slice with warp-aggregated histogram updates.
        const int k = sp.top_k;
#ifdef MG_MEGA_TRACE
        tr = (prof_on && step == p.prof_step) ? p.prof + 96 : nullptr;
#endif
        fst();                                                              // sampler start
        if (cw < GW) {
          const int s = seq_of(cw), wi = worker_of(cw), nws = workers(s);        // GW == NCW
          const int gt = wi * 32 + lane, gn = nws * 32;       // thread index / count inside the group
          const uint32_t gbar = 3 + s;                        // named barrier of the group
          float* z = logits + s * NL;
          uint32_t* gh = hist + s * 256;
          uint2* gl = local_list + s * KMAX;
          volatile int* g_rem = &misc.sel[s][0];
          volatile int* g_cnt = &misc.sel[s][1];
          volatile int* g_pre = &misc.sel[s][2];
          volatile int* g_exact = &misc.sel[s][3];
          (void)g_exact;
          // Fast path: the k-th largest of the per-thread maxima is a lower bound of the k-th largest logit, so
          // everything >= it is a (small) superset of the local top-k; rank that superset exactly.
          float* maxima = reinterpret_cast<float*>(gh);        // [gn] (the radix histogram's storage)
          uint2* clist = reinterpret_cast<uint2*>(part) + s * kCandCap;
          bool overflow = k > gn;
          if (k == 1) {
            // greedy (the reference expresses it as top_k = 1): arg-max of the slice, value descending / index ascending
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = gt; i < NL; i += gn) {
              const float zi = z[i];
              if (zi > bv) { bv = zi; bi = i; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) clist[wi] = make_uint2(__float_as_uint(bv), static_cast<uint32_t>(bi));
            ptx::named_bar_sync(gbar, gn);
            if (gt == 0) {
              for (int w = 1; w < nws; ++w) {
                const uint2 o = clist[w];
                const float ov = __uint_as_float(o.x);
                if (ov > bv || (ov == bv && static_cast<int>(o.y) < bi)) { bv = ov; bi = static_cast<int>(o.y); }
              }
              gl[0] = make_uint2(__float_as_uint(bv), static_cast<uint32_t>(r * p.VS + bi));
            }
            overflow = false;
          } else if (!overflow) {
            // every pass below keeps several independent shared-memory loads in flight (these loops are pure latency)
            float m_t;
            {
              float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY;
              int i = gt;
#pragma unroll 2
              for (; i + 2 * gn < NL; i += 3 * gn) {
                const float a = z[i], b = z[i + gn], c3 = z[i + 2 * gn];
                m0 = fmaxf(m0, a); m1 = fmaxf(m1, b); m2 = fmaxf(m2, c3);
              }
              for (; i < NL; i += gn) m0 = fmaxf(m0, z[i]);
              m_t = fmaxf(m0, fmaxf(m1, m2));
            }
            // rank the maxima on UNIQUE packed keys (monotonic float key with its low 8 bits replaced by the thread index):
            // one unsigned compare per pair, no tie-break.  The threshold is the k-th key with the index bits cleared, i.e.
            // a float <= that thread's maximum: at least k maxima (hence >= k logits) are >= it, so the superset stays exact.
            const uint32_t pk = (float_key(m_t) & 0xffffff00u) | static_cast<uint32_t>(gt);      // gn <= 256
            uint32_t* pkeys = reinterpret_cast<uint32_t*>(maxima);
            pkeys[gt] = pk;
            if (gt == 0) *g_exact = 0;
            fst();                                                          // per-thread maxima
            ptx::named_bar_sync(gbar, gn);
            fst();
            int rk = 0, rk2 = 0;
#pragma unroll 4
            for (int u = 0; u < gn; u += 8) {                               // gn is a multiple of 32
              const uint4 mu = *reinterpret_cast<const uint4*>(pkeys + u);
              const uint4 mv = *reinterpret_cast<const uint4*>(pkeys + u + 4);
              rk += (mu.x > pk) + (mu.y > pk) + (mu.z > pk) + (mu.w > pk);
              rk2 += (mv.x > pk) + (mv.y > pk) + (mv.z > pk) + (mv.w > pk);
            }
            if (rk + rk2 == k - 1) {
              const uint32_t tk = pk & 0xffffff00u;                        // inverse of float_key; keys of -inf / NaN clamp to -inf
              const float tf = tk <= 0x007fffffu ? -INFINITY : __uint_as_float((tk & 0x80000000u) ? (tk & 0x7fffffffu) : ~tk);
              *reinterpret_cast<volatile float*>(g_pre) = tf;
            }
            fst();                                                          // rank among the maxima
            ptx::named_bar_sync(gbar, gn);
            fst();
            const float tau = *reinterpret_cast<volatile float*>(g_pre);
            // gather everything >= tau without atomics: per-thread bit mask (at most NL / gn <= 64 elements per thread),
            // warp scan of the counts, warp totals through shared memory, then only the set bits are revisited
            unsigned long long take = 0ull;
            {
              int i = gt, j = 0;
#pragma unroll 2
              for (; i + 2 * gn < NL; i += 3 * gn, j += 3) {
                const float a = z[i], b = z[i + gn], c3 = z[i + 2 * gn];
                take |= (a >= tau ? 1ull : 0ull) << j;
                take |= (b >= tau ? 2ull : 0ull) << j;
                take |= (c3 >= tau ? 4ull : 0ull) << j;
              }
              for (; i < NL; i += gn, ++j) take |= (z[i] >= tau ? 1ull : 0ull) << j;
            }
            const int cnt = __popcll(take);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int up = __shfl_up_sync(0xffffffffu, inc, o);
              if (lane >= o) inc += up;
            }
            if (lane == 31) misc.wtot[s][wi] = inc;
            fst();                                                          // count + warp scan
            ptx::named_bar_sync(gbar, gn);
            fst();
            int pos = inc - cnt, total = 0;
#pragma unroll
            for (int w = 0; w < NCW; ++w) {
              const int wt = w < nws ? misc.wtot[s][w] : 0;
              pos += w < wi ? wt : 0;
              total += wt;
            }
            while (take) {
              const int j = __ffsll(static_cast<long long>(take)) - 1;
              take &= take - 1;
              const int i = gt + j * gn;
              if (pos < kCandCap) clist[pos] = make_uint2(__float_as_uint(z[i]), static_cast<uint32_t>(r * p.VS + i));
              ++pos;
            }
            if (gt == 0) *g_cnt = total;
            fst();                                                          // candidates written
            ptx::named_bar_sync(gbar, gn);
            fst();
            const int c = *g_cnt;
            overflow = c > kCandCap || c < k;
            if (!overflow) {
              const int lg = (c <= 64 && gn >= 128) ? (gn >= 256 ? 2 : 1) : 0;   // 4 / 2 / 1 threads per candidate (uniform)
              if (lg) {
                // c <= 64 candidates, gn >> lg == 64 slots: the inner loop is split over the 2 / 4 neighbouring lanes of a
                // candidate and the partial ranks are added by shuffle (the loop is a pure latency chain of c loads)
                const int j = gt >> lg, part = gt & ((1 << lg) - 1);
                const int chunk = (c + (1 << lg) - 1) >> lg, i0 = part * chunk, i1 = min(c, i0 + chunk);
                const uint2 mine = clist[min(j, c - 1)];
                const float mv = __uint_as_float(mine.x);
                int rank_j = 0;
#pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                  const uint2 o = clist[i];
                  const float ov = __uint_as_float(o.x);
                  rank_j += (ov > mv) || (ov == mv && o.y < mine.y);
                }
                rank_j += __shfl_xor_sync(0xffffffffu, rank_j, 1);
                if (lg == 2) rank_j += __shfl_xor_sync(0xffffffffu, rank_j, 2);
                if (part == 0 && j < c && rank_j < k) gl[rank_j] = mine;   // sorted: value descending, index ascending
              } else {
                for (int j = gt; j < c; j += gn) {
                  const uint2 mine = clist[j];
                  const float mv = __uint_as_float(mine.x);
                  int rank_j = 0;
#pragma unroll 4
                  for (int i = 0; i < c; ++i) {
                    const uint2 o = clist[i];
                    const float ov = __uint_as_float(o.x);
                    rank_j += (ov > mv) || (ov == mv && o.y < mine.y);
                  }
                  if (rank_j < k) gl[rank_j] = mine;          // sorted: value descending, index ascending
                }
              }
            }
            ptx::named_bar_sync(gbar, gn);
          }
          if (overflow) {
            // Exact fallback (k larger than the group, or a degenerate / heavily tied slice): 8-bit radix select.
            uint32_t prefix = 0, mask = 0;
            if (gt == 0) { *g_rem = k; *g_cnt = 0; *g_exact = 0; }
            if (gt < KMAX) gl[gt] = make_uint2(__float_as_uint(-INFINITY), 0xffffffffu);
            bool exact = false;
            for (int pass = 3; pass >= 0 && !exact; --pass) {
              const int shift = pass * 8;
              for (int i = gt; i < 256; i += gn) gh[i] = 0;
              ptx::named_bar_sync(gbar, gn);
              for (int i0 = 0; i0 < NL; i0 += gn) {
                const int i = i0 + gt;
                uint32_t bin = 256;                              // sentinel: not a candidate any more
                if (i < NL) {
                  const uint32_t u = float_key(z[i]);
                  if ((u & mask) == prefix) bin = (u >> shift) & 255u;
                }
                const unsigned same = __match_any_sync(0xffffffffu, bin);
                if (bin < 256 && (same & ((1u << lane) - 1)) == 0) atomicAdd(&gh[bin], static_cast<uint32_t>(__popc(same)));
              }
              ptx::named_bar_sync(gbar, gn);
              if (wi == 0) {
                const int remaining = *g_rem;
                int hb[8], mine = 0;
  #pragma unroll
                for (int j = 0; j < 8; ++j) { hb[j] = static_cast<int>(gh[255 - (lane * 8 + j)]); mine += hb[j]; }
                int inc = mine;
  #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                  const int t = __shfl_up_sync(0xffffffffu, inc, o);
                  if (lane >= o) inc += t;
                }
                const int before = inc - mine;
                const bool here = before < remaining && inc >= remaining;
                const unsigned who = __ballot_sync(0xffffffffu, here);
                if (here && (who & ((1u << lane) - 1)) == 0) {
                  int cum = before, j = 0;
                  for (; j < 7; ++j) { if (cum + hb[j] >= remaining) break; cum += hb[j]; }
                  *g_rem = remaining - cum;
                  *g_pre = static_cast<int>(prefix | (static_cast<uint32_t>(255 - (lane * 8 + j)) << shift));
                  *g_exact = (hb[j] == remaining - cum) ? 1 : 0;   // the whole bin is needed: no finer pass required
                }
              }
              ptx::named_bar_sync(gbar, gn);
              prefix = static_cast<uint32_t>(*g_pre);
              mask |= 255u << shift;
              exact = *g_exact != 0;
            }
            // gather: masked key above the prefix, or equal to it (all of them when `exact`, else lowest index first)
            const int need_eq = *g_rem;
            for (int i0 = 0; i0 < NL; i0 += gn) {
              const int i = i0 + gt;
              bool take = false;
              float zi = 0.f;
              if (i < NL) {
                zi = z[i];
                const uint32_t um = float_key(zi) & mask;
                take = um > prefix || (exact && um == prefix);
              }
              const unsigned bal = __ballot_sync(0xffffffffu, take);
              int base_pos = 0;
              if (lane == 0 && bal) base_pos = atomicAdd(const_cast<int*>(g_cnt), __popc(bal));
              base_pos = __shfl_sync(0xffffffffu, base_pos, 0);
              if (take) gl[base_pos + __popc(bal & ((1u << lane) - 1))] = make_uint2(__float_as_uint(zi), static_cast<uint32_t>(r * p.VS + i));
            }
            ptx::named_bar_sync(gbar, gn);
            if (!exact && wi == 0) {                             // ties at a full 32-bit threshold: index order, one warp
              int taken = 0;
              const int base_cnt = *g_cnt;
              for (int i0 = 0; i0 < NL && taken < need_eq; i0 += 32) {
                const int i = i0 + lane;
                const bool eq = i < NL && float_key(z[i]) == prefix;
                const unsigned bal = __ballot_sync(0xffffffffu, eq);
                const int my = taken + __popc(bal & ((1u << lane) - 1));
                if (eq && my < need_eq) gl[base_cnt + my] = make_uint2(__float_as_uint(z[i]), static_cast<uint32_t>(r * p.VS + i));
                taken += __popc(bal);
              }
            }
            ptx::named_bar_sync(gbar, gn);
            // the owner merges sorted lists: sort the k gathered entries (value descending, index ascending)
            for (int j = gt; j < k; j += gn) {
              const uint2 mine = gl[j];
              const float mv = __uint_as_float(mine.x);
              int rank_j = 0;
              for (int i = 0; i < k; ++i) {
                const uint2 o = gl[i];
                const float ov = __uint_as_float(o.x);
                rank_j += (ov > mv) || (ov == mv && (o.y < mine.y || (o.y == mine.y && i < j)));
              }
              clist[rank_j] = mine;
            }
            ptx::named_bar_sync(gbar, gn);
            for (int j = gt; j < k; j += gn) gl[j] = clist[j];
            ptx::named_bar_sync(gbar, gn);
          }
          // ship the k local candidates to the owner CTA of sequence s
          const uint32_t owner = static_cast<uint32_t>(s % CL);
          for (int j = gt; j < k; j += gn) {
            const uint2 e = gl[j];
            if (owner == rank) {
              cand[r * KMAX + j] = e;
            } else {
              const uint32_t ra = ptx::map_to_cta(ptx::smem_u32(&cand[r * KMAX + j]), owner);
              const uint32_t rb = ptx::map_to_cta(ptx::smem_u32(&bars.cand), owner);
              ptx::st_async_v2b32(ra, e.x, e.y, rb);
            }
          }
        }
        fst();                                                              // ranked + shipped
        bar_compute();
        fst();
        stamp(step);                                                        // local top-k done
        // ---- owner: merge CL x k candidates, draw, publish ----
        if (r < S) {
          const int s = r;
          if (ct == 0) ptx::mbar_arrive_expect_tx(&bars.cand, (CL - 1) * k * 8);
          // the uniform draw depends on nothing the wait delivers: ten Philox rounds hidden behind the DSMEM latency
          const float u01 = (cw == 0 && k != 1)
                                ? philox_uniform(sp.seed, sp.seq_base + static_cast<uint64_t>(p.st.seq_idx ? p.st.seq_idx[b0 + s] : b0 + s), static_cast<uint32_t>(misc.nnew[s]))
                                : 0.f;
          // every thread observes the arrival itself (greedy: only warp 0 needs the data): no CTA barrier behind the wait --
          // the local candidates became visible at the barrier above, the remote ones through the mbarrier
          if (k != 1 || cw == 0) ptx::mbar_wait_spin(&bars.cand, cand_use & 1);
          fst();                                                            // candidates of the peers have landed
          ++cand_use;
          int tok = 0;
          if (k == 1) {
            // greedy: the best of the CL slice maxima (the local one became visible at the barrier above)
            if (cw == 0) {
              const uint2 e = cand[(lane < CL ? lane : 0) * KMAX];
              float bv = lane < CL ? __uint_as_float(e.x) : -INFINITY;
              int bi = lane < CL ? static_cast<int>(e.y) : 0x7fffffff;
#pragma unroll
              for (int o = 2; o; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
              }
              tok = __shfl_sync(0xffffffffu, bi, 0);
            }
          } else {
            fst();                                                          // owner: merge starts
            // every CTA's list arrives sorted (value descending, index ascending): the global rank of an entry is its own
            // position plus, for each other list, the number of entries that precede it (binary search):
            // one thread per candidate slot (list = ct / KMAX, position = ct % KMAX); the lists are padded to KMAX entries
            // with (-inf, 0xffffffff) sentinels at kernel start and the senders only ever write the first k, so the three
            // searches are branch-free, have a fixed depth of 6 and advance in lock-step (their dependent loads overlap)
            static_assert(KMAX == 64 && CL * KMAX == NCT, "merge: one thread per candidate slot, 6-step searches");
            const int src = ct >> 6, j = ct & 63;
            const bool live = j < k;
            const uint2 mine = cand[src * KMAX + j];
            const float mv = __uint_as_float(mine.x);
            int pos[CL];
#pragma unroll
            for (int si = 0; si < CL; ++si) pos[si] = 0;                     // entries of list si that precede `mine`
#pragma unroll
            for (int stp = 32; stp >= 1; stp >>= 1) {
#pragma unroll
              for (int si = 0; si < CL; ++si) {
                const uint2 o = cand[si * KMAX + pos[si] + stp - 1];
                const float ov = __uint_as_float(o.x);
                const bool before = (ov > mv) || (ov == mv && (o.y < mine.y || (o.y == mine.y && si < src)));
                pos[si] += before ? stp : 0;
              }
            }
            fst();                                                          // searches done
            int rk = j;
#pragma unroll
            for (int si = 0; si < CL; ++si) {
              // the 6-step search counts at most 63 predecessors: a full list (k == KMAX) whose last entry precedes has 64
              const uint2 o = cand[si * KMAX + KMAX - 1];
              const float ov = __uint_as_float(o.x);
              const bool all = (ov > mv) || (ov == mv && (o.y < mine.y || (o.y == mine.y && si < src)));
              rk += si == src ? 0 : (all ? KMAX : pos[si]);
            }
            const int n = live ? 1 : 0;
            uint2* sorted = local_list + SMAX * KMAX;           // [KMAX] behind the per-sequence local lists
            if (n && rk < k) sorted[rk] = mine;                 // value descending, index ascending
            fst();                                                          // merged
            bar_compute();
            fst();
            if (cw == 0) {
              const float v0 = __uint_as_float(sorted[0].x);
              float w0 = lane < k ? expf(__uint_as_float(sorted[lane].x) - v0) : 0.f;
              float w1 = lane + 32 < k ? expf(__uint_as_float(sorted[lane + 32].x) - v0) : 0.f;
              float i0 = w0, i1 = w1;
#pragma unroll
              for (int o = 1; o < 32; o <<= 1) {
                const float t0 = __shfl_up_sync(0xffffffffu, i0, o), t1 = __shfl_up_sync(0xffffffffu, i1, o);
                if (lane >= o) { i0 += t0; i1 += t1; }
              }
              const float tot0 = __shfl_sync(0xffffffffu, i0, 31);
              i1 += tot0;
              const float total = __shfl_sync(0xffffffffu, i1, 31);
              const float target = u01 * total;
              const unsigned c0 = __ballot_sync(0xffffffffu, lane < k && i0 > target);
              const unsigned c1 = __ballot_sync(0xffffffffu, lane + 32 < k && i1 > target);
              const int pick = c0 ? __ffs(c0) - 1 : (c1 ? 32 + __ffs(c1) - 1 : k - 1);
              tok = static_cast<int>(sorted[pick].y);
            }
          }
          if (cw == 0) {
            if (p.forced) tok = p.forced[static_cast<size_t>(b0 + s) * p.forced_stride + step];
            tok = min(max(tok, 0), p.V - 1);                 // never index the embedding table out of range
            if (lane == 0) {
              misc.result = tok;
              if (!misc.fin[s]) {
                const int b = b0 + s;
                const int pos = misc.outlen[s];
                p.st.out_ids[static_cast<size_t>(b) * p.st.out_stride + pos] = tok;      // api_cache.py:179
                if (p.st.step_ns && b == 0) p.st.step_ns[step] = ptx::global_timer_ns();   // per-token latency read-out
                misc.outlen[s] = pos + 1;
              }
            }
            __syncwarp();
            if (lane < CL) {
              if (lane == r) {
                misc.tok[s] = misc.fin[s] ? misc.tok[s] : tok;
              } else {
                const uint32_t ra = ptx::map_to_cta(ptx::smem_u32(&misc.tok[s]), lane);
                const uint32_t rb = ptx::map_to_cta(ptx::smem_u32(&bars.tok), lane);
                ptx::st_async_b32(ra, static_cast<uint32_t>(misc.fin[s] ? misc.tok[s] : tok), rb);
              }
            }
          }
        }
        // ---- everyone: receive the tokens of the sequences owned elsewhere, advance the state ----
        {
          const int owned_elsewhere = S - (r < S ? 1 : 0);
          if (ct == 0) ptx::mbar_arrive_expect_tx(&bars.tok, owned_elsewhere * 4);
          fst();                                                            // token drawn / sent (owner warp 0)
          if (cw == 0) {
            // the polling warp also advances the (replicated) state: the tokens are visible to it through the mbarrier
            // (remote) / its own store (local), so ONE barrier publishes tokens and state to the other warps
            ptx::mbar_wait_spin(&bars.tok, tok_use & 1);
            __syncwarp();
            if (lane < S && !misc.fin[lane]) {
              const int s = lane;
              misc.len[s] += 1;
              const int n = misc.nnew[s] + 1;
              misc.nnew[s] = n;
              if (misc.tok[s] == sp.eos_id || n >= misc.maxnew[s]) misc.fin[s] = 1;        // api_cache.py:181
            }
          }
          fst();
          ++tok_use;
          bar_compute();
          stamp(step);                                                      // token published, state advanced
        }
        if (p.early_exit) {
          bool all_fin = true;
          for (int s = 0; s < S; ++s) all_fin = all_fin && misc.fin[s] != 0;
          if (ct == 0 && step + 1 < n_steps) {
            *reinterpret_cast<volatile int*>(&misc.go) = all_fin ? 0 : 1;
            __threadfence_block();
            ptx::mbar_arrive(&bars.step_go);
          }
          if (all_fin) break;                                 // identical decision in all four CTAs (replicated state)
        }
      }
      // write the decode state back (the host-side step graph / download read it)
      if (ct == r && r < S) {
        const int s = ct, b = b0 + s;
        p.st.cur_tok[b] = misc.tok[s];
        p.st.lens[b] = misc.len[s];
        p.st.n_new[b] = misc.nnew[s];
        p.st.finished[b] = static_cast<uint8_t>(misc.fin[s]);
        p.st.out_len[b] = misc.outlen[s];
      }
    }
  }

  // no CTA may exit while a peer can still write into its shared memory
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

// ---- weight re-layout: the shared-memory image of every stage, in consumption order, per rank -------
// Stage = [256 rows x 64 K] bf16 in fragment-major order (see wfrag_init).  Order per rank: for every layer {in_proj kb 0..3 | out_proj | mlp.0
// kb 0..3 | mlp.2 kb 0..3}, then the head pairs {kb 0..3}.
struct PackSrc {
  const bf16* w_in[kMegaMaxLayers];
  const bf16* w_out[kMegaMaxLayers];
  const bf16* w1[kMegaMaxLayers];
  const bf16* w2[kMegaMaxLayers];
  const bf16* head;
};

__global__ void mega_pack_kernel(PackSrc src, uint4* __restrict__ dst, int n_layer, int V, int VS, int NP, int tail) {
  const int stages_per_rank = n_layer * kMegaStagesPerLayer + 4 * NP + tail;
  const size_t total = static_cast<size_t>(CL) * stages_per_rank * (STAGE_BYTES / 16);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int chunk_pos = static_cast<int>(idx % (STAGE_BYTES / 16));
    const int stage_g = static_cast<int>(idx / (STAGE_BYTES / 16));
    const int r = stage_g / stages_per_rank, st = stage_g % stages_per_rank;
    // fragment-major: chunk_pos = ((warp * 8 + t * 4 + ks) * 32 + lane), see wfrag_init
    const int pw = chunk_pos >> 8, pf = (chunk_pos >> 5) & 7, pl = chunk_pos & 31;
    const int pt = pf >> 2, pks = pf & 3, pm = pl >> 3;
    const int i = pw * 32 + pt * 16 + (pm & 1) * 8 + (pl & 7);   // stage row 0..255
    const int c = pks * 2 + (pm >> 1);                           // logical 16-byte chunk (8 bf16 along K) of the stage's 64 K
    const bf16* w = nullptr;
    int row = -1, col = 0, ld = D;
    if (st < n_layer * kMegaStagesPerLayer) {
      const int l = st / kMegaStagesPerLayer, q = st % kMegaStagesPerLayer;
      if (q < 3) {                                          // in_proj, fragment layout of gemm_qkv (NOT the generic one)
        const int tile = q < 2 ? 4 + pw : (pw >> 1);        // m16 tile of the 192 rows q | k | v
        const int ks = q < 2 ? 8 * q + pf : 8 * (pw & 1) + pf;
        const int i192 = tile * 16 + (pm & 1) * 8 + (pl & 7);
        w = src.w_in[l]; col = ks * 16 + (pm >> 1) * 8 - c * 8;   // (c * 8 is added back below)
        row = (i192 >> 6) * D + r * FS + (i192 & 63);
      } else if (q == 3) {                                  // out_proj: all 256 output rows, this rank's 64 input columns
        w = src.w_out[l]; row = i; col = r * FS;
      } else if (q < 8) {                                   // mlp.0: this rank's 256 hidden rows
        w = src.w1[l]; row = r * HS + i; col = (q - 4) * 64;
      } else {                                              // mlp.2: all 256 output rows, this rank's 256 hidden columns
        w = src.w2[l]; row = i; col = r * HS + (q - 8) * 64; ld = 4 * D;
      }
    } else {
      const int hq = st - n_layer * kMegaStagesPerLayer;
      w = src.head;
      if (hq < 4 * NP) {
        const int pr = hq / 4, kb = hq % 4;
        const int lr = pr * 256 + i;                        // row inside this rank's vocabulary slice
        col = kb * 64;
        if (lr < VS && r * VS + lr < V) row = r * VS + lr;
      } else {                                              // tail stage, fragment layout of gemm_khalf
        const int tile = pw >> 1, ks = 8 * (pw & 1) + pf;
        const int lr = 256 * NP + tile * 16 + (pm & 1) * 8 + (pl & 7);
        col = ks * 16 + (pm >> 1) * 8 - c * 8;              // (c * 8 is added back below)
        if (lr < VS && r * VS + lr < V) row = r * VS + lr;
      }
    }
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row >= 0) v = *reinterpret_cast<const uint4*>(w + static_cast<size_t>(row) * ld + col + c * 8);
    dst[idx] = v;
  }
}

// K / V cache rows written by the prefill ([B][4][Tmax][64]) -> the caches of the persistent kernel:
//   kh [B][4][head][Tmax][hd]            (head-major rows)
//   vt [B][4][head][Tvt / 32][hd][32]    (per 32-key block transposed, key 8j + 2t + e at position 8t + 2j + e)
__global__ void mega_relayout_kv_kernel(const MegaLayer* __restrict__ layers, const int32_t* __restrict__ lens,
                                        const int32_t* __restrict__ slots, int Tmax, int Tvt, int hd) {
  // (sequence, slice), layer; `slots` (optional) = the sequences to convert (continuous batching: the newly admitted ones)
  const int bs = slots ? slots[blockIdx.x / CL] * CL + blockIdx.x % CL : blockIdx.x, l = blockIdx.y;
  const int len = lens[bs / CL];
  const bf16* ksrc = layers[l].kc + static_cast<size_t>(bs) * Tmax * FS;
  const bf16* vsrc = layers[l].vc + static_cast<size_t>(bs) * Tmax * FS;
  bf16* kdst = layers[l].kh + static_cast<size_t>(bs) * Tvt * FS;
  bf16* vdst = layers[l].vt + static_cast<size_t>(bs) * Tvt * FS;
  for (int i = threadIdx.x; i < len * FS; i += blockDim.x) {
    const int t = i >> 6, f = i & 63, h = f / hd, d = f - h * hd, ki = t & 31;
    const int pos = 8 * ((ki >> 1) & 3) + 2 * (ki >> 3) + (ki & 1);
    kdst[(static_cast<size_t>(h) * Tvt + t) * hd + d] = ksrc[i];
    vdst[(static_cast<size_t>(h) * (Tvt >> 5) + (t >> 5)) * (hd * 32) + d * 32 + pos] = vsrc[i];
  }
}

}  // namespace

int mega_relayout_kv(cudaStream_t stream, const MegaLayer* layers, const int32_t* lens, int B, int n_layer, int Tmax, int Tvt, int hd) {
  mega_relayout_kv_kernel<<<dim3(B * CL, n_layer), 256, 0, stream>>>(layers, lens, nullptr, Tmax, Tvt, hd);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int mega_relayout_kv_slots(cudaStream_t stream, const MegaLayer* layers, const int32_t* lens, const int32_t* slots, int n, int n_layer,
                           int Tmax, int Tvt, int hd) {
  mega_relayout_kv_kernel<<<dim3(n * CL, n_layer), 256, 0, stream>>>(layers, lens, slots, Tmax, Tvt, hd);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int mega_smem_bytes(int smax) {
  return (smax <= 2 ? Smem<2, kMegaStages2>::kTotal : Smem<4, kMegaStages4>::kTotal) + 1024;
}

size_t mega_packed_bytes(int n_layer, int NP, int tail) {
  return static_cast<size_t>(CL) * (n_layer * kMegaStagesPerLayer + 4 * NP + tail) * STAGE_BYTES;
}

int mega_pack_weights(cudaStream_t stream, const bf16* const* w_in, const bf16* const* w_out, const bf16* const* w1,
                      const bf16* const* w2, const bf16* head, int n_layer, int V, int VS, int NP, int tail, void* dst) {
  if (n_layer > kMegaMaxLayers) return MG_E_SHAPE;
  PackSrc src{};
  for (int l = 0; l < n_layer; ++l) { src.w_in[l] = w_in[l]; src.w_out[l] = w_out[l]; src.w1[l] = w1[l]; src.w2[l] = w2[l]; }
  src.head = head;
  mega_pack_kernel<<<148 * 4, 256, 0, stream>>>(src, reinterpret_cast<uint4*>(dst), n_layer, V, VS, NP, tail);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

template <typename F>
static auto mega_dispatch(int smax, int hd, F&& f) {
  if (smax <= 2) return hd == 32 ? f(decode_mega_kernel<2, kMegaStages2, 32>) : f(decode_mega_kernel<2, kMegaStages2, 64>);
  return hd == 32 ? f(decode_mega_kernel<4, kMegaStages4, 32>) : f(decode_mega_kernel<4, kMegaStages4, 64>);
}

int mega_init() {
  for (int smax : {2, 4})
    for (int hd : {32, 64}) {
      const cudaError_t e = mega_dispatch(smax, hd, [&](auto* k) {
        return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, mega_smem_bytes(smax));
      });
      MG_CUDA_OK(e);
    }
  return MG_OK;
}

int mega_max_clusters(int smax) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * 64);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = mega_smem_bytes(smax);
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  int n = 0;
  const cudaError_t e = mega_dispatch(smax, 32, [&](auto* k) { return cudaOccupancyMaxActiveClusters(&n, k, &cfg); });
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int launch_decode_mega(cudaStream_t stream, const MegaParams& p, int n_clusters) {
  const int smax = p.S <= 2 ? 2 : 4;
  if (p.head_dim != 32 && p.head_dim != 64) return MG_E_SHAPE;
  dim3 grid(n_clusters * CL);
  mega_dispatch(smax, p.head_dim, [&](auto* k) {
    k<<<grid, NTHREADS, mega_smem_bytes(smax), stream>>>(p);
    return cudaSuccess;
  });
  MG_LAUNCH_CHECK();
  return MG_OK;
}

}  // namespace mega
}  // namespace mg
