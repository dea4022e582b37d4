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
//     every 32 KB stage [256 weight rows x 64 K, 128B-swizzled] in consumption order, so the producer warp
//     streams them L2 -> shared memory with ONE cp.async.bulk per stage into an mbarrier ring;
//   * the contractions are warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 -> fp32) issued by all eight
//     compute warps straight from the ring with ldmatrix (weights = the M = 16 operand, the cluster's <= 8
//     sequences = the N = 8 operand).  A tcgen05/TMEM formulation of the same step (swap-AB, M = 128,
//     N = 16, single issuing thread) was built and measured first: at 2-4 sequences per cluster it is bound
//     by instruction issue (~85 cycles per tcgen05.mma, 704 per step = 31 us) and by the commit -> mbarrier
//     -> tcgen05.ld hand-offs (profiles/r1_mega_tcgen05_timeline.txt); tcgen05 stays where the contraction
//     is dense (prefill, batched multi-kernel decode, the classifier: gemm_tc.cu);
//   * row-parallel partial sums are all-gathered across the cluster with st.async (distributed shared
//     memory writes that complete_tx on the receiver's mbarrier) and summed in a fixed order;
//   * attention is flash-decoding over the CTA's own cache slice [b][slice][t][64]: 128-byte rows,
//     four rows per warp load, 8 x 2 x 16 B in flight per lane, warp-shuffle softmax reductions;
//   * sampling: every CTA radix-selects the top-k of its vocabulary slice, the candidates go to the
//     sequence's owner CTA, which ranks them, draws with Philox and broadcasts the token.
//
// Warp roles (288 threads): warp 0 bulk-copy producer, warps 1..8 compute (LayerNorm, MMA, attention,
// epilogues, exchange, sampler).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>

#include "decode_mega.cuh"
#include "mg_engine.h"
#include "ptx.cuh"

namespace mg {
namespace mega {

namespace {

constexpr int CL = kMegaCluster;       // CTAs per cluster
constexpr int FS = 64;                 // features per CTA slice
constexpr int D = 256;                 // d_model
constexpr int HS = 256;                // hidden units per CTA (d_ff / CL)
constexpr int NCW = 8;                 // compute warps (all of them stream K/V in the attention phase; 12 measured no faster)
constexpr int GW = 8;                  // ... of which the first 8 run the GEMMs, epilogues and the sampler
constexpr int NCT = NCW * 32;          // compute threads
constexpr int NTHREADS = (1 + NCW) * 32;
constexpr int XP = 264;                // activation row pitch (bf16 elements): K = 256 + 8 pad, bank-conflict free
constexpr int AP = 72;                 // attention-output row pitch: K = 64 + 8 pad
constexpr int STAGE_BYTES = kMegaStageBytes;   // one stage: two weight tiles [128 rows x 64 K] bf16 (rows 0..255)
constexpr int KMAX = kMegaMaxTopK;     // top_k limit of the in-kernel sampler
constexpr int kCandCap = 128;          // per-sequence superset capacity of the sampler's fast path
constexpr float kLog2e = 1.4426950408889634f;

enum { BAR_COMPUTE = 1, BAR_EPI = 2 };
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
  static constexpr int kTotal = kBars + 512;
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
};

struct Bars {
  uint64_t full[16];
  uint64_t empty[16];
  uint64_t xchg[2];
  uint64_t cand;
  uint64_t tok;
  uint64_t step_go;                   // compute warps -> producer: the next step will run (early-exit mode)
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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---- warp-level tensor-core helpers ---------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&a)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- ring bookkeeping shared by the producer and the compute warps -------------------------------
struct RingPos {
  int stage = 0;
  uint32_t phase = 0;
  template <int NSTAGE> __device__ __forceinline__ void advance() {
    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
  }
};

// One GEMM "pair": NKB stages of [256 weight rows x 64 K]; this warp owns weight rows [32 cw, 32 cw + 32)
// (two m16 tiles) and accumulates D[16 rows x 8 sequences] per tile over the stages.
//   act   : activations bf16 [8][pitch] (row = sequence), K contiguous
//   acc   : [2][4] fp32, thread holds rows (lane/4, lane/4 + 8) x sequences ((lane%4)*2, +1) of each tile
template <int NKB, int NSTAGE>
__device__ __forceinline__ void gemm_pair(uint8_t* ring, uint64_t* full, uint64_t* empty, RingPos& rp, const bf16* act,
                                          int pitch, int cw, int lane, bool active, float (&acc)[2][4]) {
  uint32_t breg[NKB * 4][2];
  {
    const bf16* bp = act + (lane >> 2) * pitch + (lane & 3) * 2;
#pragma unroll
    for (int ks = 0; ks < NKB * 4; ++ks) {
      breg[ks][0] = *reinterpret_cast<const uint32_t*>(bp + ks * 16);
      breg[ks][1] = *reinterpret_cast<const uint32_t*>(bp + ks * 16 + 8);
    }
  }
  float acc2[2][4];                                        // odd k-steps: halves the dependent MMA chains
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[t][e] = 0.f; acc2[t][e] = 0.f; }
  // ldmatrix row address of this lane: matrix m = lane / 8 -> rows (m & 1) * 8 + lane % 8, k chunk (m >> 1)
  const int lrow = ((lane >> 3) & 1) * 8 + (lane & 7);
  const int lchunk = lane >> 4;
#pragma unroll
  for (int kb = 0; kb < NKB; ++kb) {
    ptx::mbar_wait(&full[rp.stage], rp.phase);
    if (active) {
      const uint32_t sbase = ptx::smem_u32(ring + rp.stage * STAGE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int i = cw * 32 + t * 16 + lrow;                 // stage row 0..255
        const uint32_t rbase = sbase + (i >> 7) * (STAGE_BYTES / 2) + (i & 127) * 128;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t a[4];
          ldmatrix_x4(rbase + (((ks * 2 + lchunk) ^ (i & 7)) << 4), a);
          if (ks & 1) mma_bf16_16816(acc2[t], a, breg[kb * 4 + ks][0], breg[kb * 4 + ks][1]);
          else mma_bf16_16816(acc[t], a, breg[kb * 4 + ks][0], breg[kb * 4 + ks][1]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[rp.stage]);
    rp.template advance<NSTAGE>();
  }
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[t][e] += acc2[t][e];
}

template <int SMAX, int NSTAGE>
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
  const int n_layer = p.n_layer, NL = 2 * p.NP * 128;      // head rows per CTA, padded to tile pairs
  const int hd = p.head_dim;

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
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&bars.full[s], 1); ptx::mbar_init(&bars.empty[s], GW); }
    ptx::mbar_init(&bars.xchg[0], 1);
    ptx::mbar_init(&bars.xchg[1], 1);
    ptx::mbar_init(&bars.cand, 1);
    ptx::mbar_init(&bars.tok, 1);
    ptx::mbar_init(&bars.step_go, 1);
    ptx::fence_mbar_init();
    for (int s = 0; s < 8; ++s) {
      const bool live = s < S;
      misc.tok[s] = live ? p.st.cur_tok[b0 + s] : 0;
      misc.len[s] = live ? p.st.lens[b0 + s] : 0;
      misc.nnew[s] = live ? p.st.n_new[b0 + s] : 0;
      misc.maxnew[s] = live ? p.st.max_new[b0 + s] : 0;
      misc.fin[s] = live ? static_cast<int>(p.st.finished[b0 + s]) : 1;
    }
  }
  __syncthreads();
  // every CTA of the cluster must have initialised its barriers before any remote st.async lands
  ptx::cluster_arrive();
  ptx::cluster_wait();
  const int n_steps = p.n_steps;

  if (S > 0) {
    if (warp == 0) {
      // =========================== bulk-copy producer ===========================
      if (lane == 0) {
        RingPos rp;
        const int stages_per_step = n_layer * kMegaStagesPerLayer + 4 * p.NP;
        const uint8_t* src0 = p.packed + static_cast<size_t>(rank) * stages_per_step * STAGE_BYTES;
        for (int step = 0; step < n_steps; ++step) {
          if (p.early_exit && step > 0) {                     // do not stream weights for a step that will not run
            ptx::mbar_wait(&bars.step_go, (step - 1) & 1);
            if (*reinterpret_cast<volatile int*>(&misc.go) == 0) break;
          }
          const uint8_t* src = src0;
          for (int i = 0; i < stages_per_step; ++i, src += STAGE_BYTES) {
            ptx::mbar_wait(&bars.empty[rp.stage], rp.phase ^ 1);
            if (p.dbg_skip_loads) {
              ptx::mbar_arrive(&bars.full[rp.stage]);
            } else {
              ptx::mbar_arrive_expect_tx(&bars.full[rp.stage], STAGE_BYTES);
              ptx::bulk_load_1d(smem + L::kRing + rp.stage * STAGE_BYTES, src, STAGE_BYTES, &bars.full[rp.stage]);
            }
            rp.template advance<NSTAGE>();
          }
        }
      }
      __syncwarp();
    } else {
      // =========================== compute warps ===========================
      const int cw = warp - 1;                               // 0..7
      const int ct = threadIdx.x - 32;                       // 0..255
      uint32_t xuse = 0;
      RingPos rp;
      const SampleParams sp = *p.sp;
      const float scale_log2 = kLog2e / sqrtf(static_cast<float>(hd));
      const int cph = hd / 8;                                // 16-byte chunks per head (4 or 8)
      const int r = static_cast<int>(rank);
      uint8_t* ring = smem + L::kRing;
      // accumulator fragment coordinates: rows frow / frow + 8 of tile t (weight row 32 cw + 16 t + ...), sequences fs, fs + 1
      const int frow = lane >> 2, fs = (lane & 3) * 2;

      auto bar_compute = [&]() { ptx::named_bar_sync(BAR_COMPUTE, NCT); };
      int stamp_id = 0;
      const bool prof_on = p.prof != nullptr && cluster == 0 && rank == 0 && ct == 0;
      auto stamp = [&](int step_now) {
        if (prof_on && step_now == p.prof_step && stamp_id < 64) p.prof[stamp_id++] = ptx::global_timer_ns();
      };
      // LayerNorm of x[s] (or plain cast when w == nullptr) into xb (bf16 [8][XP]), warp s < S
      auto stage_x = [&](const float* __restrict__ w, const float* __restrict__ b) {
        if (cw < S) {
          const int s = cw;
          float v[8];
          const float4 a = *reinterpret_cast<const float4*>(xs + s * D + lane * 8);
          const float4 c = *reinterpret_cast<const float4*>(xs + s * D + lane * 8 + 4);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
          if (w) {
            float sum = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[e];
            const float mean = warp_sum(sum) * (1.0f / D);
            float sq = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float dlt = v[e] - mean; sq += dlt * dlt; }
            const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / D) + 1e-5f);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd * w[lane * 8 + e] + b[lane * 8 + e];
          }
          const uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          *reinterpret_cast<uint4*>(xb + s * XP + lane * 8) = o;
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
      auto exchange_finish = [&](const float* __restrict__ bias) {
        const int buf = xuse & 1;
        ptx::mbar_wait(&bars.xchg[buf], (xuse >> 1) & 1);
        bar_compute();                                        // local slot writes visible, everyone past the wait
        if (ct == 0) ptx::mbar_arrive_expect_tx(&bars.xchg[buf], (CL - 1) * D * SMAX * 4);   // re-arm for use + 2
        if (ct < D) {
          const int f = ct;
          const float bv = bias[f];
          float acc[SMAX];
#pragma unroll
          for (int s = 0; s < SMAX; ++s) acc[s] = 0.f;
#pragma unroll
          for (int src = 0; src < CL; ++src) {
            const float* sl = slots + ((buf * CL + src) * D + f) * SMAX;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) acc[s] += sl[s];
          }
#pragma unroll
          for (int s = 0; s < SMAX; ++s)
            if (s < S) xs[s * D + f] += acc[s] + bv;
        }
        ++xuse;
        bar_compute();
      };

      if (ct == 0) {                                          // arm the receive barriers once
        ptx::mbar_arrive_expect_tx(&bars.xchg[0], (CL - 1) * D * SMAX * 4);
        ptx::mbar_arrive_expect_tx(&bars.xchg[1], (CL - 1) * D * SMAX * 4);
      }
      uint32_t cand_use = 0, tok_use = 0;

      for (int step = 0; step < n_steps; ++step) {
        // ---- embedding: x = tok_emb[tok] + pos_emb[0]   (api_cache.py:99 with T == 1) ----
        if (ct < D) {
          const int f = ct;
          const float pe = __bfloat162float(p.pos_emb[f]);
          for (int s = 0; s < S; ++s) xs[s * D + f] = __bfloat162float(p.tok_emb[static_cast<size_t>(misc.tok[s]) * D + f]) + pe;
        }
        bar_compute();
        stamp(step);                                                        // 0: embedding done
        for (int l = 0; l < n_layer; ++l) {
          const MegaLayer& lw = p.layers[l];
          const float* pl = reinterpret_cast<const float*>(smem + L::kParams) + l * kLayerParamFloats;
          // ---- LN1 -> QKV: stage rows 0..63 = q slice, 64..127 = k slice, 128..191 = v slice ----
          stage_x(pl + P_LN1W, pl + P_LN1B);
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<4, NSTAGE>(ring, bars.full, bars.empty, rp, xb, XP, cw, lane, cw < 6, acc);
            if (cw < 6 && fs < S) {
              const int part_id = cw >> 1;                    // 0 = q, 1 = k, 2 = v
#pragma unroll
              for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  const int f = (cw & 1) * 32 + t * 16 + frow + h8 * 8;      // feature inside the 64-wide slice
                  const float bias = pl[P_BQKV + part_id * FS + f];
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const int s = fs + e;
                    if (s < S) {
                      const float v = acc[t][h8 * 2 + e] + bias;
                      if (part_id == 0) qs[s * FS + f] = v * scale_log2;
                      else if (part_id == 1) knew[s * FS + f] = __float2bfloat16_rn(v);
                      else vnew[s * FS + f] = __float2bfloat16_rn(v);
                    }
                  }
                }
            }
          }
          bar_compute();
          stamp(step);                                                      // +1: QKV done
          // ---- append the new K/V rows (api_cache.py:66-67) + flash-decoding over this CTA's slice ----
          if (cw < S && lane < 16 && !misc.fin[cw]) {
            const int s = cw, which = lane >> 3, c = lane & 7;
            bf16* dst = (which ? lw.vc : lw.kc) +
                        ((static_cast<size_t>(b0 + s) * CL + r) * p.Tmax + misc.len[s]) * FS + c * 8;
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>((which ? vnew : knew) + s * FS + c * 8);
          }
          // L2 prefetch of the NEXT layer's K/V streams of this CTA (next step's layer 0 after the last layer):
          // HBM keeps streaming while the GEMM / exchange / sampler phases run; four bulk prefetches per layer
          if (cw == NCW - 1 && lane < 2 * S) {
            const int s = lane >> 1, which = lane & 1;
            const int ln = (l + 1 < n_layer) ? l + 1 : 0;
            const int rows = misc.len[s] + (ln == 0 ? 1 : 0);
            if (!misc.fin[s] && rows > 0) {
              const MegaLayer& nx = p.layers[ln];
              const bf16* src = (which ? nx.vc : nx.kc) + (static_cast<size_t>(b0 + s) * CL + r) * p.Tmax * FS;
              ptx::prefetch_l2_bulk(src, static_cast<uint32_t>(rows) * FS * 2);
            }
          }
          {
            // warps are dealt round-robin to the sequences: warp cw serves sequence cw % S as its (cw / S)-th worker
            const int s = cw % S, wi = cw / S, nws = (NCW - s + S - 1) / S;
            const int len = misc.fin[s] ? 0 : misc.len[s];
            const bf16* kc = lw.kc + (static_cast<size_t>(b0 + s) * CL + r) * p.Tmax * FS;
            const bf16* vc = lw.vc + (static_cast<size_t>(b0 + s) * CL + r) * p.Tmax * FS;
            const int rr = lane >> 3, c = lane & 7;           // row inside a 4-row group, 16-byte chunk
            float q[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) q[e] = qs[s * FS + c * 8 + e];
            float m_run = -INFINITY, l_run = 0.f, acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
            // Batches of U four-row groups: all 2 x U 16-byte loads of a batch are issued before the first use
            // (the scoreboard tracks loads per batch, so finer-grained register pipelining serialises them).
            constexpr int U = 8;
            for (int base = wi * 4 * U; base < len; base += nws * 4 * U) {
              uint4 kr[U], vr[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int row = base + u * 4 + rr;
                if (row < len) {
                  const int arow = p.dbg_attn_hot ? (row & 31) : row;
                  kr[u] = ptx::ld_global_stream16(kc + static_cast<size_t>(arow) * FS + c * 8);
                  vr[u] = ptx::ld_global_stream16(vc + static_cast<size_t>(arow) * FS + c * 8);
                } else {
                  kr[u] = make_uint4(0, 0, 0, 0);
                  vr[u] = make_uint4(0, 0, 0, 0);
                }
              }
              // phase-ordered (all dots, then all shuffles) so the U independent chains interleave; no branches
              float sc[U];
#pragma unroll
              for (int u = 0; u < U; ++u) {
                float kf[8];
                unpack8(kr[u], kf);
                const float d0 = fmaf(q[0], kf[0], fmaf(q[2], kf[2], fmaf(q[4], kf[4], q[6] * kf[6])));
                const float d1 = fmaf(q[1], kf[1], fmaf(q[3], kf[3], fmaf(q[5], kf[5], q[7] * kf[7])));
                sc[u] = d0 + d1;
              }
#pragma unroll
              for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 1);
#pragma unroll
              for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], 2);
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const float t4 = __shfl_xor_sync(0xffffffffu, sc[u], 4);
                sc[u] += (cph == 8) ? t4 : 0.f;
              }
              float m_new = m_run;
#pragma unroll
              for (int u = 0; u < U; ++u) {
                sc[u] = (base + u * 4 + rr < len) ? sc[u] : -INFINITY;
                m_new = fmaxf(m_new, sc[u]);
              }
              if (m_new > -INFINITY) {
                const float corr = exp2f(m_run - m_new);
                float psum = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  const float pw = exp2f(sc[u] - m_new);
                  psum += pw;
                  float vf[8];
                  unpack8(vr[u], vf);
#pragma unroll
                  for (int e = 0; e < 8; ++e) acc[e] = fmaf(pw, vf[e], acc[e]);
                }
                l_run = l_run * corr + psum;
                m_run = m_new;
              }
            }
            // merge the four row residues of the warp (lanes xor 8, 16)
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              const float m_o = __shfl_xor_sync(0xffffffffu, m_run, o);
              const float l_o = __shfl_xor_sync(0xffffffffu, l_run, o);
              const float m_n = fmaxf(m_run, m_o);
              const float fa = (m_run > -INFINITY) ? exp2f(m_run - m_n) : 0.f;
              const float fb = (m_o > -INFINITY) ? exp2f(m_o - m_n) : 0.f;
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float a_o = __shfl_xor_sync(0xffffffffu, acc[e], o);
                acc[e] = acc[e] * fa + a_o * fb;
              }
              l_run = l_run * fa + l_o * fb;
              m_run = m_n;
            }
            if (lane < 8) {
              float* pp = part + (s * NCW + wi) * 68;
#pragma unroll
              for (int e = 0; e < 8; ++e) pp[c * 8 + e] = acc[e];
              if ((c % cph) == 0) { pp[64 + (c / cph) * 2] = m_run; pp[64 + (c / cph) * 2 + 1] = l_run; }
            }
          }
          bar_compute();
          // final merge over the warps + the new token's own row; writes the out_proj operand
          if (ct < S * FS) {
            const int s = ct / FS, f = ct - s * FS, hh = f / hd;
            float snew = 0.f;
            for (int e = 0; e < hd; ++e) snew = fmaf(qs[s * FS + hh * hd + e], __bfloat162float(knew[s * FS + hh * hd + e]), snew);
            const int nws = (NCW - s + S - 1) / S;
            float M = snew;
            for (int w = 0; w < nws; ++w) M = fmaxf(M, part[(s * NCW + w) * 68 + 64 + hh * 2]);
            const float pn = exp2f(snew - M);
            float Lsum = pn, o = pn * __bfloat162float(vnew[s * FS + f]);
            for (int w = 0; w < nws; ++w) {
              const float* pp = part + (s * NCW + w) * 68;
              const float mw = pp[64 + hh * 2];
              if (mw > -INFINITY) {
                const float fw = exp2f(mw - M);
                Lsum = fmaf(pp[64 + hh * 2 + 1], fw, Lsum);
                o = fmaf(pp[f], fw, o);
              }
            }
            attb[s * AP + f] = __float2bfloat16_rn(o / Lsum);
          }
          bar_compute();
          stamp(step);                                                      // +2: attention done
          // ---- out_proj (row-parallel over this CTA's 64 attention features) -> exchange -> x += attn ----
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<1, NSTAGE>(ring, bars.full, bars.empty, rp, attb, AP, cw, lane, true, acc);
            exchange_send(acc);
          }
          exchange_finish(pl + P_BOUT);
          stamp(step);                                                      // +3: out_proj + exchange done
          // ---- LN2 -> MLP1 (+GELU, this CTA's 256 hidden units) -> MLP2 (row-parallel) -> exchange ----
          stage_x(pl + P_LN2W, pl + P_LN2B);
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<4, NSTAGE>(ring, bars.full, bars.empty, rp, xb, XP, cw, lane, true, acc);
            if (fs < S) {
#pragma unroll
              for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  const int j = cw * 32 + t * 16 + frow + h8 * 8;            // hidden unit inside this CTA's slice
                  const float b1 = pl[P_B1 + j];
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (fs + e < S) hb[(fs + e) * XP + j] = __float2bfloat16_rn(gelu_erf_f(acc[t][h8 * 2 + e] + b1));
                }
            }
          }
          bar_compute();
          stamp(step);                                                      // +4: MLP1 done
          if (cw < GW) {
            float acc[2][4];
            gemm_pair<4, NSTAGE>(ring, bars.full, bars.empty, rp, hb, XP, cw, lane, true, acc);
            exchange_send(acc);
          }
          exchange_finish(pl + P_B2);
          stamp(step);                                                      // +5: MLP2 + exchange done
        }
        // ---- head: logits of this CTA's vocabulary slice (no final LayerNorm, api_cache.py:105) ----
        stage_x(nullptr, nullptr);
        {
          const int v_lo = r * p.VS, v_hi = min(p.V, (r + 1) * p.VS);
          for (int pr = 0; pr < p.NP && cw < GW; ++pr) {
            float hbv[2][2];                                   // head bias of this thread's 4 rows, loaded ahead of the MMAs
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                const int vr = v_lo + pr * 256 + cw * 32 + t * 16 + frow + h8 * 8;
                hbv[t][h8] = vr < v_hi ? __ldg(p.head_b + vr) : 0.f;
              }
            float acc[2][4];
            gemm_pair<4, NSTAGE>(ring, bars.full, bars.empty, rp, xb, XP, cw, lane, true, acc);
            if (fs < S) {
#pragma unroll
              for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                  const int lr = pr * 256 + cw * 32 + t * 16 + frow + h8 * 8;   // row inside the slice
                  const int vr = v_lo + lr;
                  const bool ok = vr < v_hi;
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const int s = fs + e;
                    if (s < S) {
                      const float lg = acc[t][h8 * 2 + e] + hbv[t][h8];
                      logits[s * NL + lr] = ok ? lg / sp.temperature : -INFINITY;
                      if (p.dbg_logits && ok) p.dbg_logits[(static_cast<size_t>(step) * p.B + b0 + s) * p.V + vr] = lg;
                    }
                  }
                }
            }
          }
        }
        bar_compute();
        stamp(step);                                                        // head done

        // ---- sampler (api_cache.py:169-181): local top-k -> owner CTA ranks, draws, broadcasts ----
        // The compute warps are dealt to the sequences (warp cw -> sequence cw % S); each group radix-selects
        // the k largest logits of this CTA's vocabulary slice with warp-aggregated histogram updates.
        const int k = sp.top_k;
        if (cw < GW) {
          const int s = cw % S, wi = cw / S, nws = (GW - s + S - 1) / S;
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
          if (!overflow) {
            float m_t = -INFINITY;
            for (int i = gt; i < NL; i += gn) m_t = fmaxf(m_t, z[i]);
            maxima[gt] = m_t;
            if (gt == 0) { *g_cnt = 0; *g_exact = 0; }
            ptx::named_bar_sync(gbar, gn);
            int rk = 0;
#pragma unroll 4
            for (int u = 0; u < gn; ++u) {
              const float mu = maxima[u];
              rk += (mu > m_t) || (mu == m_t && u < gt);
            }
            if (rk == k - 1) *reinterpret_cast<volatile float*>(g_pre) = m_t;
            ptx::named_bar_sync(gbar, gn);
            const float tau = *reinterpret_cast<volatile float*>(g_pre);
            for (int i0 = 0; i0 < NL; i0 += gn) {
              const int i = i0 + gt;
              const float zi = i < NL ? z[i] : -INFINITY;
              const bool take = i < NL && zi >= tau;
              const unsigned bal = __ballot_sync(0xffffffffu, take);
              int base_pos = 0;
              if (lane == 0 && bal) base_pos = atomicAdd(const_cast<int*>(g_cnt), __popc(bal));
              base_pos = __shfl_sync(0xffffffffu, base_pos, 0);
              const int pos = base_pos + __popc(bal & ((1u << lane) - 1));
              if (take && pos < kCandCap) clist[pos] = make_uint2(__float_as_uint(zi), static_cast<uint32_t>(r * p.VS + i));
            }
            ptx::named_bar_sync(gbar, gn);
            const int c = *g_cnt;
            overflow = c > kCandCap || c < k;
            if (!overflow) {
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
                if (rank_j < k) gl[rank_j] = mine;            // sorted: value descending, index ascending
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
        bar_compute();
        stamp(step);                                                        // local top-k done
        // ---- owner: merge CL x k candidates, draw, publish ----
        if (r < S) {
          const int s = r;
          if (ct == 0) ptx::mbar_arrive_expect_tx(&bars.cand, (CL - 1) * k * 8);
          ptx::mbar_wait(&bars.cand, cand_use & 1);
          ++cand_use;
          bar_compute();
          // every CTA's list arrives sorted (value descending, index ascending): the global rank of an entry is
          // its own position plus, for each other list, the number of entries that precede it (binary search)
          const int n = CL * k;                               // <= 256 candidates, one per thread
          uint2 mine = make_uint2(0, 0);
          int rk = 0;
          if (ct < n) {
            const int src = ct / k, j = ct - src * k;
            mine = cand[src * KMAX + j];
            const float mv = __uint_as_float(mine.x);
            rk = j;
            for (int si = 0; si < CL; ++si) {
              if (si == src) continue;
              const uint2* lst = cand + si * KMAX;
              int lo = 0, hi = k;                             // first position whose entry does NOT precede `mine`
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint2 o = lst[mid];
                const float ov = __uint_as_float(o.x);
                const bool before = (ov > mv) || (ov == mv && (o.y < mine.y || (o.y == mine.y && si < src)));
                if (before) lo = mid + 1; else hi = mid;
              }
              rk += lo;
            }
          }
          uint2* sorted = local_list + SMAX * KMAX;           // [KMAX] behind the per-sequence local lists
          if (ct < n && rk < k) sorted[rk] = mine;            // value descending, index ascending
          bar_compute();
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
            const float u01 = philox_uniform(sp.seed, sp.seq_base + static_cast<uint64_t>(b0 + s), static_cast<uint32_t>(misc.nnew[s]));
            const float target = u01 * total;
            const unsigned c0 = __ballot_sync(0xffffffffu, lane < k && i0 > target);
            const unsigned c1 = __ballot_sync(0xffffffffu, lane + 32 < k && i1 > target);
            const int pick = c0 ? __ffs(c0) - 1 : (c1 ? 32 + __ffs(c1) - 1 : k - 1);
            int tok = static_cast<int>(sorted[pick].y);
            if (p.forced) tok = p.forced[static_cast<size_t>(b0 + s) * p.forced_stride + step];
            tok = min(max(tok, 0), p.V - 1);                 // never index the embedding table out of range
            if (lane == 0) {
              misc.result = tok;
              if (!misc.fin[s]) {
                const int b = b0 + s;
                const int pos = p.st.out_len[b];
                p.st.out_ids[static_cast<size_t>(b) * p.st.out_stride + pos] = tok;      // api_cache.py:179
                p.st.out_len[b] = pos + 1;
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
          ptx::mbar_wait(&bars.tok, tok_use & 1);
          ++tok_use;
          bar_compute();
          if (ct < S && !misc.fin[ct]) {
            const int s = ct;
            misc.len[s] += 1;
            const int n = misc.nnew[s] + 1;
            misc.nnew[s] = n;
            if (misc.tok[s] == sp.eos_id || n >= misc.maxnew[s]) misc.fin[s] = 1;          // api_cache.py:181
          }
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
      }
    }
  }

  // no CTA may exit while a peer can still write into its shared memory
  __syncthreads();
  ptx::cluster_arrive();
  ptx::cluster_wait();
}

// ---- weight re-layout: the shared-memory image of every stage, in consumption order, per rank -------
// Stage = [256 rows x 64 K] bf16 as two SW128 K-major tiles; row i lives at (i / 128) * 16 KB + (i % 128) * 128 B,
// 16-byte chunk c at position c ^ (i & 7).  Order per rank: for every layer {in_proj kb 0..3 | out_proj | mlp.0
// kb 0..3 | mlp.2 kb 0..3}, then the head pairs {kb 0..3}.
struct PackSrc {
  const bf16* w_in[kMegaMaxLayers];
  const bf16* w_out[kMegaMaxLayers];
  const bf16* w1[kMegaMaxLayers];
  const bf16* w2[kMegaMaxLayers];
  const bf16* head;
};

__global__ void mega_pack_kernel(PackSrc src, uint4* __restrict__ dst, int n_layer, int V, int VS, int NP) {
  const int stages_per_rank = n_layer * kMegaStagesPerLayer + 4 * NP;
  const size_t total = static_cast<size_t>(CL) * stages_per_rank * (STAGE_BYTES / 16);
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int chunk_pos = static_cast<int>(idx % (STAGE_BYTES / 16));
    const int stage_g = static_cast<int>(idx / (STAGE_BYTES / 16));
    const int r = stage_g / stages_per_rank, st = stage_g % stages_per_rank;
    const int tile = chunk_pos / 1024, in_tile = chunk_pos % 1024;
    const int row_in_tile = in_tile / 8, cpos = in_tile % 8;
    const int i = tile * 128 + row_in_tile;               // stage row 0..255
    const int c = cpos ^ (row_in_tile & 7);                // logical 16-byte chunk (8 bf16 along K)
    const bf16* w = nullptr;
    int row = -1, col = 0, ld = D;
    if (st < n_layer * kMegaStagesPerLayer) {
      const int l = st / kMegaStagesPerLayer, q = st % kMegaStagesPerLayer;
      if (q < 4) {                                          // in_proj: q rows | k rows | v rows | unused
        w = src.w_in[l]; col = q * 64;
        if (i < 64) row = r * FS + i;
        else if (i < 128) row = D + r * FS + (i - 64);
        else if (i < 192) row = 2 * D + r * FS + (i - 128);
      } else if (q == 4) {                                  // out_proj: all 256 output rows, this rank's 64 input columns
        w = src.w_out[l]; row = i; col = r * FS;
      } else if (q < 9) {                                   // mlp.0: this rank's 256 hidden rows
        w = src.w1[l]; row = r * HS + i; col = (q - 5) * 64;
      } else {                                              // mlp.2: all 256 output rows, this rank's 256 hidden columns
        w = src.w2[l]; row = i; col = r * HS + (q - 9) * 64; ld = 4 * D;
      }
    } else {
      const int hq = st - n_layer * kMegaStagesPerLayer, pr = hq / 4, kb = hq % 4;
      const int lr = pr * 256 + i;                          // row inside this rank's vocabulary slice
      w = src.head; col = kb * 64;
      if (lr < VS && r * VS + lr < V) row = r * VS + lr;
    }
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row >= 0) v = *reinterpret_cast<const uint4*>(w + static_cast<size_t>(row) * ld + col + c * 8);
    dst[idx] = v;
  }
}

}  // namespace

int mega_smem_bytes(int smax) {
  return (smax <= 2 ? Smem<2, kMegaStages2>::kTotal : Smem<4, kMegaStages4>::kTotal) + 1024;
}

size_t mega_packed_bytes(int n_layer, int NP) {
  return static_cast<size_t>(CL) * (n_layer * kMegaStagesPerLayer + 4 * NP) * STAGE_BYTES;
}

int mega_pack_weights(cudaStream_t stream, const bf16* const* w_in, const bf16* const* w_out, const bf16* const* w1,
                      const bf16* const* w2, const bf16* head, int n_layer, int V, int VS, int NP, void* dst) {
  if (n_layer > kMegaMaxLayers) return MG_E_SHAPE;
  PackSrc src{};
  for (int l = 0; l < n_layer; ++l) { src.w_in[l] = w_in[l]; src.w_out[l] = w_out[l]; src.w1[l] = w1[l]; src.w2[l] = w2[l]; }
  src.head = head;
  mega_pack_kernel<<<148 * 4, 256, 0, stream>>>(src, reinterpret_cast<uint4*>(dst), n_layer, V, VS, NP);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int mega_init() {
  MG_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel<2, kMegaStages2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  mega_smem_bytes(2)));
  MG_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel<4, kMegaStages4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  mega_smem_bytes(4)));
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
  cudaError_t e = smax <= 2 ? cudaOccupancyMaxActiveClusters(&n, decode_mega_kernel<2, kMegaStages2>, &cfg)
                            : cudaOccupancyMaxActiveClusters(&n, decode_mega_kernel<4, kMegaStages4>, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int launch_decode_mega(cudaStream_t stream, const MegaParams& p, int n_clusters) {
  const int smax = p.S <= 2 ? 2 : 4;
  dim3 grid(n_clusters * CL);
  if (smax == 2) decode_mega_kernel<2, kMegaStages2><<<grid, NTHREADS, mega_smem_bytes(2), stream>>>(p);
  else decode_mega_kernel<4, kMegaStages4><<<grid, NTHREADS, mega_smem_bytes(4), stream>>>(p);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

}  // namespace mega
}  // namespace mg
