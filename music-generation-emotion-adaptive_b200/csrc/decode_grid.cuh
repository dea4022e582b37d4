// Host interface of the grid-synchronous persistent decode kernel (implementation: decode_grid.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.cuh"

namespace mg {
namespace grid {

constexpr int kMaxLayers = 8;
constexpr int kMaxSeqs = 64;           // 8 n-tiles of the m16n8k16 MMA = the 8 warps of a CTA
constexpr int kThreads = 256;
constexpr int kMaxSplits = 16;         // key ranges per (sequence, head) in the attention phase
constexpr int kMaxItems = 96;          // dense work items of one CTA per decode step
constexpr int kGridCtrlBytes = 4096;
constexpr int kSampMaxPer = 36;        // fast sampler: logits per thread held in registers (V <= 256 * 36)

// phase kinds of a decode step; phase index = 5 * layer + kind (kind < 5), 5 L = head, 5 L + 1 = sampler
enum Kind { K_QKV = 0, K_ATT = 1, K_OUT = 2, K_MLP1 = 3, K_MLP2 = 4, K_HEAD = 5, K_SAMPLE = 6 };

// LayerNorm is folded into the matrices that consume it: W' = W diag(gamma) (packed), c[r] = sum_k W'[r][k], d[r] = sum_k W[r][k] beta[k]
// + bias[r], so that  W LN(x) + bias = rstd (W' x - mean c) + d  with the row statistics applied in the epilogue (grid_pack_weights).
struct GridLayer {
  const float *c_in, *d_in, *c_1, *d_1, *b_out, *b2;
  bf16 *kh, *vt;                       // K head-major [B][H][Tvt][hd], V per 32-key block transposed [B][H][Tvt / 32][hd][32] (attn_tc.cuh)
  size_t w_in, w_out, w1, w2;          // byte offsets of the matrices' first tile inside `packed`
};

// Dense work item: 16 rows of one weight matrix x one group of sequences.
struct GridItem {
  int16_t phase;                       // phase index inside the step
  int16_t row_tile;                    // rows [16 row_tile, 16 row_tile + 16) of the matrix
  int16_t group;                       // sequence group: n-tiles [group * TN, group * TN + TN)
  int16_t pad;
};

struct GridParams {
  const uint8_t* packed;               // fragment-major 16-row weight tiles (grid_pack_weights)
  size_t w_head;                       // byte offset of the head's first tile
  GridLayer layers[kMaxLayers];
  const GridItem* items;               // [n_cta][kMaxItems], in phase order
  const int32_t* n_items;              // [n_cta]
  const bf16* tok_emb;
  const bf16* pos_emb;
  const float* head_b;
  const SampleParams* sp;
  DecodeState st;
  // activations of the current step, [kMaxSeqs][...] (global memory; L2 is the exchange medium between the phases)
  float* x;                            // residual stream entering a block (written by mlp.2)
  float* x1;                           // residual stream after the attention sub-layer
  bf16 *xb, *x1b;                      // bf16 copies (GEMM operands of the consumers)
  float *sx, *sx1;                     // [seq][D / 16][2] per-tile (sum, sum of squares): LayerNorm statistics assembled by the consumer
  float* q;                            // [seq][D] fp32, pre-scaled by log2(e) / sqrt(hd)
  bf16* knew;                          // [seq][D] the new token's K / V rows (folded by the first attention worker)
  bf16* vnew;
  bf16* h;                             // [seq][d_ff]
  float* logits;                       // [seq][ldl]
  float* vals;                         // [seq][ldl] scratch of the sampler's general path
  float* part;                         // attention partials [seq][head][kMaxSplits][hd + 4]: numerators | m | l
  unsigned* ctrl;                      // [0] grid barrier counter (monotonic), [1] sequences finished inside this launch, [2] status (kGridCtrlBytes, zeroed before every launch)
  int L, V, B, H, n_steps, Tvt, ldl, n_cta;
  int tn[8], ks[8];                    // per phase kind: n-tiles per item x k-splits (tn * ks == 8)
  int early_exit;
  int fence_mode;                      // debug: 1 = __threadfence() + atomicAdd (invalidates L1) instead of red.release
  float* dbg_logits;                   // parity path: [kept steps][B][V]
  const int32_t* dbg_slot;             // optional [n_steps] -> row block of dbg_logits, -1 = not kept
  const int32_t* forced;               // teacher forcing: next token of sequence b after step t = forced[b * stride + t]
  int forced_stride;
  unsigned long long* prof;            // optional [n phases + 1] %globaltimer stamps of CTA 0 in step prof_step
  int prof_step;
};

bool grid_eligible(int d_model, int d_ff, int n_head, int n_layer, int V);
size_t grid_packed_bytes(int d_model, int d_ff, int n_layer, int V);
struct GridPackSrc {
  const bf16 *w_in, *w_out, *w1, *w2;          // bf16 matrices [out, in] row-major
  const float *b_in, *b_out, *b1, *b2, *ln1w, *ln1b, *ln2w, *ln2b;
  bf16 *kh, *vt;
};
// 16-row fragment-major tiles (LayerNorm weights folded into in_proj / mlp.0), the fold vectors (`fold`: n_layer x 2 x (3 d + d_ff) floats)
// and the GridLayer table
int grid_pack_weights(cudaStream_t s, const GridPackSrc* src, const bf16* head, int n_layer, int d_model, int d_ff, int V, uint8_t* packed,
                      float* fold, GridLayer* layers, size_t* w_head);
// phase shapes for a batch of B sequences + the per-CTA item lists (host arrays: items [n_cta][kMaxItems], n_items [n_cta]);
// MG_E_SHAPE when a CTA would get more than kMaxItems items
int grid_plan(int d_model, int d_ff, int n_layer, int V, int B, int n_cta, int* tn, int* ks, GridItem* items, int32_t* n_items);
int grid_init();                                                     // cudaFuncSetAttribute; MG_OK / MG_E_CUDA
int grid_max_ctas(int d_model, int hd);                              // co-resident CTAs (cooperative launch), 0 when the query fails
// prefill caches kc / vc [B][d / 64][Tmax][64] -> kh / vt (host arrays of per-layer device pointers) for sequences 0 .. B - 1, or for the B
// sequences named by the device array `slots` (continuous batching: the newly admitted ones)
int grid_relayout_kv(cudaStream_t s, const bf16* const* kc, const bf16* const* vc, bf16* const* kh, bf16* const* vt, const int32_t* lens,
                     const int32_t* slots, int B, int n_layer, int d_model, int hd, int Tmax, int Tvt);
int launch_decode_grid(cudaStream_t s, const GridParams& p, int d_model, int hd);

}  // namespace grid
}  // namespace mg
