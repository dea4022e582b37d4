// Host interface of the weight-stationary persistent decode kernel (implementation: decode_flow.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.cuh"

namespace mg {
namespace flow {

constexpr int kMaxGroups = 8;          // sequence groups = warps per CTA (warp g of every CTA works for group g)
constexpr int kGroupSeqs = 8;          // sequences per group = the N of the m16n8k16 MMAs
constexpr int kMaxLayers = 8;
constexpr int kMaxTopK = 64;
constexpr int kMaxSM = 160;
constexpr int kMaxHeadTiles = 8;       // 16-row vocabulary tiles per SM
constexpr int kStageBytes = 8192;      // one K/V ring stage: K tile (4 KB) + V tile (4 KB) = 64 keys (head_dim 32) / 32 keys (64)
constexpr int kStages = 2;             // ring depth per warp
constexpr int kScratchBytes = 1536;    // per-warp sampler scratch
constexpr int kTileBiasBytes = 64;     // 16 fp32 appended to every weight tile
constexpr int kThreads = kMaxGroups * 32;

enum UnitType { U_NONE = -1, U_QKV = 0, U_OUT = 1, U_MLP1 = 2, U_MLP2 = 3 };

// What one SM holds and does (built on the host, read once at kernel start).
struct SmProgram {
  int32_t blob_off;                    // byte offset of this SM's weight blob in the packed buffer
  int32_t blob_bytes;
  int32_t unit_type[kMaxLayers];       // at most ONE dense unit per layer
  int32_t unit_tile[kMaxLayers];       // 16-row tile index inside that matrix
  int32_t unit_off[kMaxLayers];        // byte offset of the tile inside the blob
  int32_t n_head;                      // vocabulary tiles of the head
  int32_t head_tile[kMaxHeadTiles];
  int32_t head_off[kMaxHeadTiles];
};

struct FlowLayer {
  bf16* kc;                            // K cache [B][H][Tcap][hd], 16-byte chunks swizzled per row (see flow_kv_offset)
  bf16* vc;                            // V cache, same layout
};

// Byte sizes / offsets of the per-group exchange buffers (8-byte words: 32-bit payload + 32-bit stamp).
struct FlowExchange {
  uint64_t* base;                      // [groups][group_words]
  int64_t group_words;
  int32_t off_xin, off_xb, off_qkv, off_part, off_x1, off_x1b, off_h, off_logits, off_tmax, off_tok;   // word offsets inside a group
  int32_t nt, nt_pad;                  // vocabulary tiles, padded to a multiple of 64
};

struct FlowParams {
  const uint8_t* packed;               // all SM blobs
  const SmProgram* prog;               // [n_sm]
  const FlowLayer* layers;             // [n_layer]
  const bf16* tok_emb;
  const bf16* pos_emb;
  const SampleParams* sp;
  DecodeState st;
  FlowExchange xc;
  int n_layer, n_head, head_dim, V, B, n_groups, n_steps, Tcap, n_sm;
  int stagger_ns;                      // start offset between consecutive groups
  int early_exit;                      // EOS enabled: a group stops as soon as all of its sequences have finished
  float* dbg_logits;                   // parity path: [kept steps][B][V]
  const int32_t* dbg_slot;             // optional [n_steps] -> row block of dbg_logits, -1 = not kept
  const int32_t* forced;               // teacher forcing: next token of sequence b after step t = forced[b * stride + t]
  int forced_stride;
  int32_t* status;                     // [4]: 0 = ok; else code, SM, warp, detail (watchdog / internal error)
  unsigned long long* prof;            // optional timeline of group 0 in step `prof_steps`: [n_sm][48 slots][entry, inputs complete, done] ns
  int prof_steps;
};

struct FlowPlan {                      // host-side result of the unit -> SM assignment
  int n_sm = 0, nt = 0;
  size_t packed_bytes = 0, smem_bytes = 0, max_blob = 0;
  SmProgram prog[kMaxSM];
};

// Per-layer weight sources for packing (device pointers, fp32 masters in [out, in] row-major + fp32 vectors).
struct FlowWeightSrc {
  const float *w_in, *b_in, *w_out, *b_out, *w1, *b1, *w2, *b2, *ln1w, *ln1b, *ln2w, *ln2b;
};

int flow_init();                                                     // cudaFuncSetAttribute calls; MG_OK / MG_E_CUDA
bool flow_eligible(int d_model, int d_ff, int n_head, int n_layer, int V, int n_sm);
int flow_plan(int n_layer, int V, int n_sm, FlowPlan* plan);         // assignment + blob offsets; MG_E_OOM if it does not fit
FlowExchange flow_exchange_layout(int V, int n_head, int head_dim, int n_sm);
int flow_pack_weights(cudaStream_t s, const FlowPlan& plan, const FlowWeightSrc* layers, int n_layer, int n_head, const float* head_w,
                      const float* head_b, int V, uint8_t* packed);
int flow_tcap(int max_seq, int head_dim);                            // cache rows per (sequence, head): max_seq rounded up to whole tiles
// prefill caches [B][d/64][Tmax][64] -> flow caches, rows [0, lens[b])
int flow_relayout_kv(cudaStream_t s, const bf16* kc, const bf16* vc, bf16* fk, bf16* fv, const int32_t* lens, int B, int n_head,
                     int head_dim, int Tmax, int Tcap);
int launch_decode_flow(cudaStream_t s, const FlowParams& p, size_t smem_bytes);

}  // namespace flow
}  // namespace mg
