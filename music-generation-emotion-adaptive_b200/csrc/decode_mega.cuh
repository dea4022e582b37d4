// Host interface of the persistent cluster decode kernel (implementation: decode_mega.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.cuh"

namespace mg {
namespace mega {

constexpr int kMegaCluster = 4;       // CTAs per cluster = d_model / 64 feature slices
constexpr int kMegaMaxTopK = 64;      // the in-kernel sampler handles 1 <= top_k <= 64
constexpr int kMegaMaxNL = 2304;      // vocabulary rows per CTA, padded to 256: ceil(V / 4) <= 2304
constexpr int kMegaStageBytes = 32768;
constexpr int kMegaStagesPerLayer = 12;   // in_proj 3 + out_proj 1 + mlp.0 4 + mlp.2 4 stages of 32 KB
constexpr int kMegaMaxLayers = 8;
constexpr int kMegaMaxLayersSmem = 4;  // layers whose LN / bias parameters are kept in shared memory
#ifndef MG_MEGA_STAGES2
#define MG_MEGA_STAGES2 4
#endif
constexpr int kMegaStages2 = MG_MEGA_STAGES2;       // ring depth (32 KB stages) when <= 2 sequences per cluster
constexpr int kMegaStages4 = 2;       // ... when 3..4 sequences per cluster
constexpr int kMegaMaxSeqPerCluster = 4;

struct MegaLayer {
  const float *b_in, *b_out, *b1, *b2, *ln1w, *ln1b, *ln2w, *ln2b;
  bf16 *kc, *vc;                      // cache slices [B][4][Tmax][64], written by the prefill
  bf16 *kh, *vt;                      // caches of the persistent kernel: K head-major [B][4][head][Tvt][hd],
                                      // V per 32-key block transposed [B][4][head][Tvt / 32][hd][32] (see attn_tc)
};

struct MegaParams {
  const uint8_t* packed;              // pre-tiled weight stream (mega_pack_weights), [4 ranks][stages][32 KB]
  const MegaLayer* layers;            // device array [n_layer]
  const bf16* tok_emb;
  const bf16* pos_emb;
  const float* head_b;
  const SampleParams* sp;
  DecodeState st;
  int n_layer, head_dim, V, VS, NP;   // VS = ceil(V / 4) vocabulary rows per CTA, NP = 256-row tile pairs of the head (4 stages each)
  int head_tail;                      // 1: the last <= 64 rows of the slice are ONE extra stage (mega_head_tail)
  int B, S, Tmax, n_steps;            // S = sequences per cluster
  int Tvt;                            // keys per (sequence, head) of the V cache: Tmax rounded up to 32
  int early_exit;                     // EOS enabled: a cluster stops as soon as all of its sequences have finished
  // parity/debug (mg_step_logits): raw logits [n_steps][B][V] and teacher-forced next tokens [B][forced_stride]
  float* dbg_logits;
  const int32_t* dbg_slot;            // optional [n_steps]: row block of dbg_logits for each step, -1 = not kept
  const int32_t* forced;
  int forced_stride;
  // optional phase timeline of one step (globaltimer ns), written by cluster 0 / CTA 0: [64] entries
  unsigned long long* prof;
  int prof_step;
  int prof_thread;                    // compute thread (0..255) of cluster 0 / CTA 0 that writes the stamps
  int stagger_groups, stagger_ns;     // start offset (cluster % groups) * ns: phase de-synchronisation of the clusters
  int dbg_gemm;                       // timing experiment only: 1 = skip ldmatrix, 2 = skip the MMAs of the weight GEMMs (garbage results)
  int dbg_skip_loads;                 // timing experiment only: signal the stages without copying (results are garbage)
};

int mega_init();
// head rows per CTA -> 256-row tile pairs + an optional single-stage tail for a remainder of at most 64 rows
inline int mega_head_tail(int VS) { const int rem = VS % 256; return rem > 0 && rem <= 64 ? 1 : 0; }
inline int mega_head_pairs(int VS) { return mega_head_tail(VS) ? VS / 256 : (VS + 255) / 256; }
size_t mega_packed_bytes(int n_layer, int NP, int tail);
int mega_pack_weights(cudaStream_t stream, const bf16* const* w_in, const bf16* const* w_out, const bf16* const* w1,
                      const bf16* const* w2, const bf16* head, int n_layer, int V, int VS, int NP, int tail, void* dst);
int mega_relayout_kv(cudaStream_t stream, const MegaLayer* layers, const int32_t* lens, int B, int n_layer, int Tmax, int Tvt, int hd);
// the same for n sequences named by a device array of slot indices (continuous batching: newly admitted sequences only)
int mega_relayout_kv_slots(cudaStream_t stream, const MegaLayer* layers, const int32_t* lens, const int32_t* slots, int n, int n_layer,
                           int Tmax, int Tvt, int hd);
int mega_max_clusters(int smax);      // co-resident clusters (0 when the query fails)
int launch_decode_mega(cudaStream_t stream, const MegaParams& p, int n_clusters);

}  // namespace mega
}  // namespace mg
