// Launch wrappers of the non-tensor-core kernels (sm_100a): embedding + LayerNorm, SIMT GEMM / GEMV
// (fp32-exact mode and small batches), KV-cache append, split-K flash-decoding attention, non-causal
// attention for prefill / recompute / classifier, and the fused top-k Philox sampler.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace mg {

// Sampling parameters live in device memory so a captured decode-step graph can be reused across
// calls with different settings.
struct SampleParams {
  float temperature;
  int32_t top_k;        // 0 = whole vocabulary
  int32_t eos_id;       // -1 = never
  int32_t _pad;
  uint64_t seed;
  uint64_t seq_base;
};

// Per-sequence decode state (all device arrays of length B unless noted).
struct DecodeState {
  int32_t* cur_tok;     // token fed to the next step
  int32_t* lens;        // KV-cache length (positions already cached)
  int32_t* out_ids;     // [B][out_stride]
  int32_t* out_len;
  int32_t* n_new;       // tokens generated so far (also the Philox step index)
  int32_t* max_new;
  uint8_t* finished;
  int32_t out_stride;
  const int32_t* seq_idx;        // optional [B]: Philox sequence index of row b relative to seq_base (null: b itself) --
                                 // continuous batching gives every REQUEST its own stream whatever slot it lands in
  unsigned long long* step_ns;   // optional [max steps]: %globaltimer when the token of sequence 0 in decode step i was written
};

// x[r] = tok_emb[tok[r]] + pos_emb[pos ? pos[r] : 0]   (fp32 residual stream, optional)
// y[r] = apply_ln ? LayerNorm(x[r]) * w + b : x[r]      (activation dtype T)
template <typename T>
int launch_embed_ln(cudaStream_t s, const int32_t* tok, const int32_t* pos, const T* tok_emb, const T* pos_emb,
                    const float* w, const float* b, float* x, T* y, int M, int d, float eps, bool apply_ln);

// y[r] = LayerNorm(x[r]) * w + b  (TO);  x_out[r] = the same values in fp32 when x_out != nullptr
// (may alias x when TI == float: post-LN residual streams).
template <typename TI, typename TO>
int launch_layernorm(cudaStream_t s, const TI* x, const float* w, const float* b, TO* y, float* x_out, int M, int d,
                     float eps);

// SIMT GEMM with the same epilogue contract as the tensor-core one: D = epi(A[M,K] * W[N,K]^T).
// T = float: fp32 operands, plain fp32 FMA accumulation (no TF32) -- the bit-faithful mode.
// T = bf16 : small-batch decode (M below a tensor-core tile).  M <= 8 takes a warp-per-column GEMV.
// When T == float the typed output is epi.out_f32; when T == bf16 it is epi.out_bf16.
template <typename T>
int launch_gemm_simt(cudaStream_t s, const T* A, int lda, const T* W, int M, int N, int K, const GemmEpilogue& epi);

// Prefill: scatter the K and V thirds of qkv [M,3d] into the caches [B][d/64 slices][Tmax][64].
template <typename T>
int launch_kv_append(cudaStream_t s, const T* qkv, const int32_t* row_seq, const int32_t* row_pos, T* kcache, T* vcache,
                     int M, int d, int Tmax);

// Split-K flash-decoding: one query per sequence (all heads in one CTA), keys/values = the cached
// rows [0, lens[b]) plus the new token's own K/V (read from qkv and appended to the cache at row
// lens[b] by the same kernel).  No mask (reference api_cache.py:68).  out [B, d].
// ws_o [B][nsplit][d], ws_ml [B][nsplit][H][2], counters [B] (zero on entry, zero on exit).
template <typename T>
int launch_decode_attn(cudaStream_t s, const T* qkv, T* kcache, T* vcache, const int32_t* lens, const uint8_t* finished,
                       T* out, float* ws_o, float* ws_ml, uint32_t* counters, int B, int d, int H, int Tmax, int nsplit);

// Non-causal attention over packed variable-length sequences (prompt prefill, recompute mode and
// the classifier).  qkv [M,3d] packed rows; key_mask: optional per-row 0/1 (0 = padding key).
template <typename T>
int launch_encoder_attn(cudaStream_t s, const T* qkv, const int32_t* seq_start, const int32_t* seq_len,
                        const uint8_t* key_mask, T* out, int B, int d, int H, int max_len);

// Fused temperature -> top-k -> softmax -> Philox multinomial -> state update (one CTA / sequence).
// logits rows are `ld` floats apart.
int launch_sample_step(cudaStream_t s, const float* logits, int ld, int V, const SampleParams* sp, DecodeState st, int B);
// Stand-alone sampler on [rows, V] logits (parity tests).
int launch_sample_rows(cudaStream_t s, const float* logits, int ld, int rows, int V, const SampleParams* sp,
                       uint32_t step, int32_t* out);
// Recompute mode: x[b*Tcap + t] = tok_emb[out_ids[b][t]] + pos_emb[t] for t < out_len[b] (fp32 + T copy).
template <typename T>
int launch_nocache_embed(cudaStream_t s, const int32_t* out_ids, int out_stride, const int32_t* out_len, const T* tok_emb,
                         const T* pos_emb, float* x, T* y, int B, int Tcap, int d);
// One-time cudaFuncSetAttribute calls (must not happen inside a stream capture).
int kernels_init();
int gemm_tc_init();
// Copy prompts into the output buffer and initialise the per-sequence decode state after prefill.
int launch_decode_init(cudaStream_t s, const int32_t* prompt_ids, const int32_t* offsets, DecodeState st, int B);
int launch_slot_init(cudaStream_t s, const int32_t* prompt_ids, const int32_t* offsets, const int32_t* slots, const int32_t* max_new,
                     const int32_t* seq_idx, int32_t* seq_idx_out, DecodeState st, int n);
// Teacher forcing (mg_step_logits): cur_tok[b] = forced[b * stride + col]; lens[b] += 1.
int launch_force_next(cudaStream_t s, const int32_t* forced, int stride, int col, DecodeState st, int B);
// Count sequences still running into *active (device).
int launch_count_active(cudaStream_t s, const uint8_t* finished, int B, int32_t* active);
// out[b] = src[rows[b]] cast to TO (recompute mode / classifier: only one position feeds the head).
template <typename TI, typename TO>
int launch_gather_rows(cudaStream_t s, const TI* src, const int32_t* rows, TO* out, int B, int d);
// fp32 -> T conversion (weight upload).
template <typename T>
int launch_convert(cudaStream_t s, const float* src, T* dst, size_t n);
// argmax over [N, C] fp32 rows (classifier labels; lowest index wins ties like torch.argmax).
int launch_argmax_rows(cudaStream_t s, const float* logits, int N, int C, int32_t* out);

}  // namespace mg
