/*
 * mg_engine.h -- C ABI of the B200-native engine for the reference's hot path.
 *
 * The reference (RohitMurali18/Music-Generation-Emotion-Adaptive) has no FFI of its own: the
 * boundary is two Python call shapes plus two checkpoint layouts.  Every
 * entry point below names the reference interface it replaces.  Plain pointers and sizes only; the
 * caller owns every host buffer, the engine owns device weights, the KV arena, workspace and its
 * CUDA stream.  No pointer into engine memory is ever returned.
 *
 * Error convention: 0 = MG_OK, negative = error; mg_last_error() returns the message of the last
 * failure on the calling thread.  There is NO CPU fallback: without a usable GPU every create call
 * returns MG_E_CUDA.
 *
 * Threading: one engine per GPU; calls on one engine are serialised by an internal mutex (the
 * reference's caller is a sync FastAPI endpoint on a thread pool sharing one global model,
 * api_cache.py:186-187,108,161).  Different engines are independent (replicas).
 */
#ifndef MG_ENGINE_H_
#define MG_ENGINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 1

enum mg_status {
  MG_OK = 0,
  MG_E_SHAPE = -1,           /* tensor name/shape does not match the geometry                     */
  MG_E_PROMPT_TOO_LONG = -2, /* prompt longer than the position table (api_cache.py:99 would raise) */
  MG_E_TOPK = -3,            /* top_k > vocab (torch.topk raises in api_cache.py:172)              */
  MG_E_CUDA = -4,            /* CUDA error, no device, or wrong architecture                       */
  MG_E_OOM = -5,             /* arena/workspace allocation failed or capacity exceeded             */
  MG_E_STATE = -6,           /* call order violated (weights missing, nothing uploaded, ...)       */
  MG_E_ARG = -7,             /* invalid argument                                                   */
  MG_E_TOKEN = -8            /* token id outside [0, vocab)                                        */
};

enum mg_dtype_mode {
  MG_DTYPE_FP32 = 0, /* fp32 weights, KV and arithmetic: greedy tokens bit-identical to the reference */
  MG_DTYPE_BF16 = 1  /* bf16 weights + bf16 KV cache, fp32 accumulation                             */
};

/* Model geometry.  Replaces the shape inference of api_cache.py:31-37 plus the hard-coded n_head
 * of api_cache.py:112; d_ff is 4*d_model in every reference trainer (all scripts under train/). */
typedef struct mg_geometry {
  int32_t vocab_size;
  int32_t pos_rows;
  int32_t d_model;
  int32_t n_head;
  int32_t n_layer;
  int32_t d_ff;
} mg_geometry;

typedef struct mg_engine mg_engine; /* MIDI-token generator replica */
typedef struct mg_bert mg_bert;     /* DistilBERT emotion classifier replica */

/* ---- library ------------------------------------------------------------------------------- */
int mg_abi_version(void);
const char* mg_last_error(void);
/* Number of CUDA devices visible, or a negative mg_status. */
int mg_device_count(void);

/* ---- generator: construction and weights ---------------------------------------------------- */
/* Replaces `GPTWithKV(vocab_size, seq_len, d_model, n_head, n_layer)` (api_cache.py:108-114).
 * max_batch / max_seq size the KV arena: L * 2 * max_batch * max_seq * d_model elements. */
int mg_engine_create(const mg_geometry* geo, int device, int dtype_mode, int max_batch, int max_seq,
                     mg_engine** out);
void mg_engine_destroy(mg_engine* e);

/* Replaces `model.load_state_dict(remap_state_dict(ckpt["model"]))` (api_cache.py:118-138), one
 * tensor at a time.  `name` is the REMAPPED name (tok_emb.weight, pos_emb, head.weight, head.bias,
 * layers.N.{attn.in_proj_weight, attn.in_proj_bias, attn.out_proj.weight, attn.out_proj.bias,
 * ln1.weight, ln1.bias, ln2.weight, ln2.bias, mlp.0.weight, mlp.0.bias, mlp.2.weight, mlp.2.bias}).
 * `data` is row-major fp32 on the host ([out,in] for Linear weights). */
int mg_load_weight(mg_engine* e, const char* name, const float* data, const int64_t* shape, int ndim);
/* Verifies that all 4 + 12*n_layer tensors were loaded and makes the engine ready. */
int mg_engine_finalize(mg_engine* e);

/* ---- generator: the decode path -------------------------------------------------------------- */
/* Replaces `sample_kvcache(model, prompt, max_len, temperature, top_k, device)` (api_cache.py:159-184)
 * for a batch of B independent prompts (row b of the result == a batch-1 reference run on prompt b).
 *   prompt_ids / prompt_offsets : packed prompts, prompt b = ids[offsets[b] .. offsets[b+1])
 *   max_new_tokens              : the reference's `max_len - len(prompt)`; used for every sequence
 *                                 unless max_new_per_seq != NULL (then per sequence)
 *   temperature                 : > 0 (the reference divides by it)
 *   top_k                       : 0 = no top-k mask (reference top_k=None); 1 = greedy
 *   eos_id                      : -1 = never stop (vocab without [END_SEQUENCE], api_cache.py:181)
 *   seed                        : Philox key; stream = (seed, seq_index_base + b, step)
 *   out_ids [B][out_stride]     : prompt followed by generated ids; out_lens[b] = total length
 * Host buffers; H2D / D2H copies happen inside the call. */
int mg_generate(mg_engine* e, const int32_t* prompt_ids, const int32_t* prompt_offsets, int B,
                int max_new_tokens, const int32_t* max_new_per_seq, float temperature, int top_k,
                int eos_id, uint64_t seed, uint64_t seq_index_base, int32_t* out_ids, int out_stride,
                int32_t* out_lens);

/* The same path split at the host/device boundary (inputs resident in HBM for timing):
 *   mg_upload_prompts : H2D of the packed prompts                (pinned staging inside the engine)
 *   mg_run            : prefill + decode loop, device-resident, asynchronous on the engine stream
 *   mg_download       : D2H of the generated ids (synchronises) */
int mg_upload_prompts(mg_engine* e, const int32_t* prompt_ids, const int32_t* prompt_offsets, int B,
                      int max_new_tokens, const int32_t* max_new_per_seq);
int mg_run(mg_engine* e, float temperature, int top_k, int eos_id, uint64_t seed, uint64_t seq_index_base);
int mg_download(mg_engine* e, int32_t* out_ids, int out_stride, int32_t* out_lens);
int mg_synchronize(mg_engine* e);
/* cudaStream_t of the engine (as void*), so a host framework can record its own events on it. */
void* mg_engine_stream(mg_engine* e);

/* Parity/debug: teacher-forced logits of `GPTWithKV.forward` (api_cache.py:87-106) along the
 * reference loop.  Step 0 feeds the last prompt token against the prefilled cache (the duplicate
 * feed of api_cache.py:167-168); step i>0 feeds forced_ids[b][i-1].  logits_out is
 * [n_steps][B][vocab] fp32 on the host. */
int mg_step_logits(mg_engine* e, const int32_t* prompt_ids, const int32_t* prompt_offsets, int B,
                   const int32_t* forced_ids, int n_steps, float* logits_out);

/* The same run, keeping only the steps listed in want_steps[n_want] (distinct, each in [0, n_steps)):
 * logits_out is [n_want][B][vocab], block i = step want_steps[i].  For parity checks at the cache lengths
 * of BASELINE configs 3 and 4 (1030 / 4352 positions), where all steps x batch x vocab would be gigabytes. */
int mg_step_logits_at(mg_engine* e, const int32_t* prompt_ids, const int32_t* prompt_offsets, int B,
                      const int32_t* forced_ids, int n_steps, const int32_t* want_steps, int n_want,
                      float* logits_out);

/* Recompute mode: the no-cache twin `GPT.forward` + `sample` of generate_music/generate.py:25-61
 * (post-LN, ReLU, unmasked, true positions, whole sequence recomputed every step). Same arguments
 * as mg_generate; prompt + max_new_tokens must fit the position table. */
int mg_generate_nocache(mg_engine* e, const int32_t* prompt_ids, const int32_t* prompt_offsets, int B,
                        int max_new_tokens, float temperature, int top_k, int eos_id, uint64_t seed,
                        uint64_t seq_index_base, int32_t* out_ids, int out_stride, int32_t* out_lens);
/* Parity/debug for recompute mode: last-position logits [B][vocab] of one full forward. */
int mg_forward_nocache(mg_engine* e, const int32_t* ids, const int32_t* offsets, int B, float* logits_out);

/* Sampler on caller-provided logits (api_cache.py:169-178: /temperature, top-k, -1e10 mask,
 * softmax, multinomial).  logits [rows][vocab] fp32 host; out [rows] int32.  Used by the chi-square
 * parity test; rows use Philox streams (seed, seq_index_base + row, step). */
int mg_sample_logits(mg_engine* e, const float* logits, int rows, int vocab, float temperature, int top_k,
                     uint64_t seed, uint64_t seq_index_base, uint32_t step, int32_t* out);

/* Counters since creation: kernels launched by this library, bytes H2D, bytes D2H. */
int mg_engine_stats(mg_engine* e, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* Milliseconds of the last mg_run measured with CUDA events on the engine stream
 * (total, prefill part, decode part) and the number of decode steps it executed. */
int mg_last_run_timing(mg_engine* e, float* total_ms, float* prefill_ms, float* decode_ms, int* steps);

/* ---- slot sessions: continuous batching -------------------------------------------------------------------------------
 * Caller side in the reference: the /generate endpoint, one batch-1 sample_kvcache call per HTTP request on a shared model
 * (api_cache.py:186-204).  A slot session keeps n_slots independent sequences in flight on one engine: new requests are
 * admitted into free slots BETWEEN chunks of decode steps (their prompts are prefilled into the K/V rows of the slot, nothing
 * in flight is touched), finished ones are retired, so a stream of requests is served at batched-decode throughput.
 * Row b of the session behaves exactly like a batch-1 reference run of its request: same prefill, same loop, own Philox
 * stream (seed, seq_index) wherever and whenever it was admitted.
 *   mg_slots_begin  sampling parameters are fixed for the session (the decode path -- persistent cluster kernel or step
 *                   graph -- is chosen once); max_len = prompt + new tokens any request may reach.
 *   mg_slots_admit  n prompts (packed like mg_generate) into the free slots `slots[j]`; max_new[j] tokens each; seq_index[j] =
 *                   Philox sequence index of the request.  MG_E_STATE if a slot is still in flight.
 *   mg_slots_step   up to n_steps decode steps for every slot in flight (a slot stops at EOS or at its budget); returns
 *                   finished[b] (1 = idle or done) and out_len[b] for all slots -- ONE small D2H copy.
 *   mg_slots_fetch  token ids (prompt included) of one slot.
 *   mg_slots_end    leaves the session (any batch call does so implicitly). */
int mg_slots_begin(mg_engine* e, int n_slots, int max_len, float temperature, int top_k, int eos_id, uint64_t seed);
int mg_slots_admit(mg_engine* e, int n, const int32_t* slots, const int32_t* prompt_ids, const int32_t* prompt_offsets,
                   const int32_t* max_new, const int32_t* seq_index);
int mg_slots_step(mg_engine* e, int n_steps, uint8_t* finished, int32_t* out_len);
int mg_slots_fetch(mg_engine* e, int slot, int32_t* out_ids, int cap, int* n);
/* the rows of n slots in one go (out_ids [n][out_stride], out_stride >= the session's row stride = max_len rounded up to 8;
 * out_lens[j] = tokens of slots[j]): ONE synchronisation for all of them */
int mg_slots_fetch_many(mg_engine* e, int n, const int32_t* slots, int32_t* out_ids, int out_stride, int32_t* out_lens);
int mg_slots_end(mg_engine* e);

/* Device-side detokenisation to note events: replaces the per-token regex / float() / pretty_midi look-ups of the reference's
 * MIDI assembly loop (api_cache.py:157 note_re, :208-221) by a gather.  mg_set_note_table uploads ONE record per vocabulary entry,
 * built on the host once per checkpoint: kind 0 = any other token, 1 = "[INSTRUMENT] <name>" (value = GM program,
 * pretty_midi.instrument_name_to_program or 0, api_cache.py:211-212), 2 = a whole-line "[NOTE] [PITCH:p] [START:s] [END:e]
 * [DURATION:d]" token (value = pretty_midi.note_name_to_number(p), start / end = float(s) / float(e), api_cache.py:216-217).
 * mg_note_events walks the token ids of the LAST generation where they lie in HBM (prompt included, like the reference, which
 * iterates over all returned tokens): every instrument token opens a new instrument, a note goes to the most recent one and is
 * dropped while there is none.  Per sequence b: n_inst[b] / n_notes[b] events found (entries beyond max_inst / max_notes are
 * counted but not stored), inst_program / inst_token [B][max_inst], note_inst (index into that list) / note_pitch / note_start /
 * note_end [B][max_notes].  Caller-owned host buffers. */
int mg_set_note_table(mg_engine* e, const int32_t* kind, const int32_t* value, const float* start, const float* end, int V);
int mg_note_events(mg_engine* e, int max_inst, int max_notes, int32_t* n_inst, int32_t* inst_program, int32_t* inst_token,
                   int32_t* n_notes, int32_t* note_inst, int32_t* note_pitch, float* note_start, float* note_end);

/* Per-token latency of the last run: us_out[i] = time between the tokens of decode steps i and i + 1 of sequence 0
 * (%globaltimer stamps written on the device by whichever kernel finalised the token), at most cap values; *n = how many.
 * The reference has no counterpart (its loop is timed around api_cache.py:166-182 from Python); BASELINE's
 * "p50 ms/token @ batch 1" is the median of these. */
int mg_last_step_times(mg_engine* e, float* us_out, int cap, int* n);

/* Which decode path served the last mg_run / mg_generate / mg_step_logits of this engine: 0 = step graph (one launch per
 * kernel and step), 1 = persistent cluster kernel (decode_mega.cu), 2 = weight-stationary flow kernel (decode_flow.cu), 3 = grid-synchronous
 * persistent kernel (decode_grid.cu: the geometries the cluster kernel does not take, few sequences with long caches, or MG_GRID=1).
 * Tests assert that parity was checked on the path the benchmark runs. */
int mg_last_decode_path(mg_engine* e);

/* ---- classifier ------------------------------------------------------------------------------ */
typedef struct mg_bert_geometry {
  int32_t vocab_size, max_pos, dim, n_heads, n_layers, hidden_dim, num_labels;
} mg_bert_geometry;

/* Replaces `load_model()` (emotion_analysis/modeling.py:8-25) minus tokenizer and hub download. */
int mg_bert_create(const mg_bert_geometry* geo, int device, int max_tokens, mg_bert** out);
void mg_bert_destroy(mg_bert* b);
/* HF DistilBertForSequenceClassification tensor names; LoRA already merged (W + (alpha/r) B A). */
int mg_bert_load_weight(mg_bert* b, const char* name, const float* data, const int64_t* shape, int ndim);
int mg_bert_finalize(mg_bert* b);
/* Replaces the forward + argmax of `inference.predict` (emotion_analysis/inference.py:16-21) for N
 * texts of T tokens (row-major ids / mask; mask 1 = token, 0 = padding).  logits_out [N][num_labels],
 * label_out [N].  Host buffers. */
int mg_classify(mg_bert* b, const int32_t* ids, const uint8_t* mask, int N, int T, float* logits_out,
                int32_t* label_out);
/* Device-resident split of the same call, for timing with inputs in HBM. */
int mg_bert_upload(mg_bert* b, const int32_t* ids, const uint8_t* mask, int N, int T);
int mg_bert_run(mg_bert* b);
int mg_bert_download(mg_bert* b, float* logits_out, int32_t* label_out);
int mg_bert_synchronize(mg_bert* b);
void* mg_bert_stream(mg_bert* b);
int mg_bert_stats(mg_bert* b, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* ---- kernel-level test hooks (device work, host buffers) -------------------------------------- */
/* C[M,N] = act(A[M,K] * W[N,K]^T + bias) through the tcgen05/TMA bf16 GEMM (fp32 in/out on host,
 * converted to bf16 on device).  act: 0 none, 1 exact GELU, 2 ReLU. */
int mg_test_gemm_bf16(int device, const float* A, const float* W, const float* bias, int M, int N, int K,
                      int act, float* C);

/* Host-only: the work plan of the grid-synchronous decode kernel (decode_grid.cu) for a batch of B sequences on n_cta CTAs.
 * tn_ks[16]: n-tiles per item (8) then k-splits (8) per phase kind; items: n_cta x 96 records of 4 int16 {phase, row_tile,
 * sequence group, 0} in phase order; n_items[n_cta].  No device work (runs without a GPU). */
int mg_test_grid_plan(int d_model, int d_ff, int n_layer, int vocab, int B, int n_cta, int32_t* tn_ks, int16_t* items, int32_t* n_items);

#ifdef __cplusplus
}
#endif
#endif /* MG_ENGINE_H_ */
