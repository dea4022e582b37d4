// mg_engine: the MIDI-token generator replica behind the C ABI of include/mg_engine.h.
//
// Replaces, for one GPU (reference file:line in /root/reference):
//   model construction + weight load   api_cache.py:108-138
//   GPTWithKV.forward / GPTBlock       api_cache.py:51-74,87-106   (pre-LN, exact GELU, no final LN)
//   sample_kvcache                     api_cache.py:159-184        (prefill, duplicate last-token feed,
//                                                                   pos_emb[0] on decode, top-k sampler)
//   GPT.forward + sample (no cache)    generate_music/generate.py:25-35,46-61 (post-LN, ReLU, unmasked)
//
// HBM layout: weights in the engine dtype T ([out,in] row-major as in the checkpoint), biases and
// LayerNorm parameters fp32; KV cache per layer token-major [B_max][T_max][d] for K and for V
// (projected K/V, not the reference's LN1(x) rows: mathematically identical);
// fp32 residual stream x; T activations y / qkv / att / h; fp32 logits [B][ld_logits].
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"
#include "decode_flow.cuh"
#include "detok.cuh"
#include "decode_grid.cuh"
#include "decode_mega.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "mg_engine.h"

namespace mg {

static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
std::atomic<uint64_t> g_kernel_launches{0};

}  // namespace mg

using namespace mg;

struct LayerW {
  void *w_in = nullptr, *w_out = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *b_in = nullptr, *b_out = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *ln1w = nullptr, *ln1b = nullptr, *ln2w = nullptr, *ln2b = nullptr;
  WMaps m_in, m_out, m_w1, m_w2;
  void *kc = nullptr, *vc = nullptr;
  void *kh = nullptr, *vt = nullptr;                   // head-major K / block-transposed V caches (persistent decode kernel only)
};

struct mg_engine {
  mg_geometry geo{};
  int device = 0, dtype = 0, max_batch = 0, max_seq = 0;
  size_t esz = 4;
  bool use_tc = false, use_graph = true, ready = false;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  std::vector<void*> allocs;

  std::vector<LayerW> layers;
  void *tok_emb = nullptr, *pos_emb = nullptr, *head_w = nullptr;
  float* head_b = nullptr;
  WMaps m_head;
  std::set<std::string> loaded;
  float* stage_f32 = nullptr;          // fp32 staging for weight conversion
  size_t stage_elems = 0;

  // activations (rows_cap rows)
  int rows_cap = 0;
  float* x = nullptr;
  void *y = nullptr, *qkv = nullptr, *att = nullptr, *h = nullptr, *yl = nullptr;
  CUtensorMap tm_y, tm_att, tm_h, tm_yl;
  float* logits = nullptr;
  int ld_logits = 0;
  float *ws_o = nullptr, *ws_ml = nullptr;
  uint32_t* counters = nullptr;

  // per-call device arena (int32): prompts + row maps + decode state
  int32_t* d_arena = nullptr;
  int32_t* h_arena = nullptr;          // pinned
  cudaEvent_t ev_staged = nullptr;     // recorded behind every H2D copy out of h_arena / h_sp: waited for before they are rewritten
  size_t arena_cap = 0;
  int32_t *d_prompt = nullptr, *d_offsets = nullptr, *d_row_seq = nullptr, *d_row_pos = nullptr, *d_seq_start = nullptr,
          *d_seq_len = nullptr, *d_last_rows = nullptr;
  DecodeState st{};
  int32_t* d_out_block = nullptr;      // [out_len (B) | out_ids (B * stride)] contiguous for one D2H
  // ---- slot session (continuous batching, mg_slots_*): persistent per-slot decode state outside the per-call arena ----
  bool slots_active = false, slots_mega = false, slots_grid = false;
  int n_slots = 0, slots_eos = -1, slots_topk = 0;
  int32_t* d_slot_state = nullptr;      // cur_tok | lens | n_new | max_new | seq_idx | finished (bytes) | last_rows, n_slots each
  int32_t* d_slot_last_rows = nullptr;
  int32_t* d_slot_out = nullptr;        // out_len [n] | out_ids [n][stride]
  int32_t* h_slot_flags = nullptr;      // pinned: finished [n] (bytes, padded) | out_len [n]
  int32_t* h_slot_rows = nullptr;       // pinned staging for mg_slots_fetch_many: [n][stride]
  size_t slot_rows_cap = 0;
  size_t slot_state_cap = 0, slot_out_cap = 0;
  std::vector<uint8_t> slot_busy;       // host mirror: admitted and not yet reported finished
  int4* d_note_table = nullptr;         // device-side detokenisation: one record per vocabulary entry (detok.cu)
  int32_t* d_detok = nullptr;            // result block of mg_note_events
  int32_t* h_detok = nullptr;            // pinned
  size_t detok_cap = 0;
  unsigned long long* d_step_ns = nullptr;   // [max_seq + 1] %globaltimer per decode step of sequence 0 (mg_last_step_times)
  int32_t* h_out_block = nullptr;      // pinned
  size_t out_cap = 0;
  SampleParams* d_sp = nullptr;
  SampleParams* h_sp = nullptr;        // pinned
  int32_t* d_active = nullptr;
  int32_t* h_active = nullptr;         // pinned
  int32_t* d_forced = nullptr;
  size_t forced_cap = 0;

  int cur_B = 0, cur_M = 0, cur_max_tp = 0, cur_steps = 0;
  bool uploaded = false;

  // persistent cluster decode kernel (decode_mega.cu): bf16, d_model 256, d_ff 1024
  bool mega_ok = false, use_mega = true, last_run_mega = false;
  uint8_t* d_mega_packed = nullptr;
  mega::MegaLayer* d_mega_layers = nullptr;
  unsigned long long* d_prof = nullptr;
  int mega_clusters2 = 0, mega_clusters4 = 0;   // co-resident clusters for <= 2 / <= 4 sequences per cluster

  // weight-stationary flow kernel (decode_flow.cu): bf16, d_model 256, d_ff 1024, <= 64 sequences
  bool flow_candidate = false, flow_ok = false, use_flow = true, last_run_flow = false;
  std::map<std::string, float*> masters;     // fp32 device copies of the matrices, kept until the flow weights are packed
  flow::FlowPlan* flow_plan = nullptr;
  flow::FlowExchange flow_xc{};
  uint8_t* d_flow_packed = nullptr;
  flow::SmProgram* d_flow_prog = nullptr;
  flow::FlowLayer* d_flow_layers = nullptr;
  std::vector<flow::FlowLayer> flow_layers;
  int32_t* d_flow_status = nullptr;
  int32_t* h_flow_status = nullptr;           // pinned
  unsigned long long* d_flow_prof = nullptr;
  int flow_tcap = 0, n_sm = 0;

  // grid-synchronous decode kernel (decode_grid.cu): bf16, d_model 256 / 512, <= 64 sequences
  bool grid_ok = false, use_grid = false, last_run_grid = false;
  int grid_mode = 2;                           // MG_GRID: 0 = never, 1 = wherever eligible, unset = where the cluster kernel does not take the geometry
  uint8_t* d_grid_packed = nullptr;
  float* d_grid_fold = nullptr;
  grid::GridLayer grid_layers[grid::kMaxLayers]{};
  size_t grid_w_head = 0;
  grid::GridItem* d_grid_items = nullptr;
  int32_t* d_grid_nitems = nullptr;
  int grid_plan_B = -1, grid_ctas = 0, grid_ldl = 0;
  int grid_tn[8]{}, grid_ks[8]{};
  float *g_x = nullptr, *g_x1 = nullptr, *g_q = nullptr, *g_logits = nullptr, *g_vals = nullptr, *g_part = nullptr;
  bf16 *g_knew = nullptr, *g_vnew = nullptr, *g_h = nullptr, *g_xb = nullptr, *g_x1b = nullptr;
  float *g_sx = nullptr, *g_sx1 = nullptr;
  unsigned* d_grid_ctrl = nullptr;
  unsigned* h_grid_ctrl = nullptr;             // pinned
  unsigned long long* d_grid_prof = nullptr;

  cudaGraphExec_t graph = nullptr;
  int graph_B = -1;
  uint64_t graph_kernels = 0;          // kernels inside the captured decode step (counted per replay)

  uint64_t h2d = 0, d2h = 0, launches0 = 0;
  float t_total = 0, t_prefill = 0, t_decode = 0;
  int t_steps = 0;

  template <typename P> int dmalloc(P** p, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(MG_E_OOM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    allocs.push_back(q);
    *p = reinterpret_cast<P*>(q);
    return MG_OK;
  }
  void dfree(void* q) {
    if (!q) return;
    auto it = std::find(allocs.begin(), allocs.end(), q);
    if (it != allocs.end()) allocs.erase(it);
    cudaFree(q);
  }
};

namespace {

bool run_decode_mega(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits = nullptr,
                     const int32_t* forced = nullptr, int forced_stride = 0, const int32_t* dbg_slot = nullptr);
// flow status block: [0..3] first failure (code, SM, group, detail), [8 + 2 (sm * 8 + group)] = what every warp was waiting for
constexpr int kFlowStatusInts = 8 + 2 * flow::kMaxSM * flow::kMaxGroups;
bool run_decode_flow(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits = nullptr,
                     const int32_t* forced = nullptr, int forced_stride = 0, const int32_t* dbg_slot = nullptr);
// The persistent decode paths, best first: the weight-stationary flow kernel (decode_flow.cu), then the cluster kernel
// (decode_mega.cu).  Returns false when neither can serve the call (the caller falls back to the step-graph path).
bool run_decode_persistent(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits = nullptr,
                           const int32_t* forced = nullptr, int forced_stride = 0, const int32_t* dbg_slot = nullptr);
int persistent_status(mg_engine* e);          // after a stream sync: MG_E_CUDA if the flow kernel's watchdog fired
int setup_mega(mg_engine* e);
int setup_grid(mg_engine* e);
int setup_flow(mg_engine* e);

int decode_nsplit(const mg_engine* e, int B) {
  int n = ceil_div(2 * 148, B);
  n = std::min(n, std::max(1, e->max_seq / 64));
  return std::max(1, n);
}

// ---- GEMM dispatch: tcgen05 for bf16 with a full enough tile, SIMT otherwise -------------------
template <typename T>
int gemm(mg_engine* e, const void* A, const CUtensorMap* tmA, const void* W, const WMaps* wm, int M, int N, int K,
         GemmEpilogue epi, float* out_f32_typed, void* out_typed) {
  // typed output: float -> out_f32, bf16 -> out_bf16
  if (out_typed) {
    if (std::is_same<T, float>::value) epi.out_f32 = reinterpret_cast<float*>(out_typed);
    else epi.out_bf16 = reinterpret_cast<bf16*>(out_typed);
  }
  if (out_f32_typed) epi.out_f32 = out_f32_typed;
  if (std::is_same<T, bf16>::value && e->use_tc && M >= 32 && tmA && wm && wm->ok) {
    if (gemm_use_pair(M, N)) return launch_gemm_tc_pair(e->stream, tmA, &wm->m[bn_index(128)], M, N, K, epi);
    const int bn = pick_gemm_bn(M, N);
    return launch_gemm_tc(e->stream, tmA, &wm->m[bn_index(bn)], M, N, K, epi, bn);
  }
  return launch_gemm_simt<T>(e->stream, reinterpret_cast<const T*>(A), K, reinterpret_cast<const T*>(W), M, N, K, epi);
}

int ensure_rows(mg_engine* e, int rows) {
  if (rows <= e->rows_cap) return MG_OK;
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  const int d = e->geo.d_model, f = e->geo.d_ff;
  rows = std::max(rows, 128);
  e->dfree(e->x); e->dfree(e->y); e->dfree(e->qkv); e->dfree(e->att); e->dfree(e->h);
  e->x = nullptr; e->y = e->qkv = e->att = e->h = nullptr;
  MG_TRY(e->dmalloc(&e->x, sizeof(float) * rows * d));
  MG_TRY(e->dmalloc(&e->y, e->esz * rows * d));
  MG_TRY(e->dmalloc(&e->qkv, e->esz * rows * 3 * d));
  MG_TRY(e->dmalloc(&e->att, e->esz * rows * d));
  MG_TRY(e->dmalloc(&e->h, e->esz * rows * f));
  MG_CUDA_OK(cudaMemsetAsync(e->y, 0, e->esz * rows * d, e->stream));
  MG_CUDA_OK(cudaMemsetAsync(e->att, 0, e->esz * rows * d, e->stream));
  MG_CUDA_OK(cudaMemsetAsync(e->h, 0, e->esz * rows * f, e->stream));
  if (e->dtype == MG_DTYPE_BF16 && e->use_tc) {
    MG_TRY(make_tmap_bf16_2d(&e->tm_y, e->y, rows, d, kGemmBM));
    MG_TRY(make_tmap_bf16_2d(&e->tm_att, e->att, rows, d, kGemmBM));
    MG_TRY(make_tmap_bf16_2d(&e->tm_h, e->h, rows, f, kGemmBM));
  }
  e->rows_cap = rows;
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  e->graph_B = -1;
  return MG_OK;
}

// ---- one transformer block of model (B) on M rows; `decode` picks the attention kernel ---------
template <typename T>
int block_kv(mg_engine* e, int l, int M, bool decode, int B, int nsplit, bool ln1_done) {
  const mg_geometry& g = e->geo;
  const int d = g.d_model, f = g.d_ff;
  LayerW& w = e->layers[l];
  T* y = reinterpret_cast<T*>(e->y);
  T* qkv = reinterpret_cast<T*>(e->qkv);
  T* att = reinterpret_cast<T*>(e->att);
  if (!ln1_done) MG_TRY((launch_layernorm<float, T>(e->stream, e->x, w.ln1w, w.ln1b, y, nullptr, M, d, 1e-5f)));
  {
    GemmEpilogue epi; epi.bias = w.b_in; epi.ld_out = 3 * d;
    MG_TRY(gemm<T>(e, y, &e->tm_y, w.w_in, &w.m_in, M, 3 * d, d, epi, nullptr, qkv));
  }
  if (decode) {
    MG_TRY(launch_decode_attn<T>(e->stream, qkv, reinterpret_cast<T*>(w.kc), reinterpret_cast<T*>(w.vc), e->st.lens,
                                 e->st.finished, att, e->ws_o, e->ws_ml, e->counters, B, d, g.n_head, e->max_seq, nsplit));
  } else {
    MG_TRY(launch_kv_append<T>(e->stream, qkv, e->d_row_seq, e->d_row_pos, reinterpret_cast<T*>(w.kc),
                               reinterpret_cast<T*>(w.vc), M, d, e->max_seq));
    MG_TRY(launch_encoder_attn<T>(e->stream, qkv, e->d_seq_start, e->d_seq_len, nullptr, att, B, d, g.n_head, e->cur_max_tp));
  }
  {
    GemmEpilogue epi; epi.bias = w.b_out; epi.resid_f32 = e->x; epi.ld_out = d;
    MG_TRY(gemm<T>(e, att, &e->tm_att, w.w_out, &w.m_out, M, d, d, epi, e->x, nullptr));
  }
  MG_TRY((launch_layernorm<float, T>(e->stream, e->x, w.ln2w, w.ln2b, y, nullptr, M, d, 1e-5f)));
  {
    GemmEpilogue epi; epi.bias = w.b1; epi.act = ACT_GELU; epi.ld_out = f;
    MG_TRY(gemm<T>(e, y, &e->tm_y, w.w1, &w.m_w1, M, f, d, epi, nullptr, e->h));
  }
  {
    GemmEpilogue epi; epi.bias = w.b2; epi.resid_f32 = e->x; epi.ld_out = d;
    MG_TRY(gemm<T>(e, e->h, &e->tm_h, w.w2, &w.m_w2, M, d, f, epi, e->x, nullptr));
  }
  return MG_OK;
}

// Prompt rows -> K/V caches (B packed prompts, M rows in total; the cache row of a prompt row is named by d_row_seq / d_row_pos)
template <typename T>
int prefill_rows(mg_engine* e, int M, int B) {
  const mg_geometry& g = e->geo;
  MG_TRY(launch_embed_ln<T>(e->stream, e->d_prompt, e->d_row_pos, reinterpret_cast<const T*>(e->tok_emb),
                            reinterpret_cast<const T*>(e->pos_emb), e->layers[0].ln1w, e->layers[0].ln1b, e->x,
                            reinterpret_cast<T*>(e->y), M, g.d_model, 1e-5f, true));
  for (int l = 0; l < g.n_layer; ++l) MG_TRY(block_kv<T>(e, l, M, false, B, 1, l == 0));
  // logits of the prefill are discarded by the reference (api_cache.py:163): the head is skipped
  return MG_OK;
}

template <typename T>
int prefill(mg_engine* e) {
  MG_TRY(prefill_rows<T>(e, e->cur_M, e->cur_B));
  MG_TRY(launch_decode_init(e->stream, e->d_prompt, e->d_offsets, e->st, e->cur_B));
  return MG_OK;
}

// embed(cur_tok) + pos_emb[0] -> L blocks -> head logits            (api_cache.py:87-106 with T == 1)
template <typename T>
int decode_forward(mg_engine* e) {
  const mg_geometry& g = e->geo;
  const int B = e->cur_B, nsplit = decode_nsplit(e, B);
  MG_TRY(launch_embed_ln<T>(e->stream, e->st.cur_tok, nullptr, reinterpret_cast<const T*>(e->tok_emb),
                            reinterpret_cast<const T*>(e->pos_emb), e->layers[0].ln1w, e->layers[0].ln1b, e->x,
                            reinterpret_cast<T*>(e->y), B, g.d_model, 1e-5f, true));
  for (int l = 0; l < g.n_layer; ++l) MG_TRY(block_kv<T>(e, l, B, true, B, nsplit, l == 0));
  // no final LayerNorm (api_cache.py:105); the head reads the residual stream -> cast it into y
  MG_TRY((launch_gather_rows<float, T>(e->stream, e->x, e->d_last_rows, reinterpret_cast<T*>(e->y), B, g.d_model)));
  GemmEpilogue epi; epi.bias = e->head_b; epi.ld_out = e->ld_logits;
  MG_TRY(gemm<T>(e, e->y, &e->tm_y, e->head_w, &e->m_head, B, g.vocab_size, g.d_model, epi, e->logits, nullptr));
  return MG_OK;
}

template <typename T>
int decode_step(mg_engine* e) {
  MG_TRY(decode_forward<T>(e));
  MG_TRY(launch_sample_step(e->stream, e->logits, e->ld_logits, e->geo.vocab_size, e->d_sp, e->st, e->cur_B));
  return MG_OK;
}

template <typename T>
int run_decode_loop(mg_engine* e, int eos_id) {
  const int steps = e->cur_steps;
  bool graph_ok = false;
  if (e->use_graph && steps > 1) {
    if (!e->graph || e->graph_B != e->cur_B) {
      if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
      cudaGraph_t gr = nullptr;
      MG_CUDA_OK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
      const uint64_t before = g_kernel_launches.load();
      const int rc = decode_step<T>(e);
      e->graph_kernels = g_kernel_launches.load() - before;
      g_kernel_launches.fetch_sub(e->graph_kernels);          // capture launches nothing; replays are counted
      cudaError_t ce = cudaStreamEndCapture(e->stream, &gr);
      if (rc != MG_OK) { if (gr) cudaGraphDestroy(gr); return rc; }
      if (ce != cudaSuccess) return fail(MG_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
      ce = cudaGraphInstantiate(&e->graph, gr, 0);
      cudaGraphDestroy(gr);
      if (ce != cudaSuccess) return fail(MG_E_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
      e->graph_B = e->cur_B;
    }
    graph_ok = true;
  }
  int done_steps = 0;
  for (int i = 0; i < steps; ++i) {
    if (graph_ok) {
      MG_CUDA_OK(cudaGraphLaunch(e->graph, e->stream));
      g_kernel_launches.fetch_add(e->graph_kernels, std::memory_order_relaxed);
    } else {
      MG_TRY(decode_step<T>(e));
    }
    ++done_steps;
    if (eos_id >= 0 && (i % 32) == 31 && i + 1 < steps) {
      MG_TRY(launch_count_active(e->stream, e->st.finished, e->cur_B, e->d_active));
      MG_CUDA_OK(cudaMemcpyAsync(e->h_active, e->d_active, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      MG_CUDA_OK(cudaStreamSynchronize(e->stream));
      if (*e->h_active == 0) break;
    }
  }
  e->t_steps = done_steps;
  return MG_OK;
}

template <typename T>
int run_impl(mg_engine* e, float temperature, int top_k, int eos_id, uint64_t seed, uint64_t seq_base) {
  MG_CUDA_OK(cudaEventSynchronize(e->ev_staged));           // the previous run's copy out of h_sp (and any upload) has been consumed
  *e->h_sp = SampleParams{temperature, top_k, eos_id, 0, seed, seq_base};
  MG_CUDA_OK(cudaMemcpyAsync(e->d_sp, e->h_sp, sizeof(SampleParams), cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaEventRecord(e->ev_staged, e->stream));
  MG_CUDA_OK(cudaEventRecord(e->ev[0], e->stream));
  MG_TRY(prefill<T>(e));
  MG_CUDA_OK(cudaEventRecord(e->ev[1], e->stream));
  int mrc = MG_OK;
  e->last_run_mega = std::is_same<T, bf16>::value && run_decode_persistent(e, top_k, eos_id, &mrc);
  MG_TRY(mrc);
  if (!e->last_run_mega) MG_TRY(run_decode_loop<T>(e, eos_id));
  MG_CUDA_OK(cudaEventRecord(e->ev[2], e->stream));
  return MG_OK;
}

// ---- recompute mode: model (A), generate_music/generate.py:25-35 -------------------------------
// Row layout: sequence b owns rows [b*Tcap, b*Tcap + len_b); seq_start / seq_len describe it.
template <typename T>
int forward_nocache(mg_engine* e, int B, int Tcap, int max_len_now) {
  const mg_geometry& g = e->geo;
  const int d = g.d_model, f = g.d_ff, M = B * Tcap;
  T* y = reinterpret_cast<T*>(e->y);
  MG_TRY(launch_nocache_embed<T>(e->stream, e->st.out_ids, e->st.out_stride, e->st.out_len,
                                 reinterpret_cast<const T*>(e->tok_emb), reinterpret_cast<const T*>(e->pos_emb), e->x, y, B,
                                 Tcap, d));
  for (int l = 0; l < g.n_layer; ++l) {
    LayerW& w = e->layers[l];
    { GemmEpilogue epi; epi.bias = w.b_in; epi.ld_out = 3 * d;
      MG_TRY(gemm<T>(e, y, &e->tm_y, w.w_in, &w.m_in, M, 3 * d, d, epi, nullptr, e->qkv)); }
    MG_TRY(launch_encoder_attn<T>(e->stream, reinterpret_cast<const T*>(e->qkv), e->d_seq_start, e->st.out_len, nullptr,
                                  reinterpret_cast<T*>(e->att), B, d, g.n_head, max_len_now));
    { GemmEpilogue epi; epi.bias = w.b_out; epi.resid_f32 = e->x; epi.ld_out = d;
      MG_TRY(gemm<T>(e, e->att, &e->tm_att, w.w_out, &w.m_out, M, d, d, epi, e->x, nullptr)); }
    // post-LN: x = norm1(x + sa(x))   (nn.TransformerEncoderLayer, norm_first=False)
    MG_TRY((launch_layernorm<float, T>(e->stream, e->x, w.ln1w, w.ln1b, y, e->x, M, d, 1e-5f)));
    { GemmEpilogue epi; epi.bias = w.b1; epi.act = ACT_RELU; epi.ld_out = f;
      MG_TRY(gemm<T>(e, y, &e->tm_y, w.w1, &w.m_w1, M, f, d, epi, nullptr, e->h)); }
    { GemmEpilogue epi; epi.bias = w.b2; epi.resid_f32 = e->x; epi.ld_out = d;
      MG_TRY(gemm<T>(e, e->h, &e->tm_h, w.w2, &w.m_w2, M, d, f, epi, e->x, nullptr)); }
    MG_TRY((launch_layernorm<float, T>(e->stream, e->x, w.ln2w, w.ln2b, y, e->x, M, d, 1e-5f)));
  }
  return MG_OK;
}

// ---- persistent cluster decode kernel: eligibility, tables, launch ------------------------------
static int mega_tvt(int max_seq) { return (max_seq + 31) / 32 * 32; }

int setup_mega(mg_engine* e) {
  const mg_geometry& g = e->geo;
  e->mega_ok = false;
  const int hd = g.d_model / g.n_head;
  const int VS = ceil_div(g.vocab_size, mega::kMegaCluster);
  if (!e->use_mega || e->dtype != MG_DTYPE_BF16 || g.d_model != 256 || g.d_ff != 1024 || (hd != 32 && hd != 64) ||
      ceil_div(VS, 256) * 256 > mega::kMegaMaxNL || VS < mega::kMegaMaxTopK)
    return MG_OK;
  MG_TRY(mega::mega_init());
  e->mega_clusters2 = mega::mega_max_clusters(2);
  e->mega_clusters4 = mega::mega_max_clusters(4);
  if (e->mega_clusters2 <= 0 && e->mega_clusters4 <= 0) return MG_OK;
  const int L = g.n_layer;
  if (L > mega::kMegaMaxLayers || L > mega::kMegaMaxLayersSmem) return MG_OK;
  const int NP = mega::mega_head_pairs(VS), head_tail = mega::mega_head_tail(VS);
  std::vector<mega::MegaLayer> lay(L);
  std::vector<const bf16*> w_in(L), w_out(L), w1(L), w2(L);
  for (int l = 0; l < L; ++l) {
    LayerW& w = e->layers[l];
    w_in[l] = reinterpret_cast<const bf16*>(w.w_in); w_out[l] = reinterpret_cast<const bf16*>(w.w_out);
    w1[l] = reinterpret_cast<const bf16*>(w.w1); w2[l] = reinterpret_cast<const bf16*>(w.w2);
    if (!w.vt) {
      // zero-filled once so that V entries beyond a sequence's length are always finite (they meet probability 0)
      const size_t vt_bytes = sizeof(bf16) * static_cast<size_t>(e->max_batch) * g.d_model * mega_tvt(e->max_seq);
      MG_TRY(e->dmalloc(&w.vt, vt_bytes));
      MG_CUDA_OK(cudaMemsetAsync(w.vt, 0, vt_bytes, e->stream));
      // + one 32-row block of slack: the last block of the last sequence is read whole
      const size_t kh_bytes = vt_bytes + sizeof(bf16) * 32 * g.d_model;
      MG_TRY(e->dmalloc(&w.kh, kh_bytes));
      MG_CUDA_OK(cudaMemsetAsync(w.kh, 0, kh_bytes, e->stream));
    }
    lay[l] = mega::MegaLayer{w.b_in, w.b_out, w.b1, w.b2, w.ln1w, w.ln1b, w.ln2w, w.ln2b,
                             reinterpret_cast<bf16*>(w.kc), reinterpret_cast<bf16*>(w.vc), reinterpret_cast<bf16*>(w.kh),
                             reinterpret_cast<bf16*>(w.vt)};
  }
  if (!e->d_mega_packed) {
    MG_TRY(e->dmalloc(&e->d_mega_packed, mega::mega_packed_bytes(L, NP, head_tail)));
    MG_TRY(e->dmalloc(&e->d_mega_layers, sizeof(mega::MegaLayer) * L));
  }
  MG_TRY(mega::mega_pack_weights(e->stream, w_in.data(), w_out.data(), w1.data(), w2.data(),
                                 reinterpret_cast<const bf16*>(e->head_w), L, g.vocab_size, VS, NP, head_tail, e->d_mega_packed));
  MG_CUDA_OK(cudaMemcpyAsync(e->d_mega_layers, lay.data(), sizeof(mega::MegaLayer) * L, cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->mega_ok = true;
  return MG_OK;
}

// Returns true (and launches) when the persistent kernel can serve this call.
bool run_decode_mega(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits, const int32_t* forced, int forced_stride,
                     const int32_t* dbg_slot) {
  *rc = MG_OK;
  if (!e->mega_ok || top_k < 1 || top_k > mega::kMegaMaxTopK || e->cur_steps <= 0) return false;
  const int B = e->cur_B;
  int S = 0, n_clusters = 0;
  for (int s_try = 1; s_try <= mega::kMegaMaxSeqPerCluster; ++s_try) {
    const int avail = s_try <= 2 ? e->mega_clusters2 : e->mega_clusters4;
    const int need = ceil_div(B, s_try);
    if (avail > 0 && need <= avail) { S = s_try; n_clusters = need; break; }
  }
  if (S == 0) return false;
  const mg_geometry& g = e->geo;
  mega::MegaParams p{};
  p.packed = e->d_mega_packed; p.layers = e->d_mega_layers;
  p.tok_emb = reinterpret_cast<const bf16*>(e->tok_emb); p.pos_emb = reinterpret_cast<const bf16*>(e->pos_emb);
  p.head_b = e->head_b; p.sp = e->d_sp; p.st = e->st;
  p.n_layer = g.n_layer; p.head_dim = g.d_model / g.n_head; p.V = g.vocab_size;
  p.VS = ceil_div(g.vocab_size, mega::kMegaCluster); p.NP = mega::mega_head_pairs(p.VS); p.head_tail = mega::mega_head_tail(p.VS);
  p.B = B; p.S = S; p.Tmax = e->max_seq; p.n_steps = e->cur_steps; p.Tvt = mega_tvt(e->max_seq);
  p.dbg_logits = dbg_logits; p.dbg_slot = dbg_slot; p.forced = forced; p.forced_stride = forced_stride;
  p.early_exit = (eos_id >= 0 && forced == nullptr) ? 1 : 0;
  p.prof = nullptr; p.prof_step = -1;
  p.dbg_skip_loads = std::getenv("MG_MEGA_SKIP_LOADS") ? 1 : 0;
  p.stagger_groups = std::getenv("MG_MEGA_STAGGER_GROUPS") ? std::atoi(std::getenv("MG_MEGA_STAGGER_GROUPS")) : 0;
  p.stagger_ns = std::getenv("MG_MEGA_STAGGER_NS") ? std::atoi(std::getenv("MG_MEGA_STAGGER_NS")) : 0;
  p.dbg_gemm = std::getenv("MG_MEGA_GEMM_DBG") ? std::atoi(std::getenv("MG_MEGA_GEMM_DBG")) : 0;
  if (const char* ps = std::getenv("MG_MEGA_PROF_STEP")) {       // debug: phase timeline of one decode step -> stderr
    if (!e->d_prof) { if (e->dmalloc(&e->d_prof, 128 * sizeof(unsigned long long)) != MG_OK) e->d_prof = nullptr; }
    if (e->d_prof) {
      cudaMemsetAsync(e->d_prof, 0, 128 * sizeof(unsigned long long), e->stream);
      p.prof = e->d_prof; p.prof_step = std::atoi(ps);
      p.prof_thread = std::getenv("MG_MEGA_PROF_THREAD") ? std::atoi(std::getenv("MG_MEGA_PROF_THREAD")) : 0;
    }
  }
  if (e->slots_active) p.early_exit = forced == nullptr ? 1 : 0;    // a cluster whose slots are all idle stops at once
  // slot sessions convert the caches of newly admitted sequences themselves (mg_slots_admit); the rows a previous chunk of
  // decode steps appended exist only in the persistent layout and must not be overwritten from the prefill caches
  if (!e->slots_active)
    *rc = mega::mega_relayout_kv(e->stream, e->d_mega_layers, e->st.lens, B, g.n_layer, e->max_seq, p.Tvt, p.head_dim);
  if (*rc == MG_OK) *rc = mega::launch_decode_mega(e->stream, p, n_clusters);
  if (p.prof && *rc == MG_OK) {
    unsigned long long h[128];
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(h, e->d_prof, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[mega prof] step %d (ns since first stamp):", p.prof_step);
    for (int i = 0; i < 64 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - h[0]);
    fprintf(stderr, "\n[mega prof] layer 1 fine trace (SM cycles since layer start, warp 1 lane 0):");
    for (int i = 64; i < 96 && h[i]; ++i) fprintf(stderr, " %lld", (long long)(h[i] - h[64]));
    fprintf(stderr, "\n[mega prof] sampler fine trace (SM cycles since sampler start):");
    for (int i = 96; i < 128 && h[i]; ++i) fprintf(stderr, " %lld", (long long)(h[i] - h[96]));
    fprintf(stderr, "\n");
  }
  e->t_steps = e->cur_steps;
  return true;
}

// ---- grid-synchronous decode kernel: eligibility, packing, plan, launch -----------------------------------------------
static int alloc_persistent_caches(mg_engine* e) {
  const mg_geometry& g = e->geo;
  for (auto& w : e->layers) {
    if (w.vt) continue;
    // zero-filled once so that V entries beyond a sequence's length are always finite (they meet probability 0)
    const size_t vt_bytes = sizeof(bf16) * static_cast<size_t>(e->max_batch) * g.d_model * mega_tvt(e->max_seq);
    MG_TRY(e->dmalloc(&w.vt, vt_bytes));
    MG_CUDA_OK(cudaMemsetAsync(w.vt, 0, vt_bytes, e->stream));
    // + one 32-row block of slack: the last block of the last sequence is read whole
    const size_t kh_bytes = vt_bytes + sizeof(bf16) * 32 * g.d_model;
    MG_TRY(e->dmalloc(&w.kh, kh_bytes));
    MG_CUDA_OK(cudaMemsetAsync(w.kh, 0, kh_bytes, e->stream));
  }
  return MG_OK;
}

int setup_grid(mg_engine* e) {
  const mg_geometry& g = e->geo;
  e->grid_ok = false;
  if (e->grid_mode == 2) e->use_grid = e->use_mega;   // MG_NO_MEGA=1 means "no persistent kernel": the step graph
  if (!e->use_grid || e->dtype != MG_DTYPE_BF16 || !grid::grid_eligible(g.d_model, g.d_ff, g.n_head, g.n_layer, g.vocab_size)) return MG_OK;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
  if (!coop) return MG_OK;
  MG_TRY(grid::grid_init());
  const int hd = g.d_model / g.n_head, L = g.n_layer, D = g.d_model;
  e->grid_ctas = grid::grid_max_ctas(D, hd);
  if (e->grid_ctas <= 0) return MG_OK;
  MG_TRY(alloc_persistent_caches(e));
  std::vector<grid::GridPackSrc> src(L);
  for (int l = 0; l < L; ++l) {
    LayerW& w = e->layers[l];
    src[l] = grid::GridPackSrc{reinterpret_cast<const bf16*>(w.w_in), reinterpret_cast<const bf16*>(w.w_out), reinterpret_cast<const bf16*>(w.w1),
                               reinterpret_cast<const bf16*>(w.w2), w.b_in, w.b_out, w.b1, w.b2, w.ln1w, w.ln1b, w.ln2w, w.ln2b,
                               reinterpret_cast<bf16*>(w.kh), reinterpret_cast<bf16*>(w.vt)};
  }
  if (!e->d_grid_packed) {
    const int S = grid::kMaxSeqs;
    e->grid_ldl = (g.vocab_size + 15) / 16 * 16;
    MG_TRY(e->dmalloc(&e->d_grid_packed, grid::grid_packed_bytes(D, g.d_ff, L, g.vocab_size)));
    MG_TRY(e->dmalloc(&e->d_grid_fold, sizeof(float) * L * 2 * (3 * D + g.d_ff)));
    MG_TRY(e->dmalloc(&e->d_grid_items, sizeof(grid::GridItem) * grid::kMaxItems * e->grid_ctas));
    MG_TRY(e->dmalloc(&e->d_grid_nitems, sizeof(int32_t) * e->grid_ctas));
    const size_t part_floats = static_cast<size_t>(S) * g.n_head * grid::kMaxSplits * (hd + 4);
    struct { void** p; size_t bytes; } bufs[] = {
        {reinterpret_cast<void**>(&e->g_x), sizeof(float) * S * D}, {reinterpret_cast<void**>(&e->g_x1), sizeof(float) * S * D},
        {reinterpret_cast<void**>(&e->g_q), sizeof(float) * S * D}, {reinterpret_cast<void**>(&e->g_knew), sizeof(bf16) * S * D},
        {reinterpret_cast<void**>(&e->g_vnew), sizeof(bf16) * S * D}, {reinterpret_cast<void**>(&e->g_h), sizeof(bf16) * S * g.d_ff},
        {reinterpret_cast<void**>(&e->g_logits), sizeof(float) * S * e->grid_ldl}, {reinterpret_cast<void**>(&e->g_vals), sizeof(float) * S * e->grid_ldl},
        {reinterpret_cast<void**>(&e->g_part), sizeof(float) * part_floats},
        {reinterpret_cast<void**>(&e->g_xb), sizeof(bf16) * S * D}, {reinterpret_cast<void**>(&e->g_x1b), sizeof(bf16) * S * D},
        {reinterpret_cast<void**>(&e->g_sx), sizeof(float) * S * (D / 16) * 2}, {reinterpret_cast<void**>(&e->g_sx1), sizeof(float) * S * (D / 16) * 2}, {reinterpret_cast<void**>(&e->d_grid_ctrl), grid::kGridCtrlBytes}};
    for (auto& b : bufs) {
      MG_TRY(e->dmalloc(reinterpret_cast<uint8_t**>(b.p), b.bytes));
      MG_CUDA_OK(cudaMemsetAsync(*b.p, 0, b.bytes, e->stream));
    }
    MG_CUDA_OK(cudaMallocHost(&e->h_grid_ctrl, 64));
    std::memset(e->h_grid_ctrl, 0, 64);
  }
  MG_TRY(grid::grid_pack_weights(e->stream, src.data(), reinterpret_cast<const bf16*>(e->head_w), L, D, g.d_ff, g.vocab_size, e->d_grid_packed,
                                 e->d_grid_fold, e->grid_layers, &e->grid_w_head));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->grid_plan_B = -1;
  e->grid_ok = true;
  return MG_OK;
}

bool run_decode_grid(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits, const int32_t* forced, int forced_stride,
                     const int32_t* dbg_slot) {
  *rc = MG_OK;
  const int B = e->cur_B;
  if (!e->grid_ok || !e->use_grid || e->cur_steps <= 0 || B > grid::kMaxSeqs || (e->slots_active && !e->slots_grid)) return false;
  const mg_geometry& g = e->geo;
  const int hd = g.d_model / g.n_head;
  auto cuda_ok = [&](cudaError_t ce, const char* what) {
    if (ce == cudaSuccess) return true;
    *rc = fail(MG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    return false;
  };
  if (e->grid_plan_B != B) {
    std::vector<grid::GridItem> items(static_cast<size_t>(grid::kMaxItems) * e->grid_ctas);
    std::vector<int32_t> n_items(e->grid_ctas);
    if (grid::grid_plan(g.d_model, g.d_ff, g.n_layer, g.vocab_size, B, e->grid_ctas, e->grid_tn, e->grid_ks, items.data(), n_items.data()) != MG_OK)
      return false;
    if (!cuda_ok(cudaStreamSynchronize(e->stream), "grid plan sync")) return true;      // a previous launch may still read the tables
    if (!cuda_ok(cudaMemcpy(e->d_grid_items, items.data(), sizeof(grid::GridItem) * items.size(), cudaMemcpyHostToDevice), "grid items")) return true;
    if (!cuda_ok(cudaMemcpy(e->d_grid_nitems, n_items.data(), sizeof(int32_t) * n_items.size(), cudaMemcpyHostToDevice), "grid item counts")) return true;
    e->grid_plan_B = B;
  }
  grid::GridParams p{};
  p.packed = e->d_grid_packed; p.w_head = e->grid_w_head;
  for (int l = 0; l < g.n_layer; ++l) p.layers[l] = e->grid_layers[l];
  p.items = e->d_grid_items; p.n_items = e->d_grid_nitems;
  p.tok_emb = reinterpret_cast<const bf16*>(e->tok_emb); p.pos_emb = reinterpret_cast<const bf16*>(e->pos_emb);
  p.head_b = e->head_b; p.sp = e->d_sp; p.st = e->st;
  p.x = e->g_x; p.x1 = e->g_x1; p.q = e->g_q; p.knew = e->g_knew; p.vnew = e->g_vnew; p.h = e->g_h; p.logits = e->g_logits;
  p.vals = e->g_vals; p.part = e->g_part; p.ctrl = e->d_grid_ctrl;
  p.xb = e->g_xb; p.x1b = e->g_x1b; p.sx = e->g_sx; p.sx1 = e->g_sx1;
  p.L = g.n_layer; p.V = g.vocab_size; p.B = B; p.H = g.n_head; p.n_steps = e->cur_steps; p.Tvt = mega_tvt(e->max_seq);
  p.ldl = e->grid_ldl; p.n_cta = e->grid_ctas;
  for (int k = 0; k < 8; ++k) { p.tn[k] = e->grid_tn[k]; p.ks[k] = e->grid_ks[k]; }
  p.early_exit = (eos_id >= 0 && forced == nullptr) ? 1 : 0;
  if (e->slots_active) p.early_exit = forced == nullptr ? 1 : 0;    // a chunk ends as soon as every slot is idle
  p.dbg_logits = dbg_logits; p.dbg_slot = dbg_slot; p.forced = forced; p.forced_stride = forced_stride;
  p.fence_mode = std::getenv("MG_GRID_FENCE") ? std::atoi(std::getenv("MG_GRID_FENCE")) : 0;
  p.prof = nullptr; p.prof_step = -1;
  const int n_stamps = 2 * (5 * g.n_layer + 2) + 1;
  if (const char* ps = std::getenv("MG_GRID_PROF_STEP")) {          // debug: phase timeline of CTA 0 in one decode step -> stderr
    if (!e->d_grid_prof && e->dmalloc(&e->d_grid_prof, 128 * sizeof(unsigned long long)) != MG_OK) e->d_grid_prof = nullptr;
    if (e->d_grid_prof) {
      cudaMemsetAsync(e->d_grid_prof, 0, 128 * sizeof(unsigned long long), e->stream);
      p.prof = e->d_grid_prof; p.prof_step = std::atoi(ps);
    }
  }
  if (!cuda_ok(cudaMemsetAsync(e->d_grid_ctrl, 0, grid::kGridCtrlBytes, e->stream), "grid control reset")) return true;
  // slot sessions convert the caches of newly admitted sequences themselves (mg_slots_admit); the rows a previous chunk of decode steps
  // appended exist only in the persistent layout and must not be overwritten from the prefill caches
  if (!e->slots_active) {
    std::vector<const bf16*> kc(g.n_layer), vc(g.n_layer);
    std::vector<bf16*> kh(g.n_layer), vt(g.n_layer);
    for (int l = 0; l < g.n_layer; ++l) {
      kc[l] = reinterpret_cast<const bf16*>(e->layers[l].kc); vc[l] = reinterpret_cast<const bf16*>(e->layers[l].vc);
      kh[l] = reinterpret_cast<bf16*>(e->layers[l].kh); vt[l] = reinterpret_cast<bf16*>(e->layers[l].vt);
    }
    *rc = grid::grid_relayout_kv(e->stream, kc.data(), vc.data(), kh.data(), vt.data(), e->st.lens, nullptr, B, g.n_layer, g.d_model, hd, e->max_seq, p.Tvt);
  }
  if (*rc == MG_OK) *rc = grid::launch_decode_grid(e->stream, p, g.d_model, hd);
  if (*rc == MG_OK && !cuda_ok(cudaMemcpyAsync(e->h_grid_ctrl, e->d_grid_ctrl, 64, cudaMemcpyDeviceToHost, e->stream), "grid status copy")) return true;
  if (p.prof && *rc == MG_OK) {
    unsigned long long h[128];
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(h, e->d_grid_prof, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[grid prof] step %d, CTA 0, ns since step start: partition done | per phase (qkv attn out mlp1 mlp2 per layer, head, sampler): work done, barrier passed:", p.prof_step);
    for (int i = 0; i < n_stamps && i < 128 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - h[0]);
    fprintf(stderr, "\n[grid prof] mlp.0 of layer 1, CTA 0 thread 0, SM cycles: entry | statistics landed | operand built | MMAs done | epilogue issued (arrive) | CTA barrier | released (MEMBAR + RED) | poll done:");
    for (int i = 96; i < 112 && h[i]; ++i) fprintf(stderr, " %lld", (long long)(h[i] - h[96]));
    fprintf(stderr, "\n");
  }
  e->t_steps = e->cur_steps;
  e->last_run_grid = true;
  return true;
}

// ---- weight-stationary flow kernel: eligibility, plan, packing, launch ---------------------------------------------
int setup_flow(mg_engine* e) {
  const mg_geometry& g = e->geo;
  e->flow_ok = false;
  if (!e->flow_candidate) return MG_OK;
  const int L = g.n_layer, hd = g.d_model / g.n_head;
  for (int l = 0; l < L; ++l)
    for (const char* leaf : {"attn.in_proj_weight", "attn.out_proj.weight", "mlp.0.weight", "mlp.2.weight"})
      if (!e->masters.count("layers." + std::to_string(l) + "." + leaf)) return MG_OK;
  if (!e->masters.count("head.weight")) return MG_OK;
  MG_TRY(flow::flow_init());
  if (!e->flow_plan) e->flow_plan = new flow::FlowPlan();
  if (flow::flow_plan(L, g.vocab_size, e->n_sm, e->flow_plan) != MG_OK) return MG_OK;       // does not fit: other paths serve
  {
    uint64_t* keep = e->flow_xc.base;
    e->flow_xc = flow::flow_exchange_layout(g.vocab_size, g.n_head, hd, e->n_sm);
    e->flow_xc.base = keep;
  }
  e->flow_tcap = flow::flow_tcap(e->max_seq, hd);
  if (!e->d_flow_packed) {
    MG_TRY(e->dmalloc(&e->d_flow_packed, e->flow_plan->packed_bytes));
    MG_TRY(e->dmalloc(&e->d_flow_prog, sizeof(flow::SmProgram) * e->n_sm));
    MG_TRY(e->dmalloc(&e->d_flow_layers, sizeof(flow::FlowLayer) * L));
    MG_TRY(e->dmalloc(&e->d_flow_status, kFlowStatusInts * sizeof(int32_t)));
    MG_CUDA_OK(cudaMallocHost(&e->h_flow_status, kFlowStatusInts * sizeof(int32_t)));
    uint64_t* xbase = nullptr;
    MG_TRY(e->dmalloc(&xbase, sizeof(uint64_t) * e->flow_xc.group_words * flow::kMaxGroups));
    e->flow_xc.base = xbase;
    e->flow_layers.resize(L);
    // zero-filled once: rows past a sequence's length are masked, but must be finite (they meet probability 0)
    const size_t cache_bytes = sizeof(bf16) * static_cast<size_t>(e->max_batch) * g.d_model * e->flow_tcap;
    for (int l = 0; l < L; ++l) {
      MG_TRY(e->dmalloc(&e->flow_layers[l].kc, cache_bytes));
      MG_TRY(e->dmalloc(&e->flow_layers[l].vc, cache_bytes));
      MG_CUDA_OK(cudaMemsetAsync(e->flow_layers[l].kc, 0, cache_bytes, e->stream));
      MG_CUDA_OK(cudaMemsetAsync(e->flow_layers[l].vc, 0, cache_bytes, e->stream));
    }
    MG_CUDA_OK(cudaMemcpyAsync(e->d_flow_layers, e->flow_layers.data(), sizeof(flow::FlowLayer) * L, cudaMemcpyHostToDevice, e->stream));
  }
  std::vector<flow::FlowWeightSrc> src(L);
  for (int l = 0; l < L; ++l) {
    LayerW& w = e->layers[l];
    const std::string pfx = "layers." + std::to_string(l) + ".";
    src[l] = flow::FlowWeightSrc{e->masters[pfx + "attn.in_proj_weight"], w.b_in, e->masters[pfx + "attn.out_proj.weight"], w.b_out,
                                 e->masters[pfx + "mlp.0.weight"], w.b1, e->masters[pfx + "mlp.2.weight"], w.b2,
                                 w.ln1w, w.ln1b, w.ln2w, w.ln2b};
  }
  MG_CUDA_OK(cudaMemcpyAsync(e->d_flow_prog, e->flow_plan->prog, sizeof(flow::SmProgram) * e->n_sm, cudaMemcpyHostToDevice, e->stream));
  MG_TRY(flow::flow_pack_weights(e->stream, *e->flow_plan, src.data(), L, g.n_head, e->masters["head.weight"], e->head_b,
                                 g.vocab_size, e->d_flow_packed));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  for (auto& kv : e->masters) e->dfree(kv.second);                     // packed: the fp32 copies are not needed any more
  e->masters.clear();
  e->flow_ok = true;
  return MG_OK;
}

bool run_decode_flow(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits, const int32_t* forced, int forced_stride,
                     const int32_t* dbg_slot) {
  *rc = MG_OK;
  const int B = e->cur_B;
  if (!e->flow_ok || !e->use_flow || top_k < 1 || top_k > flow::kMaxTopK || e->cur_steps <= 0 ||
      B > flow::kMaxGroups * flow::kGroupSeqs)
    return false;
  const mg_geometry& g = e->geo;
  const int hd = g.d_model / g.n_head;
  flow::FlowParams p{};
  p.packed = e->d_flow_packed; p.prog = e->d_flow_prog; p.layers = e->d_flow_layers;
  p.tok_emb = reinterpret_cast<const bf16*>(e->tok_emb); p.pos_emb = reinterpret_cast<const bf16*>(e->pos_emb);
  p.sp = e->d_sp; p.st = e->st; p.xc = e->flow_xc;
  p.n_layer = g.n_layer; p.n_head = g.n_head; p.head_dim = hd; p.V = g.vocab_size; p.B = B;
  p.n_groups = std::min(flow::kMaxGroups, (B + flow::kGroupSeqs - 1) / flow::kGroupSeqs);
  if (const char* gs = std::getenv("MG_FLOW_GROUPS")) p.n_groups = std::max(1, std::min({flow::kMaxGroups, B, std::atoi(gs)}));
  if ((B + p.n_groups - 1) / p.n_groups > flow::kGroupSeqs) return false;
  p.n_steps = e->cur_steps; p.Tcap = e->flow_tcap; p.n_sm = e->n_sm;
  p.early_exit = (eos_id >= 0 && forced == nullptr) ? 1 : 0;
  p.stagger_ns = std::getenv("MG_FLOW_STAGGER_NS") ? std::atoi(std::getenv("MG_FLOW_STAGGER_NS")) : 6000;
  p.dbg_logits = dbg_logits; p.dbg_slot = dbg_slot; p.forced = forced; p.forced_stride = forced_stride;
  p.status = e->d_flow_status;
  p.prof = nullptr; p.prof_steps = 0;
  const size_t prof_words = static_cast<size_t>(flow::kMaxSM) * 48 * 8;
  if (const char* ps = std::getenv("MG_FLOW_PROF")) {               // debug: timeline of group 0 in one step -> stderr
    p.prof_steps = std::max(0, std::atoi(ps));
    if (!e->d_flow_prof && e->dmalloc(&e->d_flow_prof, sizeof(unsigned long long) * prof_words) != MG_OK) e->d_flow_prof = nullptr;
    if (e->d_flow_prof) { cudaMemsetAsync(e->d_flow_prof, 0, sizeof(unsigned long long) * prof_words, e->stream); p.prof = e->d_flow_prof; }
  }
  auto cuda_ok = [&](cudaError_t ce, const char* what) {
    if (ce == cudaSuccess) return true;
    *rc = fail(MG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    return false;
  };
  if (!cuda_ok(cudaMemsetAsync(e->d_flow_status, 0, kFlowStatusInts * sizeof(int32_t), e->stream), "flow status reset")) return true;
  // stamps restart at every launch: the exchange words of the previous job must not look valid
  if (!cuda_ok(cudaMemsetAsync(e->flow_xc.base, 0, sizeof(uint64_t) * e->flow_xc.group_words * flow::kMaxGroups, e->stream),
               "flow exchange reset")) return true;
  for (int l = 0; l < g.n_layer && *rc == MG_OK; ++l)
    *rc = flow::flow_relayout_kv(e->stream, reinterpret_cast<const bf16*>(e->layers[l].kc), reinterpret_cast<const bf16*>(e->layers[l].vc),
                                 e->flow_layers[l].kc, e->flow_layers[l].vc, e->st.lens, B, g.n_head, hd, e->max_seq, e->flow_tcap);
  if (*rc == MG_OK) *rc = flow::launch_decode_flow(e->stream, p, e->flow_plan->smem_bytes);
  if (*rc == MG_OK && !cuda_ok(cudaMemcpyAsync(e->h_flow_status, e->d_flow_status, kFlowStatusInts * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream),
                               "flow status copy")) return true;
  if (p.prof && *rc == MG_OK) {
    std::vector<unsigned long long> h(prof_words);
    cudaStreamSynchronize(e->stream);
    cudaMemcpy(h.data(), e->d_flow_prof, sizeof(unsigned long long) * prof_words, cudaMemcpyDeviceToHost);
    static const char* kind[5] = {"qkv", "attn", "out", "mlp1", "mlp2"};
    unsigned long long t_origin = ~0ull;
    for (size_t i = 0; i < prof_words; i += 8) if (h[i]) t_origin = std::min(t_origin, h[i]);
    fprintf(stderr, "[flow prof] step %d, group 0: per phase over its units: n | entry first..last | inputs complete first..last | done first..last | "
            "mean (done - inputs complete) ns\n", p.prof_steps);
    for (int slot = 0; slot < 5 * g.n_layer + 2; ++slot) {
      unsigned long long e0 = ~0ull, e1 = 0, r0 = ~0ull, r1 = 0, d0 = ~0ull, d1 = 0, work = 0, sen0 = ~0ull, sen1 = 0, retr = 0, rd = 0; int n = 0;
      for (int smi = 0; smi < e->n_sm; ++smi) {
        const unsigned long long* q = &h[(static_cast<size_t>(smi) * 48 + slot) * 8];
        if (!q[0]) continue;
        if (q[3]) { sen0 = std::min(sen0, q[3]); sen1 = std::max(sen1, q[3]); retr += q[4]; rd += q[5] - q[3]; }
        ++n; e0 = std::min(e0, q[0]); e1 = std::max(e1, q[0]);
        if (q[1]) { r0 = std::min(r0, q[1]); r1 = std::max(r1, q[1]); work += q[2] - q[1]; }
        d0 = std::min(d0, q[2]); d1 = std::max(d1, q[2]);
      }
      if (!n) continue;
      char name[32];
      if (slot < 5 * g.n_layer) snprintf(name, sizeof name, "L%d %s", slot / 5, kind[slot % 5]);
      else snprintf(name, sizeof name, "%s", slot == 5 * g.n_layer ? "head" : "sampler");
      fprintf(stderr, "[flow prof] %-8s n %3d | entry %7llu..%7llu | ready %7llu..%7llu | done %7llu..%7llu | work %5llu", name, n, e0 - t_origin,
              e1 - t_origin, r0 == ~0ull ? 0 : r0 - t_origin, r1 ? r1 - t_origin : 0, d0 - t_origin, d1 - t_origin, n ? work / n : 0);
      if (slot < 5 * g.n_layer && slot % 5 == 1) {
        unsigned long long wt = 0, nt = 0, tl = 0, c1 = 0, c2 = 0; int m = 0;
        for (int smi = 0; smi < e->n_sm; ++smi) {
          const unsigned long long* q = &h[(static_cast<size_t>(smi) * 48 + slot) * 8];
          if (q[0] && q[1] && q[5]) { wt += q[3]; nt += q[4]; tl += q[5] - q[1]; c1 += q[6]; c2 += q[7]; ++m; }
        }
        if (m) fprintf(stderr, " | tiles %.1f, tile loop %llu ns: waiting for tiles %llu ns, scores + softmax %llu cycles, P V %llu cycles", double(nt) / m, tl / m, wt / m, c1 / m, c2 / m);
      } else
      if (slot == 5 * g.n_layer + 1) {
        unsigned long long a = 0, b2 = 0, c2 = 0; int m = 0;
        for (int smi = 0; smi < e->n_sm; ++smi) {
          const unsigned long long* q = &h[(static_cast<size_t>(smi) * 48 + slot) * 8];
          if (q[0] && q[3] && q[1]) { a += q[3] - q[1]; b2 += q[4] - q[3]; c2 += q[5] - q[4]; ++m; }
        }
        if (m) fprintf(stderr, " | threshold + gather %llu, select + draw %llu, embed + publish %llu", a / m, b2 / m, c2 / m);
      } else
      if (sen1) fprintf(stderr, " | sentinel seen %7llu..%7llu, mean batch retries %.1f, mean read time %llu", sen0 - t_origin, sen1 - t_origin, double(retr) / n, rd / n);
      fprintf(stderr, "\n");
    }
  }
  e->t_steps = e->cur_steps;
  e->last_run_flow = true;
  return true;
}

bool run_decode_persistent(mg_engine* e, int top_k, int eos_id, int* rc, float* dbg_logits, const int32_t* forced, int forced_stride,
                           const int32_t* dbg_slot) {
  e->last_run_flow = e->last_run_grid = false;
  // Which persistent kernel: the grid-synchronous one where the cluster kernel does not take the geometry (or MG_GRID=1), and for FEW
  // sequences with LONG caches: a cluster streams its sequences' K/V at ~180 GB/s (4 SMs), the grid kernel spreads every (sequence, head)
  // over all SMs but pays ~20 us more per step for its grid barriers.  Measured crossovers (profiles/r2n_*: config 4 = 16 sequences,
  // mean cache length 2304: 91.8 vs 93.6-94.7 us per step): mean cache length >= 1000 / 1200 / 1800 for <= 4 / 8 / 16 sequences.
  bool grid_first = e->grid_mode == 1 || !e->mega_ok;
  if (!grid_first && e->grid_mode == 2 && e->grid_ok && !e->slots_active && e->cur_B > 0) {
    const int B = e->cur_B;
    const double mean_len = static_cast<double>(e->cur_M) / B + 0.5 * e->cur_steps;
    grid_first = (B <= 4 && mean_len >= 1000) || (B <= 8 && mean_len >= 1200) || (B <= 16 && mean_len >= 1800);
  }
  if (grid_first && run_decode_grid(e, top_k, eos_id, rc, dbg_logits, forced, forced_stride, dbg_slot)) return true;
  if (run_decode_flow(e, top_k, eos_id, rc, dbg_logits, forced, forced_stride, dbg_slot)) return true;
  if (run_decode_mega(e, top_k, eos_id, rc, dbg_logits, forced, forced_stride, dbg_slot)) return true;
  // what the cluster kernel refuses (top_k = None or > 64) still beats the step graph on the grid kernel (any top_k, <= 64 sequences)
  return !grid_first && e->grid_mode == 2 && !e->slots_active && run_decode_grid(e, top_k, eos_id, rc, dbg_logits, forced, forced_stride, dbg_slot);
}

// After the stream has been synchronised: did the flow kernel's watchdog fire?
int persistent_status(mg_engine* e) {
  if (e->last_run_grid && e->h_grid_ctrl && e->h_grid_ctrl[2] != 0) {
    e->grid_ok = false;                                                  // do not trust it again in this process
    return fail(MG_E_CUDA, "grid decode kernel aborted: a grid barrier was not reached within the watchdog time");
  }
  if (!e->last_run_flow || !e->h_flow_status || e->h_flow_status[0] == 0) return MG_OK;
  const int32_t* s = e->h_flow_status;
  e->flow_ok = false;                                                  // do not trust it again in this process
  if (std::getenv("MG_FLOW_WAITLOG")) {                                // debug: which hand-over every (SM, group) warp was stuck in
    for (int g = 0; g < flow::kMaxGroups; ++g) {
      std::map<std::pair<int, int>, std::vector<int>> by;
      for (int sm = 0; sm < e->n_sm; ++sm) {
        const int32_t* q = s + 8 + 2 * (sm * flow::kMaxGroups + g);
        if (q[1]) by[{q[1] - 1, q[0]}].push_back(sm);
      }
      for (auto& kv : by) {
        fprintf(stderr, "[flow waitlog] group %d step %d detail %d: %zu SMs:", g, kv.first.first, kv.first.second, kv.second.size());
        for (size_t i = 0; i < kv.second.size() && i < 12; ++i) fprintf(stderr, " %d", kv.second[i]);
        fprintf(stderr, "\n");
      }
    }
  }
  return fail(MG_E_CUDA, "flow decode kernel aborted: code " + std::to_string(s[0]) + " (1 = exchange word never arrived, 2 = K/V tile "
              "never arrived), SM " + std::to_string(s[1]) + ", group " + std::to_string(s[2]) + ", detail " + std::to_string(s[3]));
}

int parse_layer_name(const std::string& name, int* layer, std::string* leaf) {
  if (name.compare(0, 7, "layers.") != 0) return -1;
  const size_t dot = name.find('.', 7);
  if (dot == std::string::npos) return -1;
  *layer = std::atoi(name.substr(7, dot - 7).c_str());
  *leaf = name.substr(dot + 1);
  return 0;
}

int upload_tensor(mg_engine* e, void* dst, bool typed, const float* data, size_t n, float** keep_master = nullptr) {
  if (keep_master && typed && e->dtype == MG_DTYPE_BF16) {
    // the flow kernel folds LayerNorm into its packed weights: it packs from the fp32 values, not from the bf16 copy
    if (*keep_master) { e->dfree(*keep_master); *keep_master = nullptr; }
    MG_TRY(e->dmalloc(keep_master, n * sizeof(float)));
    MG_CUDA_OK(cudaMemcpyAsync(*keep_master, data, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    MG_TRY(launch_convert<bf16>(e->stream, *keep_master, reinterpret_cast<bf16*>(dst), n));
    MG_CUDA_OK(cudaStreamSynchronize(e->stream));
    e->h2d += n * sizeof(float);
    return MG_OK;
  }
  if (!typed || e->dtype == MG_DTYPE_FP32) {
    MG_CUDA_OK(cudaMemcpyAsync(dst, data, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    MG_CUDA_OK(cudaStreamSynchronize(e->stream));
    e->h2d += n * sizeof(float);
    return MG_OK;
  }
  if (n > e->stage_elems) {
    e->dfree(e->stage_f32);
    e->stage_f32 = nullptr;
    MG_TRY(e->dmalloc(&e->stage_f32, n * sizeof(float)));
    e->stage_elems = n;
  }
  MG_CUDA_OK(cudaMemcpyAsync(e->stage_f32, data, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  MG_TRY(launch_convert<bf16>(e->stream, e->stage_f32, reinterpret_cast<bf16*>(dst), n));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->h2d += n * sizeof(float);
  return MG_OK;
}

int ensure_arena(mg_engine* e, size_t ints) {
  if (ints <= e->arena_cap) return MG_OK;
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->dfree(e->d_arena);
  e->d_arena = nullptr;
  if (e->h_arena) cudaFreeHost(e->h_arena);
  e->h_arena = nullptr;
  ints = ints + ints / 2 + 256;
  MG_TRY(e->dmalloc(&e->d_arena, ints * sizeof(int32_t)));
  MG_CUDA_OK(cudaMallocHost(&e->h_arena, ints * sizeof(int32_t)));
  e->arena_cap = ints;
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  e->graph_B = -1;
  return MG_OK;
}

int ensure_out(mg_engine* e, size_t ints) {
  if (ints <= e->out_cap) return MG_OK;
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->dfree(e->d_out_block);
  e->d_out_block = nullptr;
  if (e->h_out_block) cudaFreeHost(e->h_out_block);
  e->h_out_block = nullptr;
  ints = ints + ints / 2 + 256;
  MG_TRY(e->dmalloc(&e->d_out_block, ints * sizeof(int32_t)));
  MG_CUDA_OK(cudaMallocHost(&e->h_out_block, ints * sizeof(int32_t)));
  e->out_cap = ints;
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  e->graph_B = -1;
  return MG_OK;
}

// Validates the prompts, builds the row maps on the host and uploads everything in ONE H2D copy.
// `kv` = true: KV-cache path (cache capacity checks); false: recompute mode (position-table check).
int upload_impl(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, int max_new, const int32_t* max_new_per,
                bool kv) {
  if (!e->ready) return fail(MG_E_STATE, "engine not finalized (mg_engine_finalize)");
  if (!ids || !offs || B <= 0) return fail(MG_E_ARG, "prompts: null pointer or empty batch");
  if (B > e->max_batch) return fail(MG_E_OOM, "batch larger than max_batch given at mg_engine_create");
  if (max_new < 0) return fail(MG_E_ARG, "max_new_tokens < 0");
  const mg_geometry& g = e->geo;
  if (offs[0] != 0) return fail(MG_E_ARG, "prompt_offsets[0] must be 0");
  int M = 0, max_tp = 0, max_total = 0, steps = 0;
  for (int b = 0; b < B; ++b) {
    const int tp = offs[b + 1] - offs[b];
    if (tp <= 0) return fail(MG_E_ARG, "empty prompt (the reference indexes generated[:, -1:], api_cache.py:167)");
    if (tp > g.pos_rows)
      return fail(MG_E_PROMPT_TOO_LONG, "prompt of " + std::to_string(tp) + " tokens exceeds the " +
                                            std::to_string(g.pos_rows) + "-row position table (api_cache.py:99)");
    const int mn = max_new_per ? max_new_per[b] : max_new;
    if (mn < 0) return fail(MG_E_ARG, "max_new_per_seq < 0");
    if (kv && tp + mn > e->max_seq) return fail(MG_E_OOM, "prompt + max_new_tokens exceeds max_seq given at mg_engine_create");
    if (!kv && mn > 0 && tp + mn - 1 > g.pos_rows)
      return fail(MG_E_PROMPT_TOO_LONG, "recompute mode: prompt + max_new_tokens - 1 exceeds the position table");
    M += tp;
    max_tp = std::max(max_tp, tp);
    max_total = std::max(max_total, tp + mn);
    steps = std::max(steps, mn);
  }
  for (int i = 0; i < M; ++i)
    if (ids[i] < 0 || ids[i] >= g.vocab_size) return fail(MG_E_TOKEN, "prompt token id outside [0, vocab)");

  const int stride = (max_total + 7) & ~7;
  // arena layout (int32): prompt[M] offsets[B+1] row_seq[M] row_pos[M] seq_start[B] seq_len[B] max_new[B]
  //                       last_rows[B] | device-only: cur_tok[B] lens[B] n_new[B] finished[B bytes]
  const size_t up_ints = static_cast<size_t>(M) * 3 + static_cast<size_t>(B) * 5 + 1;
  const size_t all_ints = up_ints + static_cast<size_t>(B) * 4 + 16;
  MG_TRY(ensure_arena(e, all_ints));
  MG_TRY(ensure_out(e, static_cast<size_t>(B) * (stride + 1)));
  int32_t* h = e->h_arena;
  MG_CUDA_OK(cudaEventSynchronize(e->ev_staged));           // an earlier asynchronous upload may still be reading the staging buffer
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += n; return r; };
  const size_t o_prompt = take(M), o_offs = take(B + 1), o_rseq = take(M), o_rpos = take(M), o_sstart = take(B),
               o_slen = take(B), o_maxnew = take(B), o_last = take(B);
  std::memcpy(h + o_prompt, ids, sizeof(int32_t) * M);
  std::memcpy(h + o_offs, offs, sizeof(int32_t) * (B + 1));
  for (int b = 0; b < B; ++b) {
    const int tp = offs[b + 1] - offs[b];
    for (int t = 0; t < tp; ++t) {
      h[o_rseq + offs[b] + t] = b;
      h[o_rpos + offs[b] + t] = t;
    }
    h[o_sstart + b] = offs[b];
    h[o_slen + b] = tp;
    h[o_maxnew + b] = max_new_per ? max_new_per[b] : max_new;
    h[o_last + b] = b;
  }
  MG_CUDA_OK(cudaMemcpyAsync(e->d_arena, h, up_ints * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaEventRecord(e->ev_staged, e->stream));
  e->h2d += up_ints * sizeof(int32_t);
  int32_t* da = e->d_arena;
  e->d_prompt = da + o_prompt; e->d_offsets = da + o_offs; e->d_row_seq = da + o_rseq; e->d_row_pos = da + o_rpos;
  e->d_seq_start = da + o_sstart; e->d_seq_len = da + o_slen; e->d_last_rows = da + o_last;
  int32_t* new_cur = da + take(B);
  int32_t* new_lens = da + take(B);
  int32_t* new_nnew = da + take(B);
  uint8_t* new_fin = reinterpret_cast<uint8_t*>(da + take(B));
  int32_t* new_outlen = e->d_out_block;
  int32_t* new_outids = e->d_out_block + B;
  const bool same = e->st.cur_tok == new_cur && e->st.lens == new_lens && e->st.n_new == new_nnew &&
                    e->st.finished == new_fin && e->st.out_len == new_outlen && e->st.out_ids == new_outids &&
                    e->st.max_new == da + o_maxnew && e->st.out_stride == stride && e->st.seq_idx == nullptr;
  if (!same && e->graph) {               // captured kernel arguments would be stale
    cudaGraphExecDestroy(e->graph);
    e->graph = nullptr;
    e->graph_B = -1;
  }
  e->st.cur_tok = new_cur; e->st.lens = new_lens; e->st.n_new = new_nnew; e->st.finished = new_fin;
  e->st.out_len = new_outlen; e->st.out_ids = new_outids; e->st.max_new = da + o_maxnew; e->st.out_stride = stride;
  e->st.seq_idx = nullptr;
  e->slots_active = false;                                  // a batch call ends a slot session (mg_slots_begin)
  if (!e->d_step_ns) MG_TRY(e->dmalloc(&e->d_step_ns, sizeof(unsigned long long) * (e->max_seq + 1)));
  e->st.step_ns = e->d_step_ns;
  e->cur_B = B; e->cur_M = M; e->cur_max_tp = max_tp; e->cur_steps = steps;
  MG_TRY(ensure_rows(e, std::max(M, B)));
  e->uploaded = true;
  return MG_OK;
}

int download_impl(mg_engine* e, int32_t* out_ids, int out_stride, int32_t* out_lens) {
  if (!e->uploaded) return fail(MG_E_STATE, "nothing to download (call mg_upload_prompts + mg_run first)");
  const int B = e->cur_B, stride = e->st.out_stride;
  const size_t ints = static_cast<size_t>(B) * (stride + 1);
  MG_CUDA_OK(cudaMemcpyAsync(e->h_out_block, e->d_out_block, ints * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  MG_TRY(persistent_status(e));
  e->d2h += ints * sizeof(int32_t);
  const int32_t* hl = e->h_out_block;
  const int32_t* hi = e->h_out_block + B;
  for (int b = 0; b < B; ++b) {
    const int n = hl[b];
    if (out_ids && n > out_stride) return fail(MG_E_ARG, "out_stride smaller than a generated sequence");
    if (out_lens) out_lens[b] = n;
    if (out_ids) std::memcpy(out_ids + static_cast<size_t>(b) * out_stride, hi + static_cast<size_t>(b) * stride, sizeof(int32_t) * n);
  }
  MG_CUDA_OK(cudaEventSynchronize(e->ev[2]));
  cudaEventElapsedTime(&e->t_total, e->ev[0], e->ev[2]);
  cudaEventElapsedTime(&e->t_prefill, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->t_decode, e->ev[1], e->ev[2]);
  return MG_OK;
}

int check_sampling(mg_engine* e, float temperature, int top_k) {
  if (!(temperature > 0.0f)) return fail(MG_E_ARG, "temperature must be > 0 (the reference divides by it, api_cache.py:169)");
  if (top_k < 0) return fail(MG_E_ARG, "top_k < 0");
  if (top_k > e->geo.vocab_size) return fail(MG_E_TOPK, "top_k larger than the vocabulary (torch.topk raises, api_cache.py:172)");
  return MG_OK;
}

// ---- recompute mode -------------------------------------------------------------------------------
int nocache_setup(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, int max_new, int* Tcap_out) {
  MG_TRY(upload_impl(e, ids, offs, B, max_new, nullptr, false));
  // rows: sequence b owns [b*Tcap, (b+1)*Tcap); Tcap = longest final length (multiple of 8)
  const int Tcap = e->st.out_stride;
  MG_TRY(ensure_rows(e, B * Tcap));
  // seq_start[b] = b * Tcap (overwrites the packed starts uploaded for the KV path)
  std::vector<int32_t> starts(B), last(B);
  for (int b = 0; b < B; ++b) starts[b] = b * Tcap;
  MG_CUDA_OK(cudaMemcpyAsync(e->d_seq_start, starts.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->h2d += sizeof(int32_t) * B;
  MG_TRY(launch_decode_init(e->stream, e->d_prompt, e->d_offsets, e->st, B));
  *Tcap_out = Tcap;
  return MG_OK;
}

// last_rows[b] = b*Tcap + out_len[b] - 1, computed on the device from out_len
__global__ void nocache_last_rows_kernel(const int32_t* out_len, int32_t* last_rows, int B, int Tcap) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) last_rows[b] = b * Tcap + out_len[b] - 1;
}

template <typename T>
int nocache_logits(mg_engine* e, int B, int Tcap, int max_len_now) {
  MG_TRY(forward_nocache<T>(e, B, Tcap, max_len_now));
  nocache_last_rows_kernel<<<ceil_div(B, 128), 128, 0, e->stream>>>(e->st.out_len, e->d_last_rows, B, Tcap);
  MG_LAUNCH_CHECK();
  if (!e->yl) {
    MG_TRY(e->dmalloc(&e->yl, e->esz * std::max(e->max_batch, 128) * e->geo.d_model));
    MG_CUDA_OK(cudaMemsetAsync(e->yl, 0, e->esz * std::max(e->max_batch, 128) * e->geo.d_model, e->stream));
    if (e->use_tc) MG_TRY(make_tmap_bf16_2d(&e->tm_yl, e->yl, std::max(e->max_batch, 128), e->geo.d_model, kGemmBM));
  }
  MG_TRY((launch_gather_rows<T, T>(e->stream, reinterpret_cast<const T*>(e->y), e->d_last_rows, reinterpret_cast<T*>(e->yl), B,
                                   e->geo.d_model)));
  GemmEpilogue epi; epi.bias = e->head_b; epi.ld_out = e->ld_logits;
  return gemm<T>(e, e->yl, &e->tm_yl, e->head_w, &e->m_head, B, e->geo.vocab_size, e->geo.d_model, epi, e->logits, nullptr);
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int mg_abi_version(void) { return MG_ABI_VERSION; }
const char* mg_last_error(void) { return t_last_error.c_str(); }

int mg_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(MG_E_CUDA, cudaGetErrorString(e));
  return n;
}

int mg_engine_create(const mg_geometry* geo, int device, int dtype_mode, int max_batch, int max_seq, mg_engine** out) {
  if (!geo || !out) return fail(MG_E_ARG, "null argument");
  *out = nullptr;
  if (dtype_mode != MG_DTYPE_FP32 && dtype_mode != MG_DTYPE_BF16) return fail(MG_E_ARG, "unknown dtype_mode");
  const mg_geometry g = *geo;
  if (g.vocab_size <= 0 || g.pos_rows <= 0 || g.d_model <= 0 || g.n_head <= 0 || g.n_layer <= 0 || g.d_ff <= 0)
    return fail(MG_E_SHAPE, "non-positive geometry field");
  if (g.d_model % g.n_head) return fail(MG_E_SHAPE, "d_model not divisible by n_head");
  if (g.d_model % 64 || g.d_ff % 16 || g.d_model > 1024)
    return fail(MG_E_SHAPE, "d_model must be a multiple of 64 (KV-cache slices) and <= 1024, d_ff a multiple of 16");
  const int hd = g.d_model / g.n_head;
  if (hd % 8 || hd > 64 || (hd & (hd - 1))) return fail(MG_E_SHAPE, "head_dim must be 8, 16, 32 or 64");
  if (max_batch <= 0 || max_seq <= 0) return fail(MG_E_ARG, "max_batch / max_seq must be positive");
  MG_TRY(check_device(device));

  mg_engine* e = new mg_engine();
  e->geo = g; e->device = device; e->dtype = dtype_mode; e->max_batch = max_batch; e->max_seq = max_seq;
  e->esz = dtype_mode == MG_DTYPE_BF16 ? 2 : 4;
  const char* env_gemm = std::getenv("MG_GEMM");
  e->use_tc = dtype_mode == MG_DTYPE_BF16 && !(env_gemm && std::strcmp(env_gemm, "simt") == 0);
  const char* env_graph = std::getenv("MG_NO_GRAPH");
  e->use_graph = !(env_graph && env_graph[0] == '1');
  const char* env_mega = std::getenv("MG_NO_MEGA");
  e->use_mega = !(env_mega && env_mega[0] == '1');
  // The weight-stationary flow kernel (decode_flow.cu) is OPT-IN (MG_FLOW=1): parity-green, but measured slower than the cluster
  // kernel on every BASELINE configuration (profiles/r2b_flow_*).
  const char* env_grid = std::getenv("MG_GRID");
  e->grid_mode = !env_grid ? 2 : (env_grid[0] == '1' ? 1 : 0);
  e->use_grid = e->grid_mode == 1;
  const char* env_flow = std::getenv("MG_FLOW");
  e->use_flow = env_flow && env_flow[0] == '1';
  {
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->n_sm = prop.multiProcessorCount;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    e->flow_candidate = e->use_flow && coop && dtype_mode == MG_DTYPE_BF16 &&
                        flow::flow_eligible(g.d_model, g.d_ff, g.n_head, g.n_layer, g.vocab_size, e->n_sm);
  }
  e->launches0 = g_kernel_launches.load();
  int rc = MG_OK;
  auto body = [&]() -> int {
    MG_CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    for (auto& ev : e->ev) MG_CUDA_OK(cudaEventCreate(&ev));
    MG_CUDA_OK(cudaEventCreateWithFlags(&e->ev_staged, cudaEventDisableTiming));
    MG_TRY(kernels_init());
    if (e->use_tc) MG_TRY(gemm_tc_init());
    const size_t d = g.d_model, f = g.d_ff, V = g.vocab_size;
    MG_TRY(e->dmalloc(&e->tok_emb, e->esz * V * d));
    MG_TRY(e->dmalloc(&e->pos_emb, e->esz * g.pos_rows * d));
    MG_TRY(e->dmalloc(&e->head_w, e->esz * V * d));
    MG_TRY(e->dmalloc(&e->head_b, sizeof(float) * V));
    e->layers.resize(g.n_layer);
    const size_t kv_elems = static_cast<size_t>(max_batch) * max_seq * d;
    for (auto& w : e->layers) {
      MG_TRY(e->dmalloc(&w.w_in, e->esz * 3 * d * d));
      MG_TRY(e->dmalloc(&w.w_out, e->esz * d * d));
      MG_TRY(e->dmalloc(&w.w1, e->esz * f * d));
      MG_TRY(e->dmalloc(&w.w2, e->esz * d * f));
      MG_TRY(e->dmalloc(&w.b_in, sizeof(float) * 3 * d));
      MG_TRY(e->dmalloc(&w.b_out, sizeof(float) * d));
      MG_TRY(e->dmalloc(&w.b1, sizeof(float) * f));
      MG_TRY(e->dmalloc(&w.b2, sizeof(float) * d));
      MG_TRY(e->dmalloc(&w.ln1w, sizeof(float) * d));
      MG_TRY(e->dmalloc(&w.ln1b, sizeof(float) * d));
      MG_TRY(e->dmalloc(&w.ln2w, sizeof(float) * d));
      MG_TRY(e->dmalloc(&w.ln2b, sizeof(float) * d));
      MG_TRY(e->dmalloc(&w.kc, e->esz * kv_elems));
      MG_TRY(e->dmalloc(&w.vc, e->esz * kv_elems));
    }
    e->ld_logits = (g.vocab_size + 7) & ~7;
    MG_TRY(e->dmalloc(&e->logits, sizeof(float) * static_cast<size_t>(max_batch) * e->ld_logits));
    const size_t ws_rows = static_cast<size_t>(2 * 148 + 2 * max_batch);
    MG_TRY(e->dmalloc(&e->ws_o, sizeof(float) * ws_rows * d));
    MG_TRY(e->dmalloc(&e->ws_ml, sizeof(float) * ws_rows * g.n_head * 2));
    MG_TRY(e->dmalloc(&e->counters, sizeof(uint32_t) * max_batch));
    MG_CUDA_OK(cudaMemsetAsync(e->counters, 0, sizeof(uint32_t) * max_batch, e->stream));
    MG_TRY(e->dmalloc(&e->d_sp, sizeof(SampleParams)));
    MG_TRY(e->dmalloc(&e->d_active, sizeof(int32_t)));
    MG_CUDA_OK(cudaMallocHost(&e->h_sp, sizeof(SampleParams)));
    MG_CUDA_OK(cudaMallocHost(&e->h_active, sizeof(int32_t)));
    MG_TRY(ensure_rows(e, std::max(max_batch, 128)));
    MG_CUDA_OK(cudaStreamSynchronize(e->stream));
    return MG_OK;
  };
  rc = body();
  if (rc != MG_OK) {
    const std::string keep = t_last_error;
    mg_engine_destroy(e);
    set_last_error(keep);
    return rc;
  }
  *out = e;
  return MG_OK;
}

void mg_engine_destroy(mg_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->graph) cudaGraphExecDestroy(e->graph);
  for (void* p : e->allocs) cudaFree(p);
  if (e->h_arena) cudaFreeHost(e->h_arena);
  if (e->h_out_block) cudaFreeHost(e->h_out_block);
  if (e->h_sp) cudaFreeHost(e->h_sp);
  if (e->h_active) cudaFreeHost(e->h_active);
  if (e->h_flow_status) cudaFreeHost(e->h_flow_status);
  if (e->h_grid_ctrl) cudaFreeHost(e->h_grid_ctrl);
  if (e->h_slot_flags) cudaFreeHost(e->h_slot_flags);
  if (e->h_slot_rows) cudaFreeHost(e->h_slot_rows);
  if (e->h_detok) cudaFreeHost(e->h_detok);
  delete e->flow_plan;
  for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
  if (e->ev_staged) cudaEventDestroy(e->ev_staged);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int mg_load_weight(mg_engine* e, const char* name_c, const float* data, const int64_t* shape, int ndim) {
  if (!e || !name_c || !data || !shape) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  const mg_geometry& g = e->geo;
  const int64_t d = g.d_model, f = g.d_ff, V = g.vocab_size;
  const std::string name(name_c);
  void* dst = nullptr;
  bool typed = true;
  int64_t s0 = 0, s1 = 0;                 // expected shape (s1 == 0: 1-D)
  int layer = -1;
  std::string leaf;
  if (name == "tok_emb.weight") { dst = e->tok_emb; s0 = V; s1 = d; }
  else if (name == "pos_emb") { dst = e->pos_emb; s0 = g.pos_rows; s1 = d; }
  else if (name == "head.weight") { dst = e->head_w; s0 = V; s1 = d; }
  else if (name == "head.bias") { dst = e->head_b; s0 = V; typed = false; }
  else if (parse_layer_name(name, &layer, &leaf) == 0 && layer >= 0 && layer < g.n_layer) {
    LayerW& w = e->layers[layer];
    if (leaf == "attn.in_proj_weight") { dst = w.w_in; s0 = 3 * d; s1 = d; }
    else if (leaf == "attn.in_proj_bias") { dst = w.b_in; s0 = 3 * d; typed = false; }
    else if (leaf == "attn.out_proj.weight") { dst = w.w_out; s0 = d; s1 = d; }
    else if (leaf == "attn.out_proj.bias") { dst = w.b_out; s0 = d; typed = false; }
    else if (leaf == "mlp.0.weight") { dst = w.w1; s0 = f; s1 = d; }
    else if (leaf == "mlp.0.bias") { dst = w.b1; s0 = f; typed = false; }
    else if (leaf == "mlp.2.weight") { dst = w.w2; s0 = d; s1 = f; }
    else if (leaf == "mlp.2.bias") { dst = w.b2; s0 = d; typed = false; }
    else if (leaf == "ln1.weight") { dst = w.ln1w; s0 = d; typed = false; }
    else if (leaf == "ln1.bias") { dst = w.ln1b; s0 = d; typed = false; }
    else if (leaf == "ln2.weight") { dst = w.ln2w; s0 = d; typed = false; }
    else if (leaf == "ln2.bias") { dst = w.ln2b; s0 = d; typed = false; }
  }
  if (!dst) return fail(MG_E_SHAPE, "unknown tensor name '" + name + "' for this geometry");
  const bool shape_ok = (s1 == 0) ? (ndim == 1 && shape[0] == s0) : (ndim == 2 && shape[0] == s0 && shape[1] == s1);
  if (!shape_ok) return fail(MG_E_SHAPE, "shape mismatch for '" + name + "'");
  const bool is_matrix = typed && s1 != 0 && name != "tok_emb.weight" && name != "pos_emb";
  MG_TRY(upload_tensor(e, dst, typed, data, static_cast<size_t>(s0) * (s1 ? s1 : 1),
                       (e->flow_candidate && is_matrix) ? &e->masters[name] : nullptr));
  e->loaded.insert(name);
  e->ready = false;
  return MG_OK;
}

int mg_engine_finalize(mg_engine* e) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  const size_t want = 4 + 12 * static_cast<size_t>(e->geo.n_layer);
  if (e->loaded.size() != want)
    return fail(MG_E_STATE, "weights missing: " + std::to_string(e->loaded.size()) + " of " + std::to_string(want) + " tensors loaded");
  if (e->use_tc) {
    const int d = e->geo.d_model, f = e->geo.d_ff;
    for (auto& w : e->layers) {
      MG_TRY(make_wmaps(&w.m_in, w.w_in, 3 * d, d));
      MG_TRY(make_wmaps(&w.m_out, w.w_out, d, d));
      MG_TRY(make_wmaps(&w.m_w1, w.w1, f, d));
      MG_TRY(make_wmaps(&w.m_w2, w.w2, d, f));
    }
    MG_TRY(make_wmaps(&e->m_head, e->head_w, e->geo.vocab_size, d));
  }
  MG_TRY(setup_mega(e));
  MG_TRY(setup_grid(e));
  MG_TRY(setup_flow(e));
  e->ready = true;
  return MG_OK;
}

int mg_upload_prompts(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, int max_new, const int32_t* max_new_per) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  return upload_impl(e, ids, offs, B, max_new, max_new_per, true);
}

static int run_locked(mg_engine* e, float temperature, int top_k, int eos_id, uint64_t seed, uint64_t seq_base) {
  if (!e->uploaded) return fail(MG_E_STATE, "mg_run before mg_upload_prompts");
  MG_TRY(check_sampling(e, temperature, top_k));
  return e->dtype == MG_DTYPE_BF16 ? run_impl<bf16>(e, temperature, top_k, eos_id, seed, seq_base)
                                   : run_impl<float>(e, temperature, top_k, eos_id, seed, seq_base);
}

int mg_run(mg_engine* e, float temperature, int top_k, int eos_id, uint64_t seed, uint64_t seq_base) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  return run_locked(e, temperature, top_k, eos_id, seed, seq_base);
}

int mg_download(mg_engine* e, int32_t* out_ids, int out_stride, int32_t* out_lens) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  return download_impl(e, out_ids, out_stride, out_lens);
}

int mg_synchronize(mg_engine* e) {
  if (!e) return fail(MG_E_ARG, "null engine");
  MG_CUDA_OK(cudaSetDevice(e->device));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  return persistent_status(e);
}

int mg_last_decode_path(mg_engine* e) {
  if (!e) return fail(MG_E_ARG, "null engine");
  return e->last_run_grid ? 3 : (e->last_run_flow ? 2 : (e->last_run_mega ? 1 : 0));
}

void* mg_engine_stream(mg_engine* e) { return e ? reinterpret_cast<void*>(e->stream) : nullptr; }

int mg_generate(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, int max_new, const int32_t* max_new_per,
                float temperature, int top_k, int eos_id, uint64_t seed, uint64_t seq_base, int32_t* out_ids,
                int out_stride, int32_t* out_lens) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  MG_TRY(check_sampling(e, temperature, top_k));
  MG_TRY(upload_impl(e, ids, offs, B, max_new, max_new_per, true));
  MG_TRY(run_locked(e, temperature, top_k, eos_id, seed, seq_base));
  return download_impl(e, out_ids, out_stride, out_lens);
}

// Shared body of mg_step_logits / mg_step_logits_at.  want == nullptr: every step is kept (slot = step).
static int step_logits_impl(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, const int32_t* forced, int n_steps,
                            const int32_t* want, int n_want, float* logits_out) {
  if (!e || !logits_out) return fail(MG_E_ARG, "null argument");
  if (n_steps <= 0) return fail(MG_E_ARG, "n_steps must be positive");
  std::vector<int32_t> slot;                       // step -> row block of logits_out, -1 = not kept
  int n_keep = n_steps;
  if (want) {
    if (n_want <= 0) return fail(MG_E_ARG, "n_want must be positive");
    slot.assign(n_steps, -1);
    for (int i = 0; i < n_want; ++i) {
      if (want[i] < 0 || want[i] >= n_steps || slot[want[i]] >= 0) return fail(MG_E_ARG, "want_steps must be distinct steps in [0, n_steps)");
      slot[want[i]] = i;
    }
    n_keep = n_want;
  }
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  MG_TRY(upload_impl(e, ids, offs, B, n_steps, nullptr, true));
  const int V = e->geo.vocab_size;
  if (n_steps > 1) {
    if (!forced) return fail(MG_E_ARG, "forced_ids is null");
    const size_t n = static_cast<size_t>(B) * n_steps;
    for (size_t i = 0; i < n; ++i)
      if (forced[i] < 0 || forced[i] >= V) return fail(MG_E_TOKEN, "forced token id outside [0, vocab)");
    if (n > e->forced_cap) {
      e->dfree(e->d_forced);
      e->d_forced = nullptr;
      MG_TRY(e->dmalloc(&e->d_forced, n * sizeof(int32_t)));
      e->forced_cap = n;
    }
    MG_CUDA_OK(cudaMemcpyAsync(e->d_forced, forced, n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    e->h2d += n * sizeof(int32_t);
  }
  const bool is_bf16 = e->dtype == MG_DTYPE_BF16;
  MG_TRY(is_bf16 ? prefill<bf16>(e) : prefill<float>(e));
  if (is_bf16 && (e->mega_ok || e->flow_ok || e->grid_ok)) {
    // the persistent kernels in teacher-forcing mode: same code path as mg_run, logits of the kept steps dumped
    float* d_lg = nullptr;
    int32_t* d_slot = nullptr;
    const size_t n = static_cast<size_t>(n_keep) * B * V;
    MG_TRY(e->dmalloc(&d_lg, n * sizeof(float)));
    int rc = MG_OK;
    if (want) {
      rc = e->dmalloc(&d_slot, sizeof(int32_t) * n_steps);
      if (rc == MG_OK && cudaMemcpyAsync(d_slot, slot.data(), sizeof(int32_t) * n_steps, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) rc = MG_E_CUDA;
      if (rc == MG_OK && cudaStreamSynchronize(e->stream) != cudaSuccess) rc = MG_E_CUDA;   // `slot` is pageable host memory
    }
    *e->h_sp = SampleParams{1.0f, 1, -1, 0, 0, 0};
    if (rc == MG_OK && cudaMemcpyAsync(e->d_sp, e->h_sp, sizeof(SampleParams), cudaMemcpyHostToDevice, e->stream) != cudaSuccess) rc = MG_E_CUDA;
    int mrc = MG_OK;
    e->last_run_mega = false;
    if (rc == MG_OK && run_decode_persistent(e, 1, -1, &mrc, d_lg, n_steps > 1 ? e->d_forced : nullptr, n_steps, d_slot)) {
      e->last_run_mega = true;
      rc = mrc;
      if (rc == MG_OK && cudaMemcpyAsync(logits_out, d_lg, n * sizeof(float), cudaMemcpyDeviceToHost, e->stream) != cudaSuccess) rc = MG_E_CUDA;
      if (rc == MG_OK && cudaStreamSynchronize(e->stream) != cudaSuccess) rc = fail(MG_E_CUDA, std::string("persistent decode kernel: ") + cudaGetErrorString(cudaGetLastError()));
      if (rc == MG_OK) rc = persistent_status(e);
      e->dfree(d_lg);
      e->dfree(d_slot);
      e->d2h += n * sizeof(float);
      e->uploaded = false;
      return rc;
    }
    e->dfree(d_lg);
    e->dfree(d_slot);
    MG_TRY(rc);
    MG_TRY(mrc);
  }
  e->last_run_mega = e->last_run_flow = e->last_run_grid = false;
  for (int i = 0; i < n_steps; ++i) {
    MG_TRY(is_bf16 ? decode_forward<bf16>(e) : decode_forward<float>(e));
    const int sl = want ? slot[i] : i;
    if (sl >= 0) {
      MG_CUDA_OK(cudaMemcpy2DAsync(logits_out + static_cast<size_t>(sl) * B * V, sizeof(float) * V, e->logits,
                                   sizeof(float) * e->ld_logits, sizeof(float) * V, B, cudaMemcpyDeviceToHost, e->stream));
      e->d2h += sizeof(float) * V * B;
    }
    if (i + 1 < n_steps) MG_TRY(launch_force_next(e->stream, e->d_forced, n_steps, i, e->st, B));
  }
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->uploaded = false;
  return MG_OK;
}

int mg_step_logits(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, const int32_t* forced, int n_steps,
                   float* logits_out) {
  return step_logits_impl(e, ids, offs, B, forced, n_steps, nullptr, 0, logits_out);
}

int mg_step_logits_at(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, const int32_t* forced, int n_steps,
                      const int32_t* want_steps, int n_want, float* logits_out) {
  if (!want_steps) return fail(MG_E_ARG, "want_steps is null");
  return step_logits_impl(e, ids, offs, B, forced, n_steps, want_steps, n_want, logits_out);
}

int mg_generate_nocache(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, int max_new, float temperature,
                        int top_k, int eos_id, uint64_t seed, uint64_t seq_base, int32_t* out_ids, int out_stride,
                        int32_t* out_lens) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  MG_TRY(check_sampling(e, temperature, top_k));
  int Tcap = 0;
  MG_TRY(nocache_setup(e, ids, offs, B, max_new, &Tcap));
  *e->h_sp = SampleParams{temperature, top_k, eos_id, 0, seed, seq_base};
  MG_CUDA_OK(cudaMemcpyAsync(e->d_sp, e->h_sp, sizeof(SampleParams), cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaEventRecord(e->ev[0], e->stream));
  MG_CUDA_OK(cudaEventRecord(e->ev[1], e->stream));
  const bool is_bf16 = e->dtype == MG_DTYPE_BF16;
  int done = 0;
  for (int i = 0; i < max_new; ++i) {
    const int max_len_now = std::min(Tcap, e->cur_max_tp + i);
    MG_TRY(is_bf16 ? nocache_logits<bf16>(e, B, Tcap, max_len_now) : nocache_logits<float>(e, B, Tcap, max_len_now));
    MG_TRY(launch_sample_step(e->stream, e->logits, e->ld_logits, e->geo.vocab_size, e->d_sp, e->st, B));
    ++done;
    if (eos_id >= 0 && (i % 16) == 15 && i + 1 < max_new) {
      MG_TRY(launch_count_active(e->stream, e->st.finished, B, e->d_active));
      MG_CUDA_OK(cudaMemcpyAsync(e->h_active, e->d_active, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      MG_CUDA_OK(cudaStreamSynchronize(e->stream));
      if (*e->h_active == 0) break;
    }
  }
  e->t_steps = done;
  MG_CUDA_OK(cudaEventRecord(e->ev[2], e->stream));
  const int rc = download_impl(e, out_ids, out_stride, out_lens);
  e->uploaded = false;
  return rc;
}

int mg_forward_nocache(mg_engine* e, const int32_t* ids, const int32_t* offs, int B, float* logits_out) {
  if (!e || !logits_out) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  int Tcap = 0;
  MG_TRY(nocache_setup(e, ids, offs, B, 0, &Tcap));
  const bool is_bf16 = e->dtype == MG_DTYPE_BF16;
  // logits of the LAST position of every sequence: [B][vocab]
  MG_TRY(is_bf16 ? nocache_logits<bf16>(e, B, Tcap, e->cur_max_tp) : nocache_logits<float>(e, B, Tcap, e->cur_max_tp));
  const int V = e->geo.vocab_size;
  MG_CUDA_OK(cudaMemcpy2DAsync(logits_out, sizeof(float) * V, e->logits, sizeof(float) * e->ld_logits, sizeof(float) * V, B,
                               cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->d2h += sizeof(float) * V * B;
  e->uploaded = false;
  return MG_OK;
}

int mg_sample_logits(mg_engine* e, const float* logits, int rows, int vocab, float temperature, int top_k, uint64_t seed,
                     uint64_t seq_base, uint32_t step, int32_t* out) {
  if (!e || !logits || !out) return fail(MG_E_ARG, "null argument");
  if (rows <= 0 || vocab <= 0) return fail(MG_E_ARG, "rows / vocab must be positive");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!(temperature > 0.0f)) return fail(MG_E_ARG, "temperature must be > 0");
  if (top_k < 0) return fail(MG_E_ARG, "top_k < 0");
  if (top_k > vocab) return fail(MG_E_TOPK, "top_k larger than the vocabulary");
  float* d_logits = nullptr;
  int32_t* d_out = nullptr;
  const size_t n = static_cast<size_t>(rows) * vocab;
  MG_TRY(e->dmalloc(&d_logits, n * sizeof(float)));
  MG_TRY(e->dmalloc(&d_out, rows * sizeof(int32_t)));
  int rc = MG_OK;
  auto body = [&]() -> int {
    MG_CUDA_OK(cudaMemcpyAsync(d_logits, logits, n * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    *e->h_sp = SampleParams{temperature, top_k, -1, 0, seed, seq_base};
    MG_CUDA_OK(cudaMemcpyAsync(e->d_sp, e->h_sp, sizeof(SampleParams), cudaMemcpyHostToDevice, e->stream));
    MG_TRY(launch_sample_rows(e->stream, d_logits, vocab, rows, vocab, e->d_sp, step, d_out));
    MG_CUDA_OK(cudaMemcpyAsync(out, d_out, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    MG_CUDA_OK(cudaStreamSynchronize(e->stream));
    return MG_OK;
  };
  rc = body();
  e->h2d += n * sizeof(float);
  e->d2h += rows * sizeof(int32_t);
  e->dfree(d_logits);
  e->dfree(d_out);
  return rc;
}

int mg_engine_stats(mg_engine* e, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!e) return fail(MG_E_ARG, "null engine");
  if (kernel_launches) *kernel_launches = g_kernel_launches.load() - e->launches0;
  if (h2d_bytes) *h2d_bytes = e->h2d;
  if (d2h_bytes) *d2h_bytes = e->d2h;
  return MG_OK;
}

int mg_last_run_timing(mg_engine* e, float* total_ms, float* prefill_ms, float* decode_ms, int* steps) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->ev[2]) {
    if (cudaEventQuery(e->ev[2]) == cudaSuccess) {
      cudaEventElapsedTime(&e->t_total, e->ev[0], e->ev[2]);
      cudaEventElapsedTime(&e->t_prefill, e->ev[0], e->ev[1]);
      cudaEventElapsedTime(&e->t_decode, e->ev[1], e->ev[2]);
    }
  }
  if (total_ms) *total_ms = e->t_total;
  if (prefill_ms) *prefill_ms = e->t_prefill;
  if (decode_ms) *decode_ms = e->t_decode;
  if (steps) *steps = e->t_steps;
  return MG_OK;
}

// ---- slot sessions: continuous batching (the caller side is reference api_cache.py:186-204) ---------------------
// The service is one request per HTTP call; a slot session keeps n_slots sequences in flight, admits new requests into free
// slots between chunks of decode steps and retires finished ones, so that the batch-64 throughput of the decode kernels
// serves a stream of independent requests.  Decode state and token rows live outside the per-call arena; the K/V rows of a
// newly admitted prompt are prefilled into the caches of its slot, everything already in flight is untouched.
int mg_slots_begin(mg_engine* e, int n_slots, int max_len, float temperature, int top_k, int eos_id, uint64_t seed) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!e->ready) return fail(MG_E_STATE, "engine not finalized (mg_engine_finalize)");
  if (n_slots <= 0 || n_slots > e->max_batch) return fail(MG_E_OOM, "n_slots must be in [1, max_batch]");
  if (max_len <= 1 || max_len > e->max_seq) return fail(MG_E_OOM, "max_len must be in [2, max_seq]");
  MG_TRY(check_sampling(e, temperature, top_k));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  const int stride = (max_len + 7) & ~7;
  const size_t state_ints = static_cast<size_t>(n_slots) * 7, out_ints = static_cast<size_t>(n_slots) * (stride + 1);
  if (state_ints > e->slot_state_cap) {
    e->dfree(e->d_slot_state); e->d_slot_state = nullptr;
    if (e->h_slot_flags) { cudaFreeHost(e->h_slot_flags); e->h_slot_flags = nullptr; }
    MG_TRY(e->dmalloc(&e->d_slot_state, state_ints * sizeof(int32_t)));
    MG_CUDA_OK(cudaMallocHost(&e->h_slot_flags, 2 * static_cast<size_t>(n_slots) * sizeof(int32_t)));
    e->slot_state_cap = state_ints;
  }
  if (out_ints > e->slot_out_cap) {
    e->dfree(e->d_slot_out); e->d_slot_out = nullptr;
    MG_TRY(e->dmalloc(&e->d_slot_out, out_ints * sizeof(int32_t)));
    e->slot_out_cap = out_ints;
  }
  if (out_ints > e->slot_rows_cap) {
    if (e->h_slot_rows) { cudaFreeHost(e->h_slot_rows); e->h_slot_rows = nullptr; }
    MG_CUDA_OK(cudaMallocHost(&e->h_slot_rows, out_ints * sizeof(int32_t)));
    e->slot_rows_cap = out_ints;
  }
  int32_t* d = e->d_slot_state;
  MG_CUDA_OK(cudaMemsetAsync(d, 0, state_ints * sizeof(int32_t), e->stream));
  MG_CUDA_OK(cudaMemsetAsync(d + 5 * n_slots, 1, n_slots, e->stream));                 // every slot starts idle (finished)
  MG_CUDA_OK(cudaMemsetAsync(e->d_slot_out, 0, out_ints * sizeof(int32_t), e->stream));
  {
    std::vector<int32_t> ident(n_slots);                              // row of the residual stream the head reads for slot b
    for (int b = 0; b < n_slots; ++b) ident[b] = b;
    MG_CUDA_OK(cudaMemcpyAsync(d + 6 * n_slots, ident.data(), sizeof(int32_t) * n_slots, cudaMemcpyHostToDevice, e->stream));
    MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  }
  e->d_slot_last_rows = d + 6 * n_slots;
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }                // captured state pointers change
  e->graph_B = -1;
  e->st.cur_tok = d; e->st.lens = d + n_slots; e->st.n_new = d + 2 * n_slots; e->st.max_new = d + 3 * n_slots;
  e->st.seq_idx = d + 4 * n_slots; e->st.finished = reinterpret_cast<uint8_t*>(d + 5 * n_slots);
  e->st.out_len = e->d_slot_out; e->st.out_ids = e->d_slot_out + n_slots; e->st.out_stride = stride;
  if (!e->d_step_ns) MG_TRY(e->dmalloc(&e->d_step_ns, sizeof(unsigned long long) * (e->max_seq + 1)));
  e->st.step_ns = nullptr;                                                              // per-token stamps are a batch-call read-out
  e->cur_B = n_slots; e->cur_M = 0; e->cur_max_tp = 0; e->cur_steps = 0;
  e->n_slots = n_slots; e->slots_eos = eos_id; e->slots_topk = top_k;
  e->slot_busy.assign(n_slots, 0);
  *e->h_sp = SampleParams{temperature, top_k, eos_id, 0, seed, 0};
  MG_CUDA_OK(cudaMemcpyAsync(e->d_sp, e->h_sp, sizeof(SampleParams), cudaMemcpyHostToDevice, e->stream));
  MG_TRY(ensure_rows(e, n_slots));
  // the decode path is fixed for the whole session: the persistent kernel appends K/V rows in its own layout only
  e->slots_mega = false;
  if (e->dtype == MG_DTYPE_BF16 && e->mega_ok && top_k >= 1 && top_k <= mega::kMegaMaxTopK)
    for (int s_try = 1; s_try <= mega::kMegaMaxSeqPerCluster && !e->slots_mega; ++s_try) {
      const int avail = s_try <= 2 ? e->mega_clusters2 : e->mega_clusters4;
      e->slots_mega = avail > 0 && ceil_div(n_slots, s_try) <= avail;
    }
  // ... or the grid-synchronous kernel where the cluster kernel does not take the geometry (train_large2) / MG_GRID=1
  e->slots_grid = (!e->slots_mega || e->grid_mode == 1) && e->dtype == MG_DTYPE_BF16 && e->grid_ok && e->use_grid && n_slots <= grid::kMaxSeqs;
  if (e->slots_grid) e->slots_mega = false;
  e->uploaded = false;                                                                  // batch-call read-outs do not apply
  e->slots_active = true;
  return MG_OK;
}

int mg_slots_admit(mg_engine* e, int n, const int32_t* slots, const int32_t* ids, const int32_t* offs, const int32_t* max_new,
                   const int32_t* seq_index) {
  if (!e || !slots || !ids || !offs || !max_new || !seq_index) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!e->slots_active) return fail(MG_E_STATE, "no slot session (mg_slots_begin)");
  if (n <= 0) return MG_OK;
  const mg_geometry& g = e->geo;
  if (offs[0] != 0) return fail(MG_E_ARG, "prompt_offsets[0] must be 0");
  int M = 0, max_tp = 0;
  std::vector<uint8_t> seen(e->n_slots, 0);
  for (int j = 0; j < n; ++j) {
    const int b = slots[j], tp = offs[j + 1] - offs[j];
    if (b < 0 || b >= e->n_slots) return fail(MG_E_ARG, "slot index outside the session");
    if (e->slot_busy[b] || seen[b]) return fail(MG_E_STATE, "slot " + std::to_string(b) + " is still in flight");
    seen[b] = 1;
    if (tp <= 0) return fail(MG_E_ARG, "empty prompt (the reference indexes generated[:, -1:], api_cache.py:167)");
    if (tp > g.pos_rows)
      return fail(MG_E_PROMPT_TOO_LONG, "prompt of " + std::to_string(tp) + " tokens exceeds the " + std::to_string(g.pos_rows) +
                                            "-row position table (api_cache.py:99)");
    if (max_new[j] < 0) return fail(MG_E_ARG, "max_new < 0");
    if (tp + max_new[j] > e->st.out_stride || tp + max_new[j] > e->max_seq)
      return fail(MG_E_OOM, "prompt + max_new exceeds the max_len of the slot session");
    M += tp;
    max_tp = std::max(max_tp, tp);
  }
  for (int i = 0; i < M; ++i)
    if (ids[i] < 0 || ids[i] >= g.vocab_size) return fail(MG_E_TOKEN, "prompt token id outside [0, vocab)");
  // arena (int32): prompt[M] offsets[n+1] row_seq[M] row_pos[M] seq_start[n] seq_len[n] slots[n] max_new[n] seq_index[n]
  const size_t ints = static_cast<size_t>(M) * 3 + static_cast<size_t>(n) * 6 + 1;
  MG_TRY(ensure_arena(e, ints));
  int32_t* h = e->h_arena;
  size_t o = 0;
  auto take = [&](size_t k) { size_t r = o; o += k; return r; };
  const size_t o_prompt = take(M), o_offs = take(n + 1), o_rseq = take(M), o_rpos = take(M), o_sstart = take(n), o_slen = take(n),
               o_slots = take(n), o_maxnew = take(n), o_sidx = take(n);
  MG_CUDA_OK(cudaEventSynchronize(e->ev_staged));                   // the staging buffer may still feed the previous admission
  std::memcpy(h + o_prompt, ids, sizeof(int32_t) * M);
  std::memcpy(h + o_offs, offs, sizeof(int32_t) * (n + 1));
  for (int j = 0; j < n; ++j) {
    const int tp = offs[j + 1] - offs[j];
    for (int t = 0; t < tp; ++t) { h[o_rseq + offs[j] + t] = slots[j]; h[o_rpos + offs[j] + t] = t; }
    h[o_sstart + j] = offs[j]; h[o_slen + j] = tp; h[o_slots + j] = slots[j]; h[o_maxnew + j] = max_new[j]; h[o_sidx + j] = seq_index[j];
  }
  MG_CUDA_OK(cudaMemcpyAsync(e->d_arena, h, ints * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaEventRecord(e->ev_staged, e->stream));
  e->h2d += ints * sizeof(int32_t);
  int32_t* da = e->d_arena;
  e->d_prompt = da + o_prompt; e->d_offsets = da + o_offs; e->d_row_seq = da + o_rseq; e->d_row_pos = da + o_rpos;
  e->d_seq_start = da + o_sstart; e->d_seq_len = da + o_slen;
  e->d_last_rows = e->d_slot_last_rows;
  e->cur_M = M; e->cur_max_tp = max_tp;
  MG_TRY(ensure_rows(e, std::max(M, e->n_slots)));
  MG_TRY(e->dtype == MG_DTYPE_BF16 ? prefill_rows<bf16>(e, M, n) : prefill_rows<float>(e, M, n));
  MG_TRY(launch_slot_init(e->stream, e->d_prompt, e->d_offsets, da + o_slots, da + o_maxnew, da + o_sidx,
                          const_cast<int32_t*>(e->st.seq_idx), e->st, n));
  if (e->slots_mega)
    MG_TRY(mega::mega_relayout_kv_slots(e->stream, e->d_mega_layers, e->st.lens, da + o_slots, n, g.n_layer, e->max_seq,
                                        mega_tvt(e->max_seq), g.d_model / g.n_head));
  if (e->slots_grid) {
    std::vector<const bf16*> kc(g.n_layer), vc(g.n_layer);
    std::vector<bf16*> kh(g.n_layer), vt(g.n_layer);
    for (int l = 0; l < g.n_layer; ++l) {
      kc[l] = reinterpret_cast<const bf16*>(e->layers[l].kc); vc[l] = reinterpret_cast<const bf16*>(e->layers[l].vc);
      kh[l] = reinterpret_cast<bf16*>(e->layers[l].kh); vt[l] = reinterpret_cast<bf16*>(e->layers[l].vt);
    }
    MG_TRY(grid::grid_relayout_kv(e->stream, kc.data(), vc.data(), kh.data(), vt.data(), e->st.lens, da + o_slots, n, g.n_layer, g.d_model,
                                  g.d_model / g.n_head, e->max_seq, mega_tvt(e->max_seq)));
  }
  for (int j = 0; j < n; ++j) e->slot_busy[slots[j]] = max_new[j] > 0 ? 1 : 0;
  return MG_OK;
}

int mg_slots_step(mg_engine* e, int n_steps, uint8_t* finished_out, int32_t* out_len_out) {
  if (!e || !finished_out || !out_len_out) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!e->slots_active) return fail(MG_E_STATE, "no slot session (mg_slots_begin)");
  if (n_steps <= 0 || n_steps > e->max_seq) return fail(MG_E_ARG, "n_steps must be in [1, max_seq]");
  e->cur_steps = n_steps;
  e->d_last_rows = e->d_slot_last_rows;
  e->last_run_mega = e->last_run_flow = e->last_run_grid = false;
  if (e->slots_mega) {
    int rc = MG_OK;
    if (!run_decode_mega(e, e->slots_topk, e->slots_eos, &rc, nullptr, nullptr, 0, nullptr))
      return fail(MG_E_STATE, "slot session: the persistent kernel refused the launch");
    MG_TRY(rc);
    e->last_run_mega = true;
  } else if (e->slots_grid) {
    int rc = MG_OK;
    if (!run_decode_grid(e, e->slots_topk, e->slots_eos, &rc, nullptr, nullptr, 0, nullptr))
      return fail(MG_E_STATE, "slot session: the grid kernel refused the launch");
    MG_TRY(rc);
    e->last_run_mega = true;
  } else {
    MG_TRY(e->dtype == MG_DTYPE_BF16 ? run_decode_loop<bf16>(e, e->slots_eos >= 0 ? e->slots_eos : 0)
                                     : run_decode_loop<float>(e, e->slots_eos >= 0 ? e->slots_eos : 0));
  }
  const int n = e->n_slots, fin_ints = (n + 3) / 4;
  MG_CUDA_OK(cudaMemcpyAsync(e->h_slot_flags, e->st.finished, n, cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaMemcpyAsync(e->h_slot_flags + fin_ints, e->st.out_len, n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  MG_TRY(persistent_status(e));
  e->d2h += n + n * sizeof(int32_t);
  std::memcpy(finished_out, e->h_slot_flags, n);
  std::memcpy(out_len_out, e->h_slot_flags + fin_ints, n * sizeof(int32_t));
  for (int b = 0; b < n; ++b)
    if (finished_out[b]) e->slot_busy[b] = 0;
  return MG_OK;
}

int mg_slots_fetch(mg_engine* e, int slot, int32_t* out_ids, int cap, int* n_out) {
  if (!e || !out_ids || !n_out) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!e->slots_active) return fail(MG_E_STATE, "no slot session (mg_slots_begin)");
  if (slot < 0 || slot >= e->n_slots) return fail(MG_E_ARG, "slot index outside the session");
  int32_t len = 0;
  MG_CUDA_OK(cudaMemcpyAsync(&len, e->st.out_len + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  if (len > cap) return fail(MG_E_ARG, "out buffer smaller than the sequence");
  MG_CUDA_OK(cudaMemcpyAsync(out_ids, e->st.out_ids + static_cast<size_t>(slot) * e->st.out_stride, sizeof(int32_t) * len,
                             cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  e->d2h += sizeof(int32_t) * (len + 1);
  *n_out = len;
  return MG_OK;
}

int mg_slots_fetch_many(mg_engine* e, int n, const int32_t* slots, int32_t* out_ids, int out_stride, int32_t* out_lens) {
  if (!e || !slots || !out_ids || !out_lens) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!e->slots_active) return fail(MG_E_STATE, "no slot session (mg_slots_begin)");
  if (n <= 0) return MG_OK;
  const int stride = e->st.out_stride;
  if (out_stride < stride) return fail(MG_E_ARG, "out_stride smaller than the session's row stride");
  for (int j = 0; j < n; ++j)
    if (slots[j] < 0 || slots[j] >= e->n_slots) return fail(MG_E_ARG, "slot index outside the session");
  // whole rows (stride ints each) + the lengths of all slots, then ONE synchronisation
  if (n > e->n_slots) return fail(MG_E_ARG, "more rows than slots");
  for (int j = 0; j < n; ++j)
    MG_CUDA_OK(cudaMemcpyAsync(e->h_slot_rows + static_cast<size_t>(j) * stride, e->st.out_ids + static_cast<size_t>(slots[j]) * stride,
                               sizeof(int32_t) * stride, cudaMemcpyDeviceToHost, e->stream));
  const int fin_ints = (e->n_slots + 3) / 4;
  MG_CUDA_OK(cudaMemcpyAsync(e->h_slot_flags + fin_ints, e->st.out_len, e->n_slots * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  for (int j = 0; j < n; ++j) {
    out_lens[j] = e->h_slot_flags[fin_ints + slots[j]];
    std::memcpy(out_ids + static_cast<size_t>(j) * out_stride, e->h_slot_rows + static_cast<size_t>(j) * stride, sizeof(int32_t) * stride);
  }
  e->d2h += sizeof(int32_t) * (static_cast<size_t>(n) * stride + e->n_slots);
  return MG_OK;
}

int mg_slots_end(mg_engine* e) {
  if (!e) return fail(MG_E_ARG, "null engine");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
  e->graph_B = -1;
  e->st = DecodeState{};
  e->slots_active = false;
  e->uploaded = false;
  return MG_OK;
}

// ---- device-side detokenisation (reference api_cache.py:157,208-221) ------------------------------------------------
int mg_set_note_table(mg_engine* e, const int32_t* kind, const int32_t* value, const float* start, const float* end, int V) {
  if (!e || !kind || !value || !start || !end) return fail(MG_E_ARG, "null argument");
  if (V != e->geo.vocab_size) return fail(MG_E_SHAPE, "note table must have one record per vocabulary entry");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  std::vector<int4> h(V);
  for (int i = 0; i < V; ++i) {
    if (kind[i] < DETOK_OTHER || kind[i] > DETOK_NOTE) return fail(MG_E_ARG, "note table kind must be 0 (other), 1 (instrument) or 2 (note)");
    int zs, ze;
    std::memcpy(&zs, &start[i], 4); std::memcpy(&ze, &end[i], 4);
    h[i] = make_int4(kind[i], value[i], zs, ze);
  }
  if (!e->d_note_table) MG_TRY(e->dmalloc(&e->d_note_table, sizeof(int4) * V));
  MG_CUDA_OK(cudaMemcpyAsync(e->d_note_table, h.data(), sizeof(int4) * V, cudaMemcpyHostToDevice, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));                       // `h` is pageable
  e->h2d += sizeof(int4) * V;
  return MG_OK;
}

int mg_note_events(mg_engine* e, int max_inst, int max_notes, int32_t* n_inst, int32_t* inst_program, int32_t* inst_token,
                   int32_t* n_notes, int32_t* note_inst, int32_t* note_pitch, float* note_start, float* note_end) {
  if (!e || !n_inst || !n_notes) return fail(MG_E_ARG, "null argument");
  if (max_inst < 0 || max_notes < 0) return fail(MG_E_ARG, "negative capacity");
  if ((max_inst > 0 && (!inst_program || !inst_token)) || (max_notes > 0 && (!note_inst || !note_pitch || !note_start || !note_end)))
    return fail(MG_E_ARG, "null event array");
  std::lock_guard<std::mutex> lk(e->mu);
  MG_CUDA_OK(cudaSetDevice(e->device));
  if (!e->d_note_table) return fail(MG_E_STATE, "no note table (call mg_set_note_table first)");
  if (!e->uploaded) return fail(MG_E_STATE, "nothing generated yet (call mg_generate or mg_upload_prompts + mg_run first)");
  const int B = e->cur_B;
  const size_t per = 2 + 2 * static_cast<size_t>(max_inst) + 4 * static_cast<size_t>(max_notes), ints = per * B;
  if (ints > e->detok_cap) {
    MG_CUDA_OK(cudaStreamSynchronize(e->stream));
    e->dfree(e->d_detok); e->d_detok = nullptr;
    if (e->h_detok) { cudaFreeHost(e->h_detok); e->h_detok = nullptr; }
    MG_TRY(e->dmalloc(&e->d_detok, ints * sizeof(int32_t)));
    MG_CUDA_OK(cudaMallocHost(&e->h_detok, ints * sizeof(int32_t)));
    e->detok_cap = ints;
  }
  // block layout: n_inst [B] | n_notes [B] | inst_program [B][mi] | inst_token [B][mi] | note_inst, note_pitch, note_start, note_end [B][mn]
  int32_t* d = e->d_detok;
  DetokOut o{};
  o.max_inst = max_inst; o.max_notes = max_notes;
  o.n_inst = d; o.n_notes = d + B;
  o.inst_program = d + 2 * B; o.inst_token = o.inst_program + static_cast<size_t>(B) * max_inst;
  o.note_inst = o.inst_token + static_cast<size_t>(B) * max_inst; o.note_pitch = o.note_inst + static_cast<size_t>(B) * max_notes;
  o.note_start = reinterpret_cast<float*>(o.note_pitch + static_cast<size_t>(B) * max_notes);
  o.note_end = o.note_start + static_cast<size_t>(B) * max_notes;
  MG_TRY(launch_detok(e->stream, e->st.out_ids, e->st.out_len, e->st.out_stride, e->d_note_table, e->geo.vocab_size, B, o));
  MG_CUDA_OK(cudaMemcpyAsync(e->h_detok, d, ints * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  MG_TRY(persistent_status(e));
  e->d2h += ints * sizeof(int32_t);
  const int32_t* h = e->h_detok;
  std::memcpy(n_inst, h, sizeof(int32_t) * B);
  std::memcpy(n_notes, h + B, sizeof(int32_t) * B);
  const size_t bi = static_cast<size_t>(B) * max_inst, bn = static_cast<size_t>(B) * max_notes;
  if (bi) { std::memcpy(inst_program, h + 2 * B, 4 * bi); std::memcpy(inst_token, h + 2 * B + bi, 4 * bi); }
  if (bn) {
    const int32_t* q = h + 2 * B + 2 * bi;
    std::memcpy(note_inst, q, 4 * bn); std::memcpy(note_pitch, q + bn, 4 * bn);
    std::memcpy(note_start, q + 2 * bn, 4 * bn); std::memcpy(note_end, q + 3 * bn, 4 * bn);
  }
  return MG_OK;
}

int mg_last_step_times(mg_engine* e, float* us_out, int cap, int* n_out) {
  if (!e || !n_out || (cap > 0 && !us_out)) return fail(MG_E_ARG, "null argument");
  *n_out = 0;
  const int steps = std::min(e->t_steps, e->max_seq + 1);
  if (!e->d_step_ns || steps < 2) return MG_OK;
  MG_CUDA_OK(cudaStreamSynchronize(e->stream));
  std::vector<unsigned long long> h(steps);
  MG_CUDA_OK(cudaMemcpy(h.data(), e->d_step_ns, sizeof(unsigned long long) * steps, cudaMemcpyDeviceToHost));
  const int n = std::min(cap, steps - 1);
  for (int i = 0; i < n; ++i) us_out[i] = static_cast<float>(h[i + 1] - h[i]) * 1e-3f;
  *n_out = n;
  return MG_OK;
}

int mg_test_grid_plan(int d_model, int d_ff, int n_layer, int vocab, int B, int n_cta, int32_t* tn_ks, int16_t* items, int32_t* n_items) {
  if (!tn_ks || !items || !n_items || n_cta <= 0 || B <= 0 || B > grid::kMaxSeqs) return fail(MG_E_ARG, "bad argument");
  if (!grid::grid_eligible(d_model, d_ff, d_model / 32 >= 8 ? 8 : d_model / 32, n_layer, vocab)) return fail(MG_E_SHAPE, "geometry not eligible");
  static_assert(sizeof(grid::GridItem) == 4 * sizeof(int16_t), "GridItem is four int16");
  int tn[8], ks[8];
  const int rc = grid::grid_plan(d_model, d_ff, n_layer, vocab, B, n_cta, tn, ks, reinterpret_cast<grid::GridItem*>(items), n_items);
  for (int k = 0; k < 8; ++k) { tn_ks[k] = tn[k]; tn_ks[8 + k] = ks[k]; }
  return rc == MG_OK ? MG_OK : fail(rc, "grid plan: a CTA would get more than kMaxItems items");
}

int mg_test_gemm_bf16(int device, const float* A, const float* W, const float* bias, int M, int N, int K, int act, float* C) {
  if (!A || !W || !C) return fail(MG_E_ARG, "null argument");
  if (M <= 0 || N <= 0 || K <= 0 || K % 8) return fail(MG_E_SHAPE, "M, N, K must be positive, K a multiple of 8");
  MG_TRY(check_device(device));
  MG_TRY(gemm_tc_init());
  float *dA32 = nullptr, *dW32 = nullptr, *dC = nullptr, *dbias = nullptr;
  bf16 *dA = nullptr, *dW = nullptr;
  const int Mp = ceil_div(M, 128) * 128;
  int rc = MG_OK;
  auto body = [&]() -> int {
    MG_CUDA_OK(cudaMalloc(&dA32, sizeof(float) * M * K));
    MG_CUDA_OK(cudaMalloc(&dW32, sizeof(float) * N * K));
    MG_CUDA_OK(cudaMalloc(&dA, sizeof(bf16) * Mp * K));
    MG_CUDA_OK(cudaMalloc(&dW, sizeof(bf16) * N * K));
    MG_CUDA_OK(cudaMalloc(&dC, sizeof(float) * M * N));
    MG_CUDA_OK(cudaMemset(dA, 0, sizeof(bf16) * Mp * K));
    MG_CUDA_OK(cudaMemcpy(dA32, A, sizeof(float) * M * K, cudaMemcpyHostToDevice));
    MG_CUDA_OK(cudaMemcpy(dW32, W, sizeof(float) * N * K, cudaMemcpyHostToDevice));
    if (bias) {
      MG_CUDA_OK(cudaMalloc(&dbias, sizeof(float) * N));
      MG_CUDA_OK(cudaMemcpy(dbias, bias, sizeof(float) * N, cudaMemcpyHostToDevice));
    }
    MG_TRY(launch_convert<bf16>(nullptr, dA32, dA, static_cast<size_t>(M) * K));
    MG_TRY(launch_convert<bf16>(nullptr, dW32, dW, static_cast<size_t>(N) * K));
    const bool pair = gemm_use_pair(M, N);
    const int bn = pair ? 128 : pick_gemm_bn(M, N);
    CUtensorMap ta, tw;
    MG_TRY(make_tmap_bf16_2d(&ta, dA, Mp, K, kGemmBM));
    MG_TRY(make_tmap_bf16_2d(&tw, dW, N, K, bn));
    GemmEpilogue epi; epi.bias = dbias; epi.act = act; epi.out_f32 = dC; epi.ld_out = N;
    if (pair) MG_TRY(launch_gemm_tc_pair(nullptr, &ta, &tw, M, N, K, epi));
    else MG_TRY(launch_gemm_tc(nullptr, &ta, &tw, M, N, K, epi, bn));
    MG_CUDA_OK(cudaDeviceSynchronize());
    MG_CUDA_OK(cudaMemcpy(C, dC, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    return MG_OK;
  };
  rc = body();
  cudaFree(dA32); cudaFree(dW32); cudaFree(dA); cudaFree(dW); cudaFree(dC); cudaFree(dbias);
  return rc;
}

}  // extern "C"
