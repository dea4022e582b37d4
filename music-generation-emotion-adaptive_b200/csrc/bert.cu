// mg_bert: the DistilBERT emotion-classifier replica behind the C ABI of include/mg_engine.h.
//
// Replaces the forward + argmax of the reference's `inference.predict`
// (emotion_analysis/inference.py:16-21) and the model construction of `load_model`
// (emotion_analysis/modeling.py:14-21).  The arithmetic of that path lives in the third-party
// `transformers` DistilBertForSequenceClassification (post-LN encoder, exact GELU, eps 1e-12):
//   embeddings : word[id] + position[t] -> LayerNorm
//   6 x block  : q/k/v_lin -> softmax(q k^T / sqrt(hd) + padding mask) v -> out_lin -> LN(. + x)
//                -> lin1 -> GELU -> lin2 -> LN(. + sa_out)
//   head       : hidden[:, 0] -> pre_classifier -> ReLU -> classifier -> argmax
// LoRA (q_lin, v_lin) is merged into the dense weights by the host before upload.
//
// bf16 weights and activations, fp32 accumulation / LayerNorm statistics / softmax.  q_lin, k_lin
// and v_lin are stored as one [3*dim, dim] matrix so the three projections are a single GEMM.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "mg_engine.h"

using namespace mg;

struct BertLayerW {
  bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *sa_w = nullptr, *sa_b = nullptr, *out_w = nullptr, *out_b = nullptr;
  WMaps m_qkv, m_o, m_1, m_2;
};

struct mg_bert {
  mg_bert_geometry geo{};
  int device = 0, max_tokens = 0;
  bool use_tc = true, ready = false, uploaded = false;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  std::set<std::string> loaded;

  bf16 *word = nullptr, *pos = nullptr, *wpre = nullptr, *wcls = nullptr;
  float *emb_w = nullptr, *emb_b = nullptr, *bpre = nullptr, *bcls = nullptr;
  WMaps m_pre, m_cls;
  std::vector<BertLayerW> layers;
  float* stage_f32 = nullptr;
  size_t stage_elems = 0;

  // activations: hA / hB hidden states, pre = pre-LayerNorm sums, qkv, att, h1 (FFN), cls rows
  bf16 *hA = nullptr, *hB = nullptr, *pre = nullptr, *qkv = nullptr, *att = nullptr, *h1 = nullptr, *cls = nullptr,
       *cls2 = nullptr;
  CUtensorMap tm_hA, tm_hB, tm_att, tm_h1, tm_cls, tm_cls2;
  float* logits = nullptr;
  int32_t* labels = nullptr;

  int32_t* d_arena = nullptr;          // ids[M] pos[M] seq_start[N] seq_len[N] cls_rows[N] mask[M bytes]
  int32_t* h_arena = nullptr;          // pinned
  cudaEvent_t ev_staged = nullptr;     // behind the H2D copy out of h_arena: waited for before the buffer is rewritten
  size_t arena_ints = 0;
  int32_t *d_ids = nullptr, *d_pos = nullptr, *d_seq_start = nullptr, *d_seq_len = nullptr, *d_cls_rows = nullptr;
  uint8_t* d_mask = nullptr;
  float* h_logits = nullptr;           // pinned [max_texts][labels] + labels
  int cur_N = 0, cur_T = 0;

  uint64_t h2d = 0, d2h = 0, launches0 = 0;

  // the ~50 launches of a pass replayed as ONE CUDA graph once a (N, T) shape has been seen three times in a row
  // (MG_BERT_GRAPH=0: always eager launches)
  bool use_graph = true;
  cudaGraphExec_t graph = nullptr;
  int graph_N = 0, graph_T = 0, shape_runs = 0;
  uint64_t graph_kernels = 0;
  int prof_which = 0;                       // MG_PAIR_PROF (debug)
  unsigned long long* d_prof = nullptr;

  template <typename P> int dmalloc(P** p, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
    if (e != cudaSuccess) return fail(MG_E_OOM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    allocs.push_back(q);
    *p = reinterpret_cast<P*>(q);
    return MG_OK;
  }
};

namespace {

int bgemm(mg_bert* b, const bf16* A, const CUtensorMap* tmA, const bf16* W, const WMaps* wm, int M, int N, int K,
          const GemmEpilogue& epi) {
  if (b->use_tc && M >= 32) {
    if (gemm_use_pair(M, N)) return launch_gemm_tc_pair(b->stream, tmA, &wm->m[bn_index(128)], M, N, K, epi);
    const int bn = pick_gemm_bn(M, N);
    return launch_gemm_tc(b->stream, tmA, &wm->m[bn_index(bn)], M, N, K, epi, bn);
  }
  return launch_gemm_simt<bf16>(b->stream, A, K, W, M, N, K, epi);
}

int bert_forward(mg_bert* b) {
  const mg_bert_geometry& g = b->geo;
  const int N = b->cur_N, T = b->cur_T, M = N * T, d = g.dim, f = g.hidden_dim;
  const float eps = 1e-12f;
  MG_TRY(launch_embed_ln<bf16>(b->stream, b->d_ids, b->d_pos, b->word, b->pos, b->emb_w, b->emb_b, nullptr, b->hA, M, d, eps,
                               true));
  for (int l = 0; l < g.n_layers; ++l) {
    BertLayerW& w = b->layers[l];
    { GemmEpilogue epi; epi.bias = w.bqkv; epi.out_bf16 = b->qkv; epi.ld_out = 3 * d;
      if (l == 2 && b->prof_which == 3) epi.prof = b->d_prof;
      MG_TRY(bgemm(b, b->hA, &b->tm_hA, w.wqkv, &w.m_qkv, M, 3 * d, d, epi)); }
    MG_TRY(launch_encoder_attn<bf16>(b->stream, b->qkv, b->d_seq_start, b->d_seq_len, b->d_mask, b->att, N, d, g.n_heads, T));
    { GemmEpilogue epi; epi.bias = w.bo; epi.resid_bf16 = b->hA; epi.out_bf16 = b->pre; epi.ld_out = d;
      if (l == 2 && b->prof_which == 4) epi.prof = b->d_prof;
      MG_TRY(bgemm(b, b->att, &b->tm_att, w.wo, &w.m_o, M, d, d, epi)); }
    MG_TRY((launch_layernorm<bf16, bf16>(b->stream, b->pre, w.sa_w, w.sa_b, b->hB, nullptr, M, d, eps)));
    // debug: MG_PAIR_PROF=<1 lin1 | 2 lin2 | 3 qkv | 4 out_proj> -> per-tile timeline of one CTA pair in layer 2 (see bert_prof_dump)
    { GemmEpilogue epi; epi.bias = w.b1; epi.act = ACT_GELU; epi.out_bf16 = b->h1; epi.ld_out = f;
      if (l == 2 && b->prof_which == 1) epi.prof = b->d_prof;
      MG_TRY(bgemm(b, b->hB, &b->tm_hB, w.w1, &w.m_1, M, f, d, epi)); }
    { GemmEpilogue epi; epi.bias = w.b2; epi.resid_bf16 = b->hB; epi.out_bf16 = b->pre; epi.ld_out = d;
      if (l == 2 && b->prof_which == 2) epi.prof = b->d_prof;
      MG_TRY(bgemm(b, b->h1, &b->tm_h1, w.w2, &w.m_2, M, d, f, epi)); }
    MG_TRY((launch_layernorm<bf16, bf16>(b->stream, b->pre, w.out_w, w.out_b, b->hA, nullptr, M, d, eps)));
  }
  MG_TRY((launch_gather_rows<bf16, bf16>(b->stream, b->hA, b->d_cls_rows, b->cls, N, d)));
  { GemmEpilogue epi; epi.bias = b->bpre; epi.act = ACT_RELU; epi.out_bf16 = b->cls2; epi.ld_out = d;
    MG_TRY(bgemm(b, b->cls, &b->tm_cls, b->wpre, &b->m_pre, N, d, d, epi)); }
  { GemmEpilogue epi; epi.bias = b->bcls; epi.out_f32 = b->logits; epi.ld_out = g.num_labels;
    MG_TRY(bgemm(b, b->cls2, &b->tm_cls2, b->wcls, &b->m_cls, N, g.num_labels, d, epi)); }
  MG_TRY(launch_argmax_rows(b->stream, b->logits, N, g.num_labels, b->labels));
  return MG_OK;
}

// debug read-out of MG_PAIR_PROF: per tile {MMA issue start, last MMA issued, epilogue start, epilogue end} in ns
static void bert_prof_dump(mg_bert* b) {
  unsigned long long h[128];
  cudaStreamSynchronize(b->stream);
  cudaMemcpy(h, b->d_prof, sizeof(h), cudaMemcpyDeviceToHost);
  unsigned long long t0 = ~0ull;
  for (int i = 0; i < 128; ++i) if (h[i]) t0 = std::min(t0, h[i]);
  fprintf(stderr, "[pair prof %d] tile: mma_start mma_issued | epi_start epi_end (ns)\n", b->prof_which);
  for (int t = 0; t < 32 && h[4 * t]; ++t)
    fprintf(stderr, "[pair prof] %2d: %7llu %7llu | %7llu %7llu\n", t, h[4 * t] - t0, h[4 * t + 1] - t0, h[4 * t + 2] - t0, h[4 * t + 3] - t0);
  cudaMemset(b->d_prof, 0, sizeof(h));
}

int bert_run(mg_bert* b) {
  if (b->prof_which) { const int rc = bert_forward(b); bert_prof_dump(b); return rc; }
  if (!b->use_graph) return bert_forward(b);
  if (b->graph_N != b->cur_N || b->graph_T != b->cur_T) {       // new shape: start counting again
    if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
    b->graph_N = b->cur_N;
    b->graph_T = b->cur_T;
    b->shape_runs = 0;
  }
  // capture + instantiate costs tens of milliseconds: only a shape that keeps coming back (third pass on) is worth it
  if (!b->graph && ++b->shape_runs < 3) return bert_forward(b);
  if (!b->graph) {
    cudaGraph_t gr = nullptr;
    MG_CUDA_OK(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
    const uint64_t before = g_kernel_launches.load();
    const int rc = bert_forward(b);
    b->graph_kernels = g_kernel_launches.load() - before;
    g_kernel_launches.fetch_sub(b->graph_kernels);            // capture launches nothing; replays are counted
    cudaError_t ce = cudaStreamEndCapture(b->stream, &gr);
    if (rc != MG_OK) { if (gr) cudaGraphDestroy(gr); return rc; }
    if (ce != cudaSuccess) return fail(MG_E_CUDA, std::string("classifier graph capture: ") + cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&b->graph, gr, 0);
    cudaGraphDestroy(gr);
    if (ce != cudaSuccess) return fail(MG_E_CUDA, std::string("classifier graph instantiate: ") + cudaGetErrorString(ce));
  }
  MG_CUDA_OK(cudaGraphLaunch(b->graph, b->stream));
  g_kernel_launches.fetch_add(b->graph_kernels, std::memory_order_relaxed);
  return MG_OK;
}

int bert_upload_tensor(mg_bert* b, void* dst, bool typed, const float* data, size_t n) {
  if (!typed) {
    MG_CUDA_OK(cudaMemcpyAsync(dst, data, n * sizeof(float), cudaMemcpyHostToDevice, b->stream));
  } else {
    if (n > b->stage_elems) {
      if (b->stage_f32) {
        auto it = std::find(b->allocs.begin(), b->allocs.end(), static_cast<void*>(b->stage_f32));
        if (it != b->allocs.end()) b->allocs.erase(it);
        cudaFree(b->stage_f32);
        b->stage_f32 = nullptr;
      }
      MG_TRY(b->dmalloc(&b->stage_f32, n * sizeof(float)));
      b->stage_elems = n;
    }
    MG_CUDA_OK(cudaMemcpyAsync(b->stage_f32, data, n * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    MG_TRY(launch_convert<bf16>(b->stream, b->stage_f32, reinterpret_cast<bf16*>(dst), n));
  }
  MG_CUDA_OK(cudaStreamSynchronize(b->stream));
  b->h2d += n * sizeof(float);
  return MG_OK;
}

int bert_upload_impl(mg_bert* b, const int32_t* ids, const uint8_t* mask, int N, int T) {
  if (!b->ready) return fail(MG_E_STATE, "classifier not finalized (mg_bert_finalize)");
  if (!ids || N <= 0 || T <= 0) return fail(MG_E_ARG, "classify: null ids or empty batch");
  if (T > b->geo.max_pos) return fail(MG_E_PROMPT_TOO_LONG, "text longer than the position table");
  const size_t M = static_cast<size_t>(N) * T;
  if (M > static_cast<size_t>(b->max_tokens)) return fail(MG_E_OOM, "N * T exceeds max_tokens given at mg_bert_create");
  for (size_t i = 0; i < M; ++i)
    if (ids[i] < 0 || ids[i] >= b->geo.vocab_size) return fail(MG_E_TOKEN, "token id outside [0, vocab)");
  int32_t* h = b->h_arena;
  if (b->ev_staged) MG_CUDA_OK(cudaEventSynchronize(b->ev_staged));   // an earlier asynchronous upload may still be reading it
  int32_t* h_ids = h;
  int32_t* h_pos = h_ids + M;
  int32_t* h_ss = h_pos + M;
  int32_t* h_sl = h_ss + N;
  int32_t* h_cr = h_sl + N;
  uint8_t* h_mask = reinterpret_cast<uint8_t*>(h_cr + N);
  std::memcpy(h_ids, ids, M * sizeof(int32_t));
  for (int n = 0; n < N; ++n) {
    for (int t = 0; t < T; ++t) h_pos[static_cast<size_t>(n) * T + t] = t;
    h_ss[n] = n * T;
    h_sl[n] = T;
    h_cr[n] = n * T;
  }
  if (mask) std::memcpy(h_mask, mask, M);
  else std::memset(h_mask, 1, M);
  const size_t bytes = (2 * M + 3 * static_cast<size_t>(N)) * sizeof(int32_t) + M;
  MG_CUDA_OK(cudaMemcpyAsync(b->d_arena, h, bytes, cudaMemcpyHostToDevice, b->stream));
  if (!b->ev_staged) MG_CUDA_OK(cudaEventCreateWithFlags(&b->ev_staged, cudaEventDisableTiming));
  MG_CUDA_OK(cudaEventRecord(b->ev_staged, b->stream));
  b->h2d += bytes;
  b->d_ids = b->d_arena;
  b->d_pos = b->d_ids + M;
  b->d_seq_start = b->d_pos + M;
  b->d_seq_len = b->d_seq_start + N;
  b->d_cls_rows = b->d_seq_len + N;
  b->d_mask = reinterpret_cast<uint8_t*>(b->d_cls_rows + N);
  b->cur_N = N;
  b->cur_T = T;
  b->uploaded = true;
  return MG_OK;
}

int bert_download_impl(mg_bert* b, float* logits_out, int32_t* label_out) {
  if (!b->uploaded) return fail(MG_E_STATE, "nothing to download");
  const int N = b->cur_N, C = b->geo.num_labels;
  float* hl = b->h_logits;
  int32_t* hlab = reinterpret_cast<int32_t*>(hl + static_cast<size_t>(N) * C);
  MG_CUDA_OK(cudaMemcpyAsync(hl, b->logits, sizeof(float) * N * C, cudaMemcpyDeviceToHost, b->stream));
  MG_CUDA_OK(cudaMemcpyAsync(hlab, b->labels, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, b->stream));
  MG_CUDA_OK(cudaStreamSynchronize(b->stream));
  b->d2h += sizeof(float) * N * C + sizeof(int32_t) * N;
  if (logits_out) std::memcpy(logits_out, hl, sizeof(float) * N * C);
  if (label_out) std::memcpy(label_out, hlab, sizeof(int32_t) * N);
  return MG_OK;
}

}  // namespace

extern "C" {

int mg_bert_create(const mg_bert_geometry* geo, int device, int max_tokens, mg_bert** out) {
  if (!geo || !out) return fail(MG_E_ARG, "null argument");
  *out = nullptr;
  const mg_bert_geometry g = *geo;
  if (g.vocab_size <= 0 || g.max_pos <= 0 || g.dim <= 0 || g.n_heads <= 0 || g.n_layers <= 0 || g.hidden_dim <= 0 ||
      g.num_labels <= 0)
    return fail(MG_E_SHAPE, "non-positive geometry field");
  if (g.dim % g.n_heads || g.dim % 16 || g.hidden_dim % 16 || g.dim > 1024) return fail(MG_E_SHAPE, "unsupported dim / hidden_dim");
  const int hd = g.dim / g.n_heads;
  if (hd > 64 || hd % 8) return fail(MG_E_SHAPE, "head_dim must be a multiple of 8, at most 64");
  if (max_tokens <= 0) return fail(MG_E_ARG, "max_tokens must be positive");
  MG_TRY(check_device(device));
  mg_bert* b = new mg_bert();
  b->geo = g; b->device = device;
  b->max_tokens = ceil_div(max_tokens, 128) * 128;
  const char* env_gemm = std::getenv("MG_GEMM");
  b->use_tc = !(env_gemm && std::strcmp(env_gemm, "simt") == 0);
  if (const char* pp = std::getenv("MG_PAIR_PROF")) {
    b->prof_which = std::atoi(pp);
    if (b->prof_which && cudaMalloc(&b->d_prof, 128 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(b->d_prof, 0, 128 * sizeof(unsigned long long));
    else b->prof_which = 0;
  }
  const char* env_graph = std::getenv("MG_BERT_GRAPH");
  b->use_graph = !(env_graph && std::atoi(env_graph) == 0);
  b->launches0 = g_kernel_launches.load();
  auto body = [&]() -> int {
    MG_CUDA_OK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    for (auto& ev : b->ev) MG_CUDA_OK(cudaEventCreate(&ev));
    MG_TRY(kernels_init());
    if (b->use_tc) MG_TRY(gemm_tc_init());
    const size_t d = g.dim, f = g.hidden_dim, R = b->max_tokens;
    MG_TRY(b->dmalloc(&b->word, sizeof(bf16) * g.vocab_size * d));
    MG_TRY(b->dmalloc(&b->pos, sizeof(bf16) * g.max_pos * d));
    MG_TRY(b->dmalloc(&b->emb_w, sizeof(float) * d));
    MG_TRY(b->dmalloc(&b->emb_b, sizeof(float) * d));
    MG_TRY(b->dmalloc(&b->wpre, sizeof(bf16) * d * d));
    MG_TRY(b->dmalloc(&b->bpre, sizeof(float) * d));
    MG_TRY(b->dmalloc(&b->wcls, sizeof(bf16) * g.num_labels * d));
    MG_TRY(b->dmalloc(&b->bcls, sizeof(float) * g.num_labels));
    b->layers.resize(g.n_layers);
    for (auto& w : b->layers) {
      MG_TRY(b->dmalloc(&w.wqkv, sizeof(bf16) * 3 * d * d));
      MG_TRY(b->dmalloc(&w.wo, sizeof(bf16) * d * d));
      MG_TRY(b->dmalloc(&w.w1, sizeof(bf16) * f * d));
      MG_TRY(b->dmalloc(&w.w2, sizeof(bf16) * d * f));
      MG_TRY(b->dmalloc(&w.bqkv, sizeof(float) * 3 * d));
      MG_TRY(b->dmalloc(&w.bo, sizeof(float) * d));
      MG_TRY(b->dmalloc(&w.b1, sizeof(float) * f));
      MG_TRY(b->dmalloc(&w.b2, sizeof(float) * d));
      MG_TRY(b->dmalloc(&w.sa_w, sizeof(float) * d));
      MG_TRY(b->dmalloc(&w.sa_b, sizeof(float) * d));
      MG_TRY(b->dmalloc(&w.out_w, sizeof(float) * d));
      MG_TRY(b->dmalloc(&w.out_b, sizeof(float) * d));
    }
    MG_TRY(b->dmalloc(&b->hA, sizeof(bf16) * R * d));
    MG_TRY(b->dmalloc(&b->hB, sizeof(bf16) * R * d));
    MG_TRY(b->dmalloc(&b->pre, sizeof(bf16) * R * d));
    MG_TRY(b->dmalloc(&b->qkv, sizeof(bf16) * R * 3 * d));
    MG_TRY(b->dmalloc(&b->att, sizeof(bf16) * R * d));
    MG_TRY(b->dmalloc(&b->h1, sizeof(bf16) * R * f));
    MG_TRY(b->dmalloc(&b->cls, sizeof(bf16) * R * d));
    MG_TRY(b->dmalloc(&b->cls2, sizeof(bf16) * R * d));
    MG_TRY(b->dmalloc(&b->logits, sizeof(float) * R * g.num_labels));
    MG_TRY(b->dmalloc(&b->labels, sizeof(int32_t) * R));
    MG_CUDA_OK(cudaMemsetAsync(b->att, 0, sizeof(bf16) * R * d, b->stream));
    MG_CUDA_OK(cudaMemsetAsync(b->cls, 0, sizeof(bf16) * R * d, b->stream));
    MG_CUDA_OK(cudaMemsetAsync(b->cls2, 0, sizeof(bf16) * R * d, b->stream));
    if (b->use_tc) {
      MG_TRY(make_tmap_bf16_2d(&b->tm_hA, b->hA, R, d, kGemmBM));
      MG_TRY(make_tmap_bf16_2d(&b->tm_hB, b->hB, R, d, kGemmBM));
      MG_TRY(make_tmap_bf16_2d(&b->tm_att, b->att, R, d, kGemmBM));
      MG_TRY(make_tmap_bf16_2d(&b->tm_h1, b->h1, R, f, kGemmBM));
      MG_TRY(make_tmap_bf16_2d(&b->tm_cls, b->cls, R, d, kGemmBM));
      MG_TRY(make_tmap_bf16_2d(&b->tm_cls2, b->cls2, R, d, kGemmBM));
    }
    b->arena_ints = 2 * R + 3 * R + R / 4 + 64;
    MG_TRY(b->dmalloc(&b->d_arena, b->arena_ints * sizeof(int32_t)));
    MG_CUDA_OK(cudaMallocHost(&b->h_arena, b->arena_ints * sizeof(int32_t)));
    MG_CUDA_OK(cudaMallocHost(&b->h_logits, (sizeof(float) * g.num_labels + sizeof(int32_t)) * R));
    MG_CUDA_OK(cudaStreamSynchronize(b->stream));
    return MG_OK;
  };
  const int rc = body();
  if (rc != MG_OK) {
    const std::string keep = mg_last_error();
    mg_bert_destroy(b);
    set_last_error(keep);
    return rc;
  }
  *out = b;
  return MG_OK;
}

void mg_bert_destroy(mg_bert* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (void* p : b->allocs) cudaFree(p);
  if (b->h_arena) cudaFreeHost(b->h_arena);
  if (b->ev_staged) cudaEventDestroy(b->ev_staged);
  if (b->h_logits) cudaFreeHost(b->h_logits);
  for (auto& ev : b->ev) if (ev) cudaEventDestroy(ev);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int mg_bert_load_weight(mg_bert* b, const char* name_c, const float* data, const int64_t* shape, int ndim) {
  if (!b || !name_c || !data || !shape) return fail(MG_E_ARG, "null argument");
  std::lock_guard<std::mutex> lk(b->mu);
  MG_CUDA_OK(cudaSetDevice(b->device));
  const mg_bert_geometry& g = b->geo;
  const int64_t d = g.dim, f = g.hidden_dim;
  const std::string name(name_c);
  void* dst = nullptr;
  bool typed = true;
  int64_t s0 = 0, s1 = 0;
  const std::string lp = "distilbert.transformer.layer.";
  if (name == "distilbert.embeddings.word_embeddings.weight") { dst = b->word; s0 = g.vocab_size; s1 = d; }
  else if (name == "distilbert.embeddings.position_embeddings.weight") { dst = b->pos; s0 = g.max_pos; s1 = d; }
  else if (name == "distilbert.embeddings.LayerNorm.weight") { dst = b->emb_w; s0 = d; typed = false; }
  else if (name == "distilbert.embeddings.LayerNorm.bias") { dst = b->emb_b; s0 = d; typed = false; }
  else if (name == "pre_classifier.weight") { dst = b->wpre; s0 = d; s1 = d; }
  else if (name == "pre_classifier.bias") { dst = b->bpre; s0 = d; typed = false; }
  else if (name == "classifier.weight") { dst = b->wcls; s0 = g.num_labels; s1 = d; }
  else if (name == "classifier.bias") { dst = b->bcls; s0 = g.num_labels; typed = false; }
  else if (name.compare(0, lp.size(), lp) == 0) {
    const size_t dot = name.find('.', lp.size());
    if (dot != std::string::npos) {
      const int l = std::atoi(name.substr(lp.size(), dot - lp.size()).c_str());
      const std::string leaf = name.substr(dot + 1);
      if (l >= 0 && l < g.n_layers) {
        BertLayerW& w = b->layers[l];
        if (leaf == "attention.q_lin.weight") { dst = w.wqkv; s0 = d; s1 = d; }
        else if (leaf == "attention.k_lin.weight") { dst = w.wqkv + d * d; s0 = d; s1 = d; }
        else if (leaf == "attention.v_lin.weight") { dst = w.wqkv + 2 * d * d; s0 = d; s1 = d; }
        else if (leaf == "attention.q_lin.bias") { dst = w.bqkv; s0 = d; typed = false; }
        else if (leaf == "attention.k_lin.bias") { dst = w.bqkv + d; s0 = d; typed = false; }
        else if (leaf == "attention.v_lin.bias") { dst = w.bqkv + 2 * d; s0 = d; typed = false; }
        else if (leaf == "attention.out_lin.weight") { dst = w.wo; s0 = d; s1 = d; }
        else if (leaf == "attention.out_lin.bias") { dst = w.bo; s0 = d; typed = false; }
        else if (leaf == "sa_layer_norm.weight") { dst = w.sa_w; s0 = d; typed = false; }
        else if (leaf == "sa_layer_norm.bias") { dst = w.sa_b; s0 = d; typed = false; }
        else if (leaf == "ffn.lin1.weight") { dst = w.w1; s0 = f; s1 = d; }
        else if (leaf == "ffn.lin1.bias") { dst = w.b1; s0 = f; typed = false; }
        else if (leaf == "ffn.lin2.weight") { dst = w.w2; s0 = d; s1 = f; }
        else if (leaf == "ffn.lin2.bias") { dst = w.b2; s0 = d; typed = false; }
        else if (leaf == "output_layer_norm.weight") { dst = w.out_w; s0 = d; typed = false; }
        else if (leaf == "output_layer_norm.bias") { dst = w.out_b; s0 = d; typed = false; }
      }
    }
  }
  if (!dst) return fail(MG_E_SHAPE, "unknown tensor name '" + name + "' for this classifier geometry");
  const bool shape_ok = (s1 == 0) ? (ndim == 1 && shape[0] == s0) : (ndim == 2 && shape[0] == s0 && shape[1] == s1);
  if (!shape_ok) return fail(MG_E_SHAPE, "shape mismatch for '" + name + "'");
  MG_TRY(bert_upload_tensor(b, dst, typed, data, static_cast<size_t>(s0) * (s1 ? s1 : 1)));
  b->loaded.insert(name);
  b->ready = false;
  return MG_OK;
}

int mg_bert_finalize(mg_bert* b) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  std::lock_guard<std::mutex> lk(b->mu);
  MG_CUDA_OK(cudaSetDevice(b->device));
  const size_t want = 8 + 16 * static_cast<size_t>(b->geo.n_layers);
  if (b->loaded.size() != want)
    return fail(MG_E_STATE, "weights missing: " + std::to_string(b->loaded.size()) + " of " + std::to_string(want) + " tensors loaded");
  if (b->use_tc) {
    const int d = b->geo.dim, f = b->geo.hidden_dim;
    for (auto& w : b->layers) {
      MG_TRY(make_wmaps(&w.m_qkv, w.wqkv, 3 * d, d));
      MG_TRY(make_wmaps(&w.m_o, w.wo, d, d));
      MG_TRY(make_wmaps(&w.m_1, w.w1, f, d));
      MG_TRY(make_wmaps(&w.m_2, w.w2, d, f));
    }
    MG_TRY(make_wmaps(&b->m_pre, b->wpre, d, d));
    MG_TRY(make_wmaps(&b->m_cls, b->wcls, b->geo.num_labels, d));
  }
  b->ready = true;
  return MG_OK;
}

int mg_bert_upload(mg_bert* b, const int32_t* ids, const uint8_t* mask, int N, int T) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  std::lock_guard<std::mutex> lk(b->mu);
  MG_CUDA_OK(cudaSetDevice(b->device));
  return bert_upload_impl(b, ids, mask, N, T);
}

int mg_bert_run(mg_bert* b) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  std::lock_guard<std::mutex> lk(b->mu);
  MG_CUDA_OK(cudaSetDevice(b->device));
  if (!b->uploaded) return fail(MG_E_STATE, "mg_bert_run before mg_bert_upload");
  return bert_run(b);
}

int mg_bert_download(mg_bert* b, float* logits_out, int32_t* label_out) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  std::lock_guard<std::mutex> lk(b->mu);
  MG_CUDA_OK(cudaSetDevice(b->device));
  return bert_download_impl(b, logits_out, label_out);
}

int mg_classify(mg_bert* b, const int32_t* ids, const uint8_t* mask, int N, int T, float* logits_out, int32_t* label_out) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  std::lock_guard<std::mutex> lk(b->mu);
  MG_CUDA_OK(cudaSetDevice(b->device));
  MG_TRY(bert_upload_impl(b, ids, mask, N, T));
  MG_TRY(bert_run(b));
  return bert_download_impl(b, logits_out, label_out);
}

int mg_bert_synchronize(mg_bert* b) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  MG_CUDA_OK(cudaSetDevice(b->device));
  MG_CUDA_OK(cudaStreamSynchronize(b->stream));
  return MG_OK;
}

void* mg_bert_stream(mg_bert* b) { return b ? reinterpret_cast<void*>(b->stream) : nullptr; }

int mg_bert_stats(mg_bert* b, uint64_t* kernel_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!b) return fail(MG_E_ARG, "null classifier");
  if (kernel_launches) *kernel_launches = g_kernel_launches.load() - b->launches0;
  if (h2d_bytes) *h2d_bytes = b->h2d;
  if (d2h_bytes) *d2h_bytes = b->d2h;
  return MG_OK;
}

}  // extern "C"
