// Non-tensor-core kernels of the decode / prefill / classifier paths (sm_100a).
//
// Reference operations restated here (file:line in /root/reference):
//   embed + pos        api_cache.py:99          (decode steps always add pos_emb[0])
//   LayerNorm          api_cache.py:42,44,60,73 (eps 1e-5, biased variance); DistilBERT eps 1e-12
//   Linear             api_cache.py:43,45-49,85 (nn.Linear / MHA in-proj, out-proj)
//   cache append       api_cache.py:66-67       (torch.cat of the per-layer cache)
//   attention          api_cache.py:68          (softmax(q k^T / sqrt(hd)) v, NO mask)
//   sampler            api_cache.py:169-181     (/T, top-k, -1e10 mask, softmax, multinomial, EOS)
#include "kernels.cuh"
#include "sampler.cuh"

#include <math.h>

#include <type_traits>

#include "mg_engine.h"

namespace mg {

namespace {

constexpr float kLog2e = 1.4426950408889634f;

// =================================================================================================
// Embedding + LayerNorm (one warp per row; a lane owns elements lane, lane+32, ...; d <= 1024)
// =================================================================================================
constexpr int kLnWarps = 4;
constexpr int kLnMaxPerLane = 32;

template <typename TO>
__device__ __forceinline__ void ln_normalise_store(float (&v)[kLnMaxPerLane], int d, int lane, const float* __restrict__ w,
                                                   const float* __restrict__ b, float eps, TO* __restrict__ y_row,
                                                   float* __restrict__ x_out_row) {
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    if (i < d) s += v[j];
  }
  const float mean = warp_sum(s) / static_cast<float>(d);
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    if (i < d) {
      const float c = v[j] - mean;
      q += c * c;
    }
  }
  const float var = warp_sum(q) / static_cast<float>(d);    // biased, like nn.LayerNorm
  const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    if (i < d) {
      const float o = (v[j] - mean) * rstd * __ldg(w + i) + __ldg(b + i);
      y_row[i] = from_f32<TO>(o);
      if (x_out_row) x_out_row[i] = o;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
embed_ln_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ pos, const T* __restrict__ tok_emb,
                const T* __restrict__ pos_emb, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ x, T* __restrict__ y, int M, int d, float eps, int apply_ln) {
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const T* te = tok_emb + static_cast<size_t>(tok[row]) * d;
  const T* pe = pos_emb + static_cast<size_t>(pos ? pos[row] : 0) * d;
  float v[kLnMaxPerLane];
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    v[j] = (i < d) ? to_f32<T>(te[i]) + to_f32<T>(pe[i]) : 0.0f;
  }
  if (x) {
    float* xr = x + static_cast<size_t>(row) * d;
#pragma unroll
    for (int j = 0; j < kLnMaxPerLane; ++j) {
      const int i = lane + 32 * j;
      if (i < d) xr[i] = v[j];
    }
  }
  T* yr = y + static_cast<size_t>(row) * d;
  if (apply_ln) {
    ln_normalise_store<T>(v, d, lane, w, b, eps, yr, nullptr);
  } else {
#pragma unroll
    for (int j = 0; j < kLnMaxPerLane; ++j) {
      const int i = lane + 32 * j;
      if (i < d) yr[i] = from_f32<T>(v[j]);
    }
  }
}

// Recompute mode (reference generate_music/generate.py:34-35: emb(x) + pos[:T], true positions).
// Row b*Tcap + t holds token t of sequence b; rows at or beyond the sequence length are zero-filled.
template <typename T>
__global__ void __launch_bounds__(kLnWarps * 32)
nocache_embed_kernel(const int32_t* __restrict__ out_ids, int out_stride, const int32_t* __restrict__ out_len,
                     const T* __restrict__ tok_emb, const T* __restrict__ pos_emb, float* __restrict__ x,
                     T* __restrict__ y, int B, int Tcap, int d) {
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B * Tcap) return;
  const int b = row / Tcap, t = row - b * Tcap;
  const bool live = t < out_len[b];
  const T* te = tok_emb + static_cast<size_t>(live ? out_ids[static_cast<size_t>(b) * out_stride + t] : 0) * d;
  const T* pe = pos_emb + static_cast<size_t>(live ? t : 0) * d;
  for (int i = lane; i < d; i += 32) {
    const float v = live ? to_f32<T>(te[i]) + to_f32<T>(pe[i]) : 0.0f;
    x[static_cast<size_t>(row) * d + i] = v;
    y[static_cast<size_t>(row) * d + i] = from_f32<T>(v);
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, TO* __restrict__ y,
                 float* x_out, int M, int d, float eps) {
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const TI* xr = x + static_cast<size_t>(row) * d;
  float v[kLnMaxPerLane];
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    v[j] = (i < d) ? to_f32<TI>(xr[i]) : 0.0f;
  }
  ln_normalise_store<TO>(v, d, lane, w, b, eps, y + static_cast<size_t>(row) * d,
                         x_out ? x_out + static_cast<size_t>(row) * d : nullptr);
}


// bf16 -> bf16 LayerNorm with 16-byte accesses: lane owns chunks lane, lane + 32, ... of 8 elements (d = 256 * NCH).
#ifndef MG_LN_MIN_BLOCKS
#define MG_LN_MIN_BLOCKS 1
#endif
template <int NCH>
__global__ void __launch_bounds__(kLnWarps * 32, MG_LN_MIN_BLOCKS)
layernorm_bf16_vec_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                          bf16* __restrict__ y, int M, float eps) {
  constexpr int d = 256 * NCH;
  const int row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * d);
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    Chunk16<bf16>::unpack(xr[lane + 32 * j], v[j]);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[j][e];
  }
  const float mean = warp_sum(s) * (1.0f / d);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float c = v[j][e] - mean; q += c * c; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / d) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * d);
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    const int c0 = (lane + 32 * j) * 8;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
    float o[8];
    o[0] = (v[j][0] - mean) * rstd * w0.x + b0.x; o[1] = (v[j][1] - mean) * rstd * w0.y + b0.y;
    o[2] = (v[j][2] - mean) * rstd * w0.z + b0.z; o[3] = (v[j][3] - mean) * rstd * w0.w + b0.w;
    o[4] = (v[j][4] - mean) * rstd * w1.x + b1.x; o[5] = (v[j][5] - mean) * rstd * w1.y + b1.y;
    o[6] = (v[j][6] - mean) * rstd * w1.z + b1.z; o[7] = (v[j][7] - mean) * rstd * w1.w + b1.w;
    yr[lane + 32 * j] = Chunk16<bf16>::pack(o);
  }
}

// =================================================================================================
// SIMT GEMM (64x64x16 tiles, 4x4 per thread, fp32 FMA in k order) and small-M GEMV
// =================================================================================================
constexpr int kSBM = 64, kSBN = 64, kSBK = 16;

__device__ __forceinline__ void epilogue_store(const GemmEpilogue& epi, int m, int n, float acc) {
  float v = acc + (epi.bias ? __ldg(epi.bias + n) : 0.0f);
  v = apply_act(v, epi.act);
  const size_t off = static_cast<size_t>(m) * epi.ld_out + n;
  if (epi.resid_f32) v += epi.resid_f32[off];
  if (epi.resid_bf16) v += __bfloat162float(epi.resid_bf16[off]);
  if (epi.out_f32) epi.out_f32[off] = v;
  if (epi.out_bf16) epi.out_bf16[off] = __float2bfloat16_rn(v);
}

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int M, int N, int K, GemmEpilogue epi) {
  __shared__ float As[kSBK][kSBM + 4];
  __shared__ float Ws[kSBK][kSBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kSBM, n0 = blockIdx.x * kSBN;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, each a 4 x 4 micro tile
  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // loader: row 0..63, k offset 0,4,8,12
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += kSBK) {
    {
      const int m = m0 + lrow;
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < M) {
        const T* p = A + static_cast<size_t>(m) * lda + k0 + lk;
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] = to_f32<T>(p[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) As[lk + e][lrow] = a[e];
      const int n = n0 + lrow;
      float b[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < N) {
        const T* p = W + static_cast<size_t>(n) * K + k0 + lk;
#pragma unroll
        for (int e = 0; e < 4; ++e) b[e] = to_f32<T>(p[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) Ws[lk + e][lrow] = b[e];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSBK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) epilogue_store(epi, m, n, acc[i][j]);
    }
  }
}

// GEMV for M <= 8: the activations sit in shared memory as fp32; every warp streams two weight rows
// (16-byte loads, K split over lanes) and reduces with shuffles.  Weight bytes are read exactly once.
constexpr int kGemvMaxM = 8;
constexpr int kGemvWarps = 4;
constexpr int kGemvColsPerWarp = 2;

template <typename T>
__global__ void __launch_bounds__(kGemvWarps * 32)
gemv_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int M, int N, int K, GemmEpilogue epi) {
  extern __shared__ float a_s[];                     // [M][K]
  constexpr int CH = Chunk16<T>::N;
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    a_s[i] = to_f32<T>(A[static_cast<size_t>(m) * lda + k]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_base = (blockIdx.x * kGemvWarps + warp) * kGemvColsPerWarp;
  if (n_base >= N) return;
  float acc[kGemvColsPerWarp][kGemvMaxM];
#pragma unroll
  for (int c = 0; c < kGemvColsPerWarp; ++c)
#pragma unroll
    for (int m = 0; m < kGemvMaxM; ++m) acc[c][m] = 0.0f;
  const int nchunks = K / CH;
  for (int ck = lane; ck < nchunks; ck += 32) {
    float wv[kGemvColsPerWarp][CH];
#pragma unroll
    for (int c = 0; c < kGemvColsPerWarp; ++c) {
      const int n = min(n_base + c, N - 1);
      const uint4 raw = ld_stream16(W + static_cast<size_t>(n) * K + ck * CH);
      Chunk16<T>::unpack(raw, wv[c]);
    }
#pragma unroll
    for (int m = 0; m < kGemvMaxM; ++m) {
      if (m < M) {
        const float4* ap4 = reinterpret_cast<const float4*>(a_s + m * K + ck * CH);
#pragma unroll
        for (int v4 = 0; v4 < CH / 4; ++v4) {
          const float4 t = ap4[v4];
          const float av[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int c = 0; c < kGemvColsPerWarp; ++c) acc[c][m] = fmaf(av[e], wv[c][v4 * 4 + e], acc[c][m]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kGemvColsPerWarp; ++c) {
#pragma unroll
    for (int m = 0; m < kGemvMaxM; ++m) {
      if (m < M) {
        const float r = warp_sum(acc[c][m]);
        if (lane == 0 && n_base + c < N) epilogue_store(epi, m, n_base + c, r);
      }
    }
  }
}

// =================================================================================================
// KV-cache append for prefill rows
// =================================================================================================
template <typename T>
__global__ void kv_append_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ row_seq,
                                 const int32_t* __restrict__ row_pos, T* __restrict__ kcache, T* __restrict__ vcache,
                                 int M, int d, int Tmax) {
  constexpr int CH = Chunk16<T>::N;
  const int C = d / CH;
  const int total = M * 2 * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / (2 * C);
    const int rem = i - r * 2 * C;
    const int which = rem / C, c = rem - which * C;
    const uint4 v = *reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(r) * 3 * d + (1 + which) * d + c * CH);
    // cache layout [B][d/64 slices][Tmax][64]: one (sequence, slice) stream is contiguous 64-feature rows
    const int fo = c * CH, sl = fo >> 6, within = fo & 63;
    T* dst = (which ? vcache : kcache) +
             ((static_cast<size_t>(row_seq[r]) * (d >> 6) + sl) * Tmax + row_pos[r]) * 64 + within;
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

// =================================================================================================
// Split-K flash-decoding attention (cache [B][d/64][Tmax][64]: one 16-byte chunk per lane per row pass)
// =================================================================================================
constexpr int kAttnWarps = 8;

template <typename T, int CPL>
__global__ void __launch_bounds__(kAttnWarps * 32)
decode_attn_kernel(const T* __restrict__ qkv, T* __restrict__ kcache, T* __restrict__ vcache,
                   const int32_t* __restrict__ lens, const uint8_t* __restrict__ finished, T* __restrict__ out,
                   float* __restrict__ ws_o, float* __restrict__ ws_ml, uint32_t* __restrict__ counters, int d, int H,
                   int Tmax, int nsplit, float scale_log2) {
  constexpr int CH = Chunk16<T>::N;
  constexpr int R = (CPL == 1) ? 4 : (CPL == 2 ? 2 : 1);   // rows in flight per warp
  extern __shared__ float sm[];
  float* sm_m = sm;                                   // [warps][H]
  float* sm_l = sm_m + kAttnWarps * H;                // [warps][H]
  float* sm_o = sm_l + kAttnWarps * H;                // [warps][d]
  __shared__ int s_last;

  const int b = blockIdx.x / nsplit;
  const int split = blockIdx.x - b * nsplit;
  if (finished[b]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lens[b];
  const int total = len + 1;                          // cached rows + the new token
  const int per = (total + nsplit - 1) / nsplit;
  const int r0 = split * per;
  const int r1 = min(total, r0 + per);
  const int C = d / CH;                               // 16-byte chunks per row
  const int hd = d / H;
  const int cph = hd / CH;                            // chunks per head (power of two, <= 32)

  const T* qrow = qkv + static_cast<size_t>(b) * 3 * d;
  const T* knew = qrow + d;
  const T* vnew = qrow + 2 * d;
  T* kc = kcache + static_cast<size_t>(b) * Tmax * d;
  T* vc = vcache + static_cast<size_t>(b) * Tmax * d;

  // append the new token's K/V row (reference api_cache.py:66-67) -- one CTA per sequence does it
  if (split == nsplit - 1) {
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
      const int which = c / C, cc = c - which * C;
      const uint4 v = *reinterpret_cast<const uint4*>((which ? vnew : knew) + cc * CH);
      const int fo = cc * CH;
      *reinterpret_cast<uint4*>((which ? vc : kc) + (static_cast<size_t>(fo >> 6) * Tmax + len) * 64 + (fo & 63)) = v;
    }
  }

  float q[CPL][CH], acc[CPL][CH], m_run[CPL], l_run[CPL];
  bool act[CPL];
  size_t coff[CPL];                                   // offset of this lane's chunk inside the [slice][Tmax][64] block
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    act[j] = c < C;
    coff[j] = (static_cast<size_t>((c * CH) >> 6) * Tmax) * 64 + ((c * CH) & 63);
    m_run[j] = -INFINITY;
    l_run[j] = 0.0f;
#pragma unroll
    for (int e = 0; e < CH; ++e) acc[j][e] = 0.0f;
    if (act[j]) {
      Chunk16<T>::unpack(*reinterpret_cast<const uint4*>(qrow + c * CH), q[j]);
#pragma unroll
      for (int e = 0; e < CH; ++e) q[j][e] *= scale_log2;     // scores live in the log2 domain
    } else {
#pragma unroll
      for (int e = 0; e < CH; ++e) q[j][e] = 0.0f;
    }
  }

  for (int r = r0 + warp * R; r < r1; r += kAttnWarps * R) {
    uint4 kraw[R][CPL], vraw[R][CPL];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int rr = r + i;
      // the new token's row is not in the cache from this kernel's point of view (read-only path)
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if (rr < r1 && act[j]) {
          const T* kp = (rr == len) ? knew + (lane + 32 * j) * CH : kc + coff[j] + static_cast<size_t>(rr) * 64;
          const T* vp = (rr == len) ? vnew + (lane + 32 * j) * CH : vc + coff[j] + static_cast<size_t>(rr) * 64;
          kraw[i][j] = ld_stream16(kp);
          vraw[i][j] = ld_stream16(vp);
        } else {
          kraw[i][j] = make_uint4(0, 0, 0, 0);
          vraw[i][j] = make_uint4(0, 0, 0, 0);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      float s[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        float kf[CH];
        Chunk16<T>::unpack(kraw[i][j], kf);
        float p = 0.0f;
#pragma unroll
        for (int e = 0; e < CH; ++e) p = fmaf(q[j][e], kf[e], p);
        for (int o = cph >> 1; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        s[i] = (r + i < r1) ? p : -INFINITY;
      }
      float m_new = m_run[j];
#pragma unroll
      for (int i = 0; i < R; ++i) m_new = fmaxf(m_new, s[i]);
      // the first row of a pass is always valid, so m_new is finite here
      const float corr = exp2f(m_run[j] - m_new);
      float psum = 0.0f;
      float pw[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        pw[i] = exp2f(s[i] - m_new);
        psum += pw[i];
      }
      l_run[j] = l_run[j] * corr + psum;
      m_run[j] = m_new;
#pragma unroll
      for (int e = 0; e < CH; ++e) acc[j][e] *= corr;
#pragma unroll
      for (int i = 0; i < R; ++i) {
        float vf[CH];
        Chunk16<T>::unpack(vraw[i][j], vf);
#pragma unroll
        for (int e = 0; e < CH; ++e) acc[j][e] = fmaf(pw[i], vf[e], acc[j][e]);
      }
    }
  }

  // ---- merge the warps of this CTA ----
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int c = lane + 32 * j;
    if (act[j]) {
      if ((c % cph) == 0) {
        sm_m[warp * H + c / cph] = m_run[j];
        sm_l[warp * H + c / cph] = l_run[j];
      }
#pragma unroll
      for (int e = 0; e < CH; ++e) sm_o[warp * d + c * CH + e] = acc[j][e];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < d; t += blockDim.x) {
    const int h = t / hd;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, sm_m[w * H + h]);
    float L = 0.0f, o = 0.0f;
    if (M > -INFINITY) {
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) {
        const float f = exp2f(sm_m[w * H + h] - M);
        L = fmaf(sm_l[w * H + h], f, L);
        o = fmaf(sm_o[w * d + t], f, o);
      }
    }
    if (nsplit == 1) {
      out[static_cast<size_t>(b) * d + t] = from_f32<T>(o / L);
    } else {
      ws_o[(static_cast<size_t>(b) * nsplit + split) * d + t] = o;
      if ((t % hd) == 0) {
        float* ml = ws_ml + ((static_cast<size_t>(b) * nsplit + split) * H + h) * 2;
        ml[0] = M;
        ml[1] = L;
      }
    }
  }
  if (nsplit == 1) return;

  // ---- merge the splits: the last CTA of this sequence to finish does it ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(&counters[b], 1u);
    s_last = (old == static_cast<uint32_t>(nsplit - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int t = threadIdx.x; t < d; t += blockDim.x) {
    const int h = t / hd;
    float M = -INFINITY;
    for (int s2 = 0; s2 < nsplit; ++s2)
      M = fmaxf(M, __ldcg(ws_ml + ((static_cast<size_t>(b) * nsplit + s2) * H + h) * 2));
    float L = 0.0f, o = 0.0f;
    for (int s2 = 0; s2 < nsplit; ++s2) {
      const float* ml = ws_ml + ((static_cast<size_t>(b) * nsplit + s2) * H + h) * 2;
      const float ms = __ldcg(ml);
      if (ms > -INFINITY) {
        const float f = exp2f(ms - M);
        L = fmaf(__ldcg(ml + 1), f, L);
        o = fmaf(__ldcg(ws_o + (static_cast<size_t>(b) * nsplit + s2) * d + t), f, o);
      }
    }
    out[static_cast<size_t>(b) * d + t] = from_f32<T>(o / L);
  }
  if (threadIdx.x == 0) counters[b] = 0;
}

// =================================================================================================
// Non-causal attention over packed sequences (prefill, recompute mode, classifier)
// One CTA = (sequence, head, tile of 64 queries); keys stream through shared memory 64 at a time.
// =================================================================================================
constexpr int kEncWarps = 8;
constexpr int kEncQT = 32;       // queries per CTA
constexpr int kEncKT = 64;       // keys per shared-memory tile
constexpr int kEncQPW = kEncQT / kEncWarps;
constexpr int kEncMaxHd = 64;

template <typename T>
__global__ void __launch_bounds__(kEncWarps * 32)
encoder_attn_kernel(const T* __restrict__ qkv, const int32_t* __restrict__ seq_start, const int32_t* __restrict__ seq_len,
                    const uint8_t* __restrict__ key_mask, T* __restrict__ out, int d, int H, float scale_log2) {
  __shared__ float Ks[kEncKT][kEncMaxHd + 1];
  __shared__ float Vs[kEncKT][kEncMaxHd + 1];
  __shared__ float Qs[kEncQT][kEncMaxHd];
  __shared__ float Kbias[kEncKT];                    // 0 or -inf (padding keys / beyond the sequence)

  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int len = seq_len[b];
  const int q0 = blockIdx.y * kEncQT;
  if (q0 >= len) return;
  const int start = seq_start[b];
  const int hd = d / H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = static_cast<size_t>(3) * d;

  for (int i = threadIdx.x; i < kEncQT * hd; i += blockDim.x) {
    const int qi = i / hd, j = i - qi * hd;
    const int qrow = q0 + qi;
    Qs[qi][j] = (qrow < len) ? to_f32<T>(qkv[(start + qrow) * ld + h * hd + j]) * scale_log2 : 0.0f;
  }

  float m_run[kEncQPW], l_run[kEncQPW], o_lo[kEncQPW], o_hi[kEncQPW];
#pragma unroll
  for (int qi = 0; qi < kEncQPW; ++qi) {
    m_run[qi] = -INFINITY;
    l_run[qi] = 0.0f;
    o_lo[qi] = 0.0f;
    o_hi[qi] = 0.0f;
  }

  for (int k0 = 0; k0 < len; k0 += kEncKT) {
    __syncthreads();                                   // previous tile fully consumed (and Qs written)
    for (int i = threadIdx.x; i < kEncKT * hd; i += blockDim.x) {
      const int ki = i / hd, j = i - ki * hd;
      const int krow = k0 + ki;
      float kv = 0.0f, vv = 0.0f;
      if (krow < len) {
        const T* p = qkv + (start + krow) * ld + h * hd + j;
        kv = to_f32<T>(p[d]);
        vv = to_f32<T>(p[2 * d]);
      }
      Ks[ki][j] = kv;
      Vs[ki][j] = vv;
    }
    for (int ki = threadIdx.x; ki < kEncKT; ki += blockDim.x) {
      const int krow = k0 + ki;
      const bool ok = krow < len && (key_mask == nullptr || key_mask[start + krow] != 0);
      Kbias[ki] = ok ? 0.0f : -INFINITY;
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < kEncQPW; ++qi) {
      const int ql = warp * kEncQPW + qi;
      if (q0 + ql >= len) break;                       // warp-uniform
      float s0 = 0.0f, s1 = 0.0f;
      for (int j = 0; j < hd; ++j) {
        const float qv = Qs[ql][j];
        s0 = fmaf(qv, Ks[lane][j], s0);
        s1 = fmaf(qv, Ks[lane + 32][j], s1);
      }
      s0 += Kbias[lane];
      s1 += Kbias[lane + 32];
      const float tile_max = warp_max(fmaxf(s0, s1));
      const float m_new = fmaxf(m_run[qi], tile_max);
      if (m_new == -INFINITY) continue;               // every key so far is masked
      const float corr = exp2f(m_run[qi] - m_new);
      const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
      l_run[qi] = l_run[qi] * corr + warp_sum(p0 + p1);
      m_run[qi] = m_new;
      float a_lo = o_lo[qi] * corr, a_hi = o_hi[qi] * corr;
      const bool two = hd > 32;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const float pk0 = __shfl_sync(0xffffffffu, p0, k);
        const float pk1 = __shfl_sync(0xffffffffu, p1, k);
        if (lane < hd) {
          a_lo = fmaf(pk0, Vs[k][lane], a_lo);
          a_lo = fmaf(pk1, Vs[k + 32][lane], a_lo);
        }
        if (two) {
          a_hi = fmaf(pk0, Vs[k][lane + 32], a_hi);
          a_hi = fmaf(pk1, Vs[k + 32][lane + 32], a_hi);
        }
      }
      o_lo[qi] = a_lo;
      o_hi[qi] = a_hi;
    }
  }
#pragma unroll
  for (int qi = 0; qi < kEncQPW; ++qi) {
    const int qrow = q0 + warp * kEncQPW + qi;
    if (qrow >= len) break;
    const float inv = 1.0f / l_run[qi];
    T* op = out + static_cast<size_t>(start + qrow) * d + h * hd;
    if (lane < hd) op[lane] = from_f32<T>(o_lo[qi] * inv);
    if (hd > 32) op[lane + 32] = from_f32<T>(o_hi[qi] * inv);
  }
}


// -------------------------------------------------------------------------------------------------
// Tensor-core variant for bf16 (classifier, bf16 prefill / recompute): one CTA = (sequence, head, 64
// queries), 4 warps x 16 queries, keys in tiles of 64; S = Q K^T and O += P V are mma.sync m16n8k16
// (K via ldmatrix, V via ldmatrix.trans), online softmax in the log2 domain on the accumulator
// fragments, P re-used from registers as the A operand of the second product.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int HD>
__global__ void __launch_bounds__(128)
encoder_attn_tc_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ seq_start, const int32_t* __restrict__ seq_len,
                       const uint8_t* __restrict__ key_mask, bf16* __restrict__ out, int d, int H, float scale_log2) {
  constexpr int P = HD + 8;                          // padded pitch: ldmatrix rows land in distinct banks
  constexpr int CPR = HD / 8;                        // 16-byte chunks per row
  __shared__ __align__(16) bf16 Qs[64 * P];
  __shared__ __align__(16) bf16 Ks[64 * P];
  __shared__ __align__(16) bf16 Vs[64 * P];
  __shared__ float Kbias[64];

  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int len = seq_len[b];
  const int q0 = blockIdx.y * 64;
  if (q0 >= len) return;
  const int start = seq_start[b];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = static_cast<size_t>(3) * d;
  const bf16* base = qkv + static_cast<size_t>(start) * ld + h * HD;

  // Q tile and the first K / V tile are requested together: one global round trip instead of two (T <= 64 in the classifier,
  // so this is the only tile and the CTA's lifetime is dominated by that latency)
  {
    constexpr int NI = 64 * CPR / 128;
    uint4 qv[NI], kv[NI], vv[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int i = threadIdx.x + j * 128, r = i / CPR, c = i - r * CPR;
      qv[j] = kv[j] = vv[j] = make_uint4(0, 0, 0, 0);
      if (q0 + r < len) qv[j] = *reinterpret_cast<const uint4*>(base + (q0 + r) * ld + c * 8);
      if (r < len) {
        const bf16* pk = base + r * ld + c * 8;
        kv[j] = *reinterpret_cast<const uint4*>(pk + d);
        vv[j] = *reinterpret_cast<const uint4*>(pk + 2 * d);
      }
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int i = threadIdx.x + j * 128, r = i / CPR, c = i - r * CPR;
      *reinterpret_cast<uint4*>(Qs + r * P + c * 8) = qv[j];
      *reinterpret_cast<uint4*>(Ks + r * P + c * 8) = kv[j];
      *reinterpret_cast<uint4*>(Vs + r * P + c * 8) = vv[j];
    }
    if (threadIdx.x < 64) {
      const int kr = threadIdx.x;
      Kbias[threadIdx.x] = (kr < len && (key_mask == nullptr || key_mask[start + kr] != 0)) ? 0.f : -INFINITY;
    }
  }
  __syncthreads();
  const int mi = lane >> 3;
  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks)
    ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(Qs + (warp * 16 + (lane & 7) + (mi & 1) * 8) * P + ks * 16 + (mi >> 1) * 8)), qf[ks]);

  float o[HD / 8][4];
#pragma unroll
  for (int t = 0; t < HD / 8; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[t][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};   // rows lane/4 and lane/4 + 8 (per-thread partial sums)

  for (int k0 = 0; k0 < len; k0 += 64) {
    if (k0 > 0) {
      __syncthreads();
      for (int i = threadIdx.x; i < 64 * CPR; i += 128) {
        const int r = i / CPR, c = i - r * CPR;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (k0 + r < len) {
          const bf16* pk = base + (k0 + r) * ld + c * 8;
          kv = *reinterpret_cast<const uint4*>(pk + d);
          vv = *reinterpret_cast<const uint4*>(pk + 2 * d);
        }
        *reinterpret_cast<uint4*>(Ks + r * P + c * 8) = kv;
        *reinterpret_cast<uint4*>(Vs + r * P + c * 8) = vv;
      }
      if (threadIdx.x < 64) {
        const int kr = k0 + threadIdx.x;
        Kbias[threadIdx.x] = (kr < len && (key_mask == nullptr || key_mask[start + kr] != 0)) ? 0.f : -INFINITY;
      }
      __syncthreads();
    }

    float s[8][4];
#pragma unroll
    for (int t = 0; t < 8; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[t][e] = 0.f;
#pragma unroll
    for (int nt2 = 0; nt2 < 4; ++nt2) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t kf[4];
        ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(Ks + (nt2 * 16 + (lane & 7) + (mi >> 1) * 8) * P + ks * 16 + (mi & 1) * 8)), kf);
        mma_16816(s[2 * nt2], qf[ks], kf[0], kf[1]);
        mma_16816(s[2 * nt2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float b0 = Kbias[t * 8 + (lane & 3) * 2], b1 = Kbias[t * 8 + (lane & 3) * 2 + 1];
      s[t][0] = s[t][0] * scale_log2 + b0;
      s[t][1] = s[t][1] * scale_log2 + b1;
      s[t][2] = s[t][2] * scale_log2 + b0;
      s[t][3] = s[t][3] * scale_log2 + b1;
      mx[0] = fmaxf(mx[0], fmaxf(s[t][0], s[t][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[t][2], s[t][3]));
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      msafe[r] = (m_new == -INFINITY) ? 0.f : m_new;       // every key so far masked: keep everything at zero
      corr[r] = fast_exp2(m_run[r] - msafe[r]);
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int t = 0; t < HD / 8; ++t) {
      o[t][0] *= corr[0]; o[t][1] *= corr[0];
      o[t][2] *= corr[1]; o[t][3] *= corr[1];
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      s[t][0] = fast_exp2(s[t][0] - msafe[0]); s[t][1] = fast_exp2(s[t][1] - msafe[0]);
      s[t][2] = fast_exp2(s[t][2] - msafe[1]); s[t][3] = fast_exp2(s[t][3] - msafe[1]);
      l_run[0] += s[t][0] + s[t][1];
      l_run[1] += s[t][2] + s[t][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack2_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack2_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack2_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dn2 = 0; dn2 < HD / 16; ++dn2) {
        uint32_t vf[4];
        ldsm_x4_trans(static_cast<uint32_t>(__cvta_generic_to_shared(Vs + (kk * 16 + (lane & 7) + (mi & 1) * 8) * P + dn2 * 16 + (mi >> 1) * 8)), vf);
        mma_16816(o[2 * dn2], pa, vf[0], vf[1]);
        mma_16816(o[2 * dn2 + 1], pa, vf[2], vf[3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qrow = q0 + warp * 16 + (lane >> 2) + r * 8;
    if (qrow < len) {
      const float inv = 1.0f / l_run[r];
      bf16* op = out + static_cast<size_t>(start + qrow) * d + h * HD + (lane & 3) * 2;
#pragma unroll
      for (int t = 0; t < HD / 8; ++t)
        *reinterpret_cast<uint32_t*>(op + t * 8) = pack2_bf16(o[t][2 * r] * inv, o[t][2 * r + 1] * inv);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Pipelined variant for sequences of at most 64 tokens (the classifier's shape, BASELINE config 2: 256 texts x 64 tokens x 12
// heads = 3072 independent 64 x 64 attentions).  The one-tile-per-CTA kernel above spends its life waiting for ONE global round
// trip (24 KB per CTA: 3.2 TB/s over the pass); here a persistent CTA walks over (sequence, head) items with the Q / K / V tiles
// of item i + 1 in flight (cp.async, 16-byte, zero-fill for rows past the end) while item i is computed.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(128)
encoder_attn_tc_pipe_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ seq_start, const int32_t* __restrict__ seq_len,
                            const uint8_t* __restrict__ key_mask, bf16* __restrict__ out, int d, int H, int n_items, float scale_log2) {
  constexpr int P = HD + 8;                          // padded pitch: ldmatrix rows land in distinct banks
  constexpr int CPR = HD / 8;                        // 16-byte chunks per row
  constexpr int TILE = 64 * P;                       // elements of one staged matrix
  extern __shared__ __align__(16) uint8_t attn_smem[];
  bf16* stage0 = reinterpret_cast<bf16*>(attn_smem);                 // [2 stages][Q | K | V][64][P]
  float* kbias = reinterpret_cast<float*>(attn_smem + 2 * 3 * TILE * sizeof(bf16));   // [2][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = static_cast<size_t>(3) * d;

  auto issue = [&](int item, int st) {
    const int b = item / H, h = item - b * H;
    const int len = seq_len[b], start = seq_start[b];
    const bf16* base = qkv + static_cast<size_t>(start) * ld + h * HD;
    bf16* Qs = stage0 + st * 3 * TILE;
    const uint32_t q_addr = static_cast<uint32_t>(__cvta_generic_to_shared(Qs));
#pragma unroll
    for (int j = 0; j < 64 * CPR / 128; ++j) {
      const int i = threadIdx.x + j * 128, r = i / CPR, c = i - r * CPR;
      const bool ok = r < len;
      const bf16* src = base + static_cast<size_t>(ok ? r : 0) * ld + c * 8;
      const uint32_t dst = q_addr + (r * P + c * 8) * 2;
      cp_async16_zfill(dst, src, ok);
      cp_async16_zfill(dst + TILE * 2, src + d, ok);
      cp_async16_zfill(dst + 2 * TILE * 2, src + 2 * d, ok);
    }
    if (threadIdx.x < 64) {
      const int kr = threadIdx.x;
      kbias[st * 64 + kr] = (kr < len && (key_mask == nullptr || key_mask[start + kr] != 0)) ? 0.f : -INFINITY;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int item = blockIdx.x, st = 0;
  if (item < n_items) issue(item, 0);
  for (; item < n_items; item += gridDim.x, st ^= 1) {
    const int next = item + gridDim.x;
    if (next < n_items) {
      issue(next, st ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int b = item / H, h = item - b * H;
    const int len = seq_len[b], start = seq_start[b];
    const bf16* Qs = stage0 + st * 3 * TILE;
    const bf16* Ks = Qs + TILE;
    const bf16* Vs = Ks + TILE;
    const float* Kb = kbias + st * 64;
    const int mi = lane >> 3;
    if (warp * 16 < len) {                            // warp-uniform: this warp's 16 query rows exist
      uint32_t qf[HD / 16][4];
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(Qs + (warp * 16 + (lane & 7) + (mi & 1) * 8) * P + ks * 16 + (mi >> 1) * 8)), qf[ks]);
      float s[8][4];
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[t][e] = 0.f;
#pragma unroll
      for (int nt2 = 0; nt2 < 4; ++nt2) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          uint32_t kf[4];
          ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(Ks + (nt2 * 16 + (lane & 7) + (mi >> 1) * 8) * P + ks * 16 + (mi & 1) * 8)), kf);
          mma_16816(s[2 * nt2], qf[ks], kf[0], kf[1]);
          mma_16816(s[2 * nt2 + 1], qf[ks], kf[2], kf[3]);
        }
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float b0 = Kb[t * 8 + (lane & 3) * 2], b1 = Kb[t * 8 + (lane & 3) * 2 + 1];
        s[t][0] = s[t][0] * scale_log2 + b0;
        s[t][1] = s[t][1] * scale_log2 + b1;
        s[t][2] = s[t][2] * scale_log2 + b0;
        s[t][3] = s[t][3] * scale_log2 + b1;
        mx[0] = fmaxf(mx[0], fmaxf(s[t][0], s[t][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[t][2], s[t][3]));
      }
      float l_run[2] = {0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        mx[r] = (mx[r] == -INFINITY) ? 0.f : mx[r];      // every key masked: keep everything at zero
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        s[t][0] = fast_exp2(s[t][0] - mx[0]); s[t][1] = fast_exp2(s[t][1] - mx[0]);
        s[t][2] = fast_exp2(s[t][2] - mx[1]); s[t][3] = fast_exp2(s[t][3] - mx[1]);
        l_run[0] += s[t][0] + s[t][1];
        l_run[1] += s[t][2] + s[t][3];
      }
      float o[HD / 8][4];
#pragma unroll
      for (int t = 0; t < HD / 8; ++t)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[t][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack2_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack2_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack2_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack2_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dn2 = 0; dn2 < HD / 16; ++dn2) {
          uint32_t vf[4];
          ldsm_x4_trans(static_cast<uint32_t>(__cvta_generic_to_shared(Vs + (kk * 16 + (lane & 7) + (mi & 1) * 8) * P + dn2 * 16 + (mi >> 1) * 8)), vf);
          mma_16816(o[2 * dn2], pa, vf[0], vf[1]);
          mma_16816(o[2 * dn2 + 1], pa, vf[2], vf[3]);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int qrow = warp * 16 + (lane >> 2) + r * 8;
        if (qrow < len) {
          const float inv = 1.0f / l_run[r];
          bf16* op = out + static_cast<size_t>(start + qrow) * d + h * HD + (lane & 3) * 2;
#pragma unroll
          for (int t = 0; t < HD / 8; ++t)
            *reinterpret_cast<uint32_t*>(op + t * 8) = pack2_bf16(o[t][2 * r] * inv, o[t][2 * r + 1] * inv);
        }
      }
    }
    __syncthreads();                                  // this stage is refilled by the NEXT iteration's issue()
  }
}

// Sampler device code (float_key, Philox, block scans, sample_row): sampler.cuh

__global__ void __launch_bounds__(kSampleThreads)
sample_step_kernel(const float* __restrict__ logits, int ld, int V, const SampleParams* __restrict__ sp, DecodeState st) {
  extern __shared__ float vals[];
  __shared__ SampleSmem ss;
  const int b = blockIdx.x;
  if (st.finished[b]) return;
  const SampleParams p = *sp;
  const uint32_t step = static_cast<uint32_t>(st.n_new[b]);
  const int tok = sample_row(logits + static_cast<size_t>(b) * ld, V, p.temperature, p.top_k, p.seed,
                             p.seq_base + static_cast<uint64_t>(st.seq_idx ? st.seq_idx[b] : b), step, vals, ss);
  if (threadIdx.x == 0) {
    const int pos = st.out_len[b];
    st.out_ids[static_cast<size_t>(b) * st.out_stride + pos] = tok;     // api_cache.py:179
    if (st.step_ns && b == 0) {                                          // per-token latency read-out (mg_last_step_times)
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      st.step_ns[step] = t;
    }
    st.out_len[b] = pos + 1;
    st.cur_tok[b] = tok;
    st.lens[b] += 1;
    const int n = static_cast<int>(step) + 1;
    st.n_new[b] = n;
    if (tok == p.eos_id || n >= st.max_new[b]) st.finished[b] = 1;       // api_cache.py:181
  }
}

__global__ void __launch_bounds__(kSampleThreads)
sample_rows_kernel(const float* __restrict__ logits, int ld, int V, const SampleParams* __restrict__ sp, uint32_t step,
                   int32_t* __restrict__ out) {
  extern __shared__ float vals[];
  __shared__ SampleSmem ss;
  const SampleParams p = *sp;
  const int tok = sample_row(logits + static_cast<size_t>(blockIdx.x) * ld, V, p.temperature, p.top_k, p.seed,
                             p.seq_base + blockIdx.x, step, vals, ss);
  if (threadIdx.x == 0) out[blockIdx.x] = tok;
}

__global__ void decode_init_kernel(const int32_t* __restrict__ prompt_ids, const int32_t* __restrict__ offsets,
                                   DecodeState st, int B) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const int o0 = offsets[b], n = offsets[b + 1] - o0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) st.out_ids[static_cast<size_t>(b) * st.out_stride + i] = prompt_ids[o0 + i];
  if (threadIdx.x == 0) {
    st.out_len[b] = n;
    st.cur_tok[b] = prompt_ids[o0 + n - 1];       // the last prompt token is fed again (api_cache.py:167)
    st.lens[b] = n;
    st.n_new[b] = 0;
    st.finished[b] = st.max_new[b] <= 0 ? 1 : 0;
  }
}

// continuous batching: the prompts of n newly admitted requests -> decode state of the slots they were given
__global__ void slot_init_kernel(const int32_t* __restrict__ prompt_ids, const int32_t* __restrict__ offsets,
                                 const int32_t* __restrict__ slots, const int32_t* __restrict__ max_new,
                                 const int32_t* __restrict__ seq_idx, int32_t* __restrict__ seq_idx_out, DecodeState st, int n) {
  const int j = blockIdx.x;
  if (j >= n) return;
  const int b = slots[j], o0 = offsets[j], len = offsets[j + 1] - o0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) st.out_ids[static_cast<size_t>(b) * st.out_stride + i] = prompt_ids[o0 + i];
  if (threadIdx.x == 0) {
    st.out_len[b] = len;
    st.cur_tok[b] = prompt_ids[o0 + len - 1];     // the last prompt token is fed again (api_cache.py:167)
    st.lens[b] = len;
    st.n_new[b] = 0;
    st.max_new[b] = max_new[j];
    seq_idx_out[b] = seq_idx[j];
    st.finished[b] = max_new[j] <= 0 ? 1 : 0;
  }
}

__global__ void force_next_kernel(const int32_t* __restrict__ forced, int stride, int col, DecodeState st, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  st.cur_tok[b] = forced[static_cast<size_t>(b) * stride + col];
  st.lens[b] += 1;
}

__global__ void count_active_kernel(const uint8_t* __restrict__ finished, int B, int32_t* __restrict__ active) {
  int c = 0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) c += finished[i] ? 0 : 1;
  c = static_cast<int>(warp_sum(static_cast<float>(c)));
  __shared__ int s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += s[w];
    *active = t;
  }
}

template <typename TI, typename TO>
__global__ void gather_rows_kernel(const TI* __restrict__ src, const int32_t* __restrict__ rows, TO* __restrict__ out,
                                   int B, int d) {
  const int b = blockIdx.x;
  const TI* s = src + static_cast<size_t>(rows[b]) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) out[static_cast<size_t>(b) * d + i] = from_f32<TO>(to_f32<TI>(s[i]));
}

template <typename T>
__global__ void convert_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = from_f32<T>(src[i]);
}

__global__ void argmax_rows_kernel(const float* __restrict__ logits, int N, int C, int32_t* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < C; i += 32) {
    const float v = logits[static_cast<size_t>(row) * C + i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) out[row] = bi;
}

int bad_shape(const char* what) {
  set_last_error(std::string("unsupported shape: ") + what);
  return MG_E_SHAPE;
}

}  // namespace

// =================================================================================================
// launch wrappers
// =================================================================================================
template <typename T>
int launch_embed_ln(cudaStream_t s, const int32_t* tok, const int32_t* pos, const T* tok_emb, const T* pos_emb,
                    const float* w, const float* b, float* x, T* y, int M, int d, float eps, bool apply_ln) {
  if (M <= 0) return MG_OK;
  if (d > 32 * kLnMaxPerLane) return bad_shape("d_model > 1024");
  embed_ln_kernel<T><<<ceil_div(M, kLnWarps), kLnWarps * 32, 0, s>>>(tok, pos, tok_emb, pos_emb, w, b, x, y, M, d, eps,
                                                                     apply_ln ? 1 : 0);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_embed_ln<float>(cudaStream_t, const int32_t*, const int32_t*, const float*, const float*, const float*,
                                    const float*, float*, float*, int, int, float, bool);
template int launch_embed_ln<bf16>(cudaStream_t, const int32_t*, const int32_t*, const bf16*, const bf16*, const float*,
                                   const float*, float*, bf16*, int, int, float, bool);

template <typename TI, typename TO>
int launch_layernorm(cudaStream_t s, const TI* x, const float* w, const float* b, TO* y, float* x_out, int M, int d,
                     float eps) {
  if (M <= 0) return MG_OK;
  if (d > 32 * kLnMaxPerLane) return bad_shape("d_model > 1024");
  if (std::is_same<TI, bf16>::value && std::is_same<TO, bf16>::value && x_out == nullptr && d % 256 == 0) {
    const bf16* xi = reinterpret_cast<const bf16*>(x);
    bf16* yo = reinterpret_cast<bf16*>(y);
    const int blocks = ceil_div(M, kLnWarps);
    switch (d / 256) {
      case 1: layernorm_bf16_vec_kernel<1><<<blocks, kLnWarps * 32, 0, s>>>(xi, w, b, yo, M, eps); break;
      case 2: layernorm_bf16_vec_kernel<2><<<blocks, kLnWarps * 32, 0, s>>>(xi, w, b, yo, M, eps); break;
      case 3: layernorm_bf16_vec_kernel<3><<<blocks, kLnWarps * 32, 0, s>>>(xi, w, b, yo, M, eps); break;
      default: layernorm_bf16_vec_kernel<4><<<blocks, kLnWarps * 32, 0, s>>>(xi, w, b, yo, M, eps); break;
    }
    MG_LAUNCH_CHECK();
    return MG_OK;
  }
  layernorm_kernel<TI, TO><<<ceil_div(M, kLnWarps), kLnWarps * 32, 0, s>>>(x, w, b, y, x_out, M, d, eps);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_layernorm<float, float>(cudaStream_t, const float*, const float*, const float*, float*, float*, int, int, float);
template int launch_layernorm<float, bf16>(cudaStream_t, const float*, const float*, const float*, bf16*, float*, int, int, float);
template int launch_layernorm<bf16, bf16>(cudaStream_t, const bf16*, const float*, const float*, bf16*, float*, int, int, float);

template <typename T>
int launch_gemm_simt(cudaStream_t s, const T* A, int lda, const T* W, int M, int N, int K, const GemmEpilogue& epi) {
  if (M <= 0 || N <= 0) return MG_OK;
  constexpr int CH = Chunk16<T>::N;
  if (K <= 0 || K % kSBK != 0 || K % CH != 0) return bad_shape("GEMM K must be a multiple of 16");
  const size_t gemv_smem = static_cast<size_t>(M) * K * sizeof(float);
  if (M <= kGemvMaxM && gemv_smem <= 48 * 1024 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
    const int cols_per_cta = kGemvWarps * kGemvColsPerWarp;
    gemv_kernel<T><<<ceil_div(N, cols_per_cta), kGemvWarps * 32, gemv_smem, s>>>(A, lda, W, M, N, K, epi);
  } else {
    dim3 grid(ceil_div(N, kSBN), ceil_div(M, kSBM));
    gemm_simt_kernel<T><<<grid, 256, 0, s>>>(A, lda, W, M, N, K, epi);
  }
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_gemm_simt<float>(cudaStream_t, const float*, int, const float*, int, int, int, const GemmEpilogue&);
template int launch_gemm_simt<bf16>(cudaStream_t, const bf16*, int, const bf16*, int, int, int, const GemmEpilogue&);

template <typename T>
int launch_kv_append(cudaStream_t s, const T* qkv, const int32_t* row_seq, const int32_t* row_pos, T* kcache, T* vcache,
                     int M, int d, int Tmax) {
  if (M <= 0) return MG_OK;
  constexpr int CH = Chunk16<T>::N;
  if (d % CH != 0) return bad_shape("d_model must be a multiple of 8");
  const int total = M * 2 * (d / CH);
  const int blocks = min(ceil_div(total, 256), 148 * 8);
  kv_append_kernel<T><<<blocks, 256, 0, s>>>(qkv, row_seq, row_pos, kcache, vcache, M, d, Tmax);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_kv_append<float>(cudaStream_t, const float*, const int32_t*, const int32_t*, float*, float*, int, int, int);
template int launch_kv_append<bf16>(cudaStream_t, const bf16*, const int32_t*, const int32_t*, bf16*, bf16*, int, int, int);

template <typename T>
int launch_decode_attn(cudaStream_t s, const T* qkv, T* kcache, T* vcache, const int32_t* lens, const uint8_t* finished,
                       T* out, float* ws_o, float* ws_ml, uint32_t* counters, int B, int d, int H, int Tmax, int nsplit) {
  if (B <= 0) return MG_OK;
  constexpr int CH = Chunk16<T>::N;
  if (H <= 0 || d % H != 0) return bad_shape("d_model not divisible by n_head");
  const int hd = d / H;
  if (d % CH != 0 || hd % CH != 0) return bad_shape("head_dim must be a multiple of 8");
  const int cph = hd / CH, C = d / CH;
  if ((cph & (cph - 1)) != 0 || cph > 32) return bad_shape("head_dim / 16-byte chunk must be a power of two <= 32");
  const int cpl = ceil_div(C, 32);
  const float scale_log2 = kLog2e / sqrtf(static_cast<float>(hd));
  const size_t smem = static_cast<size_t>(kAttnWarps) * (2 * H + d) * sizeof(float);
  dim3 grid(B * nsplit);
#define MG_DA(CPL)                                                                                                     \
  decode_attn_kernel<T, CPL><<<grid, kAttnWarps * 32, smem, s>>>(qkv, kcache, vcache, lens, finished, out, ws_o, ws_ml, \
                                                                 counters, d, H, Tmax, nsplit, scale_log2)
  if (cpl == 1) MG_DA(1);
  else if (cpl == 2) MG_DA(2);
  else if (cpl <= 4) MG_DA(4);
  else return bad_shape("d_model too wide for the decode attention kernel");
#undef MG_DA
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_decode_attn<float>(cudaStream_t, const float*, float*, float*, const int32_t*, const uint8_t*, float*,
                                       float*, float*, uint32_t*, int, int, int, int, int);
template int launch_decode_attn<bf16>(cudaStream_t, const bf16*, bf16*, bf16*, const int32_t*, const uint8_t*, bf16*,
                                      float*, float*, uint32_t*, int, int, int, int, int);

template <typename T>
int launch_encoder_attn(cudaStream_t s, const T* qkv, const int32_t* seq_start, const int32_t* seq_len,
                        const uint8_t* key_mask, T* out, int B, int d, int H, int max_len) {
  if (B <= 0 || max_len <= 0) return MG_OK;
  if (H <= 0 || d % H != 0) return bad_shape("d_model not divisible by n_head");
  const int hd = d / H;
  if (hd > kEncMaxHd) return bad_shape("head_dim > 64");
  const float scale_log2 = kLog2e / sqrtf(static_cast<float>(hd));
  if (std::is_same<T, bf16>::value && (hd == 32 || hd == 64) && d % 8 == 0) {
    dim3 grid_tc(B * H, ceil_div(max_len, 64));
    const bf16* q16 = reinterpret_cast<const bf16*>(qkv);
    bf16* o16 = reinterpret_cast<bf16*>(out);
    static const bool pipe_off = std::getenv("MG_ATTN_PIPE") && std::atoi(std::getenv("MG_ATTN_PIPE")) == 0;
    if (max_len <= 64 && B * H >= 1024 && !pipe_off) {
      // many short sequences (the classifier): persistent CTAs, next item's tiles in flight while this one is computed
      static int sms = 0;
      if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
      const int per_sm = hd == 64 ? 4 : 6;
      const size_t smem = static_cast<size_t>(2) * 3 * 64 * (hd + 8) * sizeof(bf16) + 2 * 64 * sizeof(float);
      const int grid_p = std::min(B * H, sms * per_sm);
      if (hd == 64) encoder_attn_tc_pipe_kernel<64><<<grid_p, 128, smem, s>>>(q16, seq_start, seq_len, key_mask, o16, d, H, B * H, scale_log2);
      else encoder_attn_tc_pipe_kernel<32><<<grid_p, 128, smem, s>>>(q16, seq_start, seq_len, key_mask, o16, d, H, B * H, scale_log2);
      MG_LAUNCH_CHECK();
      return MG_OK;
    }
    if (hd == 64) encoder_attn_tc_kernel<64><<<grid_tc, 128, 0, s>>>(q16, seq_start, seq_len, key_mask, o16, d, H, scale_log2);
    else encoder_attn_tc_kernel<32><<<grid_tc, 128, 0, s>>>(q16, seq_start, seq_len, key_mask, o16, d, H, scale_log2);
    MG_LAUNCH_CHECK();
    return MG_OK;
  }
  dim3 grid(B * H, ceil_div(max_len, kEncQT));
  encoder_attn_kernel<T><<<grid, kEncWarps * 32, 0, s>>>(qkv, seq_start, seq_len, key_mask, out, d, H, scale_log2);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_encoder_attn<float>(cudaStream_t, const float*, const int32_t*, const int32_t*, const uint8_t*, float*,
                                        int, int, int, int);
template int launch_encoder_attn<bf16>(cudaStream_t, const bf16*, const int32_t*, const int32_t*, const uint8_t*, bf16*,
                                       int, int, int, int);

static int sample_smem_ok(int V, size_t* bytes) {
  *bytes = static_cast<size_t>(V) * sizeof(float);
  if (*bytes > 200 * 1024) return bad_shape("vocabulary too large for the sampler's shared-memory row");
  return MG_OK;
}

int kernels_init() {
  MG_CUDA_OK(cudaFuncSetAttribute(encoder_attn_tc_pipe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  MG_CUDA_OK(cudaFuncSetAttribute(encoder_attn_tc_pipe_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  MG_CUDA_OK(cudaFuncSetAttribute(sample_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  MG_CUDA_OK(cudaFuncSetAttribute(sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return MG_OK;
}

int launch_sample_step(cudaStream_t s, const float* logits, int ld, int V, const SampleParams* sp, DecodeState st, int B) {
  if (B <= 0) return MG_OK;
  size_t smem;
  MG_TRY(sample_smem_ok(V, &smem));
  sample_step_kernel<<<B, kSampleThreads, smem, s>>>(logits, ld, V, sp, st);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int launch_sample_rows(cudaStream_t s, const float* logits, int ld, int rows, int V, const SampleParams* sp,
                       uint32_t step, int32_t* out) {
  if (rows <= 0) return MG_OK;
  size_t smem;
  MG_TRY(sample_smem_ok(V, &smem));
  sample_rows_kernel<<<rows, kSampleThreads, smem, s>>>(logits, ld, V, sp, step, out);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

template <typename T>
int launch_nocache_embed(cudaStream_t s, const int32_t* out_ids, int out_stride, const int32_t* out_len, const T* tok_emb,
                         const T* pos_emb, float* x, T* y, int B, int Tcap, int d) {
  if (B <= 0) return MG_OK;
  if (d > 32 * kLnMaxPerLane) return bad_shape("d_model > 1024");
  nocache_embed_kernel<T><<<ceil_div(B * Tcap, kLnWarps), kLnWarps * 32, 0, s>>>(out_ids, out_stride, out_len, tok_emb,
                                                                               pos_emb, x, y, B, Tcap, d);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_nocache_embed<float>(cudaStream_t, const int32_t*, int, const int32_t*, const float*, const float*,
                                         float*, float*, int, int, int);
template int launch_nocache_embed<bf16>(cudaStream_t, const int32_t*, int, const int32_t*, const bf16*, const bf16*,
                                        float*, bf16*, int, int, int);

int launch_decode_init(cudaStream_t s, const int32_t* prompt_ids, const int32_t* offsets, DecodeState st, int B) {
  if (B <= 0) return MG_OK;
  decode_init_kernel<<<B, 64, 0, s>>>(prompt_ids, offsets, st, B);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int launch_slot_init(cudaStream_t s, const int32_t* prompt_ids, const int32_t* offsets, const int32_t* slots, const int32_t* max_new,
                     const int32_t* seq_idx, int32_t* seq_idx_out, DecodeState st, int n) {
  if (n <= 0) return MG_OK;
  slot_init_kernel<<<n, 64, 0, s>>>(prompt_ids, offsets, slots, max_new, seq_idx, seq_idx_out, st, n);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int launch_force_next(cudaStream_t s, const int32_t* forced, int stride, int col, DecodeState st, int B) {
  if (B <= 0) return MG_OK;
  force_next_kernel<<<ceil_div(B, 128), 128, 0, s>>>(forced, stride, col, st, B);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

int launch_count_active(cudaStream_t s, const uint8_t* finished, int B, int32_t* active) {
  count_active_kernel<<<1, 256, 0, s>>>(finished, B, active);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

template <typename TI, typename TO>
int launch_gather_rows(cudaStream_t s, const TI* src, const int32_t* rows, TO* out, int B, int d) {
  if (B <= 0) return MG_OK;
  gather_rows_kernel<TI, TO><<<B, 128, 0, s>>>(src, rows, out, B, d);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_gather_rows<float, float>(cudaStream_t, const float*, const int32_t*, float*, int, int);
template int launch_gather_rows<float, bf16>(cudaStream_t, const float*, const int32_t*, bf16*, int, int);
template int launch_gather_rows<bf16, bf16>(cudaStream_t, const bf16*, const int32_t*, bf16*, int, int);

template <typename T>
int launch_convert(cudaStream_t s, const float* src, T* dst, size_t n) {
  if (n == 0) return MG_OK;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 148 * 16));
  convert_kernel<T><<<blocks, 256, 0, s>>>(src, dst, n);
  MG_LAUNCH_CHECK();
  return MG_OK;
}
template int launch_convert<float>(cudaStream_t, const float*, float*, size_t);
template int launch_convert<bf16>(cudaStream_t, const float*, bf16*, size_t);

int launch_argmax_rows(cudaStream_t s, const float* logits, int N, int C, int32_t* out) {
  if (N <= 0) return MG_OK;
  argmax_rows_kernel<<<ceil_div(N, 4), 128, 0, s>>>(logits, N, C, out);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

}  // namespace mg
