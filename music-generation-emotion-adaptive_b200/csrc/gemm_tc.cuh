// Host-side interface of the tcgen05/TMA bf16 GEMM (implementation: gemm_tc.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace mg {

// D[M,N] = epilogue(A[M,K] * W[N,K]^T): A and W are bf16, K contiguous ("K-major"), fp32 accumulate.
//   v = acc + bias[n]; v = act(v); v += resid_f32[m,n] (if set); v += resid_bf16[m,n] (if set)
//   out_f32[m,n] = v (if set; may alias resid_f32 for an in-place residual update)
//   out_bf16[m,n] = bf16(v) (if set)
struct GemmEpilogue {
  const float* bias = nullptr;
  int act = ACT_NONE;
  const float* resid_f32 = nullptr;
  const bf16* resid_bf16 = nullptr;
  float* out_f32 = nullptr;
  bf16* out_bf16 = nullptr;
  int ld_out = 0;     // row stride (elements) of out_* and resid_*
  // debug (MG_PAIR_PROF): per-tile %globaltimer stamps of cluster 0 / CTA 0 of the CTA-pair kernel: [tile][4] =
  // {MMA issue starts (accumulator buffer free), last MMA issued, epilogue starts (accumulator complete), epilogue done}
  unsigned long long* prof = nullptr;
};

// Tensor map over a row-major bf16 matrix [rows, cols] (cols contiguous) with a {64 x box_rows} box
// and 128-byte swizzle.  Out-of-bounds rows/cols read as zero.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

// Block-N choices compiled in: 32, 64, 128, 256.  `bn` == 0 picks one from the problem shape.
int launch_gemm_tc(cudaStream_t stream, const CUtensorMap* tmap_a, const CUtensorMap* tmap_w, int M, int N, int K,
                   const GemmEpilogue& epi, int bn);
int pick_gemm_bn(int M, int N);
// CTA-pair (cta_group::2) kernel for large-M GEMMs with N % 256 == 0: 256 x 256 tiles per cluster of two CTAs.
// `tmap_w_half` is the weight map with a 128-row box (each CTA loads half of the 256 weight rows of a tile).
bool gemm_use_pair(int M, int N);
int launch_gemm_tc_pair(cudaStream_t stream, const CUtensorMap* tmap_a, const CUtensorMap* tmap_w_half, int M, int N, int K,
                        const GemmEpilogue& epi);

// Weight tensor maps for every compiled block-N (32 / 64 / 128 / 256), built once per weight matrix.
struct WMaps {
  CUtensorMap m[4];
  bool ok = false;
};
inline int bn_index(int bn) { return bn == 32 ? 0 : bn == 64 ? 1 : bn == 128 ? 2 : 3; }
int make_wmaps(WMaps* w, const void* base, int N, int K);

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;

}  // namespace mg
