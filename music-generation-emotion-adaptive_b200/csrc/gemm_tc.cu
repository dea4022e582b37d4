// tcgen05 / TMEM / TMA bf16 GEMM for sm_100a:  D[M,N] = epi(A[M,K] * W[N,K]^T).
//
// Used for every dense contraction of the hot path where the batch dimension fills a tensor-core
// tile: QKV / out-proj / MLP / head GEMMs of the generator in prefill and in batched decode
// (reference api_cache.py:43,45-49,68,85 -- nn.MultiheadAttention in/out projections, the MLP and the
// head Linear), and all DistilBERT Linear layers (reference emotion_analysis/inference.py:18).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      : TMA producer  -- cp.async.bulk.tensor 2D loads of the A (128x64) and W (BNx64) bf16
//                 tiles into a kStages-deep shared-memory ring, 128B swizzle, mbarrier complete_tx
//   warp 1      : TMEM allocator + MMA issuer -- one elected thread issues tcgen05.mma
//                 (cta_group::1, kind::f16, M=128, N=BN, K=16) x4 per stage, accumulating in TMEM;
//                 tcgen05.commit releases the smem stage and finally signals the epilogue
//   warps 2..5  : epilogue -- tcgen05.ld 32x32b (each warp owns one 32-lane TMEM quarter), fused
//                 bias / exact-GELU / ReLU / residual, vectorised global stores
//
// Large M (classifier, long prefill) runs one of two persistent kernels instead (further down): one CTA per SM looping over
// 128 x BN tiles with two TMEM accumulator buffers, or -- N % 256 == 0 -- a CTA PAIR (tcgen05.mma.cta_group::2) on 256 x 256
// tiles.  Both share the coalesced bf16 epilogue (warp transpose through padded shared memory, polynomial erf GELU).
#include "gemm_tc.cuh"

#include <cstdlib>

#include "mg_engine.h"
#include "ptx.cuh"

#include <algorithm>

namespace mg {

namespace {

constexpr int kThreads = 192;

template <int BN> struct TileCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;            // 16 KB
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // keep <= ~96 KB so two CTAs share an SM (their epilogues overlap the other's main loop)
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 3 : 4);
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    int M, int N, int K, GemmEpilogue epi) {
  using Cfg = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * kGemmBM;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::kStages; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + Cfg::kABytes;
        ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
        ptx::tma_load_2d(a_dst, &tmap_a, &full_bar[s], kb * kGemmBK, m0);
        ptx::tma_load_2d(b_dst, &tmap_w, &full_bar[s], kb * kGemmBK, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(kGemmBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStageBytes);
        const uint64_t da = ptx::make_kmajor_sw128_desc(a_addr);
        const uint64_t db = ptx::make_kmajor_sw128_desc(a_addr + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128B swizzle span: +2 in 16-byte units
          ptx::umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[s]);          // frees this smem stage once the MMAs retire
      }
      ptx::umma_commit(tmem_full_bar);            // accumulator complete -> epilogue
    }
    __syncwarp();
  } else {
    // ===== epilogue (warps 2..5) =====
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after_sync();
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = m0 + quarter * 32 + lane;
    const bool row_ok = row < M;
    const size_t ld = static_cast<size_t>(epi.ld_out);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + c * 32, r);
      ptx::tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (!row_ok || col0 >= N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const bool full = (col0 + 32 <= N);
      const size_t off = static_cast<size_t>(row) * ld + col0;
      if (full && (ld % 8 == 0) && (col0 % 8 == 0)) {
        if (epi.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col0 + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (epi.act != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], epi.act);
        }
        if (epi.resid_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 x4 = *reinterpret_cast<const float4*>(epi.resid_f32 + off + j);
            v[j] += x4.x; v[j + 1] += x4.y; v[j + 2] += x4.z; v[j + 3] += x4.w;
          }
        }
        if (epi.resid_bf16) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float f[8];
            Chunk16<bf16>::unpack(*reinterpret_cast<const uint4*>(epi.resid_bf16 + off + j), f);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[j + t] += f[t];
          }
        }
        if (epi.out_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(epi.out_f32 + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (epi.out_bf16) {
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(epi.out_bf16 + off + j) = Chunk16<bf16>::pack(v + j);
        }
      } else {
        // ragged / unaligned tail: fully unrolled with predicates so that v[] keeps static indices (a dynamic index
        // would push the whole accumulator array of BOTH paths into local memory)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          if (col < N) {
            float x = v[j] + (epi.bias ? __ldg(epi.bias + col) : 0.0f);
            x = apply_act(x, epi.act);
            if (epi.resid_f32) x += epi.resid_f32[off + j];
            if (epi.resid_bf16) x += __bfloat162float(epi.resid_bf16[off + j]);
            if (epi.out_f32) epi.out_f32[off + j] = x;
            if (epi.out_bf16) epi.out_bf16[off + j] = __float2bfloat16_rn(x);
          }
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}


// Coalesced bf16 epilogue of one [32 rows x 32 columns] accumulator chunk of a warp.  tcgen05.ld 32x32b leaves ONE ROW per lane,
// so a direct 16-byte store per lane makes every store instruction touch 32 different 128-byte lines with half-written
// sectors: 4096 L1 -> L2 requests per 128 x 256 tile, which is what bounds the K = 768 GEMMs of the classifier.  Here the
// warp transposes through a private padded shared-memory buffer (32 rows x 80 bytes: the 16-byte units of eight consecutive
// rows fall into different banks) so that each instruction moves 8 rows x 64 contiguous bytes (fully written sectors, 4 x fewer
// requests); the bf16 residual is fetched the same way and transposed back into the row-per-lane layout, so the fp32 sum
// acc + bias -> act -> + residual is rounded to bf16 exactly once, as before.
constexpr int kEpiPitch = 80;                       // bytes per staged row (64 + 16 of padding)
constexpr int kEpiWarpBytes = 32 * kEpiPitch;       // per epilogue warp

__device__ __forceinline__ bool epi_can_coalesce(const GemmEpilogue& epi) {
  return epi.out_bf16 != nullptr && epi.out_f32 == nullptr && epi.resid_f32 == nullptr;
}

// Bias (8 x float4, identical in every lane) and residual (coalesced layout) of a chunk: issued between tcgen05.ld and
// tcgen05.wait::ld so that their latency runs in parallel with the accumulator read instead of behind it.
__device__ __forceinline__ void epi_prefetch(const GemmEpilogue& epi, bool fast, int row_base, int lane, int M, int col0,
                                             float4 (&bb)[8], uint4 (&u)[4]) {
  const size_t ld = static_cast<size_t>(epi.ld_out);
  const int sub = lane >> 2, piece = lane & 3;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    bb[j] = (fast && epi.bias) ? __ldg(reinterpret_cast<const float4*>(epi.bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = row_base + 8 * i + sub;
    u[i] = (fast && epi.resid_bf16 && rr < M)
               ? *reinterpret_cast<const uint4*>(epi.resid_bf16 + static_cast<size_t>(rr) * ld + col0 + 8 * piece)
               : make_uint4(0u, 0u, 0u, 0u);
  }
}

// v: this lane's row (row_base + lane), columns col0 .. col0 + 31, raw accumulators.  All 32 lanes must call.
__device__ __forceinline__ void epi_chunk_coalesced(const GemmEpilogue& epi, float (&v)[32], const float4 (&bb)[8],
                                                    const uint4 (&u)[4], int row_base, int lane, int M, int col0, uint8_t* stg) {
  const size_t ld = static_cast<size_t>(epi.ld_out);
  const int sub = lane >> 2, piece = lane & 3;       // coalesced layout: rows 8 i + sub, 16-byte piece of the 64-byte row
  if (epi.bias) {                                     // bb / u: fetched by epi_prefetch ahead of the accumulator wait
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = bb[j >> 2];
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (epi.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_poly(v[j]);
  } else if (epi.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (epi.resid_bf16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + (8 * i + sub) * kEpiPitch + 16 * piece) = u[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float f[8];
      Chunk16<bf16>::unpack(*reinterpret_cast<const uint4*>(stg + lane * kEpiPitch + 2 * j), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) v[j + q] += f[q];
    }
    __syncwarp();
  }
#pragma unroll
  for (int j = 0; j < 32; j += 8) *reinterpret_cast<uint4*>(stg + lane * kEpiPitch + 2 * j) = Chunk16<bf16>::pack(v + j);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = row_base + 8 * i + sub;
    const uint4 q = *reinterpret_cast<const uint4*>(stg + (8 * i + sub) * kEpiPitch + 16 * piece);
    if (rr < M) *reinterpret_cast<uint4*>(epi.out_bf16 + static_cast<size_t>(rr) * ld + col0 + 8 * piece) = q;
  }
  __syncwarp();
}

// -------------------------------------------------------------------------------------------------
// Persistent variant for large M (classifier, long prefill): one CTA per SM loops over 128 x BN output
// tiles; TWO TMEM accumulator buffers so the epilogue of tile i (8 warps: TMEM -> registers -> fused
// bias / GELU / residual -> global) overlaps the TMA + tcgen05.mma main loop of tile i + 1.
//   warp 0: TMA producer | warp 1: TMEM allocator + MMA issuer | warps 2..9: epilogue (two per TMEM
//   lane quarter, each half of the tile's columns)
// Tile order: consecutive CTAs take consecutive M tiles of the same N tile, so a weight tile is read
// from HBM once and then served by L2.
// -------------------------------------------------------------------------------------------------
constexpr int kPThreads = 320;

template <int BN> struct PTileCfg {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;                          // two accumulator buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + 8 * kEpiWarpBytes;
};

template <int BN>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_bf16_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                               int M, int N, int K, GemmEpilogue epi) {
  using Cfg = PTileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kGemmBM - 1) / kGemmBM, n_tiles = (N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
      for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 8); }
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (t % m_tiles) * kGemmBM, n0 = (t / m_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          ptx::tma_load_2d(a_dst, &tmap_a, &full_bar[stage], kb * kGemmBK, m0);
          ptx::tma_load_2d(a_dst + Cfg::kABytes, &tmap_w, &full_bar[stage], kb * kGemmBK, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(kGemmBM, BN);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);      // epilogue drained this accumulator buffer
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = ptx::make_kmajor_sw128_desc(a_addr);
          const uint64_t db = ptx::make_kmajor_sw128_desc(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            ptx::umma_bf16_ss(tmem_base + buf * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;                                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                               // which part (1 / kParts) of the tile's columns
    const size_t ld = static_cast<size_t>(epi.ld_out);
    uint8_t* stg = smem + Cfg::kStages * Cfg::kStageBytes + 256 + (warp - 2) * kEpiWarpBytes;   // behind the barriers
    const bool coalesce = epi_can_coalesce(epi) && (ld % 8 == 0);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int m0 = (t % m_tiles) * kGemmBM, n0 = (t / m_tiles) * BN;
      const uint32_t buf = it & 1;
      ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after_sync();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + c * 32, r);
        const int col0 = n0 + c * 32;
        const bool fast = coalesce && col0 + 32 <= N;               // warp-uniform (col0 is a multiple of 32)
        float4 bb[8];
        uint4 ru[4];
        epi_prefetch(epi, fast, m0 + quarter * 32, lane, M, col0, bb, ru);
        ptx::tmem_ld_wait();
        if (col0 >= N) continue;                                    // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (fast) {                                                 // all lanes take part in the transpose
          epi_chunk_coalesced(epi, v, bb, ru, m0 + quarter * 32, lane, M, col0, stg);
          continue;
        }
        if (!row_ok) continue;
        const size_t off = static_cast<size_t>(row) * ld + col0;
        if (col0 + 32 <= N && (ld % 8 == 0) && (col0 % 8 == 0)) {
          if (epi.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (epi.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], epi.act);
          }
          if (epi.resid_f32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 x4 = *reinterpret_cast<const float4*>(epi.resid_f32 + off + j);
              v[j] += x4.x; v[j + 1] += x4.y; v[j + 2] += x4.z; v[j + 3] += x4.w;
            }
          }
          if (epi.resid_bf16) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
              Chunk16<bf16>::unpack(*reinterpret_cast<const uint4*>(epi.resid_bf16 + off + j), f);
#pragma unroll
              for (int q = 0; q < 8; ++q) v[j + q] += f[q];
            }
          }
          if (epi.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(epi.out_f32 + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (epi.out_bf16) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(epi.out_bf16 + off + j) = Chunk16<bf16>::pack(v + j);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < N) {
              float x = v[j] + (epi.bias ? __ldg(epi.bias + col) : 0.0f);
              x = apply_act(x, epi.act);
              if (epi.resid_f32) x += epi.resid_f32[off + j];
              if (epi.resid_bf16) x += __bfloat162float(epi.resid_bf16[off + j]);
              if (epi.out_f32) epi.out_f32[off + j] = x;
              if (epi.out_bf16) epi.out_bf16[off + j] = __float2bfloat16_rn(x);
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// -------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent kernel (cta_group::2): a cluster of two CTAs computes 256 x 256 output tiles.  A single
// CTA's 128 x 256 tile moves 12 KB of operand reads + 12 KB of TMA writes through shared memory per k16 step (192 cycles at
// 128 B/clk against 128 cycles of MMA), i.e. it is bound by shared-memory bandwidth; in a pair every CTA holds its own 128
// rows of A and HALF of the B tile (128 of the 256 weight rows): 8 KB + 8 KB per step.
//   both CTAs : warp 0 = TMA producer (own A rows, own half of B; bytes counted on the LEADER's full barrier),
//               warps 2..9 = epilogue of the CTA's own 128 accumulator rows (TMEM lanes 0..127 x 256 columns x 2 buffers)
//   leader    : warp 1 = MMA issuer (tcgen05.mma.cta_group::2, M = 256); commits are multicast to both CTAs' barriers
// -------------------------------------------------------------------------------------------------
#ifndef MG_PAIR_EPI_PARTS
#define MG_PAIR_EPI_PARTS 2
#endif
struct PairCfg {
  // epilogue warps per TMEM lane quarter: each takes kBN / kParts of the tile's columns.  The epilogue of a chunk is a chain of
  // dependent latencies (tcgen05.ld -> math -> shared-memory transpose -> global store); with 2 warps per scheduler (kParts = 2)
  // the K = 768 GEMMs of the classifier were bound by it (lin1: tensor pipe 44 % active, issue slots 49 %, ncu r1j)
  static constexpr int kParts = MG_PAIR_EPI_PARTS;
  static constexpr int kEpiWarps = 4 * kParts;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kBN = 256;
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;              // 128 rows of A
  static constexpr int kBBytes = (kBN / 2) * kGemmBK * 2;            // 128 of the 256 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kParts > 2 ? 5 : 6;                 // 16 staging buffers of 2.5 KB take the sixth stage's room
  static constexpr int kTmemCols = 2 * kBN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256 + kEpiWarps * kEpiWarpBytes;
  static_assert(kSmemBytes <= 227 * 1024, "pair kernel: shared memory");
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PairCfg::kThreads, 1)
gemm_bf16_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         int M, int N, int K, GemmEpilogue epi) {
  using Cfg = PairCfg;
  constexpr int BN = Cfg::kBN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);   // used in the leader only
  uint64_t* empty_bar = full_bar + Cfg::kStages;                   // per CTA (multicast commit)
  uint64_t* tmem_full = empty_bar + Cfg::kStages;                  // [2] per CTA (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;                            // [2] used in the leader only: 16 epilogue warps arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_tiles = (M + 2 * kGemmBM - 1) / (2 * kGemmBM), n_tiles = (N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int num_kb = (K + kGemmBK - 1) / kGemmBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 2 * Cfg::kEpiWarps); }
    ptx::fence_mbar_init();
  }
  // the peer's barriers must exist before anything of this CTA can signal them (fence.mbarrier_init publishes them)
  ptx::cluster_arrive_relaxed();
  ptx::cluster_wait();
  if (warp == 1) ptx::tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < total; t += n_pairs) {
        const int m0 = (t % m_tiles) * 2 * kGemmBM + static_cast<int>(rank) * kGemmBM;
        const int n0 = (t / m_tiles) * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);   // both CTAs' bytes
          const uint32_t lead_bar = ptx::map_to_cta(ptx::smem_u32(&full_bar[stage]), 0);
          ptx::tma_load_2d_pair(a_dst, &tmap_a, lead_bar, kb * kGemmBK, m0);
          ptx::tma_load_2d_pair(a_dst + Cfg::kABytes, &tmap_w, lead_bar, kb * kGemmBK, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(2 * kGemmBM, BN);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (int t = pair; t < total; t += n_pairs, ++it) {
        const uint32_t buf = it & 1;
        ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);      // both CTAs' epilogues drained this accumulator buffer
        ptx::tc_fence_after_sync();
        if (epi.prof && pair == 0 && it < 32) epi.prof[it * 4 + 0] = ptx::global_timer_ns();
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = ptx::make_kmajor_sw128_desc(a_addr);
          const uint64_t db = ptx::make_kmajor_sw128_desc(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k)
            ptx::umma_bf16_ss_pair(tmem_base + buf * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit_pair(&empty_bar[stage]);                  // frees this stage in both CTAs
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pair(&tmem_full[buf]);
        if (epi.prof && pair == 0 && it < 32) epi.prof[it * 4 + 1] = ptx::global_timer_ns();
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;                                   // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                               // which half of the tile's columns
    const size_t ld = static_cast<size_t>(epi.ld_out);
    uint8_t* stg = smem + Cfg::kStages * Cfg::kStageBytes + 256 + (warp - 2) * kEpiWarpBytes;   // behind the barriers
    const bool coalesce = epi_can_coalesce(epi) && (ld % 8 == 0);
    const uint32_t lead_empty0 = ptx::map_to_cta(ptx::smem_u32(&tmem_empty[0]), 0);
    uint32_t it = 0;
    for (int t = pair; t < total; t += n_pairs, ++it) {
      const int m0 = (t % m_tiles) * 2 * kGemmBM + static_cast<int>(rank) * kGemmBM, n0 = (t / m_tiles) * BN;
      const uint32_t buf = it & 1;
      ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after_sync();
      const bool prof_on = epi.prof && pair == 0 && rank == 0 && warp == 2 && lane == 0 && it < 32;
      if (prof_on) epi.prof[it * 4 + 2] = ptx::global_timer_ns();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = half * (BN / (32 * Cfg::kParts)); c < (half + 1) * (BN / (32 * Cfg::kParts)); ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * BN + c * 32, r);
        const int col0 = n0 + c * 32;
        const bool fast = coalesce && col0 + 32 <= N;               // warp-uniform (col0 is a multiple of 32)
        float4 bb[8];
        uint4 ru[4];
        epi_prefetch(epi, fast, m0 + quarter * 32, lane, M, col0, bb, ru);
        ptx::tmem_ld_wait();
        if (col0 >= N) continue;                                    // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (fast) {                                                 // all lanes take part in the transpose
          epi_chunk_coalesced(epi, v, bb, ru, m0 + quarter * 32, lane, M, col0, stg);
          continue;
        }
        if (!row_ok) continue;
        const size_t off = static_cast<size_t>(row) * ld + col0;
        if (col0 + 32 <= N && (ld % 8 == 0) && (col0 % 8 == 0)) {
          if (epi.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (epi.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], epi.act);
          }
          if (epi.resid_f32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 x4 = *reinterpret_cast<const float4*>(epi.resid_f32 + off + j);
              v[j] += x4.x; v[j + 1] += x4.y; v[j + 2] += x4.z; v[j + 3] += x4.w;
            }
          }
          if (epi.resid_bf16) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
              Chunk16<bf16>::unpack(*reinterpret_cast<const uint4*>(epi.resid_bf16 + off + j), f);
#pragma unroll
              for (int q = 0; q < 8; ++q) v[j + q] += f[q];
            }
          }
          if (epi.out_f32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(epi.out_f32 + off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (epi.out_bf16) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(epi.out_bf16 + off + j) = Chunk16<bf16>::pack(v + j);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < N) {
              float x = v[j] + (epi.bias ? __ldg(epi.bias + col) : 0.0f);
              x = apply_act(x, epi.act);
              if (epi.resid_f32) x += epi.resid_f32[off + j];
              if (epi.resid_bf16) x += __bfloat162float(epi.resid_bf16[off + j]);
              if (epi.out_f32) epi.out_f32[off + j] = x;
              if (epi.out_bf16) epi.out_bf16[off + j] = __float2bfloat16_rn(x);
            }
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (prof_on) epi.prof[it * 4 + 3] = ptx::global_timer_ns();
      if (lane == 0) ptx::mbar_arrive_cluster_relaxed(lead_empty0 + buf * 8);   // the leader's tmem_empty[buf]
    }
  }

  // nobody leaves (or frees TMEM) while the peer can still signal this CTA's barriers or read its shared memory
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_arrive_relaxed();                                    // lifetime only: no data is handed over here
  ptx::cluster_wait();
  if (warp == 1) ptx::tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
}

int launch_pair(cudaStream_t stream, const CUtensorMap* ta, const CUtensorMap* tw_half, int M, int N, int K,
                const GemmEpilogue& epi) {
  const int total = ceil_div(M, 2 * kGemmBM) * ceil_div(N, PairCfg::kBN);
  gemm_bf16_tc_pair_kernel<<<2 * std::min(total, 74), PairCfg::kThreads, PairCfg::kSmemBytes, stream>>>(*ta, *tw_half, M, N, K, epi);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

template <int BN>
int launch_persistent_bn(cudaStream_t stream, const CUtensorMap* ta, const CUtensorMap* tw, int M, int N, int K,
                         const GemmEpilogue& epi) {
  const int total = ceil_div(M, kGemmBM) * ceil_div(N, BN);
  gemm_bf16_tc_persistent_kernel<BN><<<std::min(total, 148), kPThreads, PTileCfg<BN>::kSmemBytes, stream>>>(*ta, *tw, M, N, K, epi);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  }
  return fn;
}

template <int BN>
int launch_bn(cudaStream_t stream, const CUtensorMap* ta, const CUtensorMap* tw, int M, int N, int K,
              const GemmEpilogue& epi) {
  using Cfg = TileCfg<BN>;
  dim3 grid(ceil_div(N, BN), ceil_div(M, kGemmBM));
  gemm_bf16_tc_kernel<BN><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(*ta, *tw, M, N, K, epi);
  MG_LAUNCH_CHECK();
  return MG_OK;
}

template <int BN>
int init_bn() {
  MG_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TileCfg<BN>::kSmemBytes));
  return MG_OK;
}

}  // namespace

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return MG_E_CUDA;
  }
  if ((cols * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_last_error("TMA operand needs 16-byte aligned base and row pitch (K % 8 == 0)");
    return MG_E_SHAPE;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kGemmBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return MG_E_CUDA;
  }
  return MG_OK;
}

int gemm_tc_init() {
  MG_TRY(init_bn<32>());
  MG_TRY(init_bn<64>());
  MG_TRY(init_bn<128>());
  MG_TRY(init_bn<256>());
  MG_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tc_persistent_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PTileCfg<128>::kSmemBytes));
  MG_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tc_persistent_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PTileCfg<256>::kSmemBytes));
  MG_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
  return MG_OK;
}

bool gemm_use_persistent(int M, int N) { return ceil_div(M, 128) * ceil_div(N, 128) >= 2 * 148; }

// CTA-pair kernel: large-M GEMMs whose N is tiled by 256 (the classifier's 768 / 2304 / 3072); MG_GEMM_2CTA=0 disables it
bool gemm_use_pair(int M, int N) {
  static const bool enabled = !(std::getenv("MG_GEMM_2CTA") && std::atoi(std::getenv("MG_GEMM_2CTA")) == 0);
  return enabled && gemm_use_persistent(M, N) && N % 256 == 0 && M >= 1024;
}

int launch_gemm_tc_pair(cudaStream_t stream, const CUtensorMap* ta, const CUtensorMap* tw_half, int M, int N, int K,
                        const GemmEpilogue& epi) {
  if (M <= 0 || N <= 0 || K <= 0 || K % 8 != 0) {
    set_last_error("launch_gemm_tc_pair: bad shape");
    return MG_E_SHAPE;
  }
  return launch_pair(stream, ta, tw_half, M, N, K, epi);
}

int pick_gemm_bn(int M, int N) {
  // 128 x 256 tiles also for the N = 768 GEMMs of the classifier: 128 x 128 tiles (5.2 instead of 2.6 waves) measured slower,
  // 2.33 vs 2.16 ms per pass (same box)
  if (gemm_use_persistent(M, N)) return (N % 256 == 0 || N >= 1024) ? 256 : 128;
  // Few row tiles (batched decode): small BN spreads the weight stream over many SMs.
  const int m_tiles = ceil_div(M, kGemmBM);
  if (m_tiles * ceil_div(N, 128) >= 148) return 128;
  if (m_tiles * ceil_div(N, 64) >= 96) return 64;
  return 32;
}

int launch_gemm_tc(cudaStream_t stream, const CUtensorMap* ta, const CUtensorMap* tw, int M, int N, int K,
                   const GemmEpilogue& epi, int bn) {
  if (M <= 0 || N <= 0 || K <= 0 || K % 8 != 0) {
    set_last_error("launch_gemm_tc: bad shape");
    return MG_E_SHAPE;
  }
  if (gemm_use_persistent(M, N) && (bn == 128 || bn == 256))
    return bn == 256 ? launch_persistent_bn<256>(stream, ta, tw, M, N, K, epi)
                     : launch_persistent_bn<128>(stream, ta, tw, M, N, K, epi);
  switch (bn) {
    case 32: return launch_bn<32>(stream, ta, tw, M, N, K, epi);
    case 64: return launch_bn<64>(stream, ta, tw, M, N, K, epi);
    case 128: return launch_bn<128>(stream, ta, tw, M, N, K, epi);
    case 256: return launch_bn<256>(stream, ta, tw, M, N, K, epi);
    default:
      set_last_error("launch_gemm_tc: unsupported block-N");
      return MG_E_ARG;
  }
}

}  // namespace mg
