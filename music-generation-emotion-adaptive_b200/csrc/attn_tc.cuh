// Tensor-core flash-decoding of one (sequence, head, key range) by one warp straight from global memory, shared by the
// persistent cluster decode kernel (decode_mega.cu) and the grid-synchronous decode kernel (decode_grid.cu).
// Reference op: api_cache.py:66-68 (softmax(q K^T / sqrt(hd)) V over the whole cache, no mask).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "ptx.cuh"

#ifndef MG_MEGA_KVSLOT
#define MG_MEGA_KVSLOT 0
#endif

namespace mg {
namespace attn {

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_stream16(const bf16* p) { return ptx::ld_global_stream16(p); }
// L1-bypassing flavour for caches whose rows are appended by OTHER SMs during the launch (decode_grid.cu): a weak load may be
// served from a stale L1 line (observed on B200: the appending SM keeps the line it wrote); ld.volatile never looks at L1
#ifndef MG_GRID_KVLD
#define MG_GRID_KVLD 2
#endif
__device__ __forceinline__ uint4 ldg_strong16(const bf16* p) {
  uint4 r;
#if MG_GRID_KVLD == 0
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#elif MG_GRID_KVLD == 1
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#elif MG_GRID_KVLD == 2
  asm volatile("ld.relaxed.gpu.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#else
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
#endif
  return r;
}
template <int HD, bool STRONG = false>
__device__ __forceinline__ void attn_load_k(const bf16* __restrict__ kh, int b, int lane, uint4 (&kq)[4][HD / 32]) {
  const int g = lane >> 2, t = lane & 3;
  // rows past the end of the sequence are read too (finite: the cache is zero-filled once and only ever holds bf16 data; its
  // row count is rounded up to a multiple of 32) and masked in attn_tc: one pointer per block, immediate offsets per load
  const bf16* kb = kh + (static_cast<size_t>(b) * 32 + g) * HD + 8 * t;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) kq[j][c] = STRONG ? ldg_strong16(kb + j * 8 * HD + 32 * c) : ldg_stream16(kb + j * 8 * HD + 32 * c);
}
template <int HD, bool STRONG = false>
__device__ __forceinline__ void attn_load_v(const bf16* __restrict__ vt, int b, int lane, uint4 (&vq)[HD / 8]) {
  const int g = lane >> 2, t = lane & 3;
  const bf16* vb = vt + static_cast<size_t>(b) * (HD * 32) + g * 32 + 8 * t;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) vq[n] = STRONG ? ldg_strong16(vb + n * 256) : ldg_stream16(vb + n * 256);
}
template <int HD, bool STRONG = false>
__device__ __forceinline__ void attn_load_block(const bf16* __restrict__ kh, const bf16* __restrict__ vt, int len, int b, int lane,
                                                uint4 (&kq)[4][HD / 32], uint4 (&vq)[HD / 8]) {
  attn_load_k<HD, STRONG>(kh, b, lane, kq);
  attn_load_v<HD, STRONG>(vt, b, lane, vq);
}

// kq0 / vq0: block `wi` of this worker, loaded by the caller ahead of time (before the QKV GEMM, whose result the loads do not
// depend on) when PRE is set.
// One 32-key block of one head into this warp's staging slot: two bulk copies (K rows 2 KB, V^T block 2 KB) counted on the
// warp's own mbarrier.  Lane 0 issues; callers make sure every lane has finished reading the slot (__syncwarp).
template <int HD>
__device__ __forceinline__ void kvslot_issue(const bf16* __restrict__ kh, const bf16* __restrict__ vt, int b, int lane, uint8_t* slot,
                                             uint64_t* bar) {
  if (lane == 0) {
    ptx::mbar_arrive_expect_tx(bar, 2 * 32 * HD * 2);
    ptx::bulk_load_1d(slot, kh + static_cast<size_t>(b) * 32 * HD, 32 * HD * 2, bar);
    ptx::bulk_load_1d(slot + 32 * HD * 2, vt + static_cast<size_t>(b) * (HD * 32), 32 * HD * 2, bar);
  }
}

template <int HD, bool PRE, bool STRONG = false>
__device__ __forceinline__ void attn_tc(const bf16* __restrict__ kh, const bf16* __restrict__ vt, int len, int wi, int nws,
                                        int lane, const float* __restrict__ q, const bf16* __restrict__ knew,
                                        const bf16* __restrict__ vnew, bool fold_new, float* __restrict__ out,
                                        uint4 (&kq0)[4][HD / 32], uint4 (&vq0)[HD / 8], uint8_t* slot = nullptr,
                                        uint64_t* slot_bar = nullptr, uint32_t* slot_phase = nullptr, bool slot_issued = false) {
  constexpr int KS = HD / 16;              // k-steps of the score MMAs
  constexpr int NT = HD / 8;               // n-tiles (8 dims) of the output MMAs
  constexpr int KL = HD / 32;              // 16-byte K loads per lane per key row
  const int g = lane >> 2, t = lane & 3;
  // query fragments: k-step ks covers dims 32 (ks / 2) + 8 t + 4 (ks % 2) + {0..3} of this lane
  uint32_t qh[KS][2], ql[KS][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const float4 qv = *reinterpret_cast<const float4*>(q + 32 * (ks >> 1) + 8 * t + 4 * (ks & 1));
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(qv.x, qv.y), h1 = __floats2bfloat162_rn(qv.z, qv.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(qv.x - __bfloat162float(h0.x), qv.y - __bfloat162float(h0.y));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(qv.z - __bfloat162float(h1.x), qv.w - __bfloat162float(h1.y));
    const bool row0 = g == 0;
    qh[ks][0] = row0 ? *reinterpret_cast<const uint32_t*>(&h0) : 0u;
    qh[ks][1] = row0 ? *reinterpret_cast<const uint32_t*>(&h1) : 0u;
    ql[ks][0] = row0 ? *reinterpret_cast<const uint32_t*>(&l0) : 0u;
    ql[ks][1] = row0 ? *reinterpret_cast<const uint32_t*>(&l1) : 0u;
  }
  float oacc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[n][e] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;   // l_run: this lane's share of the denominator (summed over the quad at the end)

  const int nblk = (len + 31) >> 5;
  auto load_block = [&](int b, uint4 (&kq)[4][KL], uint4 (&vq)[NT]) { attn_load_block<HD, STRONG>(kh, vt, len, b, lane, kq, vq); };
  // next_b >= 0: the K rows / V block of block next_b are requested into the SAME registers as soon as this block's MMAs have consumed them
  auto compute_block_h = [&](int b, uint4 (&kq)[4][KL], uint4 (&vq)[NT], int next_b) {
    const int key0 = b << 5;
    float sc[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < KL; ++c) {
        const uint32_t a0[4] = {qh[2 * c][0], ql[2 * c][0], qh[2 * c][1], ql[2 * c][1]};
        const uint32_t a1[4] = {qh[2 * c + 1][0], ql[2 * c + 1][0], qh[2 * c + 1][1], ql[2 * c + 1][1]};
        mma_bf16_16816(c4, a0, kq[j][c].x, kq[j][c].y);
        mma_bf16_16816(c4, a1, kq[j][c].z, kq[j][c].w);
      }
      sc[j][0] = c4[0] + c4[2];                                  // row 0 (hi) + row 8 (lo)
      sc[j][1] = c4[1] + c4[3];
    }
    if (next_b >= 0) attn_load_k<HD, STRONG>(kh, next_b, lane, kq);   // the K registers are free from here on
    if (key0 + 32 > len) {                                       // only the last block of a sequence is partial
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = key0 + 8 * j + 2 * t;
        sc[j][0] = key < len ? sc[j][0] : -INFINITY;
        sc[j][1] = key + 1 < len ? sc[j][1] : -INFINITY;
      }
    }
    float mb = fmaxf(fmaxf(fmaxf(sc[0][0], sc[0][1]), fmaxf(sc[1][0], sc[1][1])),
                     fmaxf(fmaxf(sc[2][0], sc[2][1]), fmaxf(sc[3][0], sc[3][1])));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    const float m_new = fmaxf(m_run, mb);                        // finite: key0 < len, so the quad holds a valid key
    const float corr = fast_exp2(m_run - m_new);
    float psum = 0.f;
    uint32_t pa[2][2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const float p0 = fast_exp2(sc[2 * u + w][0] - m_new), p1 = fast_exp2(sc[2 * u + w][1] - m_new);
        psum += p0 + p1;
        pa[u][w] = pack_bf16(p0, p1);
      }
    l_run = fmaf(l_run, corr, psum);
    m_run = m_new;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      oacc[n][0] *= corr;
      oacc[n][1] *= corr;
      const uint32_t a0[4] = {pa[0][0], 0u, pa[0][1], 0u};
      const uint32_t a1[4] = {pa[1][0], 0u, pa[1][1], 0u};
      mma_bf16_16816(oacc[n], a0, vq[n].x, vq[n].y);
      mma_bf16_16816(oacc[n], a1, vq[n].z, vq[n].w);
    }
    if (next_b >= 0) attn_load_v<HD, STRONG>(vt, next_b, lane, vq);   // ... and the V registers from here on
  };
  auto compute_block = [&](int b, const uint4 (&kq)[4][KL], const uint4 (&vq)[NT]) {
    compute_block_h(b, const_cast<uint4(&)[4][KL]>(kq), const_cast<uint4(&)[NT]>(vq), -1);
  };
  if (HD == 32 && MG_MEGA_KVSLOT && slot != nullptr) {
    // Three blocks in flight per warp: register sets A (kq0 / vq0) and B by 16-byte global loads, the staging slot S by bulk copy.
    // Block k of this warp (wi + k nws) lives in storage k % 3; after a block is consumed its storage is refilled with block k + 3.
    uint4 kqb[4][KL], vqb[NT];
    int b = wi;
    if (!PRE && b < nblk) load_block(b, kq0, vq0);
    if (b + nws < nblk) load_block(b + nws, kqb, vqb);
    if (!slot_issued && b + 2 * nws < nblk) kvslot_issue<HD>(kh, vt, b + 2 * nws, lane, slot, slot_bar);
    uint32_t ph = *slot_phase;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t sa = ptx::smem_u32(slot) + g * 64 + t * 16;
    while (b < nblk) {
      compute_block(b, kq0, vq0);
      if (b + 3 * nws < nblk) load_block(b + 3 * nws, kq0, vq0);
      b += nws;
      if (b >= nblk) break;
      compute_block(b, kqb, vqb);
      if (b + 3 * nws < nblk) load_block(b + 3 * nws, kqb, vqb);
      b += nws;
      if (b >= nblk) break;
      {
        uint4 kqc[4][KL], vqc[NT];
        ptx::mbar_wait_spin(slot_bar, ph);
        ph ^= 1u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(kqc[j][0].x), "=r"(kqc[j][0].y), "=r"(kqc[j][0].z), "=r"(kqc[j][0].w) : "r"(sa + j * 512));
#pragma unroll
        for (int n = 0; n < NT; ++n)
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(vqc[n].x), "=r"(vqc[n].y), "=r"(vqc[n].z), "=r"(vqc[n].w) : "r"(sa + 2048 + n * 512));
        __syncwarp();                                         // every lane holds its part: the slot may be refilled
        if (b + 3 * nws < nblk) kvslot_issue<HD>(kh, vt, b + 3 * nws, lane, slot, slot_bar);
        compute_block(b, kqc, vqc);
        b += nws;
      }
    }
    *slot_phase = ph;
  } else if (HD == 32) {
    // MG_ATTN_BUFS register sets: the loads of the next block(s) are in flight while this one is computed (bytes in flight per
    // SM = warps x sets x 4 KB).  Measured on B200 (same box, config 3 / config 4): 2 sets 78.8 ms / 80.7 us per step, 3 sets
    // 83.9 / 85.8, 4 sets 106.9 / 109.7 -- with more sets the loads of different sets end up on the same scoreboard and a
    // wait for the oldest set waits for all of them, so two is the useful depth of a register pipeline.
#ifndef MG_ATTN_BUFS
#define MG_ATTN_BUFS 2
#endif
    constexpr int NB = MG_ATTN_BUFS;
    uint4 kqx[NB - 1][4][KL], vqx[NB - 1][NT];
    int b = wi;
    if (!PRE && b < nblk) load_block(b, kq0, vq0);
#pragma unroll
    for (int i = 1; i < NB; ++i)
      if (b + i * nws < nblk) load_block(b + i * nws, kqx[i - 1], vqx[i - 1]);
    while (b < nblk) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        if (b >= nblk) break;
        if (i == 0) compute_block(b, kq0, vq0); else compute_block(b, kqx[i - 1], vqx[i - 1]);
        const int bn = b + NB * nws;                         // refill the set that was just consumed
        if (bn < nblk) { if (i == 0) load_block(bn, kq0, vq0); else load_block(bn, kqx[i - 1], vqx[i - 1]); }
        b += nws;
      }
    }
  } else {
    // head_dim 64: ONE register set, load -> compute per block.  Measured and rejected on B200 (same box): a second register set
    // (spills: train_large2 251 -> 297 us per step in the grid kernel) and requesting the next block's K rows behind the score MMAs /
    // its V block behind the output MMAs into the same registers (251 -> 275; batch-1 cluster kernel 27.7 -> 29.8 us per token)
    for (int b = wi; b < nblk; b += nws) {
      if (!(PRE && b == wi)) load_block(b, kq0, vq0);
      compute_block(b, kq0, vq0);
    }
  }
  // the new token's own row (the reference's cache already contains it, api_cache.py:66-68): first worker only
  if (fold_new) {
    float sn = 0.f;
#pragma unroll
    for (int c = 0; c < KL; ++c) {
      const uint4 kn = *reinterpret_cast<const uint4*>(knew + 32 * c + 8 * t);
      float kf[8];
      unpack8(kn, kf);
      const float4 q0 = *reinterpret_cast<const float4*>(q + 32 * c + 8 * t), q1 = *reinterpret_cast<const float4*>(q + 32 * c + 8 * t + 4);
      sn = fmaf(q0.x, kf[0], fmaf(q0.y, kf[1], fmaf(q0.z, kf[2], fmaf(q0.w, kf[3], sn))));
      sn = fmaf(q1.x, kf[4], fmaf(q1.y, kf[5], fmaf(q1.z, kf[6], fmaf(q1.w, kf[7], sn))));
    }
    sn += __shfl_xor_sync(0xffffffffu, sn, 1);
    sn += __shfl_xor_sync(0xffffffffu, sn, 2);
    const float m_new = fmaxf(m_run, sn);
    const float corr = fast_exp2(m_run - m_new), pw = fast_exp2(sn - m_new);
    l_run = fmaf(l_run, corr, t == 0 ? pw : 0.f);
    m_run = m_new;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(vnew + 8 * n + 2 * t);
      oacc[n][0] = fmaf(pw, __bfloat162float(v2.x), oacc[n][0] * corr);
      oacc[n][1] = fmaf(pw, __bfloat162float(v2.y), oacc[n][1] * corr);
    }
  }
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  if (g == 0) {                                                   // out: [HD] numerators | m | l
#pragma unroll
    for (int n = 0; n < NT; ++n) *reinterpret_cast<float2*>(out + 8 * n + 2 * t) = make_float2(oacc[n][0], oacc[n][1]);
    if (t == 0) { out[64] = m_run; out[65] = l_run; }
  }
}

}  // namespace attn
}  // namespace mg
