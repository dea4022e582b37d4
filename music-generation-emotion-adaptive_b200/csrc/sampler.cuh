// Sampler device code shared by the stand-alone sampler kernels (kernels.cu) and the grid-synchronous persistent decode kernel
// (decode_grid.cu): /temperature -> top-k (radix select) -> softmax over the kept set -> Philox multinomial, one CTA of
// kSampleThreads threads per logits row (reference api_cache.py:169-178).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace mg {
namespace {

// =================================================================================================
// Sampler: /temperature -> top-k (radix select) -> softmax over the kept set -> Philox multinomial
// =================================================================================================
constexpr int kSampleThreads = 256;

__device__ __forceinline__ uint32_t float_key(float f) {       // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// uniform in [0,1) from the Philox stream (seed, sequence index, step)
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t seq, uint32_t step) {
  uint32_t c[4] = {static_cast<uint32_t>(seq), static_cast<uint32_t>(seq >> 32), step, 0u};
  philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  return static_cast<float>(c[0] >> 8) * (1.0f / 16777216.0f);
}

struct SampleSmem {
  uint32_t hist[256];
  float red_val[kSampleThreads / 32];
  int red_idx[kSampleThreads / 32];
  float scan[kSampleThreads / 32];
  int iscan[kSampleThreads / 32];
  uint32_t prefix;
  int remaining;
  int result;
  float fmax;
  int imax;
};

// exclusive block scan of one float / one int per thread (256 threads); returns exclusive prefix,
// total in *total.
__device__ __forceinline__ float block_excl_scan_f(float v, float* warp_tot, float* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  float base = 0.0f, tot = 0.0f;
#pragma unroll
  for (int w = 0; w < kSampleThreads / 32; ++w) {
    const float t = warp_tot[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}
__device__ __forceinline__ int block_excl_scan_i(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kSampleThreads / 32; ++w) {
    const int t = warp_tot[w];
    if (w < warp) base += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

// Samples one token from logits row `row` (global).  vals = dynamic smem [V].  Result valid in all
// threads.  Follows api_cache.py:169-178: z = logits / T; keep the top_k largest (everything else
// gets -1e10 added, i.e. probability exactly 0 in fp32); softmax; one multinomial draw.
__device__ int sample_row(const float* row, int V, float temperature, int top_k, uint64_t seed,
                          uint64_t seq, uint32_t step, float* vals, SampleSmem& ss) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- scaled logits into smem + argmax (lowest index wins ties) ----
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int i = tid; i < V; i += kSampleThreads) {
    const float z = row[i] / temperature;
    vals[i] = z;
    if (z > best || (z == best && i < best_i)) { best = z; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if (lane == 0) { ss.red_val[warp] = best; ss.red_idx[warp] = best_i; }
  __syncthreads();
  if (tid == 0) {
    float bv = ss.red_val[0];
    int bi = ss.red_idx[0];
    for (int w = 1; w < kSampleThreads / 32; ++w) {
      if (ss.red_val[w] > bv || (ss.red_val[w] == bv && ss.red_idx[w] < bi)) { bv = ss.red_val[w]; bi = ss.red_idx[w]; }
    }
    ss.fmax = bv;
    ss.imax = bi;
    ss.result = -1;
  }
  __syncthreads();
  const float zmax = ss.fmax;
  if (top_k == 1) return ss.imax;                     // greedy: one-hot distribution

  // ---- k-th largest key by 4 x 8-bit radix select ----
  const bool restrict_k = top_k > 0 && top_k < V;
  uint32_t thr = 0;
  int need_eq = 0x7fffffff;
  if (restrict_k) {
    uint32_t prefix = 0, mask = 0;
    if (tid == 0) ss.remaining = top_k;
    for (int pass = 3; pass >= 0; --pass) {
      const int shift = pass * 8;
      for (int i = tid; i < 256; i += kSampleThreads) ss.hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < V; i += kSampleThreads) {
        const uint32_t u = float_key(vals[i]);
        if ((u & mask) == prefix) atomicAdd(&ss.hist[(u >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        // bins 255..0 in descending order, 8 per lane: lane 0 owns bins 255..248, lane 31 owns 7..0
        const int remaining = ss.remaining;
        int hb[8], mine = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hb[j] = static_cast<int>(ss.hist[255 - (lane * 8 + j)]);
          mine += hb[j];
        }
        int inc = mine;                                  // inclusive prefix over lanes (descending bins)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const int before = inc - mine;                   // elements in strictly higher bins of earlier lanes
        const bool here = before < remaining && inc >= remaining;
        const unsigned who = __ballot_sync(0xffffffffu, here);
        // `who` has exactly one bit unless the row holds fewer than `remaining` candidates (cannot happen: k <= V)
        if (here && (who & ((1u << lane) - 1)) == 0) {
          int cum = before, j = 0;
          for (; j < 7; ++j) {
            if (cum + hb[j] >= remaining) break;
            cum += hb[j];
          }
          ss.remaining = remaining - cum;
          ss.prefix = prefix | (static_cast<uint32_t>(255 - (lane * 8 + j)) << shift);
        }
      }
      __syncthreads();
      prefix = ss.prefix;
      mask |= 255u << shift;
    }
    thr = prefix;
    need_eq = ss.remaining;                           // how many elements equal to the threshold are kept
  }

  // ---- contiguous index range per thread so that scans follow index order ----
  const int per = (V + kSampleThreads - 1) / kSampleThreads;
  const int lo = min(V, tid * per), hi = min(V, lo + per);
  int eq_before = 0;
  if (restrict_k) {
    int my_eq = 0;
    for (int i = lo; i < hi; ++i) my_eq += (float_key(vals[i]) == thr);
    int tot;
    eq_before = block_excl_scan_i(my_eq, ss.iscan, &tot);
  }
  float my_sum = 0.0f;
  {
    int eq_rank = eq_before;
    for (int i = lo; i < hi; ++i) {
      const float z = vals[i];
      bool keep = true;
      if (restrict_k) {
        const uint32_t u = float_key(z);
        keep = u > thr || (u == thr && eq_rank++ < need_eq);
      }
      const float wgt = keep ? expf(z - zmax) : 0.0f;
      vals[i] = wgt;                                  // in place: logits -> unnormalised probabilities
      my_sum += wgt;
    }
  }
  float total;
  const float excl = block_excl_scan_f(my_sum, ss.scan, &total);
  const float target = philox_uniform(seed, seq, step) * total;
  if (my_sum > 0.0f && target >= excl && target < excl + my_sum) {
    float cum = excl;
    int pick = -1, last_kept = -1;
    for (int i = lo; i < hi; ++i) {
      const float wgt = vals[i];
      if (wgt > 0.0f) {
        last_kept = i;
        cum += wgt;
        if (cum > target) { pick = i; break; }
      }
    }
    ss.result = pick >= 0 ? pick : last_kept;
  }
  __syncthreads();
  const int res = ss.result;
  return res >= 0 ? res : ss.imax;                    // rounding corner: fall back to the mode
}

}  // namespace
}  // namespace mg
