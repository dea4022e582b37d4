// Shared host/device helpers for the sm_100a engine.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

namespace mg {

// ---- error plumbing -----------------------------------------------------------------------------
void set_last_error(const std::string& msg);
extern std::atomic<uint64_t> g_kernel_launches;   // every kernel launched by this library
int fail(int code, const std::string& msg);       // records msg, returns code
int check_device(int device);                     // MG_E_CUDA unless `device` is a usable sm_100 GPU; makes it current

#define MG_CUDA_OK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::mg::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +       \
                           __FILE__ + ":" + std::to_string(__LINE__));                          \
      return MG_E_CUDA;                                                                         \
    }                                                                                           \
  } while (0)

#define MG_LAUNCH_CHECK()                                                                       \
  do {                                                                                          \
    ::mg::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      ::mg::set_last_error(std::string("kernel launch: ") + cudaGetErrorString(_e) + " at " +  \
                           __FILE__ + ":" + std::to_string(__LINE__));                          \
      return MG_E_CUDA;                                                                         \
    }                                                                                           \
  } while (0)

#define MG_TRY(expr)               \
  do {                             \
    int _s = (expr);               \
    if (_s != 0) return _s;        \
  } while (0)

// ---- dtype helpers ------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 16-byte chunk of T unpacked to floats: 4 floats (fp32) or 8 floats (bf16).
template <typename T> struct Chunk16;
template <> struct Chunk16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& r, float* f) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <> struct Chunk16<bf16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float* f) {
    // bf16 -> fp32 is a 16-bit shift: low half = element 0, high half = element 1
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
    f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
  }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
  }
};

// Streaming 128-bit load that does not pollute L1 (data read once per step: KV cache rows).
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Exact GELU (nn.GELU() default, reference api_cache.py:47; HF "gelu").
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ex2.approx.ftz (MUFU.EX2 alone, relative error 2^-22) for the soft-max of the bf16 tensor-core attention kernels: the
// arguments are <= 0 there, exp2f() without fast-math wraps the same instruction in range handling that costs more than the MMAs
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GELU with erf evaluated by ONE branch-free polynomial on the FMA pipe (no MUFU, no selects): erf(z) = z Q(u), z clamped to
// +-3.6 (1 - erf(3.6) = 3.6e-7), u = 2 z^2 / 3.6^2 - 1, Q = degree-12 Chebyshev fit of erf(z) / z converted to monomials.
// |erf error| < 1e-6, |GELU error| < 2.8e-6 absolute in fp32 evaluation -- three orders below the bf16 rounding of the
// result; used only by the coalesced bf16 epilogue of the tensor-core GEMMs (20 FMA-pipe instructions instead of erff's
// 25 + MUFU.EX2).  The fp32 (bit-identity) paths and the decode kernel keep erff.
__device__ __forceinline__ float gelu_erf_poly(float x) {
  const float z = fminf(fmaxf(x * 0.70710678118654752440f, -3.6f), 3.6f);
  const float u = fmaf(z * z, 0.15432099f, -1.0f);
  float q = 2.729401458e-03f;
  q = fmaf(q, u, -6.031278055e-03f);
  q = fmaf(q, u, 4.302724265e-03f);
  q = fmaf(q, u, -7.580903824e-03f);
  q = fmaf(q, u, 2.152218483e-02f);
  q = fmaf(q, u, -3.501450643e-02f);
  q = fmaf(q, u, 4.816431552e-02f);
  q = fmaf(q, u, -6.730011106e-02f);
  q = fmaf(q, u, 8.983867615e-02f);
  q = fmaf(q, u, -1.138864905e-01f);
  q = fmaf(q, u, 1.438089162e-01f);
  q = fmaf(q, u, -1.954871565e-01f);
  q = fmaf(q, u, 3.927121460e-01f);
  const float hx = 0.5f * x;
  return fmaf(hx, z * q, hx);
}

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };
__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_GELU) return gelu_erf(v);
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  return v;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mg
