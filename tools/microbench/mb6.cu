// Round-2 feasibility numbers for the weight-stationary "SM = weight slice, warp = sequence group" decode kernel:
//   (1) cost of one dependent hop between phases whose producers / consumers sit on different SMs and talk through L2
//       (data st.global -> fence -> red.add flag | poll flag -> ld.cg data), with the producer counts of the real step
//       (QKV 48, attention 128, out_proj 16, MLP1 64, MLP2 16 per layer, head 148, sampler 8), 8 groups in flight;
//   (2) K/V streaming rate per SM with cp.async.bulk into per-warp shared-memory rings (stage size / depth sweep);
//   (3) both at once (does the stream slow the hops down, do the hops slow the stream down).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb6 mb6.cu && ./mb6
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1;} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg16(const void* p) {
  uint4 r; asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory"); return r;
}

constexpr int NSM_MAX = 160;
constexpr int NPH = 22;                       // 4 layers x 5 + head + sampler
__constant__ int c_nprod[NPH];
__constant__ int c_off[NPH];
__constant__ int c_rbytes[NPH];              // bytes a consumer of phase p's output reads

struct Params {
  uint32_t* flags;        // [groups][NPH] one per 128-byte line
  uint8_t* act;           // [groups][NPH][16 KB]
  const uint8_t* kv;      // streaming source
  size_t kv_per_warp;     // bytes per streaming warp
  int groups, steps, nsm;
  int hop_mode;           // 0: no hops, 1: relaxed poll + fence, 2: acquire poll
  int stream_warps;       // warps 8..8+stream_warps-1 stream
  int stage_bytes, nstage, stream_reps;
  int* err;
  long long* out;         // [0] hop cycles (group 0, SM 0), [1] stream cycles
  unsigned long long* sink;
};

__global__ void __launch_bounds__(512, 1) k(Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[8][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sm = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) mbar_init(&bars[i][j], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 8) {
    if (p.hop_mode == 0 || warp >= p.groups) return;
    const int g = warp;
    uint32_t* flags = p.flags + (size_t)g * NPH * 32;
    uint8_t* act = p.act + (size_t)g * NPH * 16384;
    unsigned long long acc = 0;
    for (int step = 0; step < p.steps; ++step) {
      for (int ph = 0; ph < NPH; ++ph) {
        const int rel = (sm - c_off[ph] + 2 * p.nsm) % p.nsm;
        if (rel >= c_nprod[ph]) continue;                    // this SM has no unit in this phase
        const int prev = (ph + NPH - 1) % NPH;
        const uint32_t want = (uint32_t)(c_nprod[prev]) * (uint32_t)(ph == 0 ? step : step + 1);
        // wait for every producer of the previous phase
        if (want) {
          uint32_t tries = 0;
          while (true) {
            const uint32_t v = p.hop_mode == 2 ? ld_acquire(flags + prev * 32) : ld_relaxed(flags + prev * 32);
            if (v >= want) break;
            if (++tries > (1u << 22)) { *p.err = 1 + ph; return; }
            if ((tries & 1023) == 0 && *(volatile int*)p.err) return;
          }
          if (p.hop_mode == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        // read the previous phase's output (L2)
        const uint8_t* src = act + (size_t)prev * 16384;
        const int n16 = c_rbytes[prev] / 512;                // 16-byte loads per lane
        uint4 s = make_uint4(0, 0, 0, 0);
        for (int i = 0; i < n16; i += 4) {
          uint4 a = ld_cg16(src + (size_t)(i * 32 + lane) * 16);
          uint4 b = ld_cg16(src + (size_t)((i + 1) * 32 + lane) * 16);
          uint4 c = ld_cg16(src + (size_t)((i + 2) * 32 + lane) * 16);
          uint4 d = ld_cg16(src + (size_t)((i + 3) * 32 + lane) * 16);
          s.x += a.x + b.x + c.x + d.x; s.y += a.y + b.y + c.y + d.y; s.z += a.z + b.z + c.z + d.z; s.w += a.w + b.w + c.w + d.w;
        }
        // a little dependent arithmetic (LayerNorm statistics: two warp reductions)
        uint32_t r = s.x ^ s.y ^ s.z ^ s.w;
        for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        for (int o = 16; o; o >>= 1) r ^= __shfl_xor_sync(0xffffffffu, r, o);
        acc += r;
        // write this unit's slice (512 B) and publish
        uint8_t* dst = act + (size_t)ph * 16384 + (size_t)(rel % 32) * 512;
        *reinterpret_cast<uint4*>(dst + lane * 16) = make_uint4(r, step, ph, lane);
        __syncwarp();
        if (lane == 0) {
          if (p.hop_mode == 2) red_release(flags + ph * 32, 1);
          else { __threadfence(); red_relaxed(flags + ph * 32, 1); }
        }
      }
    }
    if (lane == 0) p.sink[sm * 16 + warp] = acc;
    if (sm == 0 && g == 0 && lane == 0) p.out[0] = clock64() - t0;
  } else {
    const int sw = warp - 8;
    if (sw >= p.stream_warps) return;
    uint8_t* ring = smem + (size_t)sw * p.nstage * p.stage_bytes;
    const uint8_t* src = p.kv + ((size_t)sm * p.stream_warps + sw) * p.kv_per_warp;
    const int nreg = (int)(p.kv_per_warp / p.stage_bytes);
    const int ntile = nreg * p.stream_reps;
    const int half = p.stage_bytes / 2;
    // prologue
    if (lane == 0)
      for (int i = 0; i < p.nstage && i < ntile; ++i) {
        expect_tx(&bars[sw][i], p.stage_bytes);
        bulk_load(ring + (size_t)i * p.stage_bytes, src + (size_t)(i % nreg) * p.stage_bytes, half, &bars[sw][i]);
        bulk_load(ring + (size_t)i * p.stage_bytes + half, src + (size_t)(i % nreg) * p.stage_bytes + half, half, &bars[sw][i]);
      }
    uint32_t acc = 0;
    int st = 0; uint32_t phs = 0;
    for (int t = 0; t < ntile; ++t) {
      uint32_t tries = 0;
      while (!try_wait(&bars[sw][st], phs)) { if (++tries > (1u << 22)) { *p.err = 100; return; } }
      const uint8_t* tile = ring + (size_t)st * p.stage_bytes;
      for (int i = lane * 16; i < p.stage_bytes; i += 512) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
      }
      __syncwarp();
      if (lane == 0 && t + p.nstage < ntile) {
        const size_t o = (size_t)((t + p.nstage) % nreg) * p.stage_bytes;
        expect_tx(&bars[sw][st], p.stage_bytes);
        bulk_load(ring + (size_t)st * p.stage_bytes, src + o, half, &bars[sw][st]);
        bulk_load(ring + (size_t)st * p.stage_bytes + half, src + o + half, half, &bars[sw][st]);
      }
      if (++st == p.nstage) { st = 0; phs ^= 1; }
    }
    if (lane == 0) p.sink[sm * 16 + warp] = acc;
    if (lane == 0 && sw == 0 && sm == 0) p.out[1] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("SMs %d, clock %d kHz\n", nsm, clk_khz);
  int nprod[NPH], off[NPH], rbytes[NPH];
  int o = 0;
  for (int l = 0; l < 4; ++l) {
    const int np[5] = {48, 128, 16, 64, 16};
    const int rb[5] = {1024, 16384, 8192, 16384, 8192};   // q slice / attention partials / x1 / hidden / x
    for (int j = 0; j < 5; ++j) { nprod[l * 5 + j] = np[j]; off[l * 5 + j] = o % nsm; rbytes[l * 5 + j] = rb[j]; o += np[j]; }
  }
  nprod[20] = nsm; off[20] = 0; rbytes[20] = 16384;        // head: every SM; the sampler reads tile maxima + candidates
  nprod[21] = 8; off[21] = 17; rbytes[21] = 8192;           // sampler units; the next QKV reads x (8 KB)
  CK(cudaMemcpyToSymbol(c_nprod, nprod, sizeof(nprod)));
  CK(cudaMemcpyToSymbol(c_off, off, sizeof(off)));
  CK(cudaMemcpyToSymbol(c_rbytes, rbytes, sizeof(rbytes)));

  Params p{};
  const int G = 8;
  CK(cudaMalloc(&p.flags, G * NPH * 128));
  CK(cudaMalloc(&p.act, (size_t)G * NPH * 16384));
  CK(cudaMemset(p.act, 1, (size_t)G * NPH * 16384));
  const size_t kv_total = (size_t)3 << 30;
  uint8_t* kv; CK(cudaMalloc(&kv, kv_total)); CK(cudaMemset(kv, 3, kv_total));
  p.kv = kv;
  CK(cudaMalloc(&p.err, 4)); CK(cudaMalloc(&p.out, 16)); CK(cudaMalloc(&p.sink, NSM_MAX * 16 * 8));
  p.nsm = nsm;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  auto run = [&](const char* name, int groups, int steps, int hop_mode, int stream_warps, int stage_bytes, int nstage, size_t per_warp, int sreps = 1) -> int {
    p.groups = groups; p.steps = steps; p.hop_mode = hop_mode; p.stream_warps = stream_warps; p.stage_bytes = stage_bytes; p.nstage = nstage; p.stream_reps = sreps;
    p.kv_per_warp = per_warp / stage_bytes * stage_bytes;
    if (stream_warps && (size_t)nsm * stream_warps * p.kv_per_warp > kv_total) { printf("%s: source too small\n", name); return 0; }
    if (stream_warps * nstage * stage_bytes > 200 * 1024) { printf("%s: ring too large\n", name); return 0; }
    float best = 1e30f; long long out[2] = {0, 0};
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(p.flags, 0, G * NPH * 128)); CK(cudaMemset(p.err, 0, 4)); CK(cudaMemset(p.out, 0, 16));
      CK(cudaEventRecord(e0));
      k<<<nsm, 512, 200 * 1024>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      int err; CK(cudaMemcpy(&err, p.err, 4, cudaMemcpyDeviceToHost));
      if (err) { printf("%s: TIMEOUT code %d\n", name, err); return 0; }
      if (ms < best) { best = ms; CK(cudaMemcpy(out, p.out, 16, cudaMemcpyDeviceToHost)); }
    }
    const double hop_us = hop_mode ? out[0] / (clk_khz * 1e-3) / ((double)steps * NPH) : 0.0;
    const double bytes = (double)nsm * stream_warps * p.kv_per_warp * sreps;
    printf("%-46s kernel %8.3f ms | hop %6.3f us (chain of %d: %6.2f us/step) | stream %7.1f GB/s total, %6.1f GB/s per SM\n", name, best,
           hop_us, NPH, hop_us * NPH, stream_warps ? bytes / (out[1] / (clk_khz * 1e3)) * 1e-9 : 0.0,
           stream_warps ? bytes / (out[1] / (clk_khz * 1e3)) * 1e-9 / nsm : 0.0);
    return 0;
  };
  const size_t MB = 1 << 20;
  run("hops relaxed+fence, 1 group", 1, 400, 1, 0, 4096, 2, 0);
  run("hops acquire/release, 1 group", 1, 400, 2, 0, 4096, 2, 0);
  run("hops relaxed+fence, 8 groups", 8, 400, 1, 0, 4096, 2, 0);
  run("hops acquire/release, 8 groups", 8, 400, 2, 0, 4096, 2, 0);
  run("stream 8 warps x 2 x 8 KB", 0, 0, 0, 8, 8192, 2, 2 * MB);
  run("stream 8 warps x 3 x 8 KB", 0, 0, 0, 8, 8192, 3, 2 * MB);
  run("stream 8 warps x 4 x 4 KB", 0, 0, 0, 8, 4096, 4, 2 * MB);
  run("stream 8 warps x 6 x 4 KB", 0, 0, 0, 8, 4096, 6, 2 * MB);
  run("stream 8 warps x 2 x 4 KB", 0, 0, 0, 8, 4096, 2, 2 * MB);
  run("stream 4 warps x 4 x 8 KB", 0, 0, 0, 4, 8192, 4, 4 * MB);
  run("stream 2 warps x 8 x 8 KB", 0, 0, 0, 2, 8192, 8, 8 * MB);
  run("stream 1 warp  x 8 x 16 KB", 0, 0, 0, 1, 16384, 8, 16 * MB);
  run("stream 8 warps x 3 x 8 KB + hops(rel) 8 groups", 8, 160, 1, 8, 8192, 3, 2 * MB, 10);
  run("stream 8 warps x 4 x 4 KB + hops(rel) 8 groups", 8, 160, 1, 8, 4096, 4, 2 * MB, 10);
  run("stream 8 warps x 3 x 8 KB + hops(acq) 8 groups", 8, 160, 2, 8, 8192, 3, 2 * MB, 10);
  printf("done\n");
  return 0;
}
