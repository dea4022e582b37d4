// Microbenchmarks for the decode kernel's two streams (weights from L2, K/V from HBM / L2): what a cp.async.bulk ring
// delivers per SM and in aggregate, and whether cp.async.bulk.prefetch.L2 issued ahead of time turns the K/V stream
// into L2 hits.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb2 mb2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ uint64_t gtimer() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// Every CTA streams `bytes_per_cta` starting at base + blockIdx.x * cta_stride (cta_stride = 0: everyone reads the same
// region), `iters` times, through a ring of NS stages of SB bytes.  Thread 0 produces, warp 1 lane 0 consumes.
// prefetch_mode: 0 none; 1 = prefetch the whole region to L2 first, wait delay_ns, then stream.
__global__ void __launch_bounds__(64, 1) ring_kernel(const uint8_t* base, size_t cta_stride, size_t bytes_per_cta, int iters, int NS, int SB,
                                                     int prefetch_mode, int delay_ns, unsigned long long* t_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) { for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint8_t* src0 = base + (size_t)blockIdx.x * cta_stride;
  const int n_st = (int)(bytes_per_cta / SB);
  uint64_t t0 = 0;
  if (prefetch_mode == 1) {
    if (threadIdx.x < 32) {
      const size_t chunk = 65536;
      for (size_t o = (size_t)threadIdx.x * chunk; o < bytes_per_cta; o += 32 * chunk) prefetch_l2(src0 + o, (uint32_t)min(chunk, bytes_per_cta - o));
    }
    const uint64_t ts = gtimer();
    while (gtimer() - ts < (uint64_t)delay_ns) {}
    __syncthreads();
  }
  if (threadIdx.x == 0) t0 = gtimer();
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it)
      for (int i = 0; i < n_st; ++i) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect(&full[st], SB);
        bulk_load(smem + (size_t)st * SB, src0 + (size_t)i * SB, SB, &full[st]);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it)
      for (int i = 0; i < n_st; ++i) {
        mbar_wait(&full[st], ph);
        mbar_arrive(&empty[st]);
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    if (t_out) t_out[blockIdx.x] = gtimer();
  }
  __syncthreads();
  if (threadIdx.x == 0 && t_out) t_out[blockIdx.x] = t_out[blockIdx.x] - t0;
}


// Two concurrent streams per CTA: stream H (cold, HBM) is a "discard prefetch" -- bulk copies into a small scratch that
// nobody reads, NH in flight; stream L re-reads a hot (L2-resident) private region through a 4 x 32 KB ring.
// mode 0: only H; 1: only L; 2: both concurrently; 3: H (discard) then L over THE SAME region (prefetch -> consume)
__global__ void __launch_bounds__(96, 1) dual_kernel(const uint8_t* hot, const uint8_t* cold, size_t cta_stride, size_t bytes_h, size_t bytes_l,
                                                     int iters_l, int mode, int NH, int CH, unsigned long long* t_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4], empty[4], hbar[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 32; ++i) mbar_init(&hbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* hsrc = cold + (size_t)blockIdx.x * cta_stride;
  const uint8_t* lsrc = (mode == 3 ? cold : hot) + (size_t)blockIdx.x * cta_stride;
  uint8_t* scratch = smem + 4 * 32768;
  const uint64_t t0 = gtimer();
  if (threadIdx.x == 64 && (mode == 0 || mode == 2 || mode == 3)) {
    const int n = (int)(bytes_h / CH);
    for (int i = 0; i < n; ++i) {
      const int slot = i % NH;
      if (i >= NH) mbar_wait(&hbar[slot], ((i / NH) - 1) & 1);
      mbar_expect(&hbar[slot], CH);
      bulk_load(scratch + (size_t)(slot & 1) * CH, hsrc + (size_t)i * CH, CH, &hbar[slot]);
    }
    for (int i = max(n - NH, 0); i < n; ++i) mbar_wait(&hbar[i % NH], (i / NH) & 1);
    t_out[blockIdx.x * 2] = gtimer() - t0;
  }
  if (mode == 3) __syncthreads();
  const uint64_t t1 = gtimer();
  if (mode >= 1) {
    if (threadIdx.x == 0) {
      int st = 0; uint32_t ph = 0;
      const int n_st = (int)(bytes_l / 32768);
      for (int it = 0; it < iters_l; ++it)
        for (int i = 0; i < n_st; ++i) {
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect(&full[st], 32768);
          bulk_load(smem + (size_t)st * 32768, lsrc + (size_t)i * 32768, 32768, &full[st]);
          if (++st == 4) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
      int st = 0; uint32_t ph = 0;
      const int n_st = (int)(bytes_l / 32768);
      for (int it = 0; it < iters_l; ++it)
        for (int i = 0; i < n_st; ++i) {
          mbar_wait(&full[st], ph);
          mbar_arrive(&empty[st]);
          if (++st == 4) { st = 0; ph ^= 1; }
        }
      t_out[blockIdx.x * 2 + 1] = gtimer() - t1;
    }
  }
}


// Path-interference tests.  hmode: 0 none, 1 bulk discard (thread 64), 2 LDG discard by warps 3..6 (uint4 per lane, 8 in flight),
// 3 prefetch.global.L2 per 128-byte line by warp 3 (then spin delay_ns).  lmode: 0 none, 1 bulk ring (4 x 32 KB), 2 LDGSTS ring
// (warps 1..2 issue cp.async 16 B, 4 x 32 KB stages).  same_region: L reads what H touched (prefetch -> consume, sequential).
__global__ void __launch_bounds__(224, 1) path_kernel(const uint8_t* hot, const uint8_t* cold, size_t cta_stride, size_t bytes_h, size_t bytes_l,
                                                      int iters_l, int hmode, int lmode, int same_region, int delay_ns, unsigned long long* t_out,
                                                      unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[4], empty[4], hbar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 8; ++i) mbar_init(&hbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* hsrc = cold + (size_t)blockIdx.x * cta_stride;
  const uint8_t* lsrc = (same_region ? cold : hot) + (size_t)blockIdx.x * cta_stride;
  uint8_t* scratch = smem + 4 * 32768;
  const uint64_t t0 = gtimer();
  if (hmode == 1 && threadIdx.x == 64) {
    const int CH = 16384, NH = 8, n = (int)(bytes_h / CH);
    for (int i = 0; i < n; ++i) {
      const int slot = i % NH;
      if (i >= NH) mbar_wait(&hbar[slot], ((i / NH) - 1) & 1);
      mbar_expect(&hbar[slot], CH);
      bulk_load(scratch + (size_t)(slot & 1) * CH, hsrc + (size_t)i * CH, CH, &hbar[slot]);
    }
    for (int i = max(n - NH, 0); i < n; ++i) mbar_wait(&hbar[i % NH], (i / NH) & 1);
    t_out[blockIdx.x * 2] = gtimer() - t0;
  }
  if (hmode == 2 && warp >= 3 && warp < 7) {
    unsigned acc = 0;
    const int w = warp - 3;
    for (size_t o = (size_t)w * 8 * 512; o < bytes_h; o += 4 * 8 * 512) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(hsrc + o + u * 512 + lane * 16));
#pragma unroll
      for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].w;
    }
    if (acc == 0x1234567u) sink[0] = acc;
    if (lane == 0 && w == 0) t_out[blockIdx.x * 2] = gtimer() - t0;
  }
  if (hmode == 3 && warp == 3) {
    for (size_t o = (size_t)lane * 128; o < bytes_h; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(hsrc + o));
    if (lane == 0) t_out[blockIdx.x * 2] = gtimer() - t0;
  }
  if (same_region) {
    const uint64_t ts = gtimer();
    while (gtimer() - ts < (uint64_t)delay_ns) {}
    __syncthreads();
  }
  const uint64_t t1 = gtimer();
  const int n_st = (int)(bytes_l / 32768);
  if (lmode == 1) {
    if (threadIdx.x == 0) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < iters_l; ++it)
        for (int i = 0; i < n_st; ++i) {
          mbar_wait(&empty[st], ph ^ 1);
          mbar_expect(&full[st], 32768);
          bulk_load(smem + (size_t)st * 32768, lsrc + (size_t)i * 32768, 32768, &full[st]);
          if (++st == 4) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < iters_l; ++it)
        for (int i = 0; i < n_st; ++i) {
          mbar_wait(&full[st], ph);
          mbar_arrive(&empty[st]);
          if (++st == 4) { st = 0; ph ^= 1; }
        }
      t_out[blockIdx.x * 2 + 1] = gtimer() - t1;
    }
  }
  if (lmode == 2 && warp >= 1 && warp < 3) {
    // 64 threads, each stage: 32768 / 16 / 64 = 32 cp.async per thread; 3 stages in flight via commit groups
    const int t = threadIdx.x - 32;
    const int total = iters_l * n_st;
    for (int g = 0; g < total + 3; ++g) {
      if (g < total) {
        const int i = g % n_st, st = g & 3;
        const uint8_t* s = lsrc + (size_t)i * 32768;
        const uint32_t d = smem_u32(smem + (size_t)st * 32768);
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
          const int off = (k * 64 + t) * 16;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + off), "l"(s + off) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 3;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (t == 0) t_out[blockIdx.x * 2 + 1] = gtimer() - t1;
  }
}

__global__ void flush_kernel(uint4* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s SMs %d L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  size_t big = (size_t)6 << 30;
  uint8_t* buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));
  uint4* fl; size_t fl_bytes = (size_t)512 << 20; CK(cudaMalloc(&fl, fl_bytes));
  unsigned long long* t_dev; CK(cudaMalloc(&t_dev, 1024 * 8));
  unsigned long long t_host[256]; static unsigned long long t_host2[1024];
  CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  auto run = [&](const char* name, int ctas, size_t stride, size_t per, int iters, int NS, int SB, int pf, int delay, bool flush) {
    if (flush) { flush_kernel<<<1184, 256>>>(fl, fl_bytes / 16); CK(cudaDeviceSynchronize()); }
    CK(cudaEventRecord(e0));
    ring_kernel<<<ctas, 64, NS * SB>>>(buf, stride, per, iters, NS, SB, pf, delay, t_dev);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    CK(cudaMemcpy(t_host, t_dev, ctas * 8, cudaMemcpyDeviceToHost));
    double mx = 0, av = 0; for (int i = 0; i < ctas; ++i) { mx = t_host[i] > mx ? t_host[i] : mx; av += t_host[i]; } av /= ctas;
    const double bytes = (double)per * iters;
    printf("%-34s ctas %3d NS %2d SB %5d per-CTA %7.0f KB x%3d : stream time avg %8.2f us max %8.2f us -> %6.1f GB/s per CTA, %7.1f GB/s total (kernel %.3f ms)\n",
           name, ctas, NS, SB, per / 1024.0, iters, av / 1e3, mx / 1e3, bytes / av, bytes * ctas / mx, ms);
  };
  if (!getenv("MB2_SKIP_AD")) {
  // A: private L2-resident regions (256 KB per CTA, re-read 64 times; first pass warms L2)
  for (int ctas : {1, 32, 128, 148})
    for (int cfg = 0; cfg < 5; ++cfg) {
      const int NSs[] = {4, 8, 2, 6, 4}, SBs[] = {32768, 16384, 32768, 32768, 16384};
      run("A private 256KB/CTA (L2 hit)", ctas, 256 << 10, 256 << 10, 64, NSs[cfg], SBs[cfg], 0, 0, false);
    }
  // B: everyone streams the SAME 2.75 MB (weights-like), 16 times
  for (int ctas : {1, 32, 128, 148}) run("B shared 2.75MB (L2 hit)", ctas, 0, 2816 << 10, 16, 4, 32768, 0, 0, false);
  // B2: 4 groups... every CTA r reads region (r % 4) of 2.75 MB (the real weight stream: 4 ranks)
  for (int ctas : {128}) run("B2 rank-shared (stride trick n/a)", ctas, 0, 2816 << 10, 16, 6, 32768, 0, 0, false);
  // C: HBM-resident private streams (8 MB per CTA, once, L2 flushed first)
  for (int ctas : {1, 32, 128, 148})
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int NSs[] = {4, 8, 6}, SBs[] = {32768, 16384, 32768};
      run("C private 8MB/CTA (HBM)", ctas, 8 << 20, 8 << 20, 1, NSs[cfg], SBs[cfg], 0, 0, true);
    }
  // D: K/V-like: 512 KB per CTA cold in HBM; no prefetch vs L2 prefetch + delay
  for (int ctas : {128}) {
    run("D cold 512KB/CTA no prefetch", ctas, 8 << 20, 512 << 10, 1, 4, 32768, 0, 0, true);
    for (int delay : {0, 2000, 5000, 10000, 20000, 40000}) {
      char nm[64]; snprintf(nm, sizeof nm, "D cold 512KB/CTA prefetch+%dus", delay / 1000);
      run(nm, ctas, 8 << 20, 512 << 10, 1, 4, 32768, 1, delay, true);
    }
    run("D cold 256KB/CTA no prefetch", ctas, 8 << 20, 256 << 10, 1, 4, 32768, 0, 0, true);
    for (int delay : {2000, 5000, 10000}) {
      char nm[64]; snprintf(nm, sizeof nm, "D cold 256KB/CTA prefetch+%dus", delay / 1000);
      run(nm, ctas, 8 << 20, 256 << 10, 1, 4, 32768, 1, delay, true);
    }
  }

  }
  {
    CK(cudaFuncSetAttribute(dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    auto run2 = [&](const char* name, int ctas, size_t bh, size_t bl, int il, int mode, int NH, int CH) {
      flush_kernel<<<1184, 256>>>(fl, fl_bytes / 16); CK(cudaDeviceSynchronize());
      // warm the hot regions (first 2 GB of buf is "hot" space, cold space starts at 3 GB)
      if (mode == 1 || mode == 2) { dual_kernel<<<ctas, 96, 4 * 32768 + 2 * CH>>>(buf, buf + ((size_t)3 << 30), 8 << 20, 0, bl, 1, 1, NH, CH, t_dev); CK(cudaDeviceSynchronize()); }
      CK(cudaMemset(t_dev, 0, 512 * 8));
      dual_kernel<<<ctas, 96, 4 * 32768 + 2 * CH>>>(buf, buf + ((size_t)3 << 30), 8 << 20, bh, bl, il, mode, NH, CH, t_dev);
      CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
      CK(cudaMemcpy(t_host2, t_dev, ctas * 16, cudaMemcpyDeviceToHost));
      double ah = 0, al = 0; for (int i = 0; i < ctas; ++i) { ah += t_host2[2 * i]; al += t_host2[2 * i + 1]; } ah /= ctas; al /= ctas;
      printf("%-30s ctas %3d NH %2d CH %5d | H %5.0f KB in %7.2f us = %6.1f GB/s/CTA | L %5.0f KB x%d in %7.2f us = %6.1f GB/s/CTA\n", name, ctas, NH, CH,
             bh / 1024.0, ah / 1e3, ah > 0 ? bh / ah : 0.0, bl / 1024.0, il, al / 1e3, al > 0 ? (double)bl * il / al : 0.0);
    };
    for (int NH : {4, 8, 16}) for (int CH : {8192, 16384}) run2("E only-H discard (HBM)", 128, 2 << 20, 0, 0, 0, NH, CH);
    run2("E only-L (L2 hit)", 128, 0, 256 << 10, 8, 1, 8, 16384);
    run2("E both concurrently", 128, 2 << 20, 256 << 10, 8, 2, 8, 16384);
    run2("E both concurrently", 128, 2 << 20, 256 << 10, 8, 2, 16, 8192);
    for (size_t kb : {256, 512, 768}) run2("E discard-prefetch then consume", 128, kb << 10, kb << 10, 1, 3, 8, 16384);
  }

  {
    CK(cudaFuncSetAttribute(path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    auto run3 = [&](const char* name, size_t bh, size_t bl, int il, int hmode, int lmode, int same, int delay) {
      const int ctas = 128;
      flush_kernel<<<1184, 256>>>(fl, fl_bytes / 16); CK(cudaDeviceSynchronize());
      if (!same && lmode) { path_kernel<<<ctas, 224, 4 * 32768 + 32768>>>(buf, buf + ((size_t)3 << 30), 8 << 20, 0, bl, 1, 0, 1, 0, 0, t_dev, sink); CK(cudaDeviceSynchronize()); }
      CK(cudaMemset(t_dev, 0, 512 * 8));
      path_kernel<<<ctas, 224, 4 * 32768 + 32768>>>(buf, buf + ((size_t)3 << 30), 8 << 20, bh, bl, il, hmode, lmode, same, delay, t_dev, sink);
      CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
      CK(cudaMemcpy(t_host2, t_dev, ctas * 16, cudaMemcpyDeviceToHost));
      double ah = 0, al = 0; for (int i = 0; i < ctas; ++i) { ah += t_host2[2 * i]; al += t_host2[2 * i + 1]; } ah /= ctas; al /= ctas;
      printf("%-44s | H %5.0f KB in %7.2f us = %6.1f GB/s/CTA | L %5.0f KB x%d in %7.2f us = %6.1f GB/s/CTA\n", name,
             bh / 1024.0, ah / 1e3, ah > 0 ? bh / ah : 0.0, bl / 1024.0, il, al / 1e3, al > 0 ? (double)bl * il / al : 0.0);
    };
    run3("F L=LDGSTS only (L2 hit)", 0, 256 << 10, 8, 0, 2, 0, 0);
    run3("F H=LDG discard only", 2 << 20, 0, 0, 2, 0, 0, 0);
    run3("F H=bulk discard + L=LDGSTS", 2 << 20, 256 << 10, 8, 1, 2, 0, 0);
    run3("F H=LDG discard + L=bulk", 2 << 20, 256 << 10, 8, 2, 1, 0, 0);
    run3("F H=LDG discard + L=LDGSTS", 2 << 20, 256 << 10, 8, 2, 2, 0, 0);
    for (int delay : {0, 3000, 6000, 12000}) {
      char nm[64]; snprintf(nm, sizeof nm, "F prefetch.global.L2 256KB, +%dus, consume bulk", delay / 1000);
      run3(nm, 256 << 10, 256 << 10, 1, 3, 1, 1, delay);
    }
    run3("F LDG-discard 256KB then consume bulk", 256 << 10, 256 << 10, 1, 2, 1, 1, 0);
  }
  printf("done\n");
  return 0;
}
