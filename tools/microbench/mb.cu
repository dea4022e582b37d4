// Microbenchmarks that size the decode-step design: L2 / HBM read bandwidth (all SMs, per SM), grid-barrier
// and cluster-barrier latency, dependent-launch latency inside a CUDA graph.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void read_kernel(const uint4* __restrict__ p, size_t n_vec, int iters, unsigned long long* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n_vec; i += 4 * stride) {
      uint4 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
      acc += a.x ^ b.y ^ c.z ^ d.w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// each CTA reads its own private contiguous region (per-SM ingest test)
__global__ void read_private_kernel(const uint4* __restrict__ p, size_t vec_per_cta, int iters, unsigned long long* sink) {
  unsigned acc = 0;
  const uint4* q = p + (size_t)blockIdx.x * vec_per_cta;
  for (int it = 0; it < iters; ++it)
    for (size_t i = threadIdx.x; i + 3 * blockDim.x < vec_per_cta; i += 4 * blockDim.x) {
      uint4 a = q[i], b = q[i + blockDim.x], c = q[i + 2 * blockDim.x], d = q[i + 3 * blockDim.x];
      acc += a.x ^ b.y ^ c.z ^ d.w;
    }
  if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void grid_barrier_kernel(unsigned* counter, int rounds, long long* cycles) {
  long long t0 = clock64();
  for (int r = 1; r <= rounds; ++r) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned target = (unsigned)r * gridDim.x;
      while (*((volatile unsigned*)counter) < target) {}
      __threadfence();
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = clock64() - t0;
}

__global__ void __cluster_dims__(8, 1, 1) cluster_barrier_kernel(int rounds, long long* cycles) {
  cg::cluster_group cl = cg::this_cluster();
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) cl.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = clock64() - t0;
}

__global__ void tiny_kernel(float* x) { if (threadIdx.x == 0 && blockIdx.x == 0) x[0] += 1.0f; }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s SMs %d clock %d kHz L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.clockRate, prop.l2CacheSize >> 20);
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  // ---- bandwidth ----
  size_t big = (size_t)4 << 30;
  uint4* buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));
  for (size_t mb : {16, 32, 64, 96, 256, 2048}) {
    size_t n_vec = (mb << 20) / 16;
    int iters = mb <= 96 ? 50 : (mb <= 256 ? 20 : 4);
    for (int ctas : {148, 296, 592}) {
      read_kernel<<<ctas, 512>>>(buf, n_vec, 2, sink);
      CK(cudaEventRecord(e0)); read_kernel<<<ctas, 512>>>(buf, n_vec, iters, sink); CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("read all-SMs   %5zu MB  ctas %4d x512 : %8.1f GB/s\n", mb, ctas, (double)(mb << 20) * iters / ms / 1e6);
    }
  }
  for (int ctas : {1, 8, 32, 64, 128, 148}) {
    size_t per = (size_t)(256 << 10) / 16;   // 256 KB private region per CTA (L2 resident overall)
    read_private_kernel<<<ctas, 512>>>(buf, per, 2, sink);
    CK(cudaEventRecord(e0)); read_private_kernel<<<ctas, 512>>>(buf, per, 200, sink); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("read private 256KB/CTA (L2 hit) ctas %4d : %8.1f GB/s total, %7.1f GB/s per CTA\n", ctas,
           (double)(256 << 10) * 200 * ctas / ms / 1e6, (double)(256 << 10) * 200 / ms / 1e6);
  }
  for (int ctas : {64, 128, 148}) {   // HBM-resident private streams
    size_t per = (size_t)(16 << 20) / 16;
    CK(cudaEventRecord(e0)); read_private_kernel<<<ctas, 512>>>(buf, per, 1, sink); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("read private 16MB/CTA (HBM)     ctas %4d : %8.1f GB/s total, %7.1f GB/s per CTA\n", ctas,
           (double)(16 << 20) * ctas / ms / 1e6, (double)(16 << 20) / ms / 1e6);
  }
  // ---- grid barrier ----
  unsigned* counter; CK(cudaMalloc(&counter, 4));
  long long* cyc; CK(cudaMalloc(&cyc, 8));
  for (int ctas : {16, 74, 128, 148}) {
    CK(cudaMemset(counter, 0, 4));
    int rounds = 2000;
    void* args[] = {&counter, &rounds, &cyc};
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((void*)grid_barrier_kernel, dim3(ctas), dim3(256), args, 0, 0));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("grid barrier %4d CTAs: %.3f us per barrier (%lld cycles)\n", ctas, ms * 1e3 / rounds, h / rounds);
  }
  // ---- cluster barrier ----
  {
    int rounds = 5000;
    cluster_barrier_kernel<<<128, 256>>>(rounds, cyc);
    CK(cudaEventRecord(e0)); cluster_barrier_kernel<<<128, 256>>>(rounds, cyc); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("cluster(8) barrier: %.3f us (%lld cycles)\n", ms * 1e3 / rounds, h / rounds);
    int nclusters = 0;
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(128); cfg.blockDim = dim3(256);
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 8; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    CK(cudaOccupancyMaxActiveClusters(&nclusters, (void*)cluster_barrier_kernel, &cfg));
    printf("max active clusters of 8 (256 thr, no smem): %d\n", nclusters);
    CK(cudaFuncSetAttribute((void*)cluster_barrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cfg.dynamicSmemBytes = 200 * 1024;
    CK(cudaOccupancyMaxActiveClusters(&nclusters, (void*)cluster_barrier_kernel, &cfg));
    printf("max active clusters of 8 (256 thr, 200 KB smem): %d\n", nclusters);
  }
  // ---- dependent launches in a graph ----
  {
    float* x; CK(cudaMalloc(&x, 4)); CK(cudaMemset(x, 0, 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < 100; ++i) tiny_kernel<<<64, 128, 0, s>>>(x);
    CK(cudaStreamEndCapture(s, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, s)); CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s)); for (int i = 0; i < 10; ++i) CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("graph of 100 dependent tiny kernels: %.3f us per kernel\n", ms * 1e3 / 1000);
    CK(cudaEventRecord(e0, s)); for (int i = 0; i < 1000; ++i) tiny_kernel<<<64, 128, 0, s>>>(x); CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("stream of 1000 tiny kernels (no graph): %.3f us per kernel\n", ms * 1e3 / 1000);
  }
  printf("done\n");
  return 0;
}
