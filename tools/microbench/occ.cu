// How many clusters of each size are co-resident with one ~200 KB CTA per SM (cluster scheduling is per GPC).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) for (int thr : {288, 416, 512}) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(thr); cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cs; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, (void*)k, &cfg);
    printf("cluster %2d threads %3d: max active clusters %d (%d SMs) %s\n", cs, thr, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
