// Host-side helpers shared by the generator and the classifier engines.
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "mg_engine.h"

namespace mg {

int fail(int code, const std::string& msg) {
  set_last_error(msg);
  return code;
}

int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(MG_E_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(MG_E_ARG, "device index out of range");
  cudaDeviceProp prop;
  MG_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(MG_E_CUDA, "this library is built for sm_100a (B200) only");
  MG_CUDA_OK(cudaSetDevice(device));
  return MG_OK;
}

int make_wmaps(WMaps* w, const void* base, int N, int K) {
  static const int bns[4] = {32, 64, 128, 256};
  for (int i = 0; i < 4; ++i) MG_TRY(make_tmap_bf16_2d(&w->m[i], base, N, K, bns[i]));
  w->ok = true;
  return MG_OK;
}

}  // namespace mg
