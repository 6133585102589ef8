// Shared definitions for libpackppi_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define PP_H 128          // hidden width (configs/model/model_cfg/MpnnNet.yaml: hidden_dim)
#define PP_KMAX 32        // neighbours per residue (configs/model/encoder_cfg/ProteinEncoder.yaml: top_k)
#define PP_NPTS 8         // IPMP points per residue (n_points)
#define PP_GEO_STRIDE 28  // per-residue geometry record: R(9) t(3) N CA C O CB (15) pad(1)
#define PP_TBL_STRIDE 136 // per-residue-type table record, see packppi_b200/tables.py packed_geometry()

#define PP_PI_F 3.14159274101257324f      // float32(np.pi)
#define PP_TWO_PI_F 6.28318548202514648f  // float32(2*np.pi)

namespace pp {

extern thread_local char g_last_error[512];

inline int fail(const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b, c);
  return 1;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  snprintf(g_last_error, sizeof(g_last_error), "%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
  return 1;
}

#define PP_REQUIRE(cond, msg)                                                      \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      snprintf(pp::g_last_error, sizeof(pp::g_last_error), "%s: %s", __func__, msg); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

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

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// torch.nan_to_num for the values that occur on this path (NaN -> 0, +-inf -> +-FLT_MAX)
__device__ __forceinline__ float nan_to_num(float v) {
  if (v != v) return 0.f;
  if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

}  // namespace pp
