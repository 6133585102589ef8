// Self-test of the tcgen05 primitives in umma.cuh: one 128 x 128 tile D = A W^T on the tensor cores (kind::f16), A
// either staged in shared memory (SS) or written to TMEM (TS), 1 pass (plain fp16) or 3 passes (split fp16, ~fp32).
// Used by tests/test_gpu_umma.py to pin the descriptor encodings before the fused kernels rely on them.
#include "common.cuh"
#include "umma.cuh"

namespace pp {

using namespace umma;

constexpr int kSelfKC = 32;  // k-columns per chunk

// Same tile through kind::f16: fp16 (hi, lo) operand pairs, 16 k-columns per instruction, packed TMEM A operand.
constexpr uint32_t kSlotBytes16 = 128 * kSelfKC * 2;

__global__ void __launch_bounds__(192, 1)
umma_selftest_f16_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ D, int K,
                         int passes, int ts_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + kSlotBytes16;
  uint8_t* b_hi = smem + 2 * kSlotBytes16;
  uint8_t* b_lo = smem + 3 * kSlotBytes16;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * kSlotBytes16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4 * kSlotBytes16 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t acc = tmem;             // columns [0,128)
  const uint32_t a_t_hi = tmem + 128;    // columns [128,144) : packed A chunk (hi) when ts_mode
  const uint32_t a_t_lo = tmem + 144;    // columns [144,160)
  const uint32_t idesc = idesc_f16(128, 128);
  uint32_t phase = 0;

  for (int k0 = 0; k0 < K; k0 += kSelfKC) {
    if (tid < 128) {
      const int row = tid;
      const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
      uint32_t vh[16], vl[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) {   // 8 k-columns = one 16-byte core-matrix row
        const int off = u * (128 * 16) + (row >> 3) * 128 + (row & 7) * 16;  // bytes
        uint4 h, l;
        float4 x = *reinterpret_cast<const float4*>(A + (size_t)row * K + k0 + u * 8);
        float4 y = *reinterpret_cast<const float4*>(A + (size_t)row * K + k0 + u * 8 + 4);
        split_f16x2(x.x, x.y, h.x, l.x); split_f16x2(x.z, x.w, h.y, l.y);
        split_f16x2(y.x, y.y, h.z, l.z); split_f16x2(y.z, y.w, h.w, l.w);
        if (ts_mode == 1) {
          vh[u * 4] = h.x; vh[u * 4 + 1] = h.y; vh[u * 4 + 2] = h.z; vh[u * 4 + 3] = h.w;
          vl[u * 4] = l.x; vl[u * 4 + 1] = l.y; vl[u * 4 + 2] = l.z; vl[u * 4 + 3] = l.w;
        } else {
          *reinterpret_cast<uint4*>(a_hi + off) = h;
          *reinterpret_cast<uint4*>(a_lo + off) = l;
        }
        x = *reinterpret_cast<const float4*>(W + (size_t)row * K + k0 + u * 8);
        y = *reinterpret_cast<const float4*>(W + (size_t)row * K + k0 + u * 8 + 4);
        split_f16x2(x.x, x.y, h.x, l.x); split_f16x2(x.z, x.w, h.y, l.y);
        split_f16x2(y.x, y.y, h.z, l.z); split_f16x2(y.z, y.w, h.w, l.w);
        *reinterpret_cast<uint4*>(b_hi + off) = h;
        *reinterpret_cast<uint4*>(b_lo + off) = l;
      }
      if (ts_mode == 1) {
        tmem_st16(a_t_hi + lane_base, vh);
        tmem_st16(a_t_lo + lane_base, vl);
        tmem_st_wait();
      }
      fence_async_smem();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (ts_mode == 2) {
      // "promotion": every K = 16 step starts a fresh accumulator (its three passes are the only accumulation the
      // tensor core does, and two of them are 2^-11 small); the steps are summed in fp32 with round-to-nearest by the
      // row threads.  Measures what the truncating accumulator costs (tests/test_gpu_umma.py).
      for (int kk = 0; kk < kSelfKC; kk += 16) {
        if (tid == 128) {
          for (int p = 0; p < passes; ++p) {
            const uint8_t* as = (p == 2) ? a_lo : a_hi;
            const uint8_t* bs = (p == 1) ? b_lo : b_hi;
            mma_f16_ss(acc, smem_desc(smem_u32(as) + (kk / 8) * 2048, 2048, 128),
                       smem_desc(smem_u32(bs) + (kk / 8) * 2048, 2048, 128), idesc, p > 0 ? 1u : 0u);
          }
          mma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        fence_after_sync();
        if (tid < 128) {
          const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
            uint32_t v[32];
            tmem_ld32(acc + lane_base + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float* d = D + (size_t)tid * 128 + c + j;
              *d = (k0 == 0 && kk == 0) ? __uint_as_float(v[j]) : *d + __uint_as_float(v[j]);
            }
          }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
      }
      continue;
    }
    if (tid == 128) {
      for (int p = 0; p < passes; ++p) {
        const uint8_t* as = (p == 2) ? a_lo : a_hi;
        const uint8_t* bs = (p == 1) ? b_lo : b_hi;
        const uint32_t at = (p == 2) ? a_t_lo : a_t_hi;
#pragma unroll
        for (int kk = 0; kk < kSelfKC; kk += 16) {
          uint64_t bd = smem_desc(smem_u32(bs) + (kk / 8) * 2048, 2048, 128);
          uint32_t accum = (k0 > 0 || p > 0 || kk > 0) ? 1u : 0u;
          if (ts_mode == 1) {
            mma_f16_ts(acc, at + kk / 2, bd, idesc, accum);
          } else {
            uint64_t ad = smem_desc(smem_u32(as) + (kk / 8) * 2048, 2048, 128);
            mma_f16_ss(acc, ad, bd, idesc, accum);
          }
        }
      }
      mma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after_sync();
  }

  if (tid < 128 && ts_mode != 2) {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t v[32];
      tmem_ld32(acc + lane_base + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c + j] = __uint_as_float(v[j]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem);
}

}  // namespace pp

// Diagnostics: D[128][128] = A[128][K] * W[128][K]^T on tcgen05 (K % 32 == 0)
// Same through kind::f16 with fp16 (hi, lo) pairs: passes 1 = plain fp16 inputs, 3 = split fp16 (~fp32).
extern "C" int pp_selftest_umma_f16(const float* A, const float* W, float* D, int64_t K, int64_t passes,
                                    int64_t ts_mode, cudaStream_t stream) {
  PP_REQUIRE(A && W && D, "null pointer");
  PP_REQUIRE(K > 0 && K % 32 == 0, "K must be a positive multiple of 32");
  PP_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
  size_t smem = 4 * pp::kSlotBytes16 + 64;
  pp::umma_selftest_f16_kernel<<<1, 192, smem, stream>>>(A, W, D, (int)K, (int)passes, (int)ts_mode);
  return pp::check_launch("pp_selftest_umma_f16");
}

// ---- TMA gather4 probe: four rows of a row-major fp32 matrix [rows][128] -> one 4 x 32-float tile in shared memory
#include <cuda.h>
namespace pp {
__global__ void gather4_probe_kernel(const __grid_constant__ CUtensorMap tm, int col, int r0, int r1, int r2, int r3,
                                     float* out) {
  __shared__ __align__(1024) uint8_t tile[1024];
  __shared__ __align__(8) uint64_t bar;
  using namespace umma;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) reinterpret_cast<float*>(tile)[i] = -1.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 4 * 128);
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
            smem_u32(tile)),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = reinterpret_cast<float*>(tile)[i];
}
}  // namespace pp

// Diagnostics: gathers rows r[0..3] (32 floats starting at column `col`) of src [rows][128] with one TMA gather4 copy
// (128-byte swizzle, tensor-map box {32, box_rows}) and returns the raw 1 KB of shared memory in out[256].
extern "C" int pp_selftest_gather4(const float* src, int64_t rows, int64_t box_rows, int64_t col, int64_t r0, int64_t r1,
                                   int64_t r2, int64_t r3, float* out, cudaStream_t stream) {
  PP_REQUIRE(src && out && rows > 0, "bad arguments");
  using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 1;
  alignas(64) CUtensorMap tm;
  const cuuint64_t gdim[2] = {128, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {512};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult rc = reinterpret_cast<Encode>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(src), gdim, gstr,
                                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    snprintf(pp::g_last_error, sizeof(pp::g_last_error), "pp_selftest_gather4: cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return 1;
  }
  pp::gather4_probe_kernel<<<1, 64, 0, stream>>>(tm, (int)col, (int)r0, (int)r1, (int)r2, (int)r3, out);
  return pp::check_launch("pp_selftest_gather4");
}
