// sm_100a primitives for the tensor-core path: tcgen05.mma (UMMA) with TMEM accumulators, mbarriers, bulk async
// copies.  Raw PTX; the descriptor bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
//
// Operand layout used throughout: K-major, no swizzle ("interleaved" canonical layout).  A core matrix is 8 rows x 16
// bytes stored as 128 contiguous bytes; for an operand tile of ROWS rows and KC fp32 columns
//     byte_offset(row, k) = (k / 4) * (ROWS * 16) + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4
// i.e. SBO (next 8 rows) = 128 B and LBO (next 4 k-columns) = ROWS * 16 B.  A thread that owns one row writes 16-byte
// units that are 16 B apart across the lanes of its warp: conflict-free 128-bit stores.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pp {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp (the MMA issuer runs its control flow on all 32 lanes and elects the lane that issues)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol error turns into a trap (reported as a launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 26)) __trap();
  }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tensor core / bulk copy reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, 16-byte aligned addresses)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)),
               "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive columns: thread t of warp w (w % 4 selects the lane quarter) receives row 32*(w%4)+t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// 32 lanes x 16 consecutive columns (packed fp16 pairs: 32 k-elements of a TMEM A operand)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors
// shared-memory matrix descriptor, K-major, no swizzle: start address, LBO, SBO in 16-byte units, version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D = fp32, A = B = fp16 (format 0), both K-major, M x N tile; one instruction covers K = 16
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// kind::f16 operands use the same canonical layout with 8 elements per 16-byte core-matrix row:
//     byte_offset(row, k) = (k / 8) * (ROWS * 16) + (row / 8) * 128 + (row % 8) * 16 + (k % 8) * 2
// A TMEM A operand holds two consecutive k-elements per 32-bit column (even k in the low half).
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The same MMAs with the shared-memory descriptors passed as (low word, high word): the low word holds the start
// address (>> 4, 14 bits) and LBO, so stepping through an operand is ONE 32-bit add of a constant per instruction
// (the start address stays below 2^14 for any shared-memory offset), and the high word (SBO, version) is constant.
// The issuing thread's dependent instruction chain per MMA, not the tensor pipe, bounded the first version.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFF) | ((lbo_bytes >> 4) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14); }
__device__ __forceinline__ void mma_f16_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread complete -> one arrival on the mbarrier
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Two fp32 values -> packed fp16 (hi, lo) pairs, x ~= hi + lo with both halves rounded to nearest: hi carries 11
// mantissa bits, lo the next 11 (its sign lets the pair resolve 22+ bits).  hi*hi + hi*lo + lo*hi with fp32
// accumulation reproduces an fp32 product to ~2^-22 at twice the tensor-core rate of kind::tf32 (K = 16 per
// instruction).  fp16 range: |x| must stay below 65504 (an overflow gives inf - inf = NaN in lo, never a silently
// wrong number); below 2^-14 the lo half is subnormal and the pair keeps an absolute 2^-25, which is why the weight
// images are pre-scaled by a power of two (weights.py) while the O(1) activations are taken as they are.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// The same, tracking the largest magnitude that was split: an operand above kF16Max becomes inf in the hi half, the
// product NaN, and a ReLU epilogue (fmaxf ignores NaN) would turn that into a plausible 0.  The kernels therefore
// report it: a thread that saw amax > kF16Max sets the caller's overflow flag before it exits (report_overflow).
constexpr float kF16Max = 65504.f;
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo, float& amax) {
  amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
  split_f16x2(a, b, hi, lo);
}
__device__ __forceinline__ void report_overflow(int* flag, float amax) {
  if (flag != nullptr && !(amax <= kF16Max)) atomicOr(flag, 1);
}

}  // namespace umma
}  // namespace pp
