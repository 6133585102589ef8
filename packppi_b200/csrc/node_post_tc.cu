// Per-residue node update of an IPMP layer on the tensor cores with fp32-grade accumulation, sm_100a.
//
// Same mathematics as node_post_kernel in mpnn.cu (reference layers.py:127-132):
//   e   = LN0(h_V + W3 mean_k(msg) + b3 mean_k(mask))
//   h_V = mask * LN1(e + W_out relu(W_in e + b_in) + b_out)
// The tensor core accumulates in fp32 with TRUNCATION (measured: -7e-7 relative bias at K = 128, growing linearly
// with K), which is harmless for the per-edge kernels but not for h_V, the state every later GEMM of the step reads
// (chi error 1.2e-4 rad after two steps with a plain TMEM accumulation).  Here every K = 16 step of a GEMM goes into
// a FRESH accumulator - the three MMAs of the split-fp16 product (hi*hi, hi*lo, lo*hi) are the only accumulation the
// tensor core performs, and two of them are 2^-11 small - and the row threads add the steps up in fp32 registers
// with round-to-nearest ("promotion").  Measured on a K = 512 product of mixed-sign data: rms error 3.1e-7 of the mean
// magnitude (sequential fp32 FMA: 5.1e-7; plain TMEM accumulation: 2.4e-6).
//
// One tile = 128 residue rows.  Two 128-column TMEM buffers alternate between "being written by the MMAs of step
// t + 1" and "being read by the row threads for step t".  TMEM also holds e in fp32 (residual) and as packed fp16
// (hi | lo), the A operand of FFN-in.  The running FFN-out sum is parked in shared memory while a hidden slice is
// accumulated, so a thread never holds more than two 64-column accumulators.
//   warps 0-7  row workers: thread (grp, m) = row m, column chunks {grp, grp + 2}
//   warp 8     MMA issue (converged warp, one elected lane)
//   warp 9     weight loader (cp.async.bulk of the path-2 operand images of weights.pack_tc_stream)
#include "common.cuh"
#include "umma.cuh"
#include "weights_layout.h"

namespace pp {
namespace post {

using namespace umma;

constexpr int kRows = 128, kKC = 32, kSA = 4, kSB = 4;
#ifndef PP_POST_PROMOTE
#define PP_POST_PROMOTE 2
#endif
// K = 16 steps accumulated in TMEM before the row threads take the partial sum over (1, 2 or 4): every hand-off is a
// round trip between the MMA warp and 256 threads, every extra step one more truncated accumulation
constexpr int kPromote = PP_POST_PROMOTE;
static_assert(kPromote == 1 || kPromote == 2 || kPromote == 4, "promotion interval");
constexpr uint32_t kImgBytes = kRows * kKC * 2;
constexpr uint32_t kSlotBytes = 2 * kImgBytes;
constexpr uint32_t kLbo = kRows * 16, kSbo = 128;
constexpr int kThreads = 320;
constexpr uint32_t kParkBytes = kRows * 128 * 4;  // the running FFN-out sum of the tile, one row per thread pair
constexpr int kNumBars = 2 * kSA + 2 * kSB + 2 + 2 + 1;
// per-column parameters: b3, LN0 gain / bias, b_in (512), b_out, LN1 gain / bias
constexpr int kP_B3 = 0, kP_LN0G = 128, kP_LN0B = 256, kP_BIN = 384, kP_BOUT = 896, kP_LN1G = 1024, kP_LN1B = 1152,
              kParamFloats = 1280;
constexpr int kRedFloats = 4 * 2 * 128;
constexpr size_t kSmem = kParkBytes + (size_t)(kSA + kSB) * kSlotBytes + kNumBars * 8 + 32 + (kParamFloats + kRedFloats) * 4;
// stream of path 2 (weights.pack_tc_stream): W3 (4 chunks), then 4-chunk blocks in0 in1 out0 in2 out1 in3 out2 out3
constexpr long long kImageFloats = 2LL * 128 * (176 + 128 + 128 + 4 * 256) * 2 / 4;

struct Args {
  int G, R;
  const float* wstream;
  const float *B3, *LN0G, *LN0B, *BIN, *BOUT, *LN1G, *LN1B;
  const float *rmask, *msum;  // [G]
  const float* accsum;        // [R][128] summed messages
  float in_scale;             // 1 / K
  float* hV;                  // [R][128], updated in place
  int* overflow;              // optional overflow flag (umma.cuh: report_overflow)
  const int* live_list;       // optional: ids of the live 128-row tiles ...
  const int* n_live;          // ... and their number (device scalar)
};

__device__ __forceinline__ void put_chunk(uint8_t* slot, int m, const float* v, float& amax) {
  const int base = (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    uint4 h, l;
    split_f16x2(v[u * 8 + 0], v[u * 8 + 1], h.x, l.x, amax); split_f16x2(v[u * 8 + 2], v[u * 8 + 3], h.y, l.y, amax);
    split_f16x2(v[u * 8 + 4], v[u * 8 + 5], h.z, l.z, amax); split_f16x2(v[u * 8 + 6], v[u * 8 + 7], h.w, l.w, amax);
    *reinterpret_cast<uint4*>(slot + u * kLbo + base) = h;
    *reinterpret_cast<uint4*>(slot + kImgBytes + u * kLbo + base) = l;
  }
}

__global__ void __launch_bounds__(kThreads, 1) node_post_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* park = smem;
  uint8_t* Aring = park + kParkBytes;
  uint8_t* Bring = Aring + kSA * kSlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bring + kSB * kSlotBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kSA;
  uint64_t* b_full = a_empty + kSA;
  uint64_t* b_empty = b_full + kSB;
  uint64_t* step_full = b_empty + kSB;   // [2] the MMAs of a K = 16 step have completed
  uint64_t* step_free = step_full + 2;   // [2] the row threads have read the buffer
  uint64_t* e_ready = step_free + 2;     // e is in TMEM (fp32 and packed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_ready + 1);
  float* prm = reinterpret_cast<float*>(e_ready + 3);
  float* red = prm + kParamFloats;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.R;
  const int ntiles = (R + kRows - 1) / kRows;
  // with a compacted list of the live 128-row tiles (tiles that hold at least one unmasked residue) every role walks
  // list positions instead of tile numbers: padding tiles of a ragged batch are never touched (their output rows keep
  // whatever they held; nothing unmasked reads them)
  const int nwork = a.live_list ? *a.n_live : ntiles;

  if (tid == 0) {
    for (int i = 0; i < kSA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&step_full[i], 1); mbar_init(&step_free[i], 256); }
    mbar_init(e_ready, 256);
    mbar_fence_init();
  }
  {
    const float* src[7] = {a.B3, a.LN0G, a.LN0B, a.BIN, a.BOUT, a.LN1G, a.LN1B};
    const int off[8] = {kP_B3, kP_LN0G, kP_LN0B, kP_BIN, kP_BOUT, kP_LN1G, kP_LN1B, kParamFloats};
    for (int t = 0; t < 7; ++t)
      for (int i = tid; i < off[t + 1] - off[t]; i += kThreads) prm[off[t] + i] = src[t][i];
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t BUF0 = tmem, BUF1 = tmem + 128, RE = tmem + 256, PK = tmem + 384;

  if (warp == 9) {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      // consumption order W3, in0, out0, in1, out1, ... as block indices of the stream (after W3)
      const int order[8] = {0, 2, 1, 4, 3, 6, 5, 7};
      int idx = 0;
      uint32_t phase = 1;
      const uint8_t* base = reinterpret_cast<const uint8_t*>(a.wstream);
      for (int tile = blockIdx.x; tile < nwork; tile += gridDim.x) {
        for (int blk = -1; blk < 8; ++blk) {
          const uint8_t* src = base + (blk < 0 ? 0 : (size_t)(1 + order[blk]) * 4 * kSlotBytes);
          for (int c = 0; c < 4; ++c) {
            mbar_wait(&b_empty[idx], phase);
            mbar_arrive_expect_tx(&b_full[idx], kSlotBytes);
            bulk_g2s(Bring + idx * kSlotBytes, src + (size_t)c * kSlotBytes, kSlotBytes, &b_full[idx]);
            if (++idx == kSB) { idx = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer (converged warp)
    constexpr uint32_t kIdesc = idesc_f16(128, 128);
    constexpr uint32_t kHi = desc_hi(kSbo);
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0, pe = 0;
    uint32_t t = 0;  // promoted groups issued so far: buffer t & 1, use number t >> 1 of that buffer
    int sub = 0;     // K = 16 steps already accumulated in the current group
    // one 32-column chunk = two promoted steps; ss: A from the ring, else from TMEM at a_tm (packed, lo at + 64)
    auto chunk = [&](bool ss, uint32_t a_tm) {
      if (ss) mbar_wait(&a_full[ia], pa);
      mbar_wait(&b_full[ib], pb);
      fence_after_sync();
      const uint32_t a_lo = desc_lo(smem_u32(Aring + ia * kSlotBytes), kLbo);
      const uint32_t b_lo = desc_lo(smem_u32(Bring + ib * kSlotBytes), kLbo);
#pragma unroll
      for (int kk = 0; kk < kKC; kk += 16) {
        const uint32_t buf = t & 1;
        if (sub == 0) {
          mbar_wait(&step_free[buf], ((t >> 1) & 1) ^ 1);
          fence_after_sync();
        }
        const bool last = sub + 1 == kPromote;
        if (elect_one_sync()) {
          const uint32_t acc = buf ? BUF1 : BUF0;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const uint32_t ad = (((p == 2) ? kImgBytes : 0u) + (kk / 8) * kLbo) >> 4;
            const uint32_t bd = (((p == 1) ? kImgBytes : 0u) + (kk / 8) * kLbo) >> 4;
            const uint32_t accum = (sub > 0 || p > 0) ? 1u : 0u;
            if (ss) mma_f16_ss2(acc, a_lo + ad, b_lo + bd, kHi, kIdesc, accum);
            else mma_f16_ts2(acc, a_tm + ((p == 2) ? 64 : 0) + kk / 2, b_lo + bd, kHi, kIdesc, accum);
          }
          if (last) mma_commit(&step_full[buf]);
          if (kk == 16) {
            mma_commit(&b_empty[ib]);
            if (ss) mma_commit(&a_empty[ia]);
          }
        }
        __syncwarp();
        if (last) { sub = 0; ++t; } else { ++sub; }
      }
      if (++ib == kSB) { ib = 0; pb ^= 1; }
      if (ss && ++ia == kSA) { ia = 0; pa ^= 1; }
    };
    for (int tile = blockIdx.x; tile < nwork; tile += gridDim.x) {
      for (int c = 0; c < 4; ++c) chunk(true, 0);  // W3
      mbar_wait(e_ready, pe); pe ^= 1;
      fence_after_sync();
      for (int j = 0; j < 4; ++j) {
        for (int c = 0; c < 4; ++c) chunk(false, PK + c * 16);  // FFN-in slice j: A = e (TMEM)
        for (int c = 0; c < 4; ++c) chunk(true, 0);             // FFN-out slice j: A = hidden slice j (ring)
      }
    }
  } else {
    // ------------------------------------------------------------------ row workers
    const int grp = tid >> 7, m = tid & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float* wsc = a.wstream + kImageFloats;  // 1 / scale of G1, G2, G3 (= W3), FFN-in, FFN-out
    const float s3 = wsc[2], sFI = wsc[3], sFO = wsc[4];
    uint32_t t = 0;
    float amax = 0.f;  // overflow report, see umma.cuh
    int q = 0;  // A chunks published so far by this thread's group schedule (kernel-wide chunk counter)
    // this thread's 16-byte units of the parked row: unit u of chunk c at row m, swizzled against bank conflicts
    uint8_t* const prow = park + m * 512;
    auto punit = [&](int c, int u) { return prow + (((c * 8 + u) ^ (m & 7)) << 4); };

    auto publish = [&](int qq, const float* vals) {
      const int slot = qq % kSA;
      mbar_wait(&a_empty[slot], ((qq / kSA) & 1) ^ 1);
      put_chunk(Aring + slot * kSlotBytes, m, vals, amax);
      fence_async_smem();
      mbar_arrive(&a_full[slot]);
    };
    // add the 8 / kPromote partial sums of one K = 128 product into acc (this thread's 2 x 32 columns)
    auto drain = [&](float (&acc)[2][32]) {
#pragma unroll 1
      for (int s = 0; s < 8 / kPromote; ++s, ++t) {
        const uint32_t buf = t & 1;
        mbar_wait(&step_full[buf], (t >> 1) & 1);
        fence_after_sync();
        const uint32_t base = (buf ? BUF1 : BUF0) + lane_base;
        uint32_t u0[32], u1[32];
        tmem_ld32(base + grp * 32, u0);
        tmem_ld32(base + (grp + 2) * 32, u1);
        tmem_ld_wait();
        fence_before_sync();
        mbar_arrive(&step_free[buf]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          acc[0][i] += __uint_as_float(u0[i]);
          acc[1][i] += __uint_as_float(u1[i]);
        }
      }
    };
    auto row_total = [&](float partial, int which) -> float {
      red[(which * 2 + grp) * 128 + m] = partial;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      return red[(which * 2) * 128 + m] + red[(which * 2 + 1) * 128 + m];
    };
    auto layer_norm = [&](float (&x)[2][32], int which, const float* gain, const float* bias) {
      float sum = 0.f;
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int i = 0; i < 32; ++i) sum += x[tt][i];
      const float mean = row_total(sum, which) * (1.f / 128.f);
      float var = 0.f;
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int i = 0; i < 32; ++i) { float d = x[tt][i] - mean; var += d * d; }
      const float rstd = rsqrtf(row_total(var, which + 1) * (1.f / 128.f) + 1e-5f);
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int c = grp + 2 * tt;
#pragma unroll
        for (int i = 0; i < 32; ++i) x[tt][i] = (x[tt][i] - mean) * rstd * gain[c * 32 + i] + bias[c * 32 + i];
      }
    };

    for (int pos = blockIdx.x; pos < nwork; pos += gridDim.x) {
      const int tile = a.live_list ? a.live_list[pos] : pos;
      const int r = tile * kRows + m;
      const bool in = r < R;
      const int rr = min(r, R - 1);
      const int g = rr % a.G;
      const bool on = in && a.rmask[g] != 0.f;
      float acc[2][32];
      float v[32];
      // ---- summed messages / K -> A operand of W3 (the mean over K commutes with W3)
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const float* src = a.accsum + (size_t)rr * 128 + (grp + 2 * tt) * 32;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 x = in ? *reinterpret_cast<const float4*>(src + u * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[u * 4] = x.x * a.in_scale; v[u * 4 + 1] = x.y * a.in_scale;
          v[u * 4 + 2] = x.z * a.in_scale; v[u * 4 + 3] = x.w * a.in_scale;
        }
        publish(q + grp + 2 * tt, v);
      }
      q += 4;
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[tt][i] = 0.f;
      drain(acc);
      // ---- e = LN0(h_V + W3 mean + b3 mean_mask)
      {
        const float ms = a.msum[g];
        const float* hv = a.hV + (size_t)rr * 128;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int c = grp + 2 * tt;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 h = *reinterpret_cast<const float4*>(hv + c * 32 + u * 4);
            const float* b = prm + kP_B3 + c * 32 + u * 4;
            acc[tt][u * 4 + 0] = h.x + fmaf(acc[tt][u * 4 + 0], s3, b[0] * ms);
            acc[tt][u * 4 + 1] = h.y + fmaf(acc[tt][u * 4 + 1], s3, b[1] * ms);
            acc[tt][u * 4 + 2] = h.z + fmaf(acc[tt][u * 4 + 2], s3, b[2] * ms);
            acc[tt][u * 4 + 3] = h.w + fmaf(acc[tt][u * 4 + 3], s3, b[3] * ms);
          }
        }
        layer_norm(acc, 0, prm + kP_LN0G, prm + kP_LN0B);
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int c = grp + 2 * tt;
          uint32_t ef[32], eh[16], el[16];
#pragma unroll
          for (int i = 0; i < 32; ++i) ef[i] = __float_as_uint(acc[tt][i]);
#pragma unroll
          for (int i = 0; i < 16; ++i) split_f16x2(acc[tt][2 * i], acc[tt][2 * i + 1], eh[i], el[i], amax);
          tmem_st32(RE + lane_base + c * 32, ef);
          tmem_st16(PK + lane_base + c * 16, eh);
          tmem_st16(PK + 64 + lane_base + c * 16, el);
        }
        tmem_st_wait();
        fence_before_sync();
        mbar_arrive(e_ready);
      }
      // ---- FFN: y = sum_j W_out[:, j] relu(W_in[j] e + b_in[j]); the running sum y is parked in shared memory
      //      while the hidden slice is accumulated
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[tt][i] = 0.f;
        drain(acc);  // FFN-in slice j
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int c = grp + 2 * tt;
          const float* b = prm + kP_BIN + j * 128 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(acc[tt][i], sFI, b[i]), 0.f);
          publish(q + c, v);
        }
        q += 4;
        // running sum back into registers (zero for the first slice)
#pragma unroll
        for (int tt = 0; tt < 2; ++tt)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 y = j == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(punit(grp + 2 * tt, u));
            acc[tt][u * 4] = y.x; acc[tt][u * 4 + 1] = y.y; acc[tt][u * 4 + 2] = y.z; acc[tt][u * 4 + 3] = y.w;
          }
        drain(acc);  // FFN-out slice j
        if (j < 3) {
#pragma unroll
          for (int tt = 0; tt < 2; ++tt)
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<float4*>(punit(grp + 2 * tt, u)) =
                  make_float4(acc[tt][u * 4], acc[tt][u * 4 + 1], acc[tt][u * 4 + 2], acc[tt][u * 4 + 3]);
        }
      }
      // ---- h_V = mask * LN1(e + y + b_out)
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int c = grp + 2 * tt;
        uint32_t ef[32];
        tmem_ld32(RE + lane_base + c * 32, ef);
        tmem_ld_wait();
        const float* b = prm + kP_BOUT + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[tt][i] = __uint_as_float(ef[i]) + fmaf(acc[tt][i], sFO, b[i]);
      }
      layer_norm(acc, 2, prm + kP_LN1G, prm + kP_LN1B);
      if (in) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          float4* o = reinterpret_cast<float4*>(a.hV + (size_t)r * 128 + (grp + 2 * tt) * 32);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            o[u] = on ? make_float4(acc[tt][u * 4], acc[tt][u * 4 + 1], acc[tt][u * 4 + 2], acc[tt][u * 4 + 3])
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    report_overflow(a.overflow, amax);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace post
}  // namespace pp

using namespace pp;

// Tensor-core node update with promoted (fp32-grade) accumulation: h_V <- mask * LN1(e + FFN(e)),
// e = LN0(h_V + W3 mean_k(msg) + b3 mean_k(mask))  (reference layers.py:127-132).  Same arguments as
// pp_ipmp_node_post; wstream = operand images of path 2 of this layer (weights.py: pack_tc_stream).
extern "C" int pp_ipmp_node_post_tc32(const float* weights, int64_t layer, const float* wstream, const float* msum,
                                      const float* residue_mask, int64_t G, int64_t K, int64_t S, const float* wsAcc,
                                      float* hV, int32_t* overflow, const int32_t* live_tiles, const int32_t* n_live,
                                      cudaStream_t stream) {
  PP_REQUIRE(weights && wstream && msum && residue_mask && wsAcc && hV, "null pointer");
  PP_REQUIRE(layer >= 0 && layer < 3, "layer out of range");
  PP_REQUIRE(G > 0 && S > 0 && K > 0 && K <= PP_KMAX, "bad sizes");
  const float* Lb = weights + layer * wl::kLayerStride;
  post::Args a{};
  a.G = (int)G; a.R = (int)(S * G);
  a.wstream = wstream;
  a.B3 = Lb + PP_OFF(L0_N_B3);
  a.LN0G = Lb + PP_OFF(L0_LN0_G); a.LN0B = Lb + PP_OFF(L0_LN0_B);
  a.BIN = Lb + PP_OFF(L0_NF_BIN); a.BOUT = Lb + PP_OFF(L0_NF_BOUT);
  a.LN1G = Lb + PP_OFF(L0_LN1_G); a.LN1B = Lb + PP_OFF(L0_LN1_B);
  a.rmask = residue_mask; a.msum = msum;
  a.accsum = wsAcc; a.in_scale = 1.f / (float)K;
  a.hV = hV;
  a.overflow = overflow;
  a.live_list = live_tiles;
  a.n_live = live_tiles ? n_live : nullptr;
  PP_REQUIRE(!live_tiles || n_live, "live_tiles needs n_live");
  cudaError_t e = cudaFuncSetAttribute(post::node_post_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)post::kSmem);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "node_post_tc_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return 1;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (a.R + post::kRows - 1) / post::kRows;
  post::node_post_tc_kernel<<<tiles < num_sms ? tiles : num_sms, post::kThreads, post::kSmem, stream>>>(a);
  return check_launch("pp_ipmp_node_post_tc32");
}
