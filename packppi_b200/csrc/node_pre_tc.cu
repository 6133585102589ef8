// Per-residue prologue of an IPMP message path on the tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same mathematics as node_pre_kernel in mpnn.cu (reference layers.py:72-77,91 and the h_V_i / h_V_j columns of the
// first Linear of node_message_fn / edge_message_fn, layers.py:119-123,134-138):
//   p_local = W_p h_V + b_p (8 points x 3), norms, points in the global frame           -> wsP [R][24]
//   A_i     = W_in[:, h_V_i | own geometry] [h_V | p_local | norms] + b1                -> wsA [R][128]
//   N_j     = W_in[:, h_V_j] h_V                                                        -> wsN [R][128]
// One tile = 128 residue rows = the M dimension of the UMMAs.  The three weight matrices (160 KB as fp16 hi / lo
// images) stay resident in shared memory for the whole persistent CTA; h_V chunks go through a 3-slot ring and feed
// three accumulators at once (P: 32 columns, A and N: 128 columns each), the geometry chunk follows once the points
// are known.  Split fp16 operand pairs, 3 MMAs per product, fp32 accumulation (see mpnn_tc.cu).
//   warps 0-7  row workers (two groups of 128 threads, thread (grp, m) = row m, column chunks {grp, grp + 2})
//   warp 8     weight copy at start-up, then MMA issue (one elected lane)
#include "common.cuh"
#include "umma.cuh"
#include "weights_layout.h"

namespace pp {
namespace pre {

using namespace umma;

constexpr int kRows = 128, kKC = 32, kSA = 3;
constexpr uint32_t kImgBytes = kRows * kKC * 2;    // fp16 image (hi or lo) of a 128-row, 32-column chunk
constexpr uint32_t kSlotBytes = 2 * kImgBytes;
constexpr uint32_t kLbo = kRows * 16, kSbo = 128;
constexpr uint32_t kPImgBytes = 32 * kKC * 2;      // the 24 (padded to 32) point outputs: 32-row B operand
constexpr uint32_t kPLbo = 32 * 16;
constexpr uint32_t kWpBytes = 4 * 2 * kPImgBytes;  // 4 chunks x (hi, lo)
constexpr uint32_t kWagBytes = 5 * kSlotBytes;     // K = 160: h_V (4 chunks) + own geometry (1 chunk)
constexpr uint32_t kWnBytes = 4 * kSlotBytes;
constexpr uint32_t kWBytes = kWpBytes + kWagBytes + kWnBytes;
constexpr long long kImageFloats = kWBytes / 4;
constexpr long long kStreamFloats = kImageFloats + 8;  // + 1 / scale of W_p, W_ag, W_n
constexpr int kThreads = 288;
constexpr int kNumBars = 1 + 2 * kSA + 3;
constexpr size_t kSmem = kWBytes + kSA * kSlotBytes + kNumBars * 8 + 16 + (32 + 128) * 4;

struct Args {
  const float* geo;      // [G][PP_GEO_STRIDE] residue frames
  int G, R;
  const float* wstream;  // operand images of this layer / path, then the inverse scales
  const float *BP, *B1;
  const float* hV;       // [R][128]
  float *A, *Nn, *P;     // wsA, wsN [R][128], wsP [R][24]
  int* overflow;         // optional overflow flag (umma.cuh: report_overflow)
  const int* live_list;  // optional: ids of the live 128-row tiles ...
  const int* n_live;     // ... and their number (device scalar)
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

__global__ void __launch_bounds__(kThreads, 1) node_pre_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Wp = smem;
  uint8_t* Wag = Wp + kWpBytes;
  uint8_t* Wn = Wag + kWagBytes;
  uint8_t* Aring = Wn + kWnBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Aring + kSA * kSlotBytes);
  uint64_t* w_full = bars;
  uint64_t* a_full = w_full + 1;
  uint64_t* a_empty = a_full + kSA;
  uint64_t* p_full = a_empty + kSA;   // the point accumulator is complete
  uint64_t* an_full = p_full + 1;     // A and N accumulators are complete
  uint64_t* tile_done = an_full + 1;  // the workers have read all three accumulators
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_done + 1);
  float* prm = reinterpret_cast<float*>(tile_done + 3);  // b_p (32), b1 (128)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.R;
  const int ntiles = (R + kRows - 1) / kRows;
  // with a compacted list of the live 128-row tiles (tiles that hold at least one unmasked residue) every role walks
  // list positions instead of tile numbers: padding tiles of a ragged batch are never touched (their output rows keep
  // whatever they held; nothing unmasked reads them)
  const int nwork = a.live_list ? *a.n_live : ntiles;

  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < kSA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    mbar_init(p_full, 1);
    mbar_init(an_full, 1);
    mbar_init(tile_done, 256);
    mbar_fence_init();
  }
  for (int i = tid; i < 32; i += kThreads) prm[i] = i < 24 ? a.BP[i] : 0.f;
  for (int i = tid; i < 128; i += kThreads) prm[32 + i] = a.B1[i];
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t ACC_A = tmem, ACC_N = tmem + 128, ACC_P = tmem + 256;

  if (warp == 8) {
    {
      // ---------------------------------------------------------------- weights in, then MMA issue (the warp stays
      // converged: all lanes run the control flow, one elected lane issues)
      if (lane == 0) {
        mbar_arrive_expect_tx(w_full, kWBytes);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wstream);
        for (uint32_t off = 0; off < kWBytes; off += 16384) bulk_g2s(smem + off, src + off, 16384, w_full);
      }
      __syncwarp();
      mbar_wait(w_full, 0);
      constexpr uint32_t kIdescP = idesc_f16(128, 32), kIdescW = idesc_f16(128, 128);
      int q = 0;
      uint32_t td_phase = 0;
      // one ring slot against one resident weight chunk: hi*hi, hi*lo, lo*hi
      auto gemm = [&](uint32_t acc, uint32_t as, uint32_t bs, uint32_t b_lo, uint32_t b_lbo, uint32_t idesc, bool fresh) {
        const uint32_t a0 = desc_lo(as, kLbo), b0 = desc_lo(bs, b_lbo);
        constexpr uint32_t kHi = desc_hi(kSbo);
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint32_t ao = (p == 2) ? kImgBytes : 0, bo = (p == 1) ? b_lo : 0;
#pragma unroll
          for (int kk = 0; kk < kKC; kk += 16)
            mma_f16_ss2(acc, a0 + ((ao + (kk / 8) * kLbo) >> 4), b0 + ((bo + (kk / 8) * b_lbo) >> 4), kHi, idesc,
                        (fresh && p == 0 && kk == 0) ? 0u : 1u);
        }
      };
      for (int tile = blockIdx.x, it = 0; tile < nwork; tile += gridDim.x, ++it) {
        if (it > 0) { mbar_wait(tile_done, td_phase); td_phase ^= 1; fence_after_sync(); }
        for (int c = 0; c < 5; ++c, ++q) {
          const int slot = q % kSA;
          mbar_wait(&a_full[slot], (q / kSA) & 1);
          fence_after_sync();
          const uint32_t as = smem_u32(Aring + slot * kSlotBytes);
          if (elect_one_sync()) {
            if (c < 4) {
              gemm(ACC_P, as, smem_u32(Wp + c * 2 * kPImgBytes), kPImgBytes, kPLbo, kIdescP, c == 0);
              gemm(ACC_N, as, smem_u32(Wn + c * kSlotBytes), kImgBytes, kLbo, kIdescW, c == 0);
            }
            gemm(ACC_A, as, smem_u32(Wag + c * kSlotBytes), kImgBytes, kLbo, kIdescW, c == 0);
            mma_commit(&a_empty[slot]);
            if (c == 3) mma_commit(p_full);
            if (c == 4) mma_commit(an_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ row workers
    const int grp = tid >> 7, m = tid & 127;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float* wsc = a.wstream + kImageFloats;
    const float sP = wsc[0], sA = wsc[1], sN = wsc[2];
    uint32_t ph = 0;
    int qbase = 0;
    float amax = 0.f;  // overflow report, see umma.cuh
    auto publish = [&](int q, const float* vals) {
      const int slot = q % kSA;
      mbar_wait(&a_empty[slot], ((q / kSA) & 1) ^ 1);
      put_chunk(Aring + slot * kSlotBytes, m, vals, amax);
      fence_async_smem();
      mbar_arrive(&a_full[slot]);
    };
    auto load_acc = [&](uint32_t acc, float (&dst)[32]) {
      uint32_t u[32];
      tmem_ld32(acc + lane_base, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(u[i]);
    };
    for (int pos = blockIdx.x; pos < nwork; pos += gridDim.x) {
      const int tile = a.live_list ? a.live_list[pos] : pos;
      const int r = tile * kRows + m;
      const bool in = r < R;
      const int rr = min(r, R - 1);
      float v[32];
      // ---- h_V row -> chunks grp, grp + 2
      const float* hrow = a.hV + (size_t)rr * 128;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 x = in ? *reinterpret_cast<const float4*>(hrow + (grp + 2 * t) * 32 + u * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[u * 4] = x.x; v[u * 4 + 1] = x.y; v[u * 4 + 2] = x.z; v[u * 4 + 3] = x.w;
        }
        publish(qbase + grp + 2 * t, v);
      }
      // ---- points (group 0): p_local, norms -> geometry chunk; points in the global frame -> wsP
      if (grp == 0) {
        mbar_wait(p_full, ph);
        fence_after_sync();
        load_acc(ACC_P, v);
#pragma unroll
        for (int i = 0; i < 24; ++i) v[i] = fmaf(v[i], sP, prm[i]);
        const float* g = a.geo + (size_t)(rr % a.G) * PP_GEO_STRIDE;
        float gl[24];
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) {
          const float x = v[pt * 3], y = v[pt * 3 + 1], z = v[pt * 3 + 2];
          gl[pt * 3 + 0] = g[0] * x + g[1] * y + g[2] * z + g[9];
          gl[pt * 3 + 1] = g[3] * x + g[4] * y + g[5] * z + g[10];
          gl[pt * 3 + 2] = g[6] * x + g[7] * y + g[8] * z + g[11];
        }
#pragma unroll
        for (int pt = 0; pt < 8; ++pt) {
          const float x = v[pt * 3], y = v[pt * 3 + 1], z = v[pt * 3 + 2];
          v[24 + pt] = sqrtf(x * x + y * y + z * z + 1e-8f);
        }
        publish(qbase + 4, v);
        if (in) {
          float4* o = reinterpret_cast<float4*>(a.P + (size_t)r * 24);
#pragma unroll
          for (int u = 0; u < 6; ++u) o[u] = make_float4(gl[u * 4], gl[u * 4 + 1], gl[u * 4 + 2], gl[u * 4 + 3]);
        }
      }
      // ---- A_i = acc + b1, N_j = acc
      mbar_wait(an_full, ph);
      fence_after_sync();
      ph ^= 1;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = grp + 2 * t;
        load_acc(ACC_A + c * 32, v);
        if (in) {
          const float* b = prm + 32 + c * 32;
          float4* o = reinterpret_cast<float4*>(a.A + (size_t)r * 128 + c * 32);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            o[u] = make_float4(fmaf(v[u * 4], sA, b[u * 4]), fmaf(v[u * 4 + 1], sA, b[u * 4 + 1]),
                               fmaf(v[u * 4 + 2], sA, b[u * 4 + 2]), fmaf(v[u * 4 + 3], sA, b[u * 4 + 3]));
        }
        load_acc(ACC_N + c * 32, v);
        if (in) {
          float4* o = reinterpret_cast<float4*>(a.Nn + (size_t)r * 128 + c * 32);
#pragma unroll
          for (int u = 0; u < 8; ++u) o[u] = make_float4(v[u * 4] * sN, v[u * 4 + 1] * sN, v[u * 4 + 2] * sN, v[u * 4 + 3] * sN);
        }
      }
      fence_before_sync();
      mbar_arrive(tile_done);
      qbase += 5;
    }
    report_overflow(a.overflow, amax);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace pre
}  // namespace pp

using namespace pp;

extern "C" int64_t pp_tc_pre_stream_floats() { return pre::kStreamFloats; }

// Tensor-core version of pp_ipmp_node_pre (path 0 = node message, 1 = edge message of `layer`): same outputs.
//   wstream: operand images of this layer and path, pp_tc_pre_stream_floats() floats (weights.py: pack_pre_stream)
extern "C" int pp_ipmp_node_pre_tc(const float* weights, int64_t layer, int64_t path, const float* wstream,
                                   const float* geo, int64_t G, int64_t S, const float* hV, float* wsA, float* wsN,
                                   float* wsP, int32_t* overflow, const int32_t* live_tiles, const int32_t* n_live,
                                   cudaStream_t stream) {
  PP_REQUIRE(weights && wstream && geo && hV && wsA && wsN && wsP, "null pointer");
  PP_REQUIRE(layer >= 0 && layer < 3 && (path == 0 || path == 1), "layer / path out of range");
  PP_REQUIRE(G > 0 && S > 0, "bad sizes");
  const float* Lb = weights + layer * wl::kLayerStride;
  pre::Args a{};
  a.geo = geo; a.G = (int)G; a.R = (int)(S * G);
  a.wstream = wstream;
  a.BP = Lb + (path ? PP_OFF(L0_E_BP) : PP_OFF(L0_N_BP));
  a.B1 = Lb + (path ? PP_OFF(L0_E_B1) : PP_OFF(L0_N_B1));
  a.hV = hV; a.A = wsA; a.Nn = wsN; a.P = wsP;
  a.overflow = overflow;
  a.live_list = live_tiles;
  a.n_live = live_tiles ? n_live : nullptr;
  PP_REQUIRE(!live_tiles || n_live, "live_tiles needs n_live");
  cudaError_t e = cudaFuncSetAttribute(pre::node_pre_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pre::kSmem);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "node_pre_tc_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return 1;
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = (a.R + pre::kRows - 1) / pre::kRows;
  pre::node_pre_tc_kernel<<<tiles < num_sms ? tiles : num_sms, pre::kThreads, pre::kSmem, stream>>>(a);
  return check_launch("pp_ipmp_node_pre_tc");
}
