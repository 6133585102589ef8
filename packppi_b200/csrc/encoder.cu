// Encoder: fused edge featurisation + 468->128 embedding + LayerNorm (once per complex) and the per-step
// 51->128 node embedding + LayerNorm.
//
// Replaces ProteinEncoder.forward (reference src/models/components/encoder.py:198-246): _af2_encoding (:34-47),
// _atomic_distances/_get_rbf/_rbf (:120-153), _pairwise_dihedrals/_dihedral_from_four_points (:164-196), the
// chain-type feature (:232-233) and edge_embedding + norm_edges (:243-244); and for nodes the one-hot / sin-cos /
// SinusoidalEmbedding concatenation (:217-229, layers.py:248-268) with node_embedding + norm_nodes (:241-242).
// The reference builds 27 dense [B,L,L] matrices per call; here every feature is computed per edge from two
// 28-float geometry records and consumed directly by the GEMM, nothing of size L^2 exists.
//
// Edge feature order inside the kernel (the weight rows are permuted to match, see weights_layout.h):
//   chunks 0..24  : atom pair (a,b) in (N,CA,C,O,CB)^2, a-major, 16 RBF each
//   chunks 25..29 : one_hot(clip(ridx_i - ridx_j + 32, 0, 64)) padded to 80
//   chunk  30     : 1 + [chain_i == chain_j], dih(C_i,N_j,CA_j,C_j), dih(N_i,CA_i,C_i,N_j), 13 zeros
#include "tile_gemm.cuh"
#include "weights_layout.h"

namespace pp {

// The two inter-residue dihedrals are raw signed angles: sign(triple product) * arccos(n1 . n2) with NaN -> 0
// (encoder.py:155-174).  At near-planar geometry the result is decided by the last bit of the cosine (|cos| > 1 ->
// NaN -> 0 instead of ~pi) and of the triple product (sign flip = 2 pi), so the kernel follows the rounding sequence
// of the reference's torch-CPU ops exactly (measured against torch 2.11 CPU, tools/check_dihedral_rounding.py:
// 0 mismatches in 1.2e5 random vectors for each step):
//   torch.cross       c_k = fma(a_k1, b_k2, -rn(a_k2 * b_k1))
//   torch.norm        sqrt(fma(z, z, fma(y, y, rn(x * x))))
//   (a * b).sum(-1)   (rn(a0 b0) + rn(a1 b1)) + rn(a2 b2)
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  o[0] = __fmaf_rn(a[1], b[2], -__fmul_rn(a[2], b[1]));
  o[1] = __fmaf_rn(a[2], b[0], -__fmul_rn(a[0], b[2]));
  o[2] = __fmaf_rn(a[0], b[1], -__fmul_rn(a[1], b[0]));
}

__device__ __forceinline__ float dot3_seq(const float* a, const float* b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}

__device__ __forceinline__ void unit_nan0(float* v) {  // encoder.py:155-162: v / |v| with NaN -> 0
  float n = __fsqrt_rn(__fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], __fmul_rn(v[0], v[0]))));
  for (int k = 0; k < 3; ++k) v[k] = nan_to_num(__fdiv_rn(v[k], n));
}

__device__ float dihedral4(const float* p0, const float* p1, const float* p2, const float* p3) {
  float u0[3], u1[3], u2[3], n1[3], n2[3], c[3];
  for (int k = 0; k < 3; ++k) {
    u0[k] = __fsub_rn(p2[k], p1[k]);
    u1[k] = __fsub_rn(p0[k], p1[k]);
    u2[k] = __fsub_rn(p3[k], p2[k]);
  }
  cross3(u0, u1, n1);
  cross3(u0, u2, n2);
  unit_nan0(n1);
  unit_nan0(n2);
  cross3(u1, u2, c);
  float s = dot3_seq(c, u0);
  float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
  float d = __fmul_rn(sg, acosf(dot3_seq(n1, n2)));
  return nan_to_num(d);
}

constexpr int kEdgeChunks = 31;
constexpr int kChunksPerPass = 10;
constexpr size_t kEncSmemFloats = (size_t)kTileRows * kLdB0 + kWbufFloats + 128 * 16 /*geo_j atoms*/ + 4 * 16 + 16;
constexpr size_t kEncSmemBytes = kEncSmemFloats * 4 + 128 * 4 * 3;

__global__ void __launch_bounds__(kThreads, 1)
edge_embed_kernel(const float* __restrict__ W, const float* __restrict__ geo, const int* __restrict__ nbr,
                  const long long* __restrict__ residue_index, const long long* __restrict__ chain_indices, int G, int K,
                  float* __restrict__ hE0) {
  extern __shared__ __align__(16) float smem_raw[];
  float* B0 = smem_raw;
  float* wbuf = B0 + kTileRows * kLdB0;
  float* aj = wbuf + kWbufFloats;      // [128][16]: N CA C O CB of the neighbour (15 used)
  float* ai = aj + 128 * 16;           // [4][16]
  int* jrow = reinterpret_cast<int*>(ai + 4 * 16 + 16);
  int* rel = jrow + 128;               // relpos class
  int* same = rel + 128;               // same chain
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int rb = blockIdx.x * 4;
  const float* mu = W + PP_OFF(RBF_MU);

  if (tid < 128) {
    int rl = tid >> 5, k = tid & 31, g = rb + rl;
    int j = 0, rc = 0, sc = 0;
    if (g < G) {
      j = (k < K) ? nbr[(size_t)g * K + k] : g;
      long long off = residue_index[g] - residue_index[j] + 32;
      rc = (int)(off < 0 ? 0 : (off > 64 ? 64 : off));
      sc = chain_indices[g] == chain_indices[j];
    }
    jrow[tid] = j;
    rel[tid] = rc;
    same[tid] = sc;
    const float* gj = geo + (size_t)j * PP_GEO_STRIDE + 12;
#pragma unroll
    for (int c = 0; c < 15; ++c) aj[tid * 16 + c] = gj[c];
  } else if (tid < 128 + 60) {
    int rl = (tid - 128) / 15, c = (tid - 128) % 15, g = min(rb + rl, G - 1);
    ai[rl * 16 + c] = geo[(size_t)g * PP_GEO_STRIDE + 12 + c];
  }
  __syncthreads();

  float acc[8][8];
  zero_acc(acc);
  for (int pass = 0; pass * kChunksPerPass < kEdgeChunks; ++pass) {
    const int c0 = pass * kChunksPerPass;
    const int nch = min(kChunksPerPass, kEdgeChunks - c0);
    for (int it = tid; it < 128 * nch; it += kThreads) {
      int m = it & 127, ch = c0 + (it >> 7);
      float* o = B0 + m * kLdB0 + (ch - c0) * 16;
      const float* Ai = ai + (m >> 5) * 16;
      const float* Aj = aj + m * 16;
      if (ch < 25) {
        int a = ch / 5, b = ch - a * 5;
        float dx = Ai[a * 3] - Aj[b * 3], dy = Ai[a * 3 + 1] - Aj[b * 3 + 1], dz = Ai[a * 3 + 2] - Aj[b * 3 + 2];
        float D = sqrtf(dx * dx + dy * dy + dz * dz + 1e-6f);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float z = (D - mu[r]) / 1.25f;
          o[r] = expf(-(z * z));
        }
      } else if (ch < 30) {
        int base = (ch - 25) * 16, rc = rel[m];
#pragma unroll
        for (int r = 0; r < 16; ++r) o[r] = (base + r == rc) ? 1.f : 0.f;
      } else {
        o[0] = same[m] ? 2.f : 1.f;
        o[1] = dihedral4(Ai + 6, Aj + 0, Aj + 3, Aj + 6);  // C_i, N_j, CA_j, C_j
        o[2] = dihedral4(Ai + 0, Ai + 3, Ai + 6, Aj + 0);  // N_i, CA_i, C_i, N_j
#pragma unroll
        for (int r = 3; r < 16; ++r) o[r] = 0.f;
      }
    }
    gemm_tile(acc, B0, kLdB0, W + PP_OFF(ENC_EDGE_WT) + (size_t)c0 * 16 * 128, 128, nch * 16, wbuf);
  }
  float bias[8], g[8], b[8];
  load_cols(bias, W + PP_OFF(ENC_EDGE_B), tx);
  load_cols(g, W + PP_OFF(ENC_EDGE_LNG), tx);
  load_cols(b, W + PP_OFF(ENC_EDGE_LNB), tx);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] += bias[j];
  layer_norm_rows(acc, g, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = tile_row(ty, i);
    int gi = rb + (m >> 5), k = m & 31;
    if (gi < G && k < K) {
      float* o = hE0 + ((size_t)gi * K + k) * 128;
      *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

// one warp per residue row r = s*G + g; lane owns output columns lane*4 .. lane*4+3
__global__ void node_embed_kernel(const float* __restrict__ W, const long long* __restrict__ residue_type,
                                  const float* __restrict__ bb_sincos /*[G][6]*/, const float* __restrict__ chi /*[R][4]*/,
                                  const float* __restrict__ chi_mask /*[G][4]*/,
                                  const float* __restrict__ sc_sincos /*[R][8] precomputed, or null*/,
                                  const float* __restrict__ t,
                                  long long t_stride, int G, int S, float* __restrict__ hV) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= S * G) return;
  int g = r % G;
  const float* Wt = W + PP_OFF(ENC_NODE_WT);
  float feat = 0.f;  // lane f holds dense feature 21 + f (30 features: 6 backbone, 8 chi, 16 time)
  if (lane < 6) {
    feat = bb_sincos[(size_t)g * 6 + lane];
  } else if (lane < 14) {
    int c = (lane - 6) >> 1;
    if (sc_sincos) {
      feat = sc_sincos[(size_t)r * 8 + (lane - 6)];
    } else {
      float a = chi[(size_t)r * 4 + c], m = chi_mask[(size_t)g * 4 + c];
      feat = (((lane - 6) & 1) ? cosf(a) : sinf(a)) * m;
    }
  } else if (lane < 30) {
    // SinusoidalEmbedding: t * 10000 (fp32), times exp(-i ln(1e4)/7), sin | cos   (layers.py:257-264)
    float ts = __fmul_rn(t[(size_t)r * t_stride], 10000.f);
    int i = (lane - 14) & 7;
    float arg = __fmul_rn(ts, W[PP_OFF(TIME_FREQ) + i]);
    feat = (lane - 14 < 8) ? sinf(arg) : cosf(arg);
  }
  int S_type = (int)residue_type[g];
  S_type = min(max(S_type, 0), 20);
  float4 acc = *reinterpret_cast<const float4*>(W + PP_OFF(ENC_NODE_B) + lane * 4);
  float4 w = *reinterpret_cast<const float4*>(Wt + (size_t)S_type * 128 + lane * 4);
  acc.x += w.x; acc.y += w.y; acc.z += w.z; acc.w += w.w;
#pragma unroll
  for (int f = 0; f < 30; ++f) {
    float v = __shfl_sync(0xffffffffu, feat, f);
    w = *reinterpret_cast<const float4*>(Wt + (size_t)(21 + f) * 128 + lane * 4);
    acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y); acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
  }
  float mean = warp_sum(acc.x + acc.y + acc.z + acc.w) * (1.f / 128.f);
  float dx = acc.x - mean, dy = acc.y - mean, dz = acc.z - mean, dw = acc.w - mean;
  float rstd = rsqrtf(warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.f / 128.f) + 1e-5f);
  float4 gm = *reinterpret_cast<const float4*>(W + PP_OFF(ENC_NODE_LNG) + lane * 4);
  float4 bt = *reinterpret_cast<const float4*>(W + PP_OFF(ENC_NODE_LNB) + lane * 4);
  *reinterpret_cast<float4*>(hV + (size_t)r * 128 + lane * 4) =
      make_float4(dx * rstd * gm.x + bt.x, dy * rstd * gm.y + bt.y, dz * rstd * gm.z + bt.z, dw * rstd * gm.w + bt.w);
}

}  // namespace pp

using namespace pp;

extern "C" int pp_edge_embed(const float* weights, const float* geo, const int32_t* nbr, const int64_t* residue_index,
                             const int64_t* chain_indices, int64_t G, int64_t K, float* hE0, cudaStream_t stream) {
  PP_REQUIRE(weights && geo && nbr && residue_index && chain_indices && hE0, "null pointer");
  PP_REQUIRE(G > 0 && K > 0 && K <= PP_KMAX, "bad sizes");
  cudaError_t e = cudaFuncSetAttribute(edge_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEncSmemBytes);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "pp_edge_embed: %s", cudaGetErrorString(e));
    return 1;
  }
  edge_embed_kernel<<<(unsigned)((G + 3) / 4), kThreads, kEncSmemBytes, stream>>>(
      weights, geo, nbr, (const long long*)residue_index, (const long long*)chain_indices, (int)G, (int)K, hE0);
  return check_launch("pp_edge_embed");
}

extern "C" int pp_node_embed(const float* weights, const int64_t* residue_type, const float* bb_sincos,
                             const float* chi, const float* chi_mask, const float* sc_sincos, const float* t,
                             int64_t t_stride, int64_t G, int64_t S, float* hV, cudaStream_t stream) {
  PP_REQUIRE(weights && residue_type && bb_sincos && t && hV, "null pointer");
  PP_REQUIRE(sc_sincos || (chi && chi_mask), "need chi + chi_mask or precomputed sc_sincos");
  PP_REQUIRE(G > 0 && S > 0, "bad sizes");
  PP_REQUIRE(t_stride == 0 || t_stride == 1, "t_stride must be 0 (scalar) or 1 (per row)");
  long long R = S * G;
  node_embed_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, stream>>>(
      weights, (const long long*)residue_type, bb_sincos, chi, chi_mask, sc_sincos, t, t_stride, (int)G, (int)S, hV);
  return check_launch("pp_node_embed");
}
