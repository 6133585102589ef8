// Residue graph: k nearest neighbours on CA (exact, lowest-index tie-break) and per-residue geometry records.
//
// Replaces ProteinEncoder._dist (reference src/models/components/encoder.py:105-118) - a dense [B,L,L] distance
// matrix + torch.topk - and Rigid.from_3_points / _impute_CB (utils/rigid_utils.py:1126-1179,
// encoder.py:137-142).  knn_bruteforce_kernel: one warp per residue scans its whole complex (O(L^2); 25 M
// distance evaluations at L = 5000, microseconds on a B200, run once per complex and not per denoising step).
// Keys are (float bits of D_adjust, j): distances are computed with the reference's fp32 rounding
// (no FMA contraction), so the selected indices are bit-exact; equal distances resolve to the lower index.
#include "common.cuh"

namespace pp {

__device__ __forceinline__ float knn_dist(float xi, float yi, float zi, float xj, float yj, float zj, float m2) {
  // D = mask2D * sqrt(((dx^2 + dy^2) + dz^2) + 1e-6), every op rounded to fp32 like torch
  float dx = __fsub_rn(xj, xi), dy = __fsub_rn(yj, yi), dz = __fsub_rn(zj, zi);
  float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return __fmul_rn(m2, __fsqrt_rn(__fadd_rn(s, 1e-6f)));
}

__device__ __forceinline__ float knn_adjust(float D, float m2, float Dmax) {
  // D_adjust = D + 2 * (1 - mask2D) * D_max
  return __fadd_rn(D, __fmul_rn(__fmul_rn(2.f, __fsub_rn(1.f, m2)), Dmax));
}

// Sorted top-K list distributed over the lanes of a warp: lane l holds the l-th smallest key.
struct WarpTopK {
  unsigned long long best;
  __device__ __forceinline__ void init() { best = ~0ull; }
  __device__ __forceinline__ unsigned long long kth(int K) const { return __shfl_sync(0xffffffffu, best, K - 1); }
  // every lane offers one candidate key (or ~0ull for none)
  __device__ __forceinline__ void offer(unsigned long long cand, int K, int lane) {
    unsigned long long thr = kth(K);
    unsigned pending = __ballot_sync(0xffffffffu, cand < thr);
    while (pending) {
      int src = __ffs(pending) - 1;
      pending &= pending - 1;
      unsigned long long c = __shfl_sync(0xffffffffu, cand, src);
      if (c >= thr) continue;  // the threshold may have tightened since the ballot (uniform branch)
      int pos = __popc(__ballot_sync(0xffffffffu, best <= c));
      unsigned long long up = __shfl_up_sync(0xffffffffu, best, 1);
      if (lane > pos) best = up;
      else if (lane == pos) best = c;
      thr = kth(K);
    }
  }
};

__global__ void knn_bruteforce_kernel(const float* __restrict__ X, const float* __restrict__ mask, int B, int L, int K,
                                      long long* __restrict__ E_idx, int* __restrict__ nbr,
                                      float* __restrict__ D_out, float* __restrict__ matt, float* __restrict__ msum) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  int b = warp / L;
  const float* Xb = X + (size_t)b * L * 42;
  const float* mb = mask + (size_t)b * L;
  const float* xi_p = X + (size_t)warp * 42 + 3;  // atom 1 = CA
  float xi = xi_p[0], yi = xi_p[1], zi = xi_p[2], mi = mask[warp];

  float dmax = 0.f;
  for (int j = lane; j < L; j += 32) {
    const float* p = Xb + (size_t)j * 42 + 3;
    float m2 = __fmul_rn(mb[j], mi);  // mask_2D[i][j] = mask[j] * mask[i]
    dmax = fmaxf(dmax, knn_dist(xi, yi, zi, p[0], p[1], p[2], m2));
  }
  dmax = warp_max(dmax);

  WarpTopK top;
  top.init();
  for (int j0 = 0; j0 < L; j0 += 32) {
    int j = j0 + lane;
    unsigned long long cand = ~0ull;
    if (j < L) {
      const float* p = Xb + (size_t)j * 42 + 3;
      float m2 = __fmul_rn(mb[j], mi);
      float d = knn_adjust(knn_dist(xi, yi, zi, p[0], p[1], p[2], m2), m2, dmax);
      cand = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
    }
    top.offer(cand, K, lane);
  }
  float ma = 0.f;
  if (lane < K) {
    int j = (int)(top.best & 0xffffffffu);
    size_t o = (size_t)warp * K + lane;
    E_idx[o] = j;
    nbr[o] = b * L + j;
    if (D_out) D_out[o] = __uint_as_float((unsigned)(top.best >> 32));
    ma = mi * mb[j];  // mask_attend = mask_i * mask_j (mpnn.py:49-50)
    if (matt) matt[o] = ma;
  }
  ma = warp_sum(ma);
  if (msum && lane == 0) msum[warp] = ma / (float)K;  // mean over K of the attention mask
}

// Per-residue geometry record (PP_GEO_STRIDE floats): backbone frame R (row-major 3x3, columns e0 e1 e2), origin CA,
// then N, CA, C, O and the virtual CB used by the edge features.
__global__ void geometry_kernel(const float* __restrict__ X, int G, float* __restrict__ geo) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= G) return;
  const float* p = X + (size_t)r * 42;
  float N[3] = {p[0], p[1], p[2]}, CA[3] = {p[3], p[4], p[5]}, C[3] = {p[6], p[7], p[8]}, O[3] = {p[9], p[10], p[11]};
  float e0[3], e1[3], e2[3];
  for (int k = 0; k < 3; ++k) { e0[k] = C[k] - CA[k]; e1[k] = N[k] - CA[k]; }
  float d = sqrtf(e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2] + 1e-8f);
  for (int k = 0; k < 3; ++k) e0[k] /= d;
  float dot = e0[0] * e1[0] + e0[1] * e1[1] + e0[2] * e1[2];
  for (int k = 0; k < 3; ++k) e1[k] -= e0[k] * dot;
  d = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2] + 1e-8f);
  for (int k = 0; k < 3; ++k) e1[k] /= d;
  e2[0] = e0[1] * e1[2] - e0[2] * e1[1];
  e2[1] = e0[2] * e1[0] - e0[0] * e1[2];
  e2[2] = e0[0] * e1[1] - e0[1] * e1[0];
  float* g = geo + (size_t)r * PP_GEO_STRIDE;
  for (int k = 0; k < 3; ++k) { g[3 * k + 0] = e0[k]; g[3 * k + 1] = e1[k]; g[3 * k + 2] = e2[k]; g[9 + k] = CA[k]; }
  // virtual CB (encoder.py:137-142)
  float b[3], c[3], a[3];
  for (int k = 0; k < 3; ++k) { b[k] = CA[k] - N[k]; c[k] = C[k] - CA[k]; }
  a[0] = b[1] * c[2] - b[2] * c[1];
  a[1] = b[2] * c[0] - b[0] * c[2];
  a[2] = b[0] * c[1] - b[1] * c[0];
  for (int k = 0; k < 3; ++k) {
    g[12 + k] = N[k]; g[15 + k] = CA[k]; g[18 + k] = C[k]; g[21 + k] = O[k];
    g[24 + k] = -0.58273431f * a[k] + 0.56802827f * b[k] - 0.54067466f * c[k] + CA[k];
  }
  g[27] = 0.f;
}

}  // namespace pp

extern "C" int pp_knn_build(const float* X, const float* residue_mask, int64_t B, int64_t L, int64_t K,
                            int64_t* E_idx, int32_t* nbr, float* D_neighbors, float* mask_attend, float* msum,
                            cudaStream_t stream) {
  PP_REQUIRE(X && residue_mask && E_idx && nbr, "null pointer");
  PP_REQUIRE(B > 0 && L > 0, "empty batch");
  PP_REQUIRE(K == (L < PP_KMAX ? L : PP_KMAX), "K must equal min(32, L)");
  PP_REQUIRE(B * L < (1ll << 31), "too many residues");
  long long warps = B * L;
  int threads = 256;
  long long blocks = (warps * 32 + threads - 1) / threads;
  pp::knn_bruteforce_kernel<<<(unsigned)blocks, threads, 0, stream>>>(X, residue_mask, (int)B, (int)L, (int)K,
                                                                      (long long*)E_idx, nbr, D_neighbors, mask_attend,
                                                                      msum);
  return pp::check_launch("pp_knn_build");
}

extern "C" int pp_geometry_build(const float* X, int64_t G, float* geo, cudaStream_t stream) {
  PP_REQUIRE(X && geo, "null pointer");
  PP_REQUIRE(G > 0, "empty batch");
  pp::geometry_kernel<<<(unsigned)((G + 127) / 128), 128, 0, stream>>>(X, (int)G, geo);
  return pp::check_launch("pp_geometry_build");
}
