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

// ------------------------------------------------------------------------------------------ cell list
// Same contract as knn_bruteforce_kernel, O(L * neighbourhood) instead of O(L^2): valid residues are binned on CA into
// cells of edge h (>= 7 A, one grid of at most kCellsMax cells per complex); a warp scans the cube of cells within
// Chebyshev radius r of its residue, row by row (cells of one x-row are contiguous in the sorted order), and stops as
// soon as its K-th best distance is <= r*h: everything outside the cube is farther than r*h, hence farther than the
// K-th neighbour.  Otherwise it restarts with a doubled radius.  Keys are the same (D bits, j) pairs, so the result is
// identical to the brute-force kernel no matter in which order candidates are visited.  Complexes with fewer than K
// valid residues (masked candidates then enter the list at 2 * rowmax(D)) and masked rows take the brute-force path.
constexpr int kCellsMax = 32768;

struct CellBox {
  float minx, miny, minz, inv_h, h;
  int nx, ny, nz;
};

__global__ void cell_box_kernel(const float* __restrict__ X, const float* __restrict__ mask, int L, float h_min,
                                CellBox* __restrict__ box, int* __restrict__ counts /*[B][kCellsMax+1]*/) {
  const int b = blockIdx.x;
  __shared__ float red[6][32];
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    if (mask[(size_t)b * L + j] > 0.f) {
      const float* p = X + ((size_t)b * L + j) * 42 + 3;
      for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], p[k]); hi[k] = fmaxf(hi[k], p[k]); }
    }
  }
  for (int k = 0; k < 3; ++k) {
    float a = lo[k], c = hi[k];
    for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o)); }
    if ((threadIdx.x & 31) == 0) { red[k][threadIdx.x >> 5] = a; red[3 + k][threadIdx.x >> 5] = c; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int k = 0; k < 3; ++k)
      for (int w = 0; w < nw; ++w) { lo[k] = fminf(lo[k], red[k][w]); hi[k] = fmaxf(hi[k], red[3 + k][w]); }
    CellBox bx;
    if (lo[0] > hi[0]) { lo[0] = lo[1] = lo[2] = 0.f; hi[0] = hi[1] = hi[2] = 0.f; }  // no valid residue
    float h = h_min;
    for (;;) {
      bx.nx = (int)floorf((hi[0] - lo[0]) / h) + 1;
      bx.ny = (int)floorf((hi[1] - lo[1]) / h) + 1;
      bx.nz = (int)floorf((hi[2] - lo[2]) / h) + 1;
      if ((long long)bx.nx * bx.ny * bx.nz <= kCellsMax) break;
      h *= 1.25f;
    }
    bx.minx = lo[0]; bx.miny = lo[1]; bx.minz = lo[2];
    bx.h = h; bx.inv_h = 1.f / h;
    box[b] = bx;
  }
  for (int c = threadIdx.x; c <= kCellsMax; c += blockDim.x) counts[(size_t)b * (kCellsMax + 1) + c] = 0;
}

__device__ __forceinline__ int cell_of(const CellBox& bx, const float* p, int& cx, int& cy, int& cz) {
  cx = min(bx.nx - 1, max(0, (int)floorf((p[0] - bx.minx) * bx.inv_h)));
  cy = min(bx.ny - 1, max(0, (int)floorf((p[1] - bx.miny) * bx.inv_h)));
  cz = min(bx.nz - 1, max(0, (int)floorf((p[2] - bx.minz) * bx.inv_h)));
  return (cz * bx.ny + cy) * bx.nx + cx;
}

__global__ void cell_count_kernel(const float* __restrict__ X, const float* __restrict__ mask, int B, int L,
                                  const CellBox* __restrict__ box, int* __restrict__ counts) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * L || !(mask[r] > 0.f)) return;
  int b = r / L, cx, cy, cz;
  int c = cell_of(box[b], X + (size_t)r * 42 + 3, cx, cy, cz);
  atomicAdd(&counts[(size_t)b * (kCellsMax + 1) + c + 1], 1);  // shifted by one: the scan turns it into cell starts
}

// one block per complex: inclusive scan of the shifted counts = exclusive cell starts; `cursor` = copy for the fill
__global__ void cell_scan_kernel(int* __restrict__ counts, int* __restrict__ cursor) {
  int* cs = counts + (size_t)blockIdx.x * (kCellsMax + 1);
  int* cu = cursor + (size_t)blockIdx.x * (kCellsMax + 1);
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int base = 0; base <= kCellsMax; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i <= kCellsMax) ? cs[i] : 0;
    int sc = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
    if (lane == 31) wsum[w] = sc;
    __syncthreads();
    if (w == 0) {
      int t = (lane < nw) ? wsum[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
      wsum[lane] = t;
    }
    __syncthreads();
    int pre = carry + (w ? wsum[w - 1] : 0) + sc;
    if (i <= kCellsMax) { cs[i] = pre; cu[i] = pre; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = pre;
    __syncthreads();
  }
}

__global__ void cell_fill_kernel(const float* __restrict__ X, const float* __restrict__ mask, int B, int L,
                                 const CellBox* __restrict__ box, int* __restrict__ cursor, int* __restrict__ order) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * L || !(mask[r] > 0.f)) return;
  int b = r / L, cx, cy, cz;
  int c = cell_of(box[b], X + (size_t)r * 42 + 3, cx, cy, cz);
  int slot = atomicAdd(&cursor[(size_t)b * (kCellsMax + 1) + c], 1);
  order[(size_t)b * L + slot] = r - b * L;
}

__global__ void knn_cells_kernel(const float* __restrict__ X, const float* __restrict__ mask, int B, int L, int K,
                                 const CellBox* __restrict__ box, const int* __restrict__ cell_start,
                                 const int* __restrict__ order, long long* __restrict__ E_idx, int* __restrict__ nbr,
                                 float* __restrict__ D_out, float* __restrict__ matt, float* __restrict__ msum) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int b = warp / L;
  const float* Xb = X + (size_t)b * L * 42;
  const float* mb = mask + (size_t)b * L;
  const float* xi_p = X + (size_t)warp * 42 + 3;
  const float xi = xi_p[0], yi = xi_p[1], zi = xi_p[2], mi = mask[warp];
  const int* cs = cell_start + (size_t)b * (kCellsMax + 1);
  const int* ord = order + (size_t)b * L;
  const CellBox bx = box[b];
  const int ncell = bx.nx * bx.ny * bx.nz;
  const int nvalid = cs[ncell];

  WarpTopK top;
  top.init();
  if (mi == 0.f || nvalid < K) {
    // brute-force path (identical to knn_bruteforce_kernel)
    float dmax = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float* p = Xb + (size_t)j * 42 + 3;
      dmax = fmaxf(dmax, knn_dist(xi, yi, zi, p[0], p[1], p[2], __fmul_rn(mb[j], mi)));
    }
    dmax = warp_max(dmax);
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
  } else {
    int cx, cy, cz;
    cell_of(bx, xi_p, cx, cy, cz);
    const int rmax = max(bx.nx, max(bx.ny, bx.nz));
    for (int r = 2;; r *= 2) {
      top.init();
      for (int dz = -r; dz <= r; ++dz) {
        int z = cz + dz;
        if (z < 0 || z >= bx.nz) continue;
        for (int dy = -r; dy <= r; ++dy) {
          int y = cy + dy;
          if (y < 0 || y >= bx.ny) continue;
          int x0 = max(cx - r, 0), x1 = min(cx + r, bx.nx - 1);
          int row = (z * bx.ny + y) * bx.nx;
          int s = cs[row + x0], e = cs[row + x1 + 1];
          for (int t0 = s; t0 < e; t0 += 32) {
            int t = t0 + lane;
            unsigned long long cand = ~0ull;
            if (t < e) {
              int j = ord[t];
              const float* p = Xb + (size_t)j * 42 + 3;
              float d = knn_dist(xi, yi, zi, p[0], p[1], p[2], 1.f);  // both valid: mask_2D = 1, D_adjust = D
              cand = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
            }
            top.offer(cand, K, lane);
          }
        }
      }
      unsigned long long kth = top.kth(K);
      float dk = __uint_as_float((unsigned)(kth >> 32));
      if ((kth != ~0ull && dk <= (float)r * bx.h * 0.999f) || r >= rmax) break;  // 0.1 % slack for cell rounding
    }
  }
  float ma = 0.f;
  if (lane < K) {
    int j = (int)(top.best & 0xffffffffu);
    size_t o = (size_t)warp * K + lane;
    E_idx[o] = j;
    nbr[o] = b * L + j;
    if (D_out) D_out[o] = __uint_as_float((unsigned)(top.best >> 32));
    ma = mi * mb[j];
    if (matt) matt[o] = ma;
  }
  ma = warp_sum(ma);
  if (msum && lane == 0) msum[warp] = ma / (float)K;
}

// Residue neighbour list of the clash term from the same cell list (cells of edge >= 2 * max reach + cutoff, so the
// 27 surrounding cells contain every partner): candidates are marked in a per-warp bitmask in shared memory and then
// emitted in ascending index order, which keeps the summation order of the pair kernel fixed from run to run.
__global__ void clash_nbr_cells_kernel(const float* __restrict__ X, const float* __restrict__ reach,
                                       const long long* __restrict__ residue_index, int B, int L, float cutoff,
                                       const CellBox* __restrict__ box, const int* __restrict__ cell_start,
                                       const int* __restrict__ order, int fill, int* __restrict__ count,
                                       const long long* __restrict__ start, int* __restrict__ list) {
  extern __shared__ unsigned bits_all[];
  const int words = (L + 31) >> 5;
  unsigned* bits = bits_all + (threadIdx.x >> 5) * words;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int b = warp / L;
  const float ri = reach[warp];
  int n = 0;
  if (ri >= 0.f) {
    for (int w = lane; w < words; w += 32) bits[w] = 0u;
    __syncwarp();
    const float* pi = X + (size_t)warp * 42 + 3;
    const float xi = pi[0], yi = pi[1], zi = pi[2];
    const long long idx_i = residue_index[warp];
    const CellBox bx = box[b];
    const int* cs = cell_start + (size_t)b * (kCellsMax + 1);
    const int* ord = order + (size_t)b * L;
    int cx, cy, cz;
    cell_of(bx, pi, cx, cy, cz);
    for (int dz = -1; dz <= 1; ++dz) {
      int z = cz + dz;
      if (z < 0 || z >= bx.nz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        int y = cy + dy;
        if (y < 0 || y >= bx.ny) continue;
        int x0 = max(cx - 1, 0), x1 = min(cx + 1, bx.nx - 1);
        int row = (z * bx.ny + y) * bx.nx;
        int s = cs[row + x0], e = cs[row + x1 + 1];
        for (int t = s + lane; t < e; t += 32) {
          int j = ord[t];
          int gj = b * L + j;
          float rj = reach[gj];
          if (residue_index[gj] != idx_i) {  // clash.py:166-169: strict '<' on the index VALUE, both ways
            const float* pj = X + (size_t)gj * 42 + 3;
            float dx = pj[0] - xi, dy2 = pj[1] - yi, dz2 = pj[2] - zi;
            float lim = ri + rj + cutoff;
            if (dx * dx + dy2 * dy2 + dz2 * dz2 < lim * lim) atomicOr(&bits[j >> 5], 1u << (j & 31));
          }
        }
      }
    }
    __syncwarp();
    const long long base = fill ? start[warp] : 0;
    for (int w0 = 0; w0 < words; w0 += 32) {
      int w = w0 + lane;
      unsigned v = (w < words) ? bits[w] : 0u;
      int c = __popc(v);
      int incl = c;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (fill) {
        int pos = n + incl - c;
        while (v) {
          int bit = __ffs(v) - 1;
          v &= v - 1;
          list[base + pos++] = b * L + (w << 5) + bit;
        }
      }
      n += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if (!fill && lane == 0) count[warp] = n;
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

// Cell-list version of pp_knn_build (same outputs, bit-identical).  Workspaces: ws_int = B * 2 * (pp_knn_cells_max() + 1)
// + B * L int32, ws_box = B * 8 floats.
extern "C" int64_t pp_knn_cells_max() { return pp::kCellsMax; }

extern "C" int pp_knn_build_cells(const float* X, const float* residue_mask, int64_t B, int64_t L, int64_t K,
                                  int64_t* E_idx, int32_t* nbr, float* D_neighbors, float* mask_attend, float* msum,
                                  int32_t* ws_int, float* ws_box, cudaStream_t stream) {
  PP_REQUIRE(X && residue_mask && E_idx && nbr && ws_int && ws_box, "null pointer");
  PP_REQUIRE(B > 0 && L > 0, "empty batch");
  PP_REQUIRE(K == (L < PP_KMAX ? L : PP_KMAX), "K must equal min(32, L)");
  PP_REQUIRE(B * L < (1ll << 31), "too many residues");
  static_assert(sizeof(pp::CellBox) == 32, "CellBox is 8 words");
  pp::CellBox* box = reinterpret_cast<pp::CellBox*>(ws_box);
  int* counts = ws_int;
  int* cursor = counts + B * (pp::kCellsMax + 1);
  int* order = cursor + B * (pp::kCellsMax + 1);
  const int G = (int)(B * L);
  pp::cell_box_kernel<<<(unsigned)B, 256, 0, stream>>>(X, residue_mask, (int)L, 7.f, box, counts);
  pp::cell_count_kernel<<<(G + 255) / 256, 256, 0, stream>>>(X, residue_mask, (int)B, (int)L, box, counts);
  pp::cell_scan_kernel<<<(unsigned)B, 1024, 0, stream>>>(counts, cursor);
  pp::cell_fill_kernel<<<(G + 255) / 256, 256, 0, stream>>>(X, residue_mask, (int)B, (int)L, box, cursor, order);
  pp::knn_cells_kernel<<<(unsigned)(((long long)G * 32 + 255) / 256), 256, 0, stream>>>(
      X, residue_mask, (int)B, (int)L, (int)K, box, counts, order, (long long*)E_idx, nbr, D_neighbors, mask_attend, msum);
  return pp::check_launch("pp_knn_build_cells");
}

// Cell-list version of pp_clash_neighbours (same counts / list).  `reach` [B*L] comes from pp_clash_reach;
// h_min >= 2 * max(reach) + cutoff.  fill = 0 bins the residues and counts; fill = 1 reuses the bins and writes the list.
extern "C" int pp_clash_neighbours_cells(const float* X, const float* reach, const int64_t* residue_index, int64_t B,
                                         int64_t L, float cutoff, float h_min, int64_t fill, int32_t* counts,
                                         const int64_t* start, int32_t* list, int32_t* ws_int, float* ws_box,
                                         cudaStream_t stream) {
  PP_REQUIRE(X && reach && residue_index && ws_int && ws_box, "null pointer");
  PP_REQUIRE(B > 0 && L > 0 && B * L < (1ll << 31), "bad sizes");
  PP_REQUIRE(fill ? (start && list) : (counts != nullptr), "missing output for this pass");
  PP_REQUIRE(h_min > 0.f, "h_min must be positive");
  pp::CellBox* box = reinterpret_cast<pp::CellBox*>(ws_box);
  int* cstart = ws_int;
  int* cursor = cstart + B * (pp::kCellsMax + 1);
  int* order = cursor + B * (pp::kCellsMax + 1);
  const int G = (int)(B * L);
  if (!fill) {
    pp::cell_box_kernel<<<(unsigned)B, 256, 0, stream>>>(X, reach, (int)L, h_min, box, cstart);
    pp::cell_count_kernel<<<(G + 255) / 256, 256, 0, stream>>>(X, reach, (int)B, (int)L, box, cstart);
    pp::cell_scan_kernel<<<(unsigned)B, 1024, 0, stream>>>(cstart, cursor);
    pp::cell_fill_kernel<<<(G + 255) / 256, 256, 0, stream>>>(X, reach, (int)B, (int)L, box, cursor, order);
  }
  const int warps_per_block = 4;
  const size_t smem = (size_t)warps_per_block * ((L + 31) / 32) * 4;
  PP_REQUIRE(smem <= 200 * 1024, "complex too long for the bitmask neighbour kernel");
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(pp::clash_nbr_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pp::clash_nbr_cells_kernel<<<(unsigned)((G + warps_per_block - 1) / warps_per_block), warps_per_block * 32, smem,
                               stream>>>(X, reach, (const long long*)residue_index, (int)B, (int)L, cutoff, box, cstart,
                                         order, (int)fill, counts, (const long long*)start, list);
  return pp::check_launch("pp_clash_neighbours_cells");
}

extern "C" int pp_geometry_build(const float* X, int64_t G, float* geo, cudaStream_t stream) {
  PP_REQUIRE(X && geo, "null pointer");
  PP_REQUIRE(G > 0, "empty batch");
  pp::geometry_kernel<<<(unsigned)((G + 127) / 128), 128, 0, stream>>>(X, (int)G, geo);
  return pp::check_launch("pp_geometry_build");
}
