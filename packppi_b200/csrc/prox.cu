// chi -> atom14 rebuild, PackPPI-Prox structural-violation loss with analytic torsion gradient, fused proximal step.
//
// Replaces get_atom14_coords (reference src/models/components/__init__.py:76-120, utils/features.py:95-194 and the
// Rigid/Rotation classes), compute_residue_clash / find_sc_violations / between_residue_clash_loss /
// within_residue_violations (models/components/clash.py:7-365) and proximal_optimizer (optimize.py:5-73).
// The reference evaluates the between-residue term on a dense [N,N,14,14] tensor (13 live copies, 22.5 GB at
// N = 1478) and differentiates it with autograd.  Here:
//   atom14_kernel       one thread per residue, rigid frames in registers; also emits the chi rotation axes
//   clash_nbr_*         residue neighbour list from the static backbone: CA distance < reach_i + reach_j + cutoff,
//                       reach = rigorous bound on |CA - atom| over all chi  (built once per complex)
//   clash_pair_kernel   one warp per residue: two half-warps of 16 lanes (one lane per atom slot) walk the even and
//                       the odd entries of the neighbour list and add their sums at the end; a bounding-sphere test
//                       on the CURRENT atoms rejects most of them; every surviving 14x14 block is evaluated from
//                       both sides, so each atom owns its loss and force: no atomics, fixed summation order.
//                       Forces are projected on the chi axes (dL/dchi_k = sum_a (u_k x (p_a - o_k)) . F_a) in the
//                       same kernel; in proximal mode the Adam update is applied there too.
// Only pairs closer than r_a + r_b - tol can contribute (relu), so the result equals the dense sum exactly.
#include "common.cuh"

namespace pp {

struct Frame {
  float R[9];  // row-major
  float t[3];
};

__device__ __forceinline__ void mat_mul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void mat_vec(const float* A, const float* v, float* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

// Optional tail of an atom14 launch in the proximal loop: blocks first .. first + items - 1 reduce the objective of the
// PREVIOUS step from its per-row terms (one launch less per step; the rows are rewritten only by the next pair kernel).
struct ReduceArgs {
  const float* loss_rows;  // null: no reduction in this launch
  int first, B, L;
  const int* n_res;
  float lamda, inv_n_total;
  float* out;
};
__device__ void prox_reduce_item(const float* __restrict__ loss_rows, int it, int B, int L, const int* __restrict__ n_res,
                                 float lamda, float inv_n_total, float* __restrict__ out, float (*sh)[128]);

// tbl record: [0:48) chi1..4 default frames 3x4 | [48:90) literature positions | [90:104) group | [104:118) ideal mask
//             | [118:132) clash radius | [132] reach
__global__ void __launch_bounds__(128) atom14_kernel(const float* __restrict__ tbl, const float* __restrict__ X,
                              const long long* __restrict__ residue_type, const float* __restrict__ chi,
                              const float* __restrict__ chi_alt, const unsigned char* __restrict__ use_alt, int G, int S,
                              float* __restrict__ xyz_out /*[R][14][3] or null*/,
                              const float* __restrict__ atom_exists /*[G][14] or null*/,
                              float4* __restrict__ atoms4 /*[R][14] or null*/, float* __restrict__ axes /*[R][4][6] or null*/,
                              float* __restrict__ bound /*[R] current max |CA - atom| or null*/, ReduceArgs red) {
  if (red.loss_rows && (int)blockIdx.x >= red.first) {
    __shared__ float sh[2][128];
    prox_reduce_item(red.loss_rows, (int)blockIdx.x - red.first, red.B, red.L, red.n_res, red.lamda, red.inv_n_total,
                     red.out, sh);
    return;
  }
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S * G) return;
  int g = r % G;
  int type = (int)residue_type[g];
  type = min(max(type, 0), 20);
  const float* T = tbl + (size_t)type * PP_TBL_STRIDE;
  const float* p = X + (size_t)g * 42;
  float N[3] = {p[0], p[1], p[2]}, CA[3] = {p[3], p[4], p[5]}, C[3] = {p[6], p[7], p[8]};

  Frame bb;  // Rigid.from_3_points(N, CA, C, fixed=True): e0 ~ C-CA, e1 ~ N-CA orthogonalised, origin CA
  {
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
    for (int k = 0; k < 3; ++k) { bb.R[3 * k] = e0[k]; bb.R[3 * k + 1] = e1[k]; bb.R[3 * k + 2] = e2[k]; bb.t[k] = CA[k]; }
  }

  // atom positions; groups 1-3 (omega/phi/psi frames) only own slots that are overwritten by the input backbone
  float pos[14][3];
#pragma unroll
  for (int a = 0; a < 14; ++a) {
    if (a < 4) {
      pos[a][0] = p[a * 3]; pos[a][1] = p[a * 3 + 1]; pos[a][2] = p[a * 3 + 2];
    } else {
      float q[3];
      mat_vec(bb.R, T + 48 + a * 3, q);  // group 0 (CB); overwritten below if the slot belongs to a chi group
      float im = T[104 + a];
      for (int i = 0; i < 3; ++i) pos[a][i] = (q[i] + bb.t[i]) * im;
    }
  }
  {
    // chi1..chi4 frames, chained (features.py:137-156); each frame lives in registers only while its atoms are placed
    float Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tc[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = chi[(size_t)r * 4 + k];
      if (use_alt && !use_alt[(size_t)r * 4 + k]) a = chi_alt[(size_t)r * 4 + k];  // x' = where(mask, x, SC_D)
      float s = sinf(a), c = cosf(a);
      float nrm = sqrtf(fmaxf(s * s + c * c, 1e-12f));
      s /= nrm;
      c /= nrm;
      const float* D = T + k * 12;  // default frame, 3x4 row-major: [R | t]
      float Rl[9], tl[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {  // Default * Rx(chi), Rx = [[1,0,0],[0,c,-s],[0,s,c]]   (features.py:127-135)
        float d0 = D[i * 4], d1 = D[i * 4 + 1], d2 = D[i * 4 + 2];
        Rl[i * 3] = d0;
        Rl[i * 3 + 1] = d1 * c + d2 * s;
        Rl[i * 3 + 2] = d2 * c - d1 * s;
        tl[i] = D[i * 4 + 3];
      }
      float Rn[9], tn[3];
      mat_vec(Rc, tl, tn);
      for (int i = 0; i < 3; ++i) tn[i] += tc[i];
      mat_mul(Rc, Rl, Rn);
      for (int i = 0; i < 9; ++i) Rc[i] = Rn[i];
      for (int i = 0; i < 3; ++i) tc[i] = tn[i];
      Frame f;
      mat_mul(bb.R, Rc, f.R);
      mat_vec(bb.R, tc, f.t);
      for (int i = 0; i < 3; ++i) f.t[i] += bb.t[i];
#pragma unroll
      for (int a2 = 4; a2 < 14; ++a2) {
        if ((int)T[90 + a2] == 4 + k) {
          float q[3];
          mat_vec(f.R, T + 48 + a2 * 3, q);
          float im = T[104 + a2];
          for (int i = 0; i < 3; ++i) pos[a2][i] = (q[i] + f.t[i]) * im;
        }
      }
      if (axes) {
        float* o = axes + ((size_t)r * 4 + k) * 6;
        o[0] = f.R[0]; o[1] = f.R[3]; o[2] = f.R[6];  // x axis of the chi_k frame = rotation axis
        o[3] = f.t[0]; o[4] = f.t[1]; o[5] = f.t[2];
      }
    }
  }

  float reach2 = 0.f;
#pragma unroll
  for (int a = 0; a < 14; ++a) {
    if (xyz_out) {
      float* o = xyz_out + ((size_t)r * 14 + a) * 3;
      o[0] = pos[a][0]; o[1] = pos[a][1]; o[2] = pos[a][2];
    }
    if (atoms4) {
      float ex = atom_exists[(size_t)g * 14 + a];
      atoms4[(size_t)r * 14 + a] = make_float4(pos[a][0], pos[a][1], pos[a][2], ex * T[118 + a]);  // clash.py:286-289
      if (ex != 0.f) {
        float dx = pos[a][0] - CA[0], dy = pos[a][1] - CA[1], dz = pos[a][2] - CA[2];
        reach2 = fmaxf(reach2, dx * dx + dy * dy + dz * dz);
      }
    }
  }
  if (bound) bound[r] = sqrtf(reach2) + 1e-4f;
}

// ------------------------------------------------------------------------------------------ neighbour list
// static reach of residue g: max(table bound over all chi, actual backbone atoms), 0 if the residue has no atoms
__global__ void clash_reach_kernel(const float* __restrict__ tbl, const float* __restrict__ X,
                                   const long long* __restrict__ residue_type, const float* __restrict__ atom_exists, int G,
                                   float* __restrict__ reach) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  int type = min(max((int)residue_type[g], 0), 20);
  const float* p = X + (size_t)g * 42;
  float any = 0.f, bbmax = 0.f;
  for (int a = 0; a < 14; ++a) any += atom_exists[(size_t)g * 14 + a];
  for (int a = 0; a < 4; ++a) {
    float dx = p[a * 3] - p[3], dy = p[a * 3 + 1] - p[4], dz = p[a * 3 + 2] - p[5];
    bbmax = fmaxf(bbmax, sqrtf(dx * dx + dy * dy + dz * dz));
  }
  reach[g] = (any > 0.f) ? fmaxf(tbl[(size_t)type * PP_TBL_STRIDE + 132], bbmax + 1e-3f) : -1.f;
}

// one warp per residue; pass 0 counts, pass 1 fills (ascending j, so the summation order is fixed)
__global__ void clash_nbr_kernel(const float* __restrict__ X, const float* __restrict__ reach,
                                 const long long* __restrict__ residue_index, int B, int L, float cutoff, int fill,
                                 int* __restrict__ count, const long long* __restrict__ start, int* __restrict__ list) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  int b = warp / L;
  float ri = reach[warp];
  const float* pi = X + (size_t)warp * 42 + 3;
  float xi = pi[0], yi = pi[1], zi = pi[2];
  long long idx_i = residue_index[warp];
  int n = 0;
  long long base = fill ? start[warp] : 0;
  if (ri >= 0.f) {
    for (int j0 = 0; j0 < L; j0 += 32) {
      int j = j0 + lane;
      bool hit = false;
      if (j < L) {
        int gj = b * L + j;
        float rj = reach[gj];
        if (rj >= 0.f && residue_index[gj] != idx_i) {  // clash.py:166-169: strict '<' on the index VALUE, both ways
          const float* pj = X + (size_t)gj * 42 + 3;
          float dx = pj[0] - xi, dy = pj[1] - yi, dz = pj[2] - zi;
          float lim = ri + rj + cutoff;
          hit = dx * dx + dy * dy + dz * dz < lim * lim;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (fill && hit) list[base + n + __popc(m & ((1u << lane) - 1))] = b * L + j;
      n += __popc(m);
    }
  }
  if (!fill && lane == 0) count[warp] = n;
}

// ------------------------------------------------------------------------------------------ pair kernel
struct AdamArgs {
  float* x;             // [R][4] optimised variable (in/out)
  float* m;             // [R][4]
  float* v;             // [R][4]
  const float* z;       // [R][4] proximal anchor  z = SC_D * mask
  const float* sc_d;    // [R][4] starting angles
  const unsigned char* mask;  // [R][4] clash mask
  float* snapshot;      // [R][4] where(mask, x_new, SC_D)
  const unsigned char* owned;  // [R] or null: residues this rank owns (slab-partitioned complex); others are halo
  float step_size, bc2_sqrt, beta1, beta2, eps;
  float lamda;
  float inv_n_total;    // > 0: 1 / residues of the whole complex (slab-partitioned); else 1 / n_res[complex of the row]
  const int* n_res;     // [B] residues per complex of a padded batch (rows l >= n_res[b] are padding)
  int L;                // padded length: complex of residue g is g / L
  float* loss_rows;     // [R][2] per-row terms of the objective: (clash per residue, sum_k (x' - z)^2), 0 if not counted
};

template <int MODE>  // 0: loss only   1: loss + dL/dchi   2: proximal step (loss, gradient, Adam update)
__global__ void __launch_bounds__(128, 6)
clash_pair_kernel(const float4* __restrict__ atoms, const float* __restrict__ bound, const float* __restrict__ axes,
                  const float* __restrict__ X, const long long* __restrict__ residue_type,
                  const float* __restrict__ atom_exists, const long long* __restrict__ nbr_start,
                  const int* __restrict__ nbr_list, const float* __restrict__ lower, const float* __restrict__ upper,
                  const float* __restrict__ tbl, const float* __restrict__ res_w /*[R] upstream weight or null*/,
                  float tol, float max_cut, int G, int S, float* __restrict__ per_res /*[R]*/,
                  float* __restrict__ grad_chi /*[R][4]*/, AdamArgs ad) {
  const int lane16 = threadIdx.x & 15;
  const int half = (threadIdx.x >> 4) & 1;  // which entries of the neighbour list this half-warp takes
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int R = S * G;
  const bool live = r < R;
  const int rr = live ? r : R - 1;
  const int s = rr / G, g = rr - s * G;
  const int a = lane16;
  const bool atom_ok = live && a < 14;

  float4 pa = atom_ok ? atoms[(size_t)rr * 14 + a] : make_float4(0.f, 0.f, 0.f, 0.f);
  const bool ea = pa.w != 0.f;
  // number of side-chain atoms present (clash.py:342-344)
  float nsc = (atom_ok && a >= 4) ? atom_exists[(size_t)g * 14 + a] : 0.f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) nsc += __shfl_xor_sync(0xffffffffu, nsc, o);
  const float inv_i = 1.f / (1e-10f + nsc);
  // proximal step: the clash term enters the objective as lamda * mean over the residues of the row's own complex
  float inv_n = 0.f;
  if (MODE == 2) {
    const int cb = g / ad.L;
    inv_n = ad.inv_n_total > 0.f ? ad.inv_n_total : 1.f / (float)(ad.n_res ? ad.n_res[cb] : ad.L);
  }
  const float w_uniform = ad.lamda * inv_n;
  const float wi = (MODE == 0) ? 0.f : ((res_w ? res_w[rr] : w_uniform) * inv_i);

  float loss = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
  const float cax = X[(size_t)g * 42 + 3], cay = X[(size_t)g * 42 + 4], caz = X[(size_t)g * 42 + 5];
  const float bi = bound[rr];

  // ---- between residues (clash.py:102-254)
  // The neighbour list (static, built from worst-case extents) is pruned against the CURRENT bounding spheres 32
  // entries at a time, one entry per lane (the loads of a chunk are independent); the survivors are then visited in
  // list order, the two half-warps taking them alternately, with all 14 atom records of a survivor fetched at once.
  const long long n0 = nbr_start[g], n1 = nbr_start[g + 1];
  const int lane = threadIdx.x & 31;
  for (long long nb = n0; nb < n1; nb += 32) {
    const long long idx = nb + lane;
    int gj_l = 0;
    float wj_l = 0.f;
    bool pass = false;
    if (idx < n1) {
      gj_l = nbr_list[idx];
      const float dx = X[(size_t)gj_l * 42 + 3] - cax, dy = X[(size_t)gj_l * 42 + 4] - cay, dz = X[(size_t)gj_l * 42 + 5] - caz;
      const float lim = bi + bound[s * G + gj_l] + max_cut;
      pass = dx * dx + dy * dy + dz * dz < lim * lim;
      if (MODE != 0 && pass) {
        float nj = 0.f;
#pragma unroll
        for (int b = 4; b < 14; ++b) nj += atom_exists[(size_t)gj_l * 14 + b];
        wj_l = (res_w ? res_w[s * G + gj_l] : w_uniform) / (1e-10f + nj);
      }
    }
    unsigned surv = __ballot_sync(0xffffffffu, pass);
    while (surv) {
      // two survivors per round, one for each half-warp: uniform control flow (only a trailing odd survivor leaves
      // the second half-warp idle)
      const int s0 = __ffs(surv) - 1;
      surv &= surv - 1;
      const int s1 = surv ? __ffs(surv) - 1 : -1;
      if (surv) surv &= surv - 1;
      const int src = half ? s1 : s0;
      const int gj = __shfl_sync(0xffffffffu, gj_l, src < 0 ? 0 : src);
      const float wj = __shfl_sync(0xffffffffu, wj_l, src < 0 ? 0 : src);
      if (src < 0) continue;
      const float4* pj = atoms + (size_t)(s * G + gj) * 14;
#pragma unroll
      for (int b0 = 0; b0 < 14; b0 += 7) {
        float4 q[7];
#pragma unroll
        for (int b = 0; b < 7; ++b) q[b] = pj[b0 + b];
#pragma unroll
        for (int bb = 0; bb < 7; ++bb) {
          const int b = b0 + bb;
          const float ex = pa.x - q[bb].x, ey = pa.y - q[bb].y, ez = pa.z - q[bb].z;
          const float d2 = ex * ex + ey * ey + ez * ez;
          const float lim = pa.w + q[bb].w - tol;
          // cheap conservative reject (no sqrt): e = lim - sqrt(1e-10 + d2) can only be positive if d2 < lim^2
          if (d2 > lim * lim * 1.0001f || lim <= 0.f) continue;
          const bool ok = ea && q[bb].w != 0.f && !(a < 4 && b < 4) && !(a == 5 && b == 5);
          const float d = sqrtf(1e-10f + d2);
          const float e = __fsub_rn(__fsub_rn(__fadd_rn(pa.w, q[bb].w), tol), d);  // (r_a + r_b) - tol - d
          if (ok && e > 0.f) {
            loss += e;
            if (MODE != 0) {
              const float c = ((a >= 4) ? wi : 0.f) + ((b >= 4) ? wj : 0.f);
              const float sc = -c / d;  // d e / d p_a = -(p_a - p_b) / d
              fx += sc * ex; fy += sc * ey; fz += sc * ez;
            }
          }
        }
      }
    }
  }
  // the two halves of the survivors (fixed order: even + odd); both half-warps continue with the totals
  loss += __shfl_xor_sync(0xffffffffu, loss, 16);
  fx += __shfl_xor_sync(0xffffffffu, fx, 16);
  fy += __shfl_xor_sync(0xffffffffu, fy, 16);
  fz += __shfl_xor_sync(0xffffffffu, fz, 16);
  // ---- within the residue (clash.py:7-99): every ordered pair adds its error to both atoms
  {
    int type = min(max((int)residue_type[g], 0), 20);
    const float* lo = lower + ((size_t)type * 14 + a) * 14;
    const float* hi = upper + ((size_t)type * 14 + a) * 14;
    float wl = 0.f;
#pragma unroll
    for (int b = 0; b < 14; ++b) {
      float qx = __shfl_sync(0xffffffffu, pa.x, (threadIdx.x & 16) + b);
      float qy = __shfl_sync(0xffffffffu, pa.y, (threadIdx.x & 16) + b);
      float qz = __shfl_sync(0xffffffffu, pa.z, (threadIdx.x & 16) + b);
      float qw = __shfl_sync(0xffffffffu, pa.w, (threadIdx.x & 16) + b);
      bool ok = atom_ok && ea && qw != 0.f && a != b && !(a < 4 && b < 4);
      float ex = pa.x - qx, ey = pa.y - qy, ez = pa.z - qz;
      float d = sqrtf(1e-10f + (ex * ex + ey * ey + ez * ez));
      if (ok) {
        float el = lo[b] - d, eh = d - hi[b];
        float e = fmaxf(el, 0.f) + fmaxf(eh, 0.f);
        wl += e;
        if (MODE != 0) {
          float c = 2.f * (((a >= 4) ? wi : 0.f) + ((b >= 4) ? wi : 0.f));
          float sg = ((eh > 0.f) ? 1.f : 0.f) - ((el > 0.f) ? 1.f : 0.f);
          float sc = c * sg / d;
          fx += sc * ex; fy += sc * ey; fz += sc * ez;
        }
      }
    }
    loss += wl + wl;  // sum over rows + sum over columns of the symmetric error matrix
  }

  // ---- per-residue loss (clash.py:356-363)
  float lsum = (atom_ok && a >= 4) ? loss : 0.f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  const float pr = lsum * inv_i;
  if (live && lane16 == 0 && half == 0 && per_res) per_res[r] = pr;

  float gk[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE != 0) {
    // ---- torque about each chi axis; slots 0-3 are overwritten by the input backbone and carry no dependence
    int grp = atom_ok ? (int)tbl[(size_t)min(max((int)residue_type[g], 0), 20) * PP_TBL_STRIDE + 90 + a] : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = 0.f;
      if (atom_ok && a >= 4 && grp >= 4 + k) {
        const float* ax = axes + ((size_t)rr * 4 + k) * 6;
        float rx = pa.x - ax[3], ry = pa.y - ax[4], rz = pa.z - ax[5];
        float cx = ax[1] * rz - ax[2] * ry, cy = ax[2] * rx - ax[0] * rz, cz = ax[0] * ry - ax[1] * rx;
        t = cx * fx + cy * fy + cz * fz;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      gk[k] = t;
    }
    if (MODE == 1 && live && lane16 < 4 && half == 0) grad_chi[(size_t)r * 4 + lane16] = gk[lane16];
  }

  float sc_term = 0.f;
  if (MODE == 2) {
    // f(x) = mean_res |x' - z|^2 + lamda * mean_res clash(x'),  x' = where(mask, x, SC_D)   (optimize.py:33-45)
    const bool mine = ad.owned == nullptr || ad.owned[rr] != 0;
    if (live && lane16 < 4 && half == 0) {
      size_t o = (size_t)r * 4 + lane16;
      bool mk = ad.mask[o] != 0 && mine;
      float x = ad.x[o], z = ad.z[o], s0 = ad.sc_d[o];
      float xp = mk ? x : s0;
      float diff = xp - z;
      sc_term = diff * diff;
      float gcl = (lane16 == 0) ? gk[0] : (lane16 == 1) ? gk[1] : (lane16 == 2) ? gk[2] : gk[3];
      float grad = mk ? (2.f * diff * inv_n + gcl) : 0.f;
      float m = ad.beta1 * ad.m[o] + (1.f - ad.beta1) * grad;
      float v = ad.beta2 * ad.v[o] + (1.f - ad.beta2) * grad * grad;
      float xn = x - ad.step_size * m / (sqrtf(v) / ad.bc2_sqrt + ad.eps);
      ad.m[o] = m;
      ad.v[o] = v;
      ad.x[o] = xn;
      ad.snapshot[o] = mk ? xn : s0;
      if (!mine) sc_term = 0.f;
    }
    sc_term += __shfl_xor_sync(0xffffffffu, sc_term, 1);
    sc_term += __shfl_xor_sync(0xffffffffu, sc_term, 2);
  }
  if (MODE != 1 && ad.loss_rows && live && lane16 == 0 && half == 0) {
    const bool counted = ad.owned == nullptr || ad.owned[rr] != 0;
    ad.loss_rows[(size_t)r * 2] = counted ? pr : 0.f;
    ad.loss_rows[(size_t)r * 2 + 1] = counted ? sc_term : 0.f;
  }
}

// Objective of one (sample, complex) item from its per-row terms, fixed summation order (128 threads):
//   out[0] = sum_rows |x' - z|^2 / n + lamda * sum_rows clash / n   (optimize.py:37-45),  out[1] = mean clash.
// Item it = s * B + b owns rows s * G + b * L + [0, n_res[b]).
__device__ void prox_reduce_item(const float* __restrict__ loss_rows, int it, int B, int L, const int* __restrict__ n_res,
                                 float lamda, float inv_n_total, float* __restrict__ out, float (*sh)[128]) {
  const int s = it / B, b = it - s * B;
  const int n = n_res ? n_res[b] : L;
  const float* rows = loss_rows + ((size_t)s * B * L + (size_t)b * L) * 2;
  float c = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) { c += rows[2 * i]; q += rows[2 * i + 1]; }
  sh[0][threadIdx.x] = c;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float inv_n = inv_n_total > 0.f ? inv_n_total : 1.f / (float)n;
    out[(size_t)it * 2] = sh[1][0] * inv_n + lamda * (sh[0][0] * inv_n);
    out[(size_t)it * 2 + 1] = sh[0][0] * inv_n;
  }
}

__global__ void __launch_bounds__(128)
prox_reduce_kernel(const float* __restrict__ loss_rows, int B, int L, const int* __restrict__ n_res, float lamda,
                   float inv_n_total, float* __restrict__ out) {
  __shared__ float sh[2][128];
  prox_reduce_item(loss_rows, blockIdx.x, B, L, n_res, lamda, inv_n_total, out, sh);
}

// mask = per_res > mean(per_res of the row's item), expanded to the 4 chi; z = SC_D * mask; x = z; m = v = 0
// (optimize.py:5-31,47).  mean [items][2] (slot 1 = mean clash) or, with one_mean, a single pair for every row.
__global__ void prox_init_kernel(const float* __restrict__ per_res, const float* __restrict__ mean, int one_mean,
                                 const float* __restrict__ sc_d, int R, int B, int L, const unsigned char* __restrict__ owned,
                                 unsigned char* __restrict__ mask, float* __restrict__ z, float* __restrict__ x,
                                 float* __restrict__ m, float* __restrict__ v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 4) return;
  const int r = i >> 2;
  const int it = one_mean ? 0 : r / L;  // = s * B + b
  bool mk = per_res[r] > mean[(size_t)it * 2 + 1] && (owned == nullptr || owned[r] != 0);
  mask[i] = mk;
  float zz = mk ? sc_d[i] : 0.f * sc_d[i];
  z[i] = zz;
  x[i] = zz;
  m[i] = 0.f;
  v[i] = 0.f;
}

}  // namespace pp

using namespace pp;

static const ReduceArgs kNoReduce{nullptr, 0, 0, 0, nullptr, 0.f, 0.f, nullptr};

extern "C" int pp_atom14_fwd(const float* tables, const float* X, const int64_t* residue_type, const float* chi,
                             int64_t G, int64_t S, float* xyz_out, cudaStream_t stream) {
  PP_REQUIRE(tables && X && residue_type && chi && xyz_out, "null pointer");
  PP_REQUIRE(G > 0 && S > 0, "bad sizes");
  long long R = S * G;
  atom14_kernel<<<(unsigned)((R + 127) / 128), 128, 0, stream>>>(tables, X, (const long long*)residue_type, chi, nullptr,
                                                                nullptr, (int)G, (int)S, xyz_out, nullptr, nullptr,
                                                                nullptr, nullptr, kNoReduce);
  return check_launch("pp_atom14_fwd");
}

// Residue neighbour list for the clash term.  Call with fill = 0 to get counts[B*L] (and reach[B*L]), build
// start[B*L+1] as their exclusive prefix sum, then call with fill = 1 to write list[start[B*L]].
extern "C" int pp_clash_neighbours(const float* tables, const float* X, const int64_t* residue_type,
                                   const float* atom_exists, const int64_t* residue_index, int64_t B, int64_t L,
                                   float cutoff, int64_t fill, float* reach, int32_t* counts, const int64_t* start,
                                   int32_t* list, cudaStream_t stream) {
  PP_REQUIRE(tables && X && residue_type && atom_exists && residue_index && reach, "null pointer");
  PP_REQUIRE(B > 0 && L > 0, "bad sizes");
  PP_REQUIRE(fill ? (start && list) : (counts != nullptr), "missing output for this pass");
  long long G = B * L;
  if (!fill) clash_reach_kernel<<<(unsigned)((G + 127) / 128), 128, 0, stream>>>(tables, X, (const long long*)residue_type,
                                                                                atom_exists, (int)G, reach);
  clash_nbr_kernel<<<(unsigned)((G * 32 + 255) / 256), 256, 0, stream>>>(X, reach, (const long long*)residue_index, (int)B,
                                                                        (int)L, cutoff, (int)fill, counts,
                                                                        (const long long*)start, list);
  return check_launch("pp_clash_neighbours");
}

// reach [G] of every residue (bound on |CA - atom| over all chi; -1 if the residue has no atoms).
extern "C" int pp_clash_reach(const float* tables, const float* X, const int64_t* residue_type, const float* atom_exists,
                              int64_t G, float* reach, cudaStream_t stream) {
  PP_REQUIRE(tables && X && residue_type && atom_exists && reach, "null pointer");
  PP_REQUIRE(G > 0, "bad sizes");
  clash_reach_kernel<<<(unsigned)((G + 127) / 128), 128, 0, stream>>>(tables, X, (const long long*)residue_type,
                                                                     atom_exists, (int)G, reach);
  return check_launch("pp_clash_reach");
}

// mode 0: per_res only.  mode 1: per_res and grad_chi = d(sum_r res_w[r] * per_res[r]) / d chi.
// Workspaces: atoms4 [S*G][14] float4, axes [S*G][4][6], bound [S*G].
extern "C" int pp_clash_fwd_bwd(const float* tables, const float* lower, const float* upper, const float* X,
                                const int64_t* residue_type, const float* atom_exists, const int64_t* nbr_start,
                                const int32_t* nbr_list, const float* chi, int64_t G, int64_t S, float tol, float max_cut,
                                int64_t mode, const float* res_w, float* per_res, float* grad_chi, float* atoms4,
                                float* axes, float* bound, cudaStream_t stream) {
  PP_REQUIRE(tables && lower && upper && X && residue_type && atom_exists && nbr_start && nbr_list && chi, "null pointer");
  PP_REQUIRE(per_res && atoms4 && axes && bound, "null output/workspace");
  PP_REQUIRE(mode == 0 || (mode == 1 && res_w && grad_chi), "mode 1 needs res_w and grad_chi");
  PP_REQUIRE(G > 0 && S > 0, "bad sizes");
  long long R = S * G;
  atom14_kernel<<<(unsigned)((R + 127) / 128), 128, 0, stream>>>(tables, X, (const long long*)residue_type, chi, nullptr,
                                                                nullptr, (int)G, (int)S, nullptr, atom_exists,
                                                                (float4*)atoms4, axes, bound, kNoReduce);
  AdamArgs ad{};
  unsigned blocks = (unsigned)((R + 3) / 4);
  if (mode == 0)
    clash_pair_kernel<0><<<blocks, 128, 0, stream>>>((const float4*)atoms4, bound, axes, X, (const long long*)residue_type,
                                                     atom_exists, (const long long*)nbr_start, nbr_list, lower, upper,
                                                     tables, nullptr, tol, max_cut, (int)G, (int)S, per_res, nullptr, ad);
  else
    clash_pair_kernel<1><<<blocks, 128, 0, stream>>>((const float4*)atoms4, bound, axes, X, (const long long*)residue_type,
                                                     atom_exists, (const long long*)nbr_start, nbr_list, lower, upper,
                                                     tables, res_w, tol, max_cut, (int)G, (int)S, per_res, grad_chi, ad);
  return check_launch("pp_clash_fwd_bwd");
}

// ---- proximal_optimizer (optimize.py:5-73) for S samples of a padded batch of B complexes: items (s, b) are
//      independent problems that share one launch; item (s, b) owns rows s*B*L + b*L + [0, n_res[b]).

// Clash mask and optimiser state from the starting angles (optimize.py:5-31,47-51): one loss evaluation, its mean
// per item, mask = per_res > mean, z = SC_D*mask, x = z, Adam moments zero.  mean_out [S*B][2] = {unused, mean}.
extern "C" int pp_prox_init(const float* tables, const float* lower, const float* upper, const float* X,
                            const int64_t* residue_type, const float* atom_exists, const int64_t* nbr_start,
                            const int32_t* nbr_list, const float* sc_d, int64_t B, int64_t L, int64_t S,
                            const int32_t* n_res, float tol, float max_cut, uint8_t* mask, float* z, float* x, float* m,
                            float* v, float* per_res, float* mean_out, float* atoms4, float* axes, float* bound,
                            float* loss_rows, cudaStream_t stream) {
  PP_REQUIRE(tables && lower && upper && X && residue_type && atom_exists && nbr_start && nbr_list && sc_d, "null pointer");
  PP_REQUIRE(mask && z && x && m && v && per_res && mean_out && atoms4 && axes && bound && loss_rows, "null output");
  PP_REQUIRE(B > 0 && L > 0 && S > 0, "bad sizes");
  const long long G = B * L, R = S * G;
  atom14_kernel<<<(unsigned)((R + 127) / 128), 128, 0, stream>>>(tables, X, (const long long*)residue_type, sc_d, nullptr,
                                                                nullptr, (int)G, (int)S, nullptr, atom_exists,
                                                                (float4*)atoms4, axes, bound, kNoReduce);
  AdamArgs ad{};
  ad.L = (int)L;
  ad.loss_rows = loss_rows;
  clash_pair_kernel<0><<<(unsigned)((R + 3) / 4), 128, 0, stream>>>(
      (const float4*)atoms4, bound, axes, X, (const long long*)residue_type, atom_exists, (const long long*)nbr_start,
      nbr_list, lower, upper, tables, nullptr, tol, max_cut, (int)G, (int)S, per_res, nullptr, ad);
  prox_reduce_kernel<<<(unsigned)(S * B), 128, 0, stream>>>(loss_rows, (int)B, (int)L, n_res, 1.f, 0.f, mean_out);
  prox_init_kernel<<<(unsigned)((R * 4 + 255) / 256), 256, 0, stream>>>(per_res, mean_out, 0, sc_d, (int)R, (int)B, (int)L,
                                                                        nullptr, mask, z, x, m, v);
  return check_launch("pp_prox_init");
}

// Slab-partitioned variant of the second half of pp_prox_init: per_res and mean[1] (the mean over the WHOLE complex,
// all-reduced by the caller) are given; residues with owned == 0 are halo copies and are never optimised here.
extern "C" int pp_prox_init_from_mean(const float* per_res, const float* mean, const float* sc_d, const uint8_t* owned,
                                      int64_t G, uint8_t* mask, float* z, float* x, float* m, float* v,
                                      cudaStream_t stream) {
  PP_REQUIRE(per_res && mean && sc_d && mask && z && x && m && v, "null pointer");
  PP_REQUIRE(G > 0, "bad sizes");
  prox_init_kernel<<<(unsigned)((G * 4 + 255) / 256), 256, 0, stream>>>(per_res, mean, 1, sc_d, (int)G, 1, (int)G, owned, mask,
                                                                        z, x, m, v);
  return check_launch("pp_prox_init_from_mean");
}

// One proximal step (optimize.py:60-71): loss and gradient at the current x, Adam update, snapshot.  Two launches:
// the atom14 rebuild - whose tail blocks reduce the objective of the PREVIOUS step into prev_loss_out [S*B][2] when it
// is non-null - and the pair kernel, which leaves this step's per-row terms in loss_rows.  pp_prox_loss reduces the
// last step.  loss[it][0] = objective BEFORE the update (what the reference appends to loss_list), [1] = mean clash.
extern "C" int pp_prox_step(const float* tables, const float* lower, const float* upper, const float* X,
                            const int64_t* residue_type, const float* atom_exists, const int64_t* nbr_start,
                            const int32_t* nbr_list, const float* sc_d, const uint8_t* mask, const float* z, float* x,
                            float* m, float* v, int64_t B, int64_t L, int64_t S, const int32_t* n_res, float tol,
                            float max_cut, float lamda, float step_size, float bc2_sqrt, float beta1, float beta2,
                            float eps, float* snapshot, float* prev_loss_out, float* per_res, float* atoms4, float* axes,
                            float* bound, float* loss_rows, const uint8_t* owned, int64_t n_total, cudaStream_t stream) {
  PP_REQUIRE(tables && lower && upper && X && residue_type && atom_exists && nbr_start && nbr_list && sc_d, "null pointer");
  PP_REQUIRE(mask && z && x && m && v && snapshot && per_res && atoms4 && axes && bound && loss_rows, "null output");
  PP_REQUIRE(B > 0 && L > 0 && S > 0, "bad sizes");
  const long long G = B * L, R = S * G;
  const float inv_n_total = n_total > 0 ? 1.f / (float)n_total : 0.f;  // slab: mean over the WHOLE complex
  const unsigned nb = (unsigned)((R + 127) / 128);
  ReduceArgs red{prev_loss_out ? loss_rows : nullptr, (int)nb, (int)B, (int)L, n_res, lamda, inv_n_total, prev_loss_out};
  atom14_kernel<<<nb + (prev_loss_out ? (unsigned)(S * B) : 0u), 128, 0, stream>>>(
      tables, X, (const long long*)residue_type, x, sc_d, mask, (int)G, (int)S, nullptr, atom_exists, (float4*)atoms4, axes,
      bound, red);
  AdamArgs ad{x, m, v, z, sc_d, mask, snapshot, owned, step_size, bc2_sqrt, beta1, beta2, eps, lamda, inv_n_total, n_res,
              (int)L, loss_rows};
  clash_pair_kernel<2><<<(unsigned)((R + 3) / 4), 128, 0, stream>>>(
      (const float4*)atoms4, bound, axes, X, (const long long*)residue_type, atom_exists, (const long long*)nbr_start,
      nbr_list, lower, upper, tables, nullptr, tol, max_cut, (int)G, (int)S, per_res, nullptr, ad);
  return check_launch("pp_prox_step");
}

// Objective of the step whose per-row terms are in loss_rows (the last step of a loop) -> loss_out [S*B][2].
extern "C" int pp_prox_loss(const float* loss_rows, int64_t B, int64_t L, int64_t S, const int32_t* n_res, float lamda,
                            int64_t n_total, float* loss_out, cudaStream_t stream) {
  PP_REQUIRE(loss_rows && loss_out, "null pointer");
  PP_REQUIRE(B > 0 && L > 0 && S > 0, "bad sizes");
  prox_reduce_kernel<<<(unsigned)(S * B), 128, 0, stream>>>(loss_rows, (int)B, (int)L, n_res, lamda,
                                                            n_total > 0 ? 1.f / (float)n_total : 0.f, loss_out);
  return check_launch("pp_prox_loss");
}
