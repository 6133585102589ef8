// Batch featurisation on the device: raw atom records of a padded batch -> the per-residue tensors the sampling
// path consumes.  Restates ComplexDataset.prot_to_data (reference src/datamodules/components/complex_dataset.py:
// 64-148) with calc_dihedrals / calc_bb_dihedrals / calc_sc_dihedrals (src/datamodules/components/helper.py:20-101)
// and the zero padding of collate_fn (src/datamodules/complex_datamodule.py:196-226); the host version is
// packppi_b200/featurize.py.  One thread per residue slot of the padded batch.
//
// Dihedral of four points as the reference computes it: unit bond vectors u (a zero or non-finite bond -> 0),
// n2 = unit(u2 x u1), n1 = unit(u1 x u0), angle = sign(u2 . n1) * acos(clamp(n2 . n1, -1, 1)).
#include "common.cuh"

namespace pp {

struct V3 {
  float x, y, z;
};
// Every operation below is rounded where torch-CPU rounds it (measured, see csrc/encoder.cu and
// tests/test_dihedral_rounding.py): cross = fma(a1, b2, -rn(a2 b1)), norm = sqrt(fma(z, z, fma(y, y, rn(x x)))),
// (a * b).sum(-1) = (rn(a0 b0) + rn(a1 b1)) + rn(a2 b2).  acos is ill-conditioned near +-1 (d acos = d c / sin), so a
// cosine that differs in the last bit moves a near-planar dihedral by ~3e-4 rad; with the same bits it cannot.
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
          __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x))};
}
__device__ __forceinline__ float dot(V3 a, V3 b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
}
// t / |t| with nan_to_num: NaN -> 0 (helper.py:16-18); an infinite component cannot occur for finite / NaN input
__device__ __forceinline__ V3 unit(V3 t) {
  const float n = __fsqrt_rn(__fmaf_rn(t.z, t.z, __fmaf_rn(t.y, t.y, __fmul_rn(t.x, t.x))));
  V3 r = {__fdiv_rn(t.x, n), __fdiv_rn(t.y, n), __fdiv_rn(t.z, n)};
  if (!(isfinite(r.x) && isfinite(r.y) && isfinite(r.z))) {
    r.x = isfinite(r.x) ? r.x : 0.f;
    r.y = isfinite(r.y) ? r.y : 0.f;
    r.z = isfinite(r.z) ? r.z : 0.f;
  }
  return r;
}
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
// dihedral from three consecutive unit bonds b0 (p0->p1), b1, b2   (helper.py:20-36: u2 = b0, u1 = b1, u0 = b2)
__device__ __forceinline__ float dihedral(V3 b0, V3 b1, V3 b2) {
  const V3 n2 = unit(cross(b0, b1)), n1 = unit(cross(b1, b2));
  const float c = fminf(fmaxf(dot(n2, n1), -1.f), 1.f);  // 1 - 1e-8 == 1 in fp32
  return sgn(dot(b0, n1)) * acosf(c);
}
__device__ __forceinline__ V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }

__global__ void featurize_kernel(const float* __restrict__ Xin, const long long* __restrict__ aatype,
                                 const float* __restrict__ atom_mask_in, const long long* __restrict__ ridx_in,
                                 const long long* __restrict__ chain_in, const int* __restrict__ length, int B, int L,
                                 const int* __restrict__ chi_atoms /*[21][7]*/, const float* __restrict__ chi_mask /*[21][4]*/,
                                 const float* __restrict__ chi_pi /*[21][4]*/, float* __restrict__ X, float* __restrict__ atom_mask,
                                 long long* __restrict__ rtype, float* __restrict__ rmask, long long* __restrict__ ridx,
                                 long long* __restrict__ chain, float* __restrict__ bb_d, float* __restrict__ bb_sc,
                                 float* __restrict__ bb_m, float* __restrict__ sc_d, float* __restrict__ sc_sc,
                                 float* __restrict__ sc_m, unsigned char* __restrict__ p1, unsigned char* __restrict__ p2) {
  const int gidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= B * L) return;
  const int b = gidx / L, i = gidx - b * L;
  const int n = length[b];
  const size_t g = (size_t)gidx;
  if (i >= n) {  // padding slot: zeros everywhere (collate_fn)
    for (int q = 0; q < 42; ++q) X[g * 42 + q] = 0.f;
    for (int q = 0; q < 14; ++q) atom_mask[g * 14 + q] = 0.f;
    rtype[g] = 0; rmask[g] = 0.f; ridx[g] = 0; chain[g] = 0;
    for (int q = 0; q < 3; ++q) { bb_d[g * 3 + q] = 0.f; bb_m[g * 3 + q] = 0.f; bb_sc[g * 6 + 2 * q] = 0.f; bb_sc[g * 6 + 2 * q + 1] = 0.f; }
    for (int q = 0; q < 4; ++q) {
      sc_d[g * 4 + q] = 0.f; sc_m[g * 4 + q] = 0.f; sc_sc[g * 8 + 2 * q] = 0.f; sc_sc[g * 8 + 2 * q + 1] = 0.f;
      p1[g * 4 + q] = 0; p2[g * 4 + q] = 0;
    }
    return;
  }
  const float* x = Xin + g * 42;
  // residue mask: N, CA, C, O all finite (complex_dataset.py:94)
  float s4 = 0.f;
  for (int q = 0; q < 12; ++q) s4 += x[q];
  const float rm = isfinite(s4) ? 1.f : 0.f;
  const long long aa = aatype[g];
  // ---- backbone dihedrals (pre-omega, phi, psi) over the residues of the array, masks from residue_index continuity
  const V3 N = ld3(x), CA = ld3(x + 3), C = ld3(x + 6);
  const bool has_prev = i > 0, has_next = i + 1 < n;
  float d[3] = {0.f, 0.f, 0.f}, m[3] = {0.f, 0.f, 0.f};
  const V3 u_nca = unit(sub(CA, N)), u_cac = unit(sub(C, CA));
  if (has_prev) {
    const float* xp = x - 42;
    const V3 CAp = ld3(xp + 3), Cp = ld3(xp + 6);
    const V3 u_cap_cp = unit(sub(Cp, CAp)), u_cp_n = unit(sub(N, Cp));
    d[0] = dihedral(u_cap_cp, u_cp_n, u_nca);  // CA-1, C-1, N, CA
    d[1] = dihedral(u_cp_n, u_nca, u_cac);     // C-1, N, CA, C
    const float pre = (ridx_in[g] - 1 == ridx_in[g - 1]) ? 1.f : 0.f;
    m[0] = pre; m[1] = pre;
  }
  if (has_next) {
    const float* xn = x + 42;
    const V3 Nn = ld3(xn);
    d[2] = dihedral(u_nca, u_cac, unit(sub(Nn, C)));  // N, CA, C, N+1
    m[2] = (ridx_in[g] + 1 == ridx_in[g + 1]) ? 1.f : 0.f;
  }
  for (int q = 0; q < 3; ++q) {
    const float dv = d[q] * rm, mv = m[q] * rm;
    bb_d[g * 3 + q] = dv;
    bb_m[g * 3 + q] = mv;
    // sincos of the unmasked angle times the mask, then the residue mask (complex_dataset.py:102-105,131-133)
    bb_sc[g * 6 + 2 * q] = sinf(d[q]) * m[q] * rm;
    bb_sc[g * 6 + 2 * q + 1] = cosf(d[q]) * m[q] * rm;
  }
  // ---- side-chain dihedrals along the 7-atom chi path of the residue type
  const int* path = chi_atoms + aa * 7;
  V3 ub[6];
  {
    V3 prev = ld3(x + path[0] * 3);
    for (int q = 0; q < 6; ++q) {
      const V3 cur = ld3(x + path[q + 1] * 3);
      ub[q] = unit(sub(cur, prev));
      prev = cur;
    }
  }
  for (int q = 0; q < 4; ++q) {
    float v = dihedral(ub[q], ub[q + 1], ub[q + 2]);
    v = (isfinite(v) ? v : 0.f) * chi_mask[aa * 4 + q];
    const float mk = v != 0.f ? 1.f : 0.f;
    sc_d[g * 4 + q] = v * rm;
    sc_m[g * 4 + q] = mk * rm;
    sc_sc[g * 8 + 2 * q] = sinf(v) * mk * rm;
    sc_sc[g * 8 + 2 * q + 1] = cosf(v) * mk * rm;
    const bool pi = chi_pi[aa * 4 + q] != 0.f;
    p1[g * 4 + q] = (mk * rm != 0.f) && pi && rm != 0.f;
    p2[g * 4 + q] = (mk * rm != 0.f) && !pi && rm != 0.f;
  }
  // ---- masked copies; NaN (missing atoms) -> 0
  for (int q = 0; q < 42; ++q) {
    const float v = x[q] * rm;
    X[g * 42 + q] = isfinite(v) ? v : 0.f;
  }
  for (int q = 0; q < 14; ++q) atom_mask[g * 14 + q] = atom_mask_in[g * 14 + q] * rm;
  rtype[g] = rm != 0.f ? aa : 0;
  rmask[g] = rm;
  ridx[g] = rm != 0.f ? ridx_in[g] : 0;
  chain[g] = rm != 0.f ? chain_in[g] : 0;
}

}  // namespace pp

// Device featurisation of a padded batch [B][L].  Inputs: atom14 coordinates (NaN = missing atom),
// residue types, atom masks, residue indices (already offset per chain, complex_dataset.py:86-92), 1-based chain
// numbers, the residue count of every complex; chi tables from packppi_b200/data/tables.npz.  Outputs: every tensor
// field of the batch contract (packppi_b200/batch.py TENSOR_FIELDS), zero in the padding.
extern "C" int pp_featurize(const float* X_in, const int64_t* aatype, const float* atom_mask_in, const int64_t* ridx_in,
                            const int64_t* chain_in, const int32_t* length, int64_t B, int64_t L, const int32_t* chi_atoms,
                            const float* chi_mask, const float* chi_pi, float* X, float* atom_mask, int64_t* residue_type,
                            float* residue_mask, int64_t* residue_index, int64_t* chain_indices, float* bb_d, float* bb_sincos,
                            float* bb_mask, float* sc_d, float* sc_sincos, float* sc_mask, uint8_t* chi_1pi, uint8_t* chi_2pi,
                            cudaStream_t stream) {
  PP_REQUIRE(X_in && aatype && atom_mask_in && ridx_in && chain_in && length && chi_atoms && chi_mask && chi_pi,
             "null input");
  PP_REQUIRE(X && atom_mask && residue_type && residue_mask && residue_index && chain_indices && bb_d && bb_sincos &&
                 bb_mask && sc_d && sc_sincos && sc_mask && chi_1pi && chi_2pi,
             "null output");
  PP_REQUIRE(B > 0 && L > 0, "bad sizes");
  const long long n = B * L;
  pp::featurize_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
      X_in, reinterpret_cast<const long long*>(aatype), atom_mask_in, reinterpret_cast<const long long*>(ridx_in),
      reinterpret_cast<const long long*>(chain_in), length, (int)B, (int)L, chi_atoms, chi_mask, chi_pi, X, atom_mask,
      reinterpret_cast<long long*>(residue_type), residue_mask, reinterpret_cast<long long*>(residue_index),
      reinterpret_cast<long long*>(chain_indices), bb_d, bb_sincos, bb_mask, sc_d, sc_sincos, sc_mask, chi_1pi, chi_2pi);
  return pp::check_launch("pp_featurize");
}
