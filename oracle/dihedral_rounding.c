/* CPU ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's inter-residue dihedral feature
 * (/root/reference/src/models/components/encoder.py:155-174, `_normalize` + `_dihedral_from_four_points`)
 * with the fp32 rounding sequence torch 2.11 CPU executes for it, one explicit operation per rounding:
 *   torch.cross      c_k = fma(a_k1, b_k2, -rn(a_k2 * b_k1))       (aten cross kernel, contracted by the compiler)
 *   torch.norm       sqrt(fma(z, z, fma(y, y, rn(x * x))))          (norm_reduce scalar tail, contracted)
 *   (a * b).sum(-1)  (rn(a0 b0) + rn(a1 b1)) + rn(a2 b2)
 *   nan_to_num       NaN -> 0, +-inf -> +-FLT_MAX
 * tests/test_dihedral_rounding.py pins this file against torch itself (bit-exact cosine and sign) and the CUDA
 * kernel (csrc/encoder.cu dihedral4) repeats the same sequence with __fmaf_rn / __fmul_rn / __fadd_rn.
 * Compile with -ffp-contract=off so that only the fmaf() calls below fuse.
 */
#include <float.h>
#include <math.h>

static void cross3(const float* a, const float* b, float* o) {
  o[0] = fmaf(a[1], b[2], -(a[2] * b[1]));
  o[1] = fmaf(a[2], b[0], -(a[0] * b[2]));
  o[2] = fmaf(a[0], b[1], -(a[1] * b[0]));
}

static float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

static float nan_to_num(float v) {
  if (v != v) return 0.f;
  if (isinf(v)) return v > 0 ? FLT_MAX : -FLT_MAX;
  return v;
}

static void unit_nan0(float* v) {
  float n = sqrtf(fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0])));
  for (int k = 0; k < 3; ++k) v[k] = nan_to_num(v[k] / n);
}

/* p0..p3: [n][3]; out_cos, out_sign, out_angle: [n] */
void pp_oracle_dihedral(const float* p0, const float* p1, const float* p2, const float* p3, long n, float* out_cos,
                        float* out_sign, float* out_angle) {
  for (long i = 0; i < n; ++i) {
    float u0[3], u1[3], u2[3], n1[3], n2[3], c[3];
    for (int k = 0; k < 3; ++k) {
      u0[k] = p2[3 * i + k] - p1[3 * i + k];
      u1[k] = p0[3 * i + k] - p1[3 * i + k];
      u2[k] = p3[3 * i + k] - p2[3 * i + k];
    }
    cross3(u0, u1, n1);
    cross3(u0, u2, n2);
    unit_nan0(n1);
    unit_nan0(n2);
    cross3(u1, u2, c);
    float s = dot3(c, u0);
    float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
    float cs = dot3(n1, n2);
    out_cos[i] = cs;
    out_sign[i] = sg;
    out_angle[i] = nan_to_num(sg * acosf(cs));
  }
}
