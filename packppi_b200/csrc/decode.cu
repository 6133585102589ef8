// Score decoder fused with the reverse-ODE update of the chi angles.
//
// Replaces decoder_score (reference src/models/TorsionalDiffusion.py:62-68,106-108: Linear 128-64 relu 64-32, ReLU,
// 32-16 relu 16-4) and, when `do_step` is set, the two SO2VESchedule.step calls plus the wrap and mask of the
// sampling loop (schedule.py:198-235 ode branch, TorsionalDiffusion.py:272-280):
//     chi <- wrap(chi + [step_mask] * c * (score * w)) * SC_D_mask ,  wrap(x) = (x + pi) mod 2pi - pi
// with c = 0.5 g(t)^2 dt and w = annealed weight, both evaluated on the host in fp32 exactly as the reference's
// 0-dim tensor arithmetic does, so the kernel only multiplies.  SDE branch (schedule.py:224-228), selected by passing
// the two noise tensors the reference draws with torch.normal (one per schedule.step call):
//     chi <- wrap(chi + [step_mask] (c (score * w) + d * noise)) * SC_D_mask,  c = g^2 dt, d = g sqrt(dt),
// noise = noise_1pi where the chi is pi-periodic, else noise_2pi.  Without injected tensors (d != 0, noise pointers
// NULL) the standard normals come from a counter-based generator inside the kernel: Philox4x32-10 keyed by the call's
// seed, counter = (residue row, step), Box-Muller on the four words - one independent stream per (item, step), no
// [steps, 2, rows, 4] noise tensor in memory.
#include "common.cuh"
#include "weights_layout.h"

namespace pp {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// standard normal number `which` (0..3) of the block (row, step) of stream `seed`
__device__ __forceinline__ float philox_normal(unsigned long long seed, uint32_t row, uint32_t step, int which) {
  uint32_t w[4];
  philox4x32_10(row, step, 0x5043504Bu /* "PCPK" */, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const int p = which & 2;  // words (0,1) -> normals 0,1 ; words (2,3) -> normals 2,3
  const float u1 = (float)(w[p] >> 8) * (1.f / 16777216.f) + (0.5f / 16777216.f);      // (0, 1)
  const float u2 = (float)(w[p + 1] >> 8) * (1.f / 16777216.f);                         // [0, 1)
  const float rad = sqrtf(-2.f * logf(u1));
  float sn, cs;
  sincosf(6.283185307179586f * u2, &sn, &cs);
  return rad * ((which & 1) ? sn : cs);
}

// one warp per residue row; activations are exchanged through shuffles
__global__ void decode_step_kernel(const float* __restrict__ W, const float* __restrict__ hV, int G, int S,
                                   float* __restrict__ score_out, int do_step, float c_ode, float w_anneal,
                                   const unsigned char* __restrict__ step_mask /*[G][4]*/,
                                   const float* __restrict__ chi_mask /*[G][4]*/, float* __restrict__ chi /*[R][4]*/,
                                   const float* __restrict__ noise_1pi, const float* __restrict__ noise_2pi /*[R][4]*/,
                                   const unsigned char* __restrict__ mask_1pi /*[G][4]*/, float d_sde,
                                   unsigned long long seed, int step_index) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= S * G) return;
  const float* W0 = W + PP_OFF(DEC_W0);
  const float* W1 = W + PP_OFF(DEC_W1);
  const float* W2 = W + PP_OFF(DEC_W2);
  const float* W3 = W + PP_OFF(DEC_W3);
  float4 h = *reinterpret_cast<const float4*>(hV + (size_t)r * 128 + lane * 4);  // inputs lane*4 .. lane*4+3

  // 128 -> 64 : lane owns outputs 2*lane, 2*lane+1
  float2 a0 = *reinterpret_cast<const float2*>(W + PP_OFF(DEC_B0) + lane * 2);
#pragma unroll 4
  for (int src = 0; src < 32; ++src) {
    float x = __shfl_sync(0xffffffffu, h.x, src), y = __shfl_sync(0xffffffffu, h.y, src);
    float z = __shfl_sync(0xffffffffu, h.z, src), w = __shfl_sync(0xffffffffu, h.w, src);
    const float* wr = W0 + (size_t)(src * 4) * 64 + lane * 2;
    float2 w0 = *reinterpret_cast<const float2*>(wr), w1 = *reinterpret_cast<const float2*>(wr + 64);
    float2 w2 = *reinterpret_cast<const float2*>(wr + 128), w3 = *reinterpret_cast<const float2*>(wr + 192);
    a0.x = fmaf(x, w0.x, a0.x); a0.y = fmaf(x, w0.y, a0.y);
    a0.x = fmaf(y, w1.x, a0.x); a0.y = fmaf(y, w1.y, a0.y);
    a0.x = fmaf(z, w2.x, a0.x); a0.y = fmaf(z, w2.y, a0.y);
    a0.x = fmaf(w, w3.x, a0.x); a0.y = fmaf(w, w3.y, a0.y);
  }
  a0.x = fmaxf(a0.x, 0.f);
  a0.y = fmaxf(a0.y, 0.f);
  // 64 -> 32 : lane owns output lane; then the nn.ReLU between the two MLPs
  float a1 = W[PP_OFF(DEC_B1) + lane];
#pragma unroll 8
  for (int src = 0; src < 32; ++src) {
    float x = __shfl_sync(0xffffffffu, a0.x, src), y = __shfl_sync(0xffffffffu, a0.y, src);
    a1 = fmaf(x, W1[(size_t)(src * 2) * 32 + lane], a1);
    a1 = fmaf(y, W1[(size_t)(src * 2 + 1) * 32 + lane], a1);
  }
  a1 = fmaxf(a1, 0.f);
  // 32 -> 16 (relu) : lanes 0..15
  float a2 = W[PP_OFF(DEC_B2) + (lane & 15)];
#pragma unroll 8
  for (int src = 0; src < 32; ++src) a2 = fmaf(__shfl_sync(0xffffffffu, a1, src), W2[src * 16 + (lane & 15)], a2);
  a2 = fmaxf(a2, 0.f);
  // 16 -> 4 : lanes 0..3
  float a3 = W[PP_OFF(DEC_B3) + (lane & 3)];
#pragma unroll
  for (int src = 0; src < 16; ++src) a3 = fmaf(__shfl_sync(0xffffffffu, a2, src), W3[src * 4 + (lane & 3)], a3);

  if (lane < 4) {
    size_t o = (size_t)r * 4 + lane;
    if (score_out) score_out[o] = a3;
    if (do_step) {
      int g = r % G;
      float x = chi[o];
      if (step_mask[(size_t)g * 4 + lane]) {
        float drift = __fmul_rn(c_ode, __fmul_rn(a3, w_anneal));
        if (noise_1pi) {  // SDE: drift + diffusion, rounded like (drift + diffusion) then x += ...
          float n = mask_1pi[(size_t)g * 4 + lane] ? noise_1pi[o] : noise_2pi[o];
          drift = __fadd_rn(drift, __fmul_rn(d_sde, n));
        } else if (d_sde != 0.f) {  // SDE with the in-kernel counter-based generator
          drift = __fadd_rn(drift, __fmul_rn(d_sde, philox_normal(seed, (uint32_t)r, (uint32_t)step_index, lane)));
        }
        x = __fadd_rn(x, drift);
      }
      float y = fmodf(__fadd_rn(x, PP_PI_F), PP_TWO_PI_F);
      if (y != 0.f && y < 0.f) y = __fadd_rn(y, PP_TWO_PI_F);  // python-style remainder (torch %)
      chi[o] = __fmul_rn(__fsub_rn(y, PP_PI_F), chi_mask[(size_t)g * 4 + lane]);
    }
  }
}

}  // namespace pp

extern "C" int pp_decode_step(const float* weights, const float* hV, int64_t G, int64_t S, float* score_out,
                              int64_t do_step, float c_ode, float w_anneal, const uint8_t* step_mask,
                              const float* chi_mask, float* chi, const float* noise_1pi, const float* noise_2pi,
                              const uint8_t* mask_1pi, float d_sde, int64_t seed, int64_t step_index,
                              cudaStream_t stream) {
  PP_REQUIRE(weights && hV, "null pointer");
  PP_REQUIRE(G > 0 && S > 0, "bad sizes");
  PP_REQUIRE(score_out || do_step, "nothing to do");
  PP_REQUIRE(!do_step || (step_mask && chi_mask && chi), "step needs step_mask, chi_mask and chi");
  PP_REQUIRE(!noise_1pi || (noise_2pi && mask_1pi), "the SDE step needs both noise tensors and the pi-periodic mask");
  long long R = S * G;
  pp::decode_step_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, stream>>>(
      weights, hV, (int)G, (int)S, score_out, (int)do_step, c_ode, w_anneal, step_mask, chi_mask, chi, noise_1pi,
      noise_2pi, mask_1pi, d_sde, (unsigned long long)seed, (int)step_index);
  return pp::check_launch("pp_decode_step");
}
