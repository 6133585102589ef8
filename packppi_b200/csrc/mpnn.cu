// Invariant-point message passing (IPMP) layers, fp32 CUDA-core path.
//
// Replaces InvariantPointMessagePassing.forward / _get_message_input and MpnnNet.forward
// (reference src/models/components/layers.py:36-148, mpnn.py:47-62), which materialise a [B,L,K,456] message
// tensor twice per layer and run it through torch.gather + nn.Linear.  Here the 456-wide first Linear is split by
// operand:
//     W_in * [h_V_i | h_E_ik | h_V_j | own points 32 | pair geometry 40]
//   = (W_a h_V_i + W_gs own_i + b1)  +  W_n h_V_j  +  [W_e | W_gp] [h_E_ik | pair_ik]
//     `------- per residue: A_i ----'   `- per residue: N_j, gathered -'   `-- per edge, K = 168 --'
// so only a 168-wide GEMM remains per edge.  For the node update the last Linear commutes with the masked mean over
// K (mean_k m_k (W3 x_k + b3) = W3 mean_k(m_k x_k) + b3 mean_k(m_k)), so it runs once per residue.
//
// Kernels (one CTA = 256 threads = one tile of 128 rows):
//   node_pre_kernel   per residue: IPMP points, A_i, N_i                      [S*G rows]
//   edge_node_kernel  per edge:    x2 = relu(W2 relu(A_i+N_j+W_eg[h_E|pair])+b2); masked sum over K
//   node_post_kernel  per residue: W3, residual+LN0, FFN+LN1, mask           -> h_V
//   edge_edge_kernel  per edge:    full 3-layer message MLP, residual+LN2, FFN+LN3, mask -> h_E
// Rows: r = s*G + g (s = diffusion sample, g = residue of the padded batch).  The step-invariant h_E0 is shared by
// all samples of a complex (he_shared != 0).
#include "tile_gemm.cuh"
#include "weights_layout.h"

namespace pp {

struct MpnnCtx {
  const float* geo;    // [G][PP_GEO_STRIDE]
  const int* nbr;      // [G][K] neighbour residue (global row of the padded batch)
  const float* matt;   // [G][K] mask_attend (mpnn.py:49-50)
  const float* rmask;  // [G]
  int G, K, S;
};

struct PathWeights {  // one message path (node or edge) of one layer
  const float *WP, *BP, *WAG, *B1, *WN, *WEG, *W2, *B2, *W3, *B3;
};

constexpr size_t kSmemFloats = (size_t)kTileRows * kLdB0 + (size_t)kTileRows * kLdB1 + kWbufFloats + 128 /*matt*/ +
                               4 * 12 /*frames*/ + 4 * 24 /*own global points*/ + 16;
constexpr size_t kSmemBytes = kSmemFloats * 4 + 128 * 4 /*j rows*/;

struct Smem {
  float *B0, *B1, *wbuf, *matt, *frame, *pg;
  int* jrow;
  __device__ explicit Smem(float* base) {
    B0 = base;
    B1 = B0 + kTileRows * kLdB0;
    wbuf = B1 + kTileRows * kLdB1;
    matt = wbuf + kWbufFloats;
    frame = matt + 128;
    pg = frame + 48;
    jrow = reinterpret_cast<int*>(pg + 96 + 16);
  }
};

// ---------------------------------------------------------------------------------------- node_pre
__global__ void __launch_bounds__(kThreads, 1)
node_pre_kernel(MpnnCtx cx, PathWeights w, const float* __restrict__ hV, float* __restrict__ A_out,
                float* __restrict__ N_out, float* __restrict__ pglob_out) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int R = cx.S * cx.G;
  const int r0 = blockIdx.x * kTileRows;

  // h_V tile -> B0[:, 0:128]
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    int f = tid + q * kThreads;
    int row = f >> 5, c4 = f & 31;
    float* dst = sm.B0 + row * kLdB0 + c4 * 4;
    if (r0 + row < R) cp_async16(dst, hV + (size_t)(r0 + row) * 128 + c4 * 4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // points weights [128][24] -> wbuf
  for (int f = tid; f < 128 * 24 / 4; f += kThreads) cp_async16(sm.wbuf + f * 4, w.WP + f * 4);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  {  // p_local = W_p h + b_p : thread -> (row, 12 of the 24 outputs)
    int row = tid >> 1, half = tid & 1;
    float p[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) p[q] = w.BP[half * 12 + q];
    const float* hrow = sm.B0 + row * kLdB0;
    for (int k = 0; k < 128; ++k) {
      float a = hrow[k];
      const float* wr = sm.wbuf + k * 24 + half * 12;
#pragma unroll
      for (int q = 0; q < 12; ++q) p[q] = fmaf(a, wr[q], p[q]);
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) sm.B0[row * kLdB0 + 128 + half * 12 + q] = p[q];
  }
  __syncthreads();
  // norms of the local points and the points in the global frame (layers.py:72-77,91)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int it = tid + q * kThreads;
    int row = it >> 3, pt = it & 7;
    const float* p = sm.B0 + row * kLdB0 + 128 + pt * 3;
    float x = p[0], y = p[1], z = p[2];
    sm.B0[row * kLdB0 + 152 + pt] = sqrtf(x * x + y * y + z * z + 1e-8f);
    int r = r0 + row;
    if (r < R) {
      const float* g = cx.geo + (size_t)(r % cx.G) * PP_GEO_STRIDE;
      float* o = pglob_out + (size_t)r * 24 + pt * 3;
      o[0] = g[0] * x + g[1] * y + g[2] * z + g[9];
      o[1] = g[3] * x + g[4] * y + g[5] * z + g[10];
      o[2] = g[6] * x + g[7] * y + g[8] * z + g[11];
    }
  }
  // (the barrier at the top of gemm_tile orders the writes above before the GEMM reads)

  float acc[8][8], bias[8];
  zero_acc(acc);
  gemm_tile(acc, sm.B0, kLdB0, w.WAG, 128, 160, sm.wbuf);
  load_cols(bias, w.B1, tx);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = r0 + tile_row(ty, i);
    if (r < R) {
      float* o = A_out + (size_t)r * 128;
      *reinterpret_cast<float4*>(o + tx * 4) =
          make_float4(acc[i][0] + bias[0], acc[i][1] + bias[1], acc[i][2] + bias[2], acc[i][3] + bias[3]);
      *reinterpret_cast<float4*>(o + 64 + tx * 4) =
          make_float4(acc[i][4] + bias[4], acc[i][5] + bias[5], acc[i][6] + bias[6], acc[i][7] + bias[7]);
    }
  }
  zero_acc(acc);
  gemm_tile(acc, sm.B0, kLdB0, w.WN, 128, 128, sm.wbuf);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = r0 + tile_row(ty, i);
    if (r < R) {
      float* o = N_out + (size_t)r * 128;
      *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
}

// ---------------------------------------------------------------------------------------- edge front
// Loads the tile's h_E rows and pair geometry, runs the first two Linear layers of the message MLP.
// Returns x2 = relu(W2 relu(A_i + N_j + W_eg [h_E | pair]) + b2) in registers.  B0[:, 0:128] still holds h_E.
// Returns false (uniformly) if every residue of the tile is masked; nothing has been computed then.
__device__ __forceinline__ bool edge_front(float (&x2)[8][8], Smem& sm, const MpnnCtx& cx, const PathWeights& w,
                                           const float* hE_in, int he_shared,
                                           const float* __restrict__ A, const float* __restrict__ Nn,
                                           const float* __restrict__ pglob) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int R = cx.S * cx.G, K = cx.K;
  const int rb = blockIdx.x * 4;

  if (tid < 128) {
    int rl = tid >> 5, k = tid & 31;
    int r = rb + rl;
    float m = 0.f;
    int j = 0;
    if (r < R) {
      int s = r / cx.G, g = r - s * cx.G;
      j = r;
      if (k < K) {
        m = cx.matt[(size_t)g * K + k];
        j = s * cx.G + cx.nbr[(size_t)g * K + k];
      }
    }
    sm.matt[tid] = m;
    sm.jrow[tid] = j;
  } else if (tid < 128 + 48) {
    int rl = (tid - 128) / 12, c = (tid - 128) % 12;
    int r = rb + rl;
    sm.frame[rl * 12 + c] = (r < R) ? cx.geo[(size_t)(r % cx.G) * PP_GEO_STRIDE + c] : 0.f;
  }
  for (int f = tid; f < 96; f += kThreads) {
    int rl = f / 24, c = f % 24;
    int r = rb + rl;
    sm.pg[f] = (r < R) ? pglob[(size_t)r * 24 + c] : 0.f;
  }
  // any live edge in this tile?
  int live = 0;
  if (tid < 128) live = sm.matt[tid] != 0.f;  // own write, no barrier needed
  if (!__syncthreads_or(live)) return false;

  // h_E rows -> B0[:, 0:128]
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    int f = tid + q * kThreads;
    int m = f >> 5, c4 = f & 31;
    int rl = m >> 5, k = m & 31, r = rb + rl;
    float* dst = sm.B0 + m * kLdB0 + c4 * 4;
    if (r < R && k < K) {
      size_t row = he_shared ? (size_t)(r % cx.G) : (size_t)r;
      cp_async16(dst, hE_in + (row * K + k) * 128 + c4 * 4);
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  cp_async_commit();
  // pair geometry -> B0[:, 128:168]  (layers.py:93-103)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int it = tid + q * kThreads;
    int m = it >> 3, pt = it & 7, rl = m >> 5;
    const float* pj = pglob + (size_t)sm.jrow[m] * 24 + pt * 3;
    const float* fr = sm.frame + rl * 12;
    const float* pi = sm.pg + rl * 24 + pt * 3;
    float jx = pj[0], jy = pj[1], jz = pj[2];
    float dx = jx - fr[9], dy = jy - fr[10], dz = jz - fr[11];
    float qx = fr[0] * dx + fr[3] * dy + fr[6] * dz;  // R^T d
    float qy = fr[1] * dx + fr[4] * dy + fr[7] * dz;
    float qz = fr[2] * dx + fr[5] * dy + fr[8] * dz;
    float gx = pi[0] - jx, gy = pi[1] - jy, gz = pi[2] - jz;
    float* o = sm.B0 + m * kLdB0 + 128;
    o[pt * 3 + 0] = qx;
    o[pt * 3 + 1] = qy;
    o[pt * 3 + 2] = qz;
    o[24 + pt] = sqrtf(qx * qx + qy * qy + qz * qz + 1e-8f);
    o[32 + pt] = sqrtf(gx * gx + gy * gy + gz * gz + 1e-8f);
  }

  float acc[8][8];
  zero_acc(acc);
  gemm_tile(acc, sm.B0, kLdB0, w.WEG, 128, 168, sm.wbuf);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = tile_row(ty, i);
    int r = min(rb + (m >> 5), R - 1);
    const float* a = A + (size_t)r * 128;
    const float* n = Nn + (size_t)sm.jrow[m] * 128;
    float4 a0 = *reinterpret_cast<const float4*>(a + tx * 4), a1 = *reinterpret_cast<const float4*>(a + 64 + tx * 4);
    float4 n0 = *reinterpret_cast<const float4*>(n + tx * 4), n1 = *reinterpret_cast<const float4*>(n + 64 + tx * 4);
    acc[i][0] = fmaxf(acc[i][0] + a0.x + n0.x, 0.f);
    acc[i][1] = fmaxf(acc[i][1] + a0.y + n0.y, 0.f);
    acc[i][2] = fmaxf(acc[i][2] + a0.z + n0.z, 0.f);
    acc[i][3] = fmaxf(acc[i][3] + a0.w + n0.w, 0.f);
    acc[i][4] = fmaxf(acc[i][4] + a1.x + n1.x, 0.f);
    acc[i][5] = fmaxf(acc[i][5] + a1.y + n1.y, 0.f);
    acc[i][6] = fmaxf(acc[i][6] + a1.z + n1.z, 0.f);
    acc[i][7] = fmaxf(acc[i][7] + a1.w + n1.w, 0.f);
  }
  store_tile_smem(acc, sm.B1, kLdB1, tx, ty);
  zero_acc(x2);
  gemm_tile(x2, sm.B1, kLdB1, w.W2, 128, 128, sm.wbuf);
  float b2[8];
  load_cols(b2, w.B2, tx);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x2[i][j] = fmaxf(x2[i][j] + b2[j], 0.f);
  return true;
}

// ---------------------------------------------------------------------------------------- edge_node
__global__ void __launch_bounds__(kThreads, 1)
edge_node_kernel(MpnnCtx cx, PathWeights w, const float* __restrict__ hE_in, int he_shared,
                 const float* __restrict__ A, const float* __restrict__ Nn, const float* __restrict__ pglob,
                 float* __restrict__ accsum) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int R = cx.S * cx.G;
  const int rb = blockIdx.x * 4;
  float x2[8][8];
  if (!edge_front(x2, sm, cx, w, hE_in, he_shared, A, Nn, pglob)) {
    for (int f = tid; f < 512; f += kThreads) {
      int r = rb + (f >> 7);
      if (r < R) accsum[(size_t)r * 128 + (f & 127)] = 0.f;
    }
    return;
  }
  // masked sum over the 32 edges of each residue: thread-local over its 4 rows, then across the 8 ty groups
  float* red = sm.B1;  // [2 halves][16 ty][128]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float m = sm.matt[tile_row(ty, h * 4 + i)];
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] += (m != 0.f) ? x2[h * 4 + i][j] : 0.f;
    }
    float* o = red + (h * 16 + ty) * 128;
    *reinterpret_cast<float4*>(o + tx * 4) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(p[4], p[5], p[6], p[7]);
  }
  __syncthreads();
  for (int f = tid; f < 512; f += kThreads) {
    int rl = f >> 7, col = f & 127;
    int h = rl >> 1, t0 = (rl & 1) * 8;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += red[(h * 16 + t0 + t) * 128 + col];
    int r = rb + rl;
    if (r < R) accsum[(size_t)r * 128 + col] = s;
  }
}

// ---------------------------------------------------------------------------------------- node_post
struct NodePostWeights {
  const float *W3, *B3, *LN0G, *LN0B, *WIN, *BIN, *WOUT, *BOUT, *LN1G, *LN1B;
};

__global__ void __launch_bounds__(kThreads, 1)
node_post_kernel(MpnnCtx cx, NodePostWeights w, const float* __restrict__ accsum, const float* __restrict__ msum,
                 const float* hV_in, float* hV_out /* may alias hV_in: a tile reads only the rows it writes */) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int R = cx.S * cx.G;
  const int r0 = blockIdx.x * kTileRows;
  const float invK = 1.f / (float)cx.K;

#pragma unroll
  for (int q = 0; q < 16; ++q) {
    int f = tid + q * kThreads;
    int row = f >> 5, c4 = f & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + row < R) v = *reinterpret_cast<const float4*>(accsum + (size_t)(r0 + row) * 128 + c4 * 4);
    v.x *= invK; v.y *= invK; v.z *= invK; v.w *= invK;
    *reinterpret_cast<float4*>(sm.B1 + row * kLdB1 + c4 * 4) = v;
  }
  float e[8][8];
  zero_acc(e);
  gemm_tile(e, sm.B1, kLdB1, w.W3, 128, 128, sm.wbuf);
  {
    float b3[8], g[8], b[8];
    load_cols(b3, w.B3, tx);
    load_cols(g, w.LN0G, tx);
    load_cols(b, w.LN0B, tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int r = min(r0 + tile_row(ty, i), R - 1);
      float ms = msum[r % cx.G];
      const float* h = hV_in + (size_t)r * 128;
      float4 h0 = *reinterpret_cast<const float4*>(h + tx * 4), h1 = *reinterpret_cast<const float4*>(h + 64 + tx * 4);
      e[i][0] += h0.x + b3[0] * ms; e[i][1] += h0.y + b3[1] * ms; e[i][2] += h0.z + b3[2] * ms; e[i][3] += h0.w + b3[3] * ms;
      e[i][4] += h1.x + b3[4] * ms; e[i][5] += h1.y + b3[5] * ms; e[i][6] += h1.z + b3[6] * ms; e[i][7] += h1.w + b3[7] * ms;
    }
    layer_norm_rows(e, g, b);
  }
  store_tile_smem(e, sm.B1, kLdB1, tx, ty);  // gemm_tile ended with a barrier: B1 is free
  float y[8][8];
  ffn_residual_ln(y, sm.B0, sm.B1, sm.wbuf, w.WIN, w.BIN, w.WOUT, w.BOUT, w.LN1G, w.LN1B);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = r0 + tile_row(ty, i);
    if (r < R) {
      bool on = cx.rmask[r % cx.G] != 0.f;
      float* o = hV_out + (size_t)r * 128;
      *reinterpret_cast<float4*>(o + tx * 4) =
          on ? make_float4(y[i][0], y[i][1], y[i][2], y[i][3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(o + 64 + tx * 4) =
          on ? make_float4(y[i][4], y[i][5], y[i][6], y[i][7]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------- edge_edge
struct EdgePostWeights {
  const float *LN2G, *LN2B, *WIN, *BIN, *WOUT, *BOUT, *LN3G, *LN3B;
};

__global__ void __launch_bounds__(kThreads, 1)
edge_edge_kernel(MpnnCtx cx, PathWeights w, EdgePostWeights pw, const float* hE_in, int he_shared,
                 const float* __restrict__ A, const float* __restrict__ Nn, const float* __restrict__ pglob,
                 float* hE_out /* may alias hE_in (he_shared == 0): a tile reads only the rows it writes */) {
  extern __shared__ __align__(16) float smem_raw[];
  Smem sm(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int R = cx.S * cx.G, K = cx.K;
  const int rb = blockIdx.x * 4;
  float x[8][8];
  if (!edge_front(x, sm, cx, w, hE_in, he_shared, A, Nn, pglob)) {
    // every edge of the tile is masked: h_E * mask_attend = 0 (layers.py:145-146)
    for (int f = tid; f < 128 * 32; f += kThreads) {
      int m = f >> 5, c4 = f & 31, r = rb + (m >> 5), k = m & 31;
      if (r < R && k < K)
        *reinterpret_cast<float4*>(hE_out + ((size_t)r * K + k) * 128 + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  store_tile_smem(x, sm.B1, kLdB1, tx, ty);
  zero_acc(x);
  gemm_tile(x, sm.B1, kLdB1, w.W3, 128, 128, sm.wbuf);
  {
    float b3[8], g[8], b[8];
    load_cols(b3, w.B3, tx);
    load_cols(g, pw.LN2G, tx);
    load_cols(b, pw.LN2B, tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = tile_row(ty, i);
      bool on = sm.matt[m] != 0.f;
      const float* h = sm.B0 + m * kLdB0;
      float4 h0 = *reinterpret_cast<const float4*>(h + tx * 4), h1 = *reinterpret_cast<const float4*>(h + 64 + tx * 4);
      x[i][0] = h0.x + (on ? x[i][0] + b3[0] : 0.f); x[i][1] = h0.y + (on ? x[i][1] + b3[1] : 0.f);
      x[i][2] = h0.z + (on ? x[i][2] + b3[2] : 0.f); x[i][3] = h0.w + (on ? x[i][3] + b3[3] : 0.f);
      x[i][4] = h1.x + (on ? x[i][4] + b3[4] : 0.f); x[i][5] = h1.y + (on ? x[i][5] + b3[5] : 0.f);
      x[i][6] = h1.z + (on ? x[i][6] + b3[6] : 0.f); x[i][7] = h1.w + (on ? x[i][7] + b3[7] : 0.f);
    }
    layer_norm_rows(x, g, b);
  }
  store_tile_smem(x, sm.B1, kLdB1, tx, ty);
  float y[8][8];
  ffn_residual_ln(y, sm.B0, sm.B1, sm.wbuf, pw.WIN, pw.BIN, pw.WOUT, pw.BOUT, pw.LN3G, pw.LN3B);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = tile_row(ty, i);
    int r = rb + (m >> 5), k = m & 31;
    if (r < R && k < K) {
      bool on = sm.matt[m] != 0.f;
      float* o = hE_out + ((size_t)r * K + k) * 128;
      *reinterpret_cast<float4*>(o + tx * 4) =
          on ? make_float4(y[i][0], y[i][1], y[i][2], y[i][3]) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(o + 64 + tx * 4) =
          on ? make_float4(y[i][4], y[i][5], y[i][6], y[i][7]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------- host side
static PathWeights path_weights(const float* W, int layer, bool edge) {
  const float* b = W + (long long)layer * wl::kLayerStride;
  PathWeights p;
  if (!edge) {
    p.WP = b + PP_OFF(L0_N_WP); p.BP = b + PP_OFF(L0_N_BP); p.WAG = b + PP_OFF(L0_N_WAG); p.B1 = b + PP_OFF(L0_N_B1);
    p.WN = b + PP_OFF(L0_N_WN); p.WEG = b + PP_OFF(L0_N_WEG); p.W2 = b + PP_OFF(L0_N_W2); p.B2 = b + PP_OFF(L0_N_B2);
    p.W3 = b + PP_OFF(L0_N_W3); p.B3 = b + PP_OFF(L0_N_B3);
  } else {
    p.WP = b + PP_OFF(L0_E_WP); p.BP = b + PP_OFF(L0_E_BP); p.WAG = b + PP_OFF(L0_E_WAG); p.B1 = b + PP_OFF(L0_E_B1);
    p.WN = b + PP_OFF(L0_E_WN); p.WEG = b + PP_OFF(L0_E_WEG); p.W2 = b + PP_OFF(L0_E_W2); p.B2 = b + PP_OFF(L0_E_B2);
    p.W3 = b + PP_OFF(L0_E_W3); p.B3 = b + PP_OFF(L0_E_B3);
  }
  return p;
}

template <typename KernelT>
static int opt_in_smem(KernelT k) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "cudaFuncSetAttribute(smem=%zu): %s", kSmemBytes, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

}  // namespace pp

using namespace pp;

namespace {

struct LayerArgs {
  const float* weights;
  int layer;
  MpnnCtx cx;
  unsigned row_tiles, edge_tiles;
  const float* Lb;
};

int make_args(LayerArgs& a, const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
              const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S) {
  PP_REQUIRE(weights && geo && nbr && mask_attend && residue_mask, "null pointer");
  PP_REQUIRE(layer >= 0 && layer < 3, "layer out of range");
  PP_REQUIRE(G > 0 && S > 0 && K > 0 && K <= PP_KMAX, "bad sizes");
  PP_REQUIRE(S * G < (1ll << 31) / 128, "too many rows");
  a.weights = weights;
  a.layer = (int)layer;
  a.cx = MpnnCtx{geo, nbr, mask_attend, residue_mask, (int)G, (int)K, (int)S};
  const long long R = S * G;
  a.row_tiles = (unsigned)((R + kTileRows - 1) / kTileRows);
  a.edge_tiles = (unsigned)((R + 3) / 4);
  a.Lb = weights + layer * wl::kLayerStride;
  return 0;
}

int launch_node_pre(const LayerArgs& a, bool edge_path, const float* hV, float* wsA, float* wsN, float* wsP,
                    cudaStream_t stream) {
  if (opt_in_smem(node_pre_kernel)) return 1;
  node_pre_kernel<<<a.row_tiles, kThreads, kSmemBytes, stream>>>(a.cx, path_weights(a.weights, a.layer, edge_path), hV,
                                                                wsA, wsN, wsP);
  return 0;
}

int launch_edge_node(const LayerArgs& a, const float* hE_in, int he_shared, const float* wsA, const float* wsN,
                     const float* wsP, float* wsAcc, cudaStream_t stream) {
  if (opt_in_smem(edge_node_kernel)) return 1;
  edge_node_kernel<<<a.edge_tiles, kThreads, kSmemBytes, stream>>>(a.cx, path_weights(a.weights, a.layer, false), hE_in,
                                                                  he_shared, wsA, wsN, wsP, wsAcc);
  return 0;
}

int launch_node_post(const LayerArgs& a, const float* wsAcc, const float* msum, float* hV, cudaStream_t stream) {
  if (opt_in_smem(node_post_kernel)) return 1;
  PathWeights pn = path_weights(a.weights, a.layer, false);
  const float* Lb = a.Lb;
  NodePostWeights np{pn.W3, pn.B3, Lb + PP_OFF(L0_LN0_G), Lb + PP_OFF(L0_LN0_B), Lb + PP_OFF(L0_NF_WIN),
                     Lb + PP_OFF(L0_NF_BIN), Lb + PP_OFF(L0_NF_WOUT), Lb + PP_OFF(L0_NF_BOUT), Lb + PP_OFF(L0_LN1_G),
                     Lb + PP_OFF(L0_LN1_B)};
  node_post_kernel<<<a.row_tiles, kThreads, kSmemBytes, stream>>>(a.cx, np, wsAcc, msum, hV, hV);
  return 0;
}

int launch_edge_edge(const LayerArgs& a, const float* hE_in, int he_shared, const float* wsA, const float* wsN,
                     const float* wsP, float* hE_out, cudaStream_t stream) {
  if (opt_in_smem(edge_edge_kernel)) return 1;
  const float* Lb = a.Lb;
  EdgePostWeights ep{Lb + PP_OFF(L0_LN2_G), Lb + PP_OFF(L0_LN2_B), Lb + PP_OFF(L0_EF_WIN), Lb + PP_OFF(L0_EF_BIN),
                     Lb + PP_OFF(L0_EF_WOUT), Lb + PP_OFF(L0_EF_BOUT), Lb + PP_OFF(L0_LN3_G), Lb + PP_OFF(L0_LN3_B)};
  edge_edge_kernel<<<a.edge_tiles, kThreads, kSmemBytes, stream>>>(a.cx, path_weights(a.weights, a.layer, true), ep,
                                                                  hE_in, he_shared, wsA, wsN, wsP, hE_out);
  return 0;
}

}  // namespace

// One IPMP layer on S*G residue rows (reference layers.py:119-148).
//   hV [S*G][128] is updated in place; hE_in -> hE_out ([rows][K][128]; hE_in has G rows if he_shared else S*G).
//   edge_update = 0 skips the edge half (the reference computes and discards it for the last layer, mpnn.py:53-62).
//   Workspaces (caller-allocated, fp32): wsA, wsN [S*G][128], wsP [S*G][24], wsAcc [S*G][128].
extern "C" int pp_ipmp_layer(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                             const float* mask_attend, const float* msum, const float* residue_mask, int64_t G,
                             int64_t K, int64_t S, float* hV, const float* hE_in, int64_t he_shared, float* hE_out,
                             int64_t edge_update, float* wsA, float* wsN, float* wsP, float* wsAcc,
                             cudaStream_t stream) {
  LayerArgs a;
  if (int rc = make_args(a, weights, layer, geo, nbr, mask_attend, residue_mask, G, K, S)) return rc;
  PP_REQUIRE(msum && hV && hE_in && wsA && wsN && wsP && wsAcc, "null pointer");
  PP_REQUIRE(!edge_update || hE_out, "hE_out required when edge_update is set");
  if (launch_node_pre(a, false, hV, wsA, wsN, wsP, stream)) return 1;
  if (launch_edge_node(a, hE_in, (int)he_shared, wsA, wsN, wsP, wsAcc, stream)) return 1;
  if (launch_node_post(a, wsAcc, msum, hV, stream)) return 1;
  if (edge_update) {
    if (launch_node_pre(a, true, hV, wsA, wsN, wsP, stream)) return 1;
    if (launch_edge_edge(a, hE_in, (int)he_shared, wsA, wsN, wsP, hE_out, stream)) return 1;
  }
  return check_launch("pp_ipmp_layer");
}

// The four kernels of pp_ipmp_layer as separate entry points (same arguments), so that a caller can time or
// re-order them.  path: 0 = node message path, 1 = edge message path.
extern "C" int pp_ipmp_node_pre(const float* weights, int64_t layer, int64_t path, const float* geo, const int32_t* nbr,
                                const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S,
                                const float* hV, float* wsA, float* wsN, float* wsP, cudaStream_t stream) {
  LayerArgs a;
  if (int rc = make_args(a, weights, layer, geo, nbr, mask_attend, residue_mask, G, K, S)) return rc;
  PP_REQUIRE(hV && wsA && wsN && wsP, "null pointer");
  if (launch_node_pre(a, path != 0, hV, wsA, wsN, wsP, stream)) return 1;
  return check_launch("pp_ipmp_node_pre");
}

extern "C" int pp_ipmp_edge_node(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                                 const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S,
                                 const float* hE_in, int64_t he_shared, const float* wsA, const float* wsN,
                                 const float* wsP, float* wsAcc, cudaStream_t stream) {
  LayerArgs a;
  if (int rc = make_args(a, weights, layer, geo, nbr, mask_attend, residue_mask, G, K, S)) return rc;
  PP_REQUIRE(hE_in && wsA && wsN && wsP && wsAcc, "null pointer");
  if (launch_edge_node(a, hE_in, (int)he_shared, wsA, wsN, wsP, wsAcc, stream)) return 1;
  return check_launch("pp_ipmp_edge_node");
}

extern "C" int pp_ipmp_node_post(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                                 const float* mask_attend, const float* msum, const float* residue_mask, int64_t G,
                                 int64_t K, int64_t S, const float* wsAcc, float* hV, cudaStream_t stream) {
  LayerArgs a;
  if (int rc = make_args(a, weights, layer, geo, nbr, mask_attend, residue_mask, G, K, S)) return rc;
  PP_REQUIRE(msum && wsAcc && hV, "null pointer");
  if (launch_node_post(a, wsAcc, msum, hV, stream)) return 1;
  return check_launch("pp_ipmp_node_post");
}

extern "C" int pp_ipmp_edge_edge(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                                 const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S,
                                 const float* hE_in, int64_t he_shared, const float* wsA, const float* wsN,
                                 const float* wsP, float* hE_out, cudaStream_t stream) {
  LayerArgs a;
  if (int rc = make_args(a, weights, layer, geo, nbr, mask_attend, residue_mask, G, K, S)) return rc;
  PP_REQUIRE(hE_in && wsA && wsN && wsP && hE_out, "null pointer");
  if (launch_edge_edge(a, hE_in, (int)he_shared, wsA, wsN, wsP, hE_out, stream)) return 1;
  return check_launch("pp_ipmp_edge_edge");
}
