// fp32 tile GEMM building blocks for the message-passing kernels (CUDA-core FFMA path, exact fp32).
//
// A CTA of 256 threads owns a tile of 128 rows (4 residues x 32 edges, or 128 residue rows) and all 128 output
// columns.  Activations live in shared memory, row-major with a padded leading dimension (lda % 32 == 4, so the
// two row groups of a warp fall on disjoint banks); weights stream from global/L2 in K-chunks of 32 rows through
// a cp.async double buffer.  Thread (tx = tid % 16, ty = tid / 16) accumulates an 8 x 8 register tile:
//   rows  ty*4 + {0..3} and 64 + ty*4 + {0..3}      cols  tx*4 + {0..3} and 64 + tx*4 + {0..3}
// so that every shared-memory read is a conflict-free 128-bit access.
#pragma once
#ifndef PP_USE_FFMA2
#define PP_USE_FFMA2 1
#endif
#include "common.cuh"

namespace pp {

constexpr int kTileRows = 128;
constexpr int kThreads = 256;
constexpr int kKC = 32;                    // weight rows per pipeline stage
constexpr int kLdB0 = 172;                 // [h_E 128 | geo 40] + 4 pad
constexpr int kLdB1 = 132;                 // 128 + 4 pad
constexpr int kWbufFloats = 2 * kKC * 128;

__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)); }
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)); }

__device__ __forceinline__ void zero_acc(float (&acc)[8][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ void load_w_chunk(float* dst, const float* Wg, int ldw, int k0, int KD, int tid) {
  // kKC rows x 128 floats = 1024 float4, 4 per thread; rows past KD are skipped (never read by the consumer)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int f = tid + q * kThreads;
    int row = f >> 5, c4 = f & 31;
    if (k0 + row < KD) cp_async16(dst + row * 128 + c4 * 4, Wg + (size_t)(k0 + row) * ldw + c4 * 4);
  }
}

// acc += As[128 x KD] * Wg[KD x 128]   (Wg row stride ldw; KD % 4 == 0).  All 256 threads must call.
// Contains the barriers that order (a) earlier writes to As by other threads before the first read and
// (b) all reads of As / wbuf before the caller overwrites them afterwards.
__device__ __forceinline__ void gemm_tile(float (&acc)[8][8], const float* As, int lda, const float* __restrict__ Wg,
                                          int ldw, int KD, float* wbuf) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nchunk = (KD + kKC - 1) / kKC;
  load_w_chunk(wbuf, Wg, ldw, 0, KD, tid);
  cp_async_commit();
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) load_w_chunk(wbuf + ((c + 1) & 1) * kKC * 128, Wg, ldw, (c + 1) * kKC, KD, tid);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* wb = wbuf + (c & 1) * kKC * 128;
    const int k0 = c * kKC;
    const int kmax = min(kKC, KD - k0);
    for (int kk = 0; kk < kmax; kk += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(As + tile_row(ty, i) * lda + k0 + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 b0 = *reinterpret_cast<const float4*>(wb + (kk + q) * 128 + tx * 4);
        float4 b1 = *reinterpret_cast<const float4*>(wb + (kk + q) * 128 + 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = (q == 0) ? a[i].x : (q == 1) ? a[i].y : (q == 2) ? a[i].z : a[i].w;
#if PP_USE_FFMA2
          // packed fp32 FMA (sm_100): two IEEE fmas per instruction, same results, half the issue slots
          const float2 a2 = make_float2(av, av);
          float2 r;
          r = __ffma2_rn(a2, make_float2(b0.x, b0.y), make_float2(acc[i][0], acc[i][1])); acc[i][0] = r.x; acc[i][1] = r.y;
          r = __ffma2_rn(a2, make_float2(b0.z, b0.w), make_float2(acc[i][2], acc[i][3])); acc[i][2] = r.x; acc[i][3] = r.y;
          r = __ffma2_rn(a2, make_float2(b1.x, b1.y), make_float2(acc[i][4], acc[i][5])); acc[i][4] = r.x; acc[i][5] = r.y;
          r = __ffma2_rn(a2, make_float2(b1.z, b1.w), make_float2(acc[i][6], acc[i][7])); acc[i][6] = r.x; acc[i][7] = r.y;
#else
          acc[i][0] = fmaf(av, b0.x, acc[i][0]);
          acc[i][1] = fmaf(av, b0.y, acc[i][1]);
          acc[i][2] = fmaf(av, b0.z, acc[i][2]);
          acc[i][3] = fmaf(av, b0.w, acc[i][3]);
          acc[i][4] = fmaf(av, b1.x, acc[i][4]);
          acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]);
          acc[i][7] = fmaf(av, b1.w, acc[i][7]);
#endif
        }
      }
    }
    __syncthreads();
  }
}

// per-column vector (bias, LayerNorm gain ...) for this thread's 8 columns
__device__ __forceinline__ void load_cols(float (&v)[8], const float* __restrict__ g, int tx) {
  float4 a = *reinterpret_cast<const float4*>(g + tx * 4);
  float4 b = *reinterpret_cast<const float4*>(g + 64 + tx * 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void store_tile_smem(const float (&acc)[8][8], float* Bs, int ldb, int tx, int ty) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* p = Bs + tile_row(ty, i) * ldb;
    *reinterpret_cast<float4*>(p + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(p + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// sum over the 16 threads (tx = 0..15, consecutive lanes) that share a row
__device__ __forceinline__ float row_sum16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// in-place LayerNorm(128, eps 1e-5) of every row of the register tile
__device__ __forceinline__ void layer_norm_rows(float (&x)[8][8], const float (&g)[8], const float (&b)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[i][j];
    float mean = row_sum16(s) * (1.f / 128.f);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { float d = x[i][j] - mean; v += d * d; }
    float rstd = rsqrtf(row_sum16(v) * (1.f / 128.f) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[i][j] = (x[i][j] - mean) * rstd * g[j] + b[j];
  }
}

__device__ __forceinline__ void load_tile_smem(float (&acc)[8][8], const float* Bs, int ldb, int tx, int ty) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float* p = Bs + tile_row(ty, i) * ldb;
    float4 a = *reinterpret_cast<const float4*>(p + tx * 4);
    float4 b = *reinterpret_cast<const float4*>(p + 64 + tx * 4);
    acc[i][0] = a.x; acc[i][1] = a.y; acc[i][2] = a.z; acc[i][3] = a.w;
    acc[i][4] = b.x; acc[i][5] = b.y; acc[i][6] = b.z; acc[i][7] = b.w;
  }
}

// Position-wise feed-forward with residual and LayerNorm, shared by the node and the edge update
// (reference layers.py:129-130 and :143-144):   y = LN(e + W_out relu(W_in e + b_in) + b_out)
// On entry e is in B1 (row-major, ld kLdB1, visible to all threads after the barrier inside gemm_tile);
// B0's first 128 columns are scratch.  Result y in registers.
__device__ __forceinline__ void ffn_residual_ln(float (&y)[8][8], float* B0, float* B1, float* wbuf,
                                                const float* __restrict__ Win, const float* __restrict__ bin,
                                                const float* __restrict__ Wout, const float* __restrict__ bout,
                                                const float* __restrict__ lng, const float* __restrict__ lnb) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  zero_acc(y);
  for (int c = 0; c < 4; ++c) {
    float h[8][8];
    zero_acc(h);
    gemm_tile(h, B1, kLdB1, Win + c * 128, 512, 128, wbuf);
    float bi[8];
    load_cols(bi, bin + c * 128, tx);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) h[i][j] = fmaxf(h[i][j] + bi[j], 0.f);
    store_tile_smem(h, B0, kLdB0, tx, ty);
    gemm_tile(y, B0, kLdB0, Wout + (size_t)c * 128 * 128, 128, 128, wbuf);
  }
  float bo[8], g[8], b[8];
  load_cols(bo, bout, tx);
  load_cols(g, lng, tx);
  load_cols(b, lnb, tx);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float* p = B1 + tile_row(ty, i) * kLdB1;
    float4 e0 = *reinterpret_cast<const float4*>(p + tx * 4);
    float4 e1 = *reinterpret_cast<const float4*>(p + 64 + tx * 4);
    y[i][0] += e0.x + bo[0]; y[i][1] += e0.y + bo[1]; y[i][2] += e0.z + bo[2]; y[i][3] += e0.w + bo[3];
    y[i][4] += e1.x + bo[4]; y[i][5] += e1.y + bo[5]; y[i][6] += e1.z + bo[6]; y[i][7] += e1.w + bo[7];
  }
  layer_norm_rows(y, g, b);
}

}  // namespace pp
