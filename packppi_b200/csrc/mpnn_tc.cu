// Per-edge message MLPs of the IPMP layers on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same mathematics as edge_node_kernel / edge_edge_kernel in mpnn.cu (reference layers.py:119-148), same operand
// split (only the 168-wide [h_E | pair geometry] part of the first Linear is per edge).  One CTA = one tile of 128
// edges (4 residues x 32 neighbours) = the M dimension of a 128 x 128 UMMA; the whole GEMM chain of the tile stays on
// chip:
//   workers    kGroups (4) groups of 128 threads = 16 warps: thread (group, m) owns edge row m = TMEM lane m and
//              one of the four 32-column chunks (kGroups = 2: two chunks each, 8 warps).  They build the first A operand, then act as the epilogue of every GEMM:
//              tcgen05.ld 32 columns -> bias / ReLU / LayerNorm -> next A operand, written into a shared-memory ring
//              (or into TMEM for the FFN input); gathered rows are fetched before the wait on the accumulator
//   warp 8     MMA issuer (one elected lane): tcgen05.mma kind::f16 (K = 16 per instruction), accumulators ping-pong
//              between two 128-column TMEM regions, completion signalled with tcgen05.commit on mbarriers
//   warp 9     weight loader: cp.async.bulk of pre-packed operand images (K-major core-matrix layout, consumption
//              order, see pack_tc_stream in packppi_b200/weights.py) into a ring; with CLUSTER > 1 every CTA fetches
//              1/CLUSTER of each image and multicasts it to the whole cluster, dividing the L2 -> SM weight traffic
//   warp 10    h_E tile loader: one TMA tensor copy (cp.async.bulk.tensor, 128-byte
//              swizzle) per residue and 32-column chunk brings the tile after the next one into a 64 KB staging
//              buffer, from which each worker reads its own row without bank conflicts; the edge update writes its
//              result rows back through the same buffer with TMA stores (a row-per-thread global access would cost
//              32 L1 wavefronts per instruction, the staged path 4)
// A chunks are consumed by the MMA warp as soon as they are written, so the epilogue of GEMM n overlaps the MMAs of
// GEMM n+1.  The FFN input e = LayerNorm(...) is kept in TMEM twice: as packed fp16 (hi, lo) pairs, the TMEM A
// operand of the four 128-wide slices of the 128 -> 512 Linear, and in fp32 for the residual.
//
// Precision (PASSES): 3 = split fp16, x ~= hi + lo with both halves rounded to nearest fp16: hi*hi + hi*lo + lo*hi
// keeps 22 mantissa bits per product with fp32 accumulation (parity mode) and runs on kind::f16, twice the rate of
// kind::tf32; 1 = hi only (11 bits, the precision of TF32; fast mode, looser tolerance).  Weight images are
// pre-scaled by a power of two per matrix (weights.py: pack_tc_stream) so that their lo halves stay in the normal
// fp16 range; the epilogues multiply the accumulator by the inverse scale (exact).
#include <type_traits>

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "umma.cuh"
#include "weights_layout.h"

namespace pp {
namespace tc {

using namespace umma;

constexpr int kRows = 128;
constexpr int kKC = 32;
#ifndef PP_TC_SA
#define PP_TC_SA 4
#endif
#ifndef PP_TC_SB
#define PP_TC_SB 4
#endif
constexpr int kSA = PP_TC_SA;   // A ring: slots of 32 k-columns (fp16 hi + lo images, 16 KB)
constexpr int kSB = PP_TC_SB;   // B ring: slots of one weight chunk (<= 32 k-columns, hi + lo images, 16 KB)
constexpr uint32_t kImgBytes = kRows * kKC * 2;  // one fp16 operand image (hi or lo) of a 32-column chunk: 8 KB
constexpr uint32_t kSlotBytes = 2 * kImgBytes;   // hi + lo
#ifndef PP_TC_GROUPS
#define PP_TC_GROUPS 2
#endif
constexpr int kGroups = PP_TC_GROUPS;      // worker threads per row: 2 (each owns two 32-column chunks) or 4 (one chunk)
constexpr int kCPT = 4 / kGroups;          // chunks per thread: thread (grp, m) owns chunks grp, grp + kGroups, ...
constexpr int kWorkers = kGroups * 128;
constexpr int kWarpMMA = kWorkers / 32, kWarpWeights = kWarpMMA + 1, kWarpTiles = kWarpMMA + 2;
constexpr int kThreadsTC = kWorkers + 96;  // worker warps, MMA warp, weight loader, tile loader
constexpr uint32_t kLbo = kRows * 16, kSbo = 128;
constexpr uint32_t kIdesc = idesc_f16(128, 128);
constexpr int kPairKC = 16;                      // last G1 chunk: 8 pair distances padded to one K = 16 instruction

constexpr int kChunksNode = 10;                  // G1 (5x32 + 16), G2 (4x32)
constexpr int kChunksEdge = 6 + 4 + 4 + 4 * 8;   // + G3, 4 x (FFN-in slice, FFN-out slice)
// operand images (fp16 hi + lo) of one layer / path, followed by 8 floats: 1 / scale of G1, G2, G3, FFN-in, FFN-out
constexpr long long kImageFloats = 2LL * 128 * (176 + 128 + 128 + 4 * 256) * 2 / 4;
constexpr long long kStreamFloats = kImageFloats + 8;

constexpr size_t kBarBytes = (2 * kSA + 2 * kSB + 2 + 2 + 2) * 8;
constexpr uint32_t kStageBytes = kRows * 128 * 4;  // one tile of h_E rows: 16 boxes (4 chunks x 4 residues) of 4 KB
// per-column parameters staged in shared memory (floats): b2, b3, LN2 gain/bias, FFN b_in (512), b_out, LN3 gain/bias
constexpr int kP_B2 = 0, kP_B3 = 128, kP_LN2G = 256, kP_LN2B = 384, kP_BIN = 512, kP_BOUT = 1024, kP_LN3G = 1152,
              kP_LN3B = 1280, kParamFloats = 1408;
constexpr int kRedFloats = 4 * kGroups * 128;  // four row reductions x groups
constexpr size_t kSmemTC = kStageBytes + ((size_t)kSA + kSB) * kSlotBytes + kBarBytes + 16 +
                           (kParamFloats + kRedFloats) * 4;

struct Args {
  const float* geo; const int* nbr; const float* matt;
  int G, K, S;
  const float* wstream;   // operand images of this layer / path, then the inverse scales
  const float *B2, *B3, *LNG, *LNB, *BIN, *BOUT, *LN3G, *LN3B;
  const float* hE_in; int he_shared;
  const float *A, *Nn, *pglob;
  const float* msum;                 // mean attention mask [G]: 0 marks a padding residue
  const int* live_list;              // optional: compacted ids of the live tiles [n_live (+ padding)] ...
  const int* n_live;                 // ... and their number (device scalar)
  float* out;             // node path: accsum [R][128]; edge path: hE_out [R][K][128]
  int* overflow;              // optional: set to 1 if an activation left the fp16 range (see umma.cuh)
  unsigned long long* trace;  // optional: clock64 stamps of one tile of CTA 0 (pp_set_tc_trace), else null
  int trace_it;               // which tile of the CTA (0 = first: the stamps include the prologue)
};

struct Ring {
  int idx; uint32_t phase;
  __device__ __forceinline__ void next(int n) { if (++idx == n) { idx = 0; phase ^= 1; } }
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// TMA tensor copies of one box {32 floats, K rows, 1 residue} of a [rows][K][128] fp32 tensor (128-byte swizzle)
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* tm, int c0, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(0), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, int c0, int c2, const void* src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(c0), "r"(0), "r"(c2), "r"(smem_u32(src))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// write this thread's row of a chunk (kc columns, v[0..kc)) into an A ring slot in the UMMA core-matrix layout:
// 8 fp16 values = one 16-byte core-matrix row; the 32 lanes of a warp write 512 contiguous bytes (conflict-free)
template <int PASSES>
__device__ __forceinline__ void put_chunk(uint8_t* slot, int m, const float* v, int kc, float& amax) {
  const int base = (m >> 3) * 128 + (m & 7) * 16;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (u * 8 < kc) {
      uint4 h, l;
      split_f16x2(v[u * 8 + 0], v[u * 8 + 1], h.x, l.x, amax); split_f16x2(v[u * 8 + 2], v[u * 8 + 3], h.y, l.y, amax);
      split_f16x2(v[u * 8 + 4], v[u * 8 + 5], h.z, l.z, amax); split_f16x2(v[u * 8 + 6], v[u * 8 + 7], h.w, l.w, amax);
      *reinterpret_cast<uint4*>(slot + u * kLbo + base) = h;
      if (PASSES == 3) *reinterpret_cast<uint4*>(slot + kImgBytes + u * kLbo + base) = l;
    }
  }
}

// 32 consecutive per-column parameters (bias, LayerNorm gain ...): the address is warp-uniform, 8 broadcast loads
__device__ __forceinline__ void ld32(const float* __restrict__ p, float (&d)[32]) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    float4 x = *reinterpret_cast<const float4*>(p + u * 4);
    d[u * 4] = x.x; d[u * 4 + 1] = x.y; d[u * 4 + 2] = x.z; d[u * 4 + 3] = x.w;
  }
}

// MODE 0: node message (G1, G2, masked sum over K)      tile = 4 residues x 32 edges
// MODE 1: edge update  (G1, G2, G3, LN, FFN, LN)        tile = 4 residues x 32 edges
// (the per-residue node update W3 / LN0 / FFN / LN1 lives in node_post_tc.cu: it needs promoted accumulation)
template <int MODE, int PASSES, int CLUSTER>
__global__ void __launch_bounds__(kThreadsTC, 1)
edge_tc_kernel(const Args a, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out) {
  constexpr bool EDGE = MODE != 0;   // runs the G3 / LayerNorm / FFN part
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage = smem;  // h_E rows of one tile: box (chunk c, residue rl) at ((c * 4 + rl) << 12), row k at k << 7
  uint8_t* Aring = stage + kStageBytes;
  uint8_t* Bring = Aring + kSA * kSlotBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Bring + kSB * kSlotBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kSA;
  uint64_t* b_full = a_empty + kSA;
  uint64_t* b_empty = b_full + kSB;
  uint64_t* acc_full = b_empty + kSB;  // [2]
  uint64_t* wk_done = acc_full + 2;     // FFN hand-offs within a tile (e in TMEM, slice j read out)
  uint64_t* tile_done = wk_done + 1;    // the workers have read out everything a tile left in TMEM
  uint64_t* stage_full = tile_done + 1;  // TMA loads of a tile have landed in the staging buffer
  uint64_t* stage_free = stage_full + 1; // the workers are done with the staging buffer: rows read / result rows written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stage_free + 1);
  float* prm = reinterpret_cast<float*>(stage_free + 3);
  float* red = prm + kParamFloats;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int R = a.S * a.G, K = a.K;
  const int ntiles = (R + 3) / 4;
  // persistent CTAs: every CTA runs the same number of iterations (a cluster shares one weight stream in lockstep);
  // iterations past the last tile work on fully masked rows
  const int niter = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1);
  // Tiles whose residues are all padding (mean attention mask 0) are skipped: every role walks the same sequence of
  // live tiles.  In a cluster of 2 the CTAs share one weight stream in lockstep, so the unit of skipping is the PAIR of
  // tiles (2t, 2t + 1) the cluster works on - both CTAs take the same decision (padding is contiguous, so pairs are
  // almost always uniformly live or dead).  The output rows of skipped tiles are not written: the caller keeps them
  // zero (nothing but the skipped tiles themselves ever reads or writes them).
#ifndef PP_TC_SKIP
#define PP_TC_SKIP 1
#endif
  constexpr bool SKIP = PP_TC_SKIP && CLUSTER <= 2;
  const int tstep = (int)gridDim.x;  // a multiple of CLUSTER
  // tiles of this CTA: blockIdx.x, + tstep, ... < tend; with a cluster the bound is rounded up so that both CTAs of a
  // pair run the same number of iterations (a tile past the end works on fully masked rows)
  // With a compacted list of the live tiles (a.live_list, built once per graph and sample count) the loop variable
  // `tile` of every role is a POSITION in that list and tile_id() maps it to the tile: CTA b takes positions b,
  // b + grid, ... so the live tiles are spread evenly (+-1) over the CTAs.  Walking tile = b, b + grid, ... and testing
  // liveness instead (the fallback without a list) leaves up to 18 % more tiles on the fullest CTA than on the average
  // one when a ragged micro-batch is heavily padded, because runs of live tiles alias with the grid size.
  const int* const ll = a.live_list;
  const int n_live = ll ? *a.n_live : 0;
  const int tend = ll ? (n_live + CLUSTER - 1) / CLUSTER * CLUSTER
                      : (SKIP ? (ntiles + CLUSTER - 1) / CLUSTER * CLUSTER : niter * tstep);
  auto tile_id = [&](int t) { return ll ? (t < n_live ? ll[t] : ntiles) : t; };  // past the end: fully masked rows
  auto live = [&](int tile) {
    bool any = false;
    const int base = (CLUSTER == 2) ? (tile & ~1) * 4 : tile * 4;
#pragma unroll
    for (int i = 0; i < 4 * (CLUSTER == 2 ? 2 : 1); ++i) {
      const int r = base + i;
      if (r < R) any |= a.msum[r % a.G] != 0.f;
    }
    return any;
  };
  auto next_tile = [&](int tile) {
    if (SKIP && !ll) while (tile < tend && !live(tile)) tile += tstep;
    return tile;
  };

  if (tid == 0) {
    for (int i = 0; i < kSA; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kSB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], CLUSTER); }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(wk_done, kWorkers);
    mbar_init(tile_done, kWorkers);
    mbar_init(stage_full, 1);
    mbar_init(stage_free, kWorkers);
    mbar_fence_init();
  }
  {  // stage the per-column parameters
    const float* src[8] = {a.B2, a.B3, a.LNG, a.LNB, a.BIN, a.BOUT, a.LN3G, a.LN3B};
    const int off[9] = {kP_B2, kP_B3, kP_LN2G, kP_LN2B, kP_BIN, kP_BOUT, kP_LN3G, kP_LN3B, kParamFloats};
    for (int t = 0; t < 8; ++t)
      if (EDGE || t == 0)
        for (int i = tid; i < off[t + 1] - off[t]; i += kThreadsTC) prm[off[t] + i] = src[t][i];
  }
  if (warp == kWarpMMA) tmem_alloc<512>(tmem_slot);
  fence_before_sync();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // every CTA's barriers exist before any remote arrive / multicast
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: two accumulators and two operand regions; per tile one region holds the raw h_E row and later e
  // in fp32 (residual), the other e as packed fp16 (hi: 64 columns, lo: 64 columns); the roles swap every tile
  const uint32_t ACC0 = tmem, ACC1 = tmem + 128, R0 = tmem + 256, R1 = tmem + 384;

  if (warp == kWarpWeights) {
    // ------------------------------------------------------------------ weight loader
    if (lane == 0) {
      Ring rb_{0, 1};
      const uint32_t crank = (CLUSTER > 1) ? cluster_rank() : 0;
      for (int tile = next_tile((int)blockIdx.x); tile < tend; tile = next_tile(tile + tstep)) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wstream);
        constexpr int NCHUNK = EDGE ? kChunksEdge : kChunksNode;
        for (int i = 0; i < NCHUNK; ++i) {
          const int kc = (i == 5) ? kPairKC : kKC;
          const uint32_t img = (uint32_t)kRows * kc * 2;          // hi image; the lo image follows it in the stream
          const uint32_t bytes = img * (PASSES == 3 ? 2 : 1);
          mbar_wait(&b_empty[rb_.idx], rb_.phase);
          mbar_arrive_expect_tx(&b_full[rb_.idx], bytes);
          uint8_t* dst = Bring + rb_.idx * kSlotBytes;
          if (CLUSTER == 1) {
            bulk_g2s(dst, src, bytes, &b_full[rb_.idx]);
          } else {
            const uint32_t piece = bytes / CLUSTER, po = crank * piece;
            bulk_g2s_mc(dst + po, src + po, piece, &b_full[rb_.idx], kMask);
          }
          rb_.next(kSB);
          src += 2 * img;
        }
      }
    }
  } else if (warp == kWarpTiles) {
    if (lane == 0) {
      // ---------------------------------------------------------------- h_E tile loader (TMA)
      // This thread owns the staging buffer.  Node message path: tile t+1 is fetched as soon as the workers have
      // read tile t.  Edge update: the workers read tile it+1, then write the result rows of tile it into the buffer;
      // those are stored from here and tile it+2 is fetched once the stores have read the buffer.
      const uint32_t box_bytes = (uint32_t)min(K, 32) * 128;
      uint32_t free_phase = 0;
      auto load_tile = [&](int tile) {
        const int r0 = tile_id(tile) * 4;
        const int nv = max(0, min(4, R - r0));  // residues of the tile that exist
        mbar_arrive_expect_tx(stage_full, (uint32_t)nv * 4 * box_bytes);
        for (int i = 0; i < nv; ++i) {
          const int row = a.he_shared ? (r0 + i) % a.G : r0 + i;
          for (int c = 0; c < 4; ++c) tma_load_box(stage + ((c * 4 + i) << 12), &tm_in, c * 32, row, stage_full);
        }
      };
      auto wait_workers = [&]() { mbar_wait(stage_free, free_phase); free_phase ^= 1; };
      int cur = next_tile((int)blockIdx.x);
      if (cur < tend) {
        load_tile(cur);
        if (MODE == 0) {
          for (int t = next_tile(cur + tstep); t < tend; t = next_tile(t + tstep)) { wait_workers(); load_tile(t); }
        } else {
          wait_workers();  // the first tile has been read
          int nxt = next_tile(cur + tstep);
          if (nxt < tend) load_tile(nxt);
          while (cur < tend) {
            wait_workers();  // result rows of tile `cur` are in the buffer (and tile `nxt` has been read out of it)
            const int r0 = tile_id(cur) * 4;
            const int nv = max(0, min(4, R - r0));  // rows k >= K are clipped by the tensor map
            for (int i = 0; i < nv; ++i)
              for (int c = 0; c < 4; ++c) tma_store_box(&tm_out, c * 32, r0 + i, stage + ((c * 4 + i) << 12));
            bulk_commit();
            const int after = nxt < tend ? next_tile(nxt + tstep) : tend;
#ifdef PP_TC_ALWAYS_WAIT_READ
            bulk_wait_read();
            if (after < tend) load_tile(after);
#else
            if (after < tend) { bulk_wait_read(); load_tile(after); }
#endif
            cur = nxt;
            nxt = after;
          }
          bulk_wait_all();  // the result rows are in global memory before the CTA retires
        }
      }
    }
  } else if (warp == kWarpMMA) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the control flow (waits included) so that it stays converged and the descriptor arithmetic
    // lives in uniform registers; one elected lane issues the MMAs and commits of a chunk.
    {
      Ring ra{0, 0}, rbq{0, 0};
      uint32_t wk_phase = 0, td_phase = 0;
      // ss: A from the shared-memory ring; otherwise a_tm = TMEM address of the packed hi half of the chunk (lo: + 64)
      auto chunk_kc = [&](auto KC, bool ss, uint32_t acc, uint32_t a_tm, bool fresh) {
        // pass 0: hi * hi, pass 1: hi * lo, pass 2: lo * hi.  Descriptors: low word of the slot start + a constant
        constexpr int kc = decltype(KC)::value;
        if (ss) mbar_wait(&a_full[ra.idx], ra.phase);
        const uint32_t a_lo = desc_lo(smem_u32(Aring + ra.idx * kSlotBytes), kLbo);
        mbar_wait(&b_full[rbq.idx], rbq.phase);
        fence_after_sync();
        const uint32_t b_lo = desc_lo(smem_u32(Bring + rbq.idx * kSlotBytes), kLbo);
        constexpr uint32_t kHi = desc_hi(kSbo);
        if (elect_one_sync()) {
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
#pragma unroll
            for (int kk = 0; kk < kc; kk += 16) {
              const uint32_t ad = (((p == 2) ? kImgBytes : 0u) + (kk / 8) * kLbo) >> 4;
              const uint32_t bd = (((p == 1) ? (uint32_t)kRows * kc * 2 : 0u) + (kk / 8) * kLbo) >> 4;
              const uint32_t accum = (fresh && p == 0 && kk == 0) ? 0u : 1u;
              if (ss) mma_f16_ss2(acc, a_lo + ad, b_lo + bd, kHi, kIdesc, accum);
              else mma_f16_ts2(acc, a_tm + ((p == 2) ? 64 : 0) + kk / 2, b_lo + bd, kHi, kIdesc, accum);
            }
          }
          if (CLUSTER == 1) mma_commit(&b_empty[rbq.idx]); else mma_commit_mc(&b_empty[rbq.idx], kMask);
          if (ss) mma_commit(&a_empty[ra.idx]);
        }
        __syncwarp();
        rbq.next(kSB);
        if (ss) ra.next(kSA);
      };
      auto commit_acc = [&](int bar) {
        if (elect_one_sync()) mma_commit(&acc_full[bar]);
        __syncwarp();
      };
      auto chunk = [&](bool ss, uint32_t acc, uint32_t a_tm, int kc, bool fresh) {
        if (kc == kKC) chunk_kc(std::integral_constant<int, kKC>{}, ss, acc, a_tm, fresh);
        else chunk_kc(std::integral_constant<int, kPairKC>{}, ss, acc, a_tm, fresh);
      };
      auto wait_workers = [&]() { mbar_wait(wk_done, wk_phase); wk_phase ^= 1; fence_after_sync(); };
      // first GEMM of a tile: G1 = [h_E | pair] (176 wide) of the message MLP
      auto head = [&](uint32_t acc, int bar) {
        for (int c = 0; c < 6; ++c) chunk(true, acc, 0, c == 5 ? kPairKC : kKC, c == 0);
        commit_acc(bar);
      };
      // Tiles are software-pipelined: the head GEMM of tile it+1 is issued before the workers run the last epilogue of
      // tile it.  In the EDGE modes the roles of the two accumulators (X: head, G3, FFN-out; Y: G2, FFN-in) and of
      // the two TMEM operand regions swap with the parity of the tile so that nothing live is overwritten.
      int tile = next_tile((int)blockIdx.x);
      if (tile < tend) head(ACC0, 0);
      for (int it = 0; tile < tend; ++it) {
        const int nxt_tile = next_tile(tile + tstep);
        const bool more = nxt_tile < tend;
        tile = nxt_tile;
        const int par = EDGE ? (it & 1) : 0;
        const uint32_t X = par ? ACC1 : ACC0, Y = par ? ACC0 : ACC1;
        const int xb = par, yb = par ^ 1;
        const uint32_t PK = par ? R0 : R1;  // packed fp16 (hi | lo) image of e
        if (it > 0) {  // the previous tile has been read out completely: its X (this tile's Y) may be overwritten
          mbar_wait(tile_done, td_phase); td_phase ^= 1; fence_after_sync();
        }
        for (int c = 0; c < 4; ++c) chunk(true, Y, 0, kKC, c == 0);  // G2
        commit_acc(yb);
        if (!EDGE) {
          // G2's operand chunks exist only once every worker has drained ACC0, so the next G1 may follow directly
          if (more) head(ACC0, 0);
          continue;
        }
        for (int c = 0; c < 4; ++c) chunk(true, X, 0, kKC, c == 0);  // G3
        commit_acc(xb);
        // FFN, software-pipelined by one slice: while the workers turn slice j into the A operand of FFN-out j,
        // the tensor pipe runs FFN-out j-1 and FFN-in j+1
        wait_workers();  // e is in TMEM, X / Y are drained
        for (int c = 0; c < 4; ++c) chunk(false, Y, PK + c * (kKC / 2), kKC, c == 0);  // FFN-in slice 0: A = e (TMEM)
        commit_acc(yb);
        for (int j = 0; j < 4; ++j) {
          if (j + 1 < 4) {
            wait_workers();  // slice j has been read out of Y
            for (int c = 0; c < 4; ++c) chunk(false, Y, PK + c * (kKC / 2), kKC, c == 0);  // FFN-in slice j+1
            commit_acc(yb);
          }
          for (int c = 0; c < 4; ++c) chunk(true, X, 0, kKC, j == 0 && c == 0);  // FFN-out slice j accumulates
        }
        commit_acc(xb);
        if (more) {
          wait_workers();  // slice 3 has been read out of Y, which becomes the X of the next tile
          head(Y, yb);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ row workers
    // kGroups groups of 128 threads; thread (grp, m) owns edge row m (= TMEM lane m) and the 32-column chunks
    // grp, grp + kGroups, ... of every 128-wide activation, so the A chunks of an operand are produced side by side.
    const int grp = tid >> 7, m = tid & 127, rl = m >> 5, k = lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t accph[2] = {0, 0};
    const float* wsc = a.wstream + kImageFloats;  // 1 / scale of the weight images
    const float sG1 = wsc[0], sG2 = wsc[1], sG3 = wsc[2], sFI = wsc[3], sFO = wsc[4];
    constexpr int kChunksTile = EDGE ? 30 : 10;  // A chunks per tile
    const bool tracer = a.trace && blockIdx.x == 0 && tid == 0;
    int stamp_i = 0;
    auto stamp = [&](bool first_tile) {
      if (tracer && first_tile) a.trace[stamp_i++] = clock64();
    };

    struct RowCtx {  // where this thread's edge row of a tile lives
      int r, rr, g, jrow;
      bool in_range, on;
      const float* hrow;
    };
    auto row_ctx = [&](int tile) {
      RowCtx c;
      c.r = tile_id(tile) * 4 + rl;
      c.in_range = c.r < R && k < K;
      c.rr = min(c.r, R - 1);
      const int s = c.rr / a.G;
      c.g = c.rr - s * a.G;
      c.on = c.in_range && a.matt[(size_t)c.g * K + k] != 0.f;
      c.jrow = c.in_range ? s * a.G + a.nbr[(size_t)c.g * K + k] : c.rr;
      c.hrow = a.hE_in + ((size_t)(a.he_shared ? c.g : c.rr) * K + (c.in_range ? k : 0)) * 128;
      return c;
    };
    auto row_total = [&](float partial, int which) -> float {  // sum over the threads that share a row
      red[(which * kGroups + grp) * 128 + m] = partial;
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");
      float tot = red[(which * kGroups) * 128 + m];
#pragma unroll
      for (int g2 = 1; g2 < kGroups; ++g2) tot += red[(which * kGroups + g2) * 128 + m];
      return tot;
    };
    uint32_t sf_phase = 0;
    float amax = 0.f;  // largest magnitude this thread split into fp16 halves (overflow report)
    // this thread's row in the staging buffer; 16-byte unit u of chunk c sits at (c << 14) + ((u ^ (k & 7)) << 4)
    uint8_t* const srow = stage + (rl << 12) + (k << 7);
    const int swz = k & 7;
    // A chunk number q of the kernel-wide schedule -> ring slot q % kSA, (q / kSA)-th use of that slot
    auto publish = [&](int q, const float* vals, int kc) {
      const int slot = q % kSA;
      mbar_wait(&a_empty[slot], ((q / kSA) & 1) ^ 1);
      put_chunk<PASSES>(Aring + slot * kSlotBytes, m, vals, kc, amax);
      fence_async_smem();
      mbar_arrive(&a_full[slot]);
    };
    auto load_acc = [&](uint32_t acc, int c, float (&dst)[32]) {
      uint32_t u[32];
      tmem_ld32(acc + lane_base + c * 32, u);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[i] = __uint_as_float(u[i]);
    };
    auto store_tmem = [&](uint32_t col, const float* vals) {
      uint32_t u[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(vals[i]);
      tmem_st32(col + lane_base, u);
    };

    // ---- first A operand of a tile (chunks qb .. qb+5): h_E row (4 chunks) and the pair geometry (32 + 16 columns).
    //      Edge update: the raw row is also parked in the TMEM region `stash` for the residual, instead of reading
    //      it from global memory a second time (a row-per-thread read costs 32 L1 wavefronts per instruction).
    //      `release`: the staging buffer is handed back to the loader once the rows have been consumed (node message
    //      path, first tile); otherwise the edge update's result rows of the previous tile do that.
    auto first_operand = [&](const RowCtx& c, int qb, uint32_t stash, bool release) {
      float v[32];
      float4 h[kCPT][8];
      mbar_wait(stage_full, sf_phase); sf_phase ^= 1;
#pragma unroll
      for (int t = 0; t < kCPT; ++t)
#pragma unroll
        for (int u = 0; u < 8; ++u)
          h[t][u] = c.in_range ? *reinterpret_cast<const float4*>(srow + ((grp + kGroups * t) << 14) + ((u ^ swz) << 4))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* fr4 = reinterpret_cast<const float4*>(a.geo + (size_t)c.g * PP_GEO_STRIDE);
      const float4* pi4 = reinterpret_cast<const float4*>(a.pglob + (size_t)c.rr * 24);
      const float4* pj4 = reinterpret_cast<const float4*>(a.pglob + (size_t)c.jrow * 24);
      float fr[12], pi[24], pj[24];
      if (grp < 2) {
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          float4 x = pj4[u];
          pj[u * 4] = x.x; pj[u * 4 + 1] = x.y; pj[u * 4 + 2] = x.z; pj[u * 4 + 3] = x.w;
        }
      }
      if (grp == 0) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          float4 x = fr4[u];
          fr[u * 4] = x.x; fr[u * 4 + 1] = x.y; fr[u * 4 + 2] = x.z; fr[u * 4 + 3] = x.w;
        }
      } else if (grp == 1) {
#pragma unroll
        for (int u = 0; u < 6; ++u) {
          float4 x = pi4[u];
          pi[u * 4] = x.x; pi[u * 4 + 1] = x.y; pi[u * 4 + 2] = x.z; pi[u * 4 + 3] = x.w;
        }
      }
#pragma unroll
      for (int t = 0; t < kCPT; ++t) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u * 4] = h[t][u].x; v[u * 4 + 1] = h[t][u].y; v[u * 4 + 2] = h[t][u].z; v[u * 4 + 3] = h[t][u].w;
        }
        if (MODE == 1) store_tmem(stash + (grp + kGroups * t) * 32, v);
        publish(qb + grp + kGroups * t, v, kKC);
      }
      // Hand the buffer back only now: the rows have provably left shared memory (their values were consumed by the
      // stores above).  Arriving right after issuing the loads let the next TMA copy overtake reads still in flight
      // (measured: one corrupted residue in ~1000 tiles when the copy hits in L2).
      if (release) mbar_arrive(stage_free);
      if (MODE == 1) tmem_st_wait();
      if (grp >= 2) return;  // the pair geometry is built by groups 0 (frames) and 1 (distances)
      float geo[32];
#pragma unroll
      for (int i = 8; i < kPairKC; ++i) geo[i] = 0.f;  // zero padding of the distance chunk (group 1)
#pragma unroll
      for (int pt = 0; pt < 8; ++pt) {
        const float jx = pj[pt * 3], jy = pj[pt * 3 + 1], jz = pj[pt * 3 + 2];
        if (grp == 0) {  // neighbour points in the local frame and their norms   (layers.py:93-97)
          float dx = jx - fr[9], dy = jy - fr[10], dz = jz - fr[11];
          float qx = fr[0] * dx + fr[3] * dy + fr[6] * dz;
          float qy = fr[1] * dx + fr[4] * dy + fr[7] * dz;
          float qz = fr[2] * dx + fr[5] * dy + fr[8] * dz;
          geo[pt * 3] = qx; geo[pt * 3 + 1] = qy; geo[pt * 3 + 2] = qz;
          geo[24 + pt] = sqrtf(qx * qx + qy * qy + qz * qz + 1e-8f);
        } else {         // distances between the global points   (layers.py:99-103)
          float gx = pi[pt * 3] - jx, gy = pi[pt * 3 + 1] - jy, gz = pi[pt * 3 + 2] - jz;
          geo[pt] = sqrtf(gx * gx + gy * gy + gz * gz + 1e-8f);
        }
      }
      if (grp == 0) publish(qb + 4, geo, kKC); else publish(qb + 5, geo, kPairKC);
    };

    int tile = next_tile((int)blockIdx.x);
    RowCtx cx = row_ctx(tile);
    int qbase = 0;  // first A chunk of the current tile (ring positions persist across tiles)
    stamp(a.trace_it == 0);  // 0: start
    if (tile < tend) first_operand(cx, 0, R0, true);
    stamp(a.trace_it == 0);  // 1: first operand of the first tile published

    for (int it = 0; tile < tend; ++it) {
      const bool t0 = it == a.trace_it;
      if (a.trace_it > 0) stamp(t0);  // tile start (a later tile: its first operand was built during the previous one)
      const int nxt_tile = next_tile(tile + tstep);
      const bool more = nxt_tile < tend;
      const int par = EDGE ? (it & 1) : 0;
      const uint32_t X = par ? ACC1 : ACC0, Y = par ? ACC0 : ACC1;
      const int xb = par, yb = par ^ 1;
      const uint32_t RE = par ? R1 : R0;   // raw h_E row, later e in fp32 (residual)
      const uint32_t PK = par ? R0 : R1;   // e as packed fp16: hi in columns [0, 64), lo in [64, 128)
      const bool on = cx.on, in_range = cx.in_range;
      const int r = cx.r, rr = cx.rr;
      RowCtx nx = cx;
      if (more) nx = row_ctx(nxt_tile);
      float v[32];

      // ---- epilogue of the head GEMM G1: x1 = relu(acc + A_i + N_j); the two gathered rows are fetched before the wait
      {
        const float* Ai = a.A + (size_t)rr * 128;
        const float* Nj = a.Nn + (size_t)cx.jrow * 128;
        float4 an[kCPT][8];
#pragma unroll
        for (int t = 0; t < kCPT; ++t)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 x = *reinterpret_cast<const float4*>(Ai + (grp + kGroups * t) * 32 + u * 4);
            float4 y = *reinterpret_cast<const float4*>(Nj + (grp + kGroups * t) * 32 + u * 4);
            an[t][u] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
          }
        mbar_wait(&acc_full[xb], accph[xb]); accph[xb] ^= 1;
        fence_after_sync();
        stamp(t0);  // 2: G1 complete
#pragma unroll
        for (int t = 0; t < kCPT; ++t) {
          load_acc(X, grp + kGroups * t, v);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            v[u * 4 + 0] = fmaxf(fmaf(v[u * 4 + 0], sG1, an[t][u].x), 0.f);
            v[u * 4 + 1] = fmaxf(fmaf(v[u * 4 + 1], sG1, an[t][u].y), 0.f);
            v[u * 4 + 2] = fmaxf(fmaf(v[u * 4 + 2], sG1, an[t][u].z), 0.f);
            v[u * 4 + 3] = fmaxf(fmaf(v[u * 4 + 3], sG1, an[t][u].w), 0.f);
          }
          publish(qbase + 6 + grp + kGroups * t, v, kKC);
        }
        stamp(t0);  // 3: x1 published
      }

      if (!EDGE) {
        // node message path: the next tile's first operand is built while G2 runs; its G1 then overlaps the reduction
        if (more) first_operand(nx, qbase + kChunksTile, 0, true);
        stamp(t0);  // next first operand published
        mbar_wait(&acc_full[1], accph[1]); accph[1] ^= 1;
        fence_after_sync();
        stamp(t0);  // 4: G2 complete
        // masked sum over the 32 edges of the residue (= the 32 lanes of this warp), layers.py:125-127
#pragma unroll
        for (int t = 0; t < kCPT; ++t) {
          const int c = grp + kGroups * t;
          load_acc(ACC1, c, v);
          const float* b = prm + kP_B2 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = on ? fmaxf(fmaf(v[i], sG2, b[i]), 0.f) : 0.f;
          // transpose-reduce: after the 5 steps lane l holds the sum of column l over the 32 lanes
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const bool upper = (lane & step) != 0;
#pragma unroll
            for (int i = 0; i < step; ++i) {
              float keep = upper ? v[i + step] : v[i];
              float send = upper ? v[i] : v[i + step];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
            }
          }
          if (r < R) a.out[(size_t)r * 128 + c * 32 + lane] = v[0];
        }
        stamp(t0);  // 5: sums written
      } else {
        // ---- epilogue of G2: x2 = relu(acc + b2)
        {
          mbar_wait(&acc_full[yb], accph[yb]); accph[yb] ^= 1;
          fence_after_sync();
          stamp(t0);  // 4: G2 complete
#pragma unroll
          for (int t = 0; t < kCPT; ++t) {
            const int c = grp + kGroups * t;
            load_acc(Y, c, v);
            const float* b = prm + kP_B2 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(v[i], sG2, b[i]), 0.f);
            publish(qbase + 10 + c, v, kKC);
          }
          stamp(t0);  // 5: x2 published
        }
        // ---- epilogue of G3: e = LN2(h_E + mask * (acc + b3))   (layers.py:139-142); e -> TMEM (fp32 and packed fp16)
        {
          const bool gate = on;
          mbar_wait(&acc_full[xb], accph[xb]); accph[xb] ^= 1;
          fence_after_sync();
          stamp(t0);  // 6: G3 complete
          float x[kCPT][32];
          float sum = 0.f;
#pragma unroll
          for (int t = 0; t < kCPT; ++t) {
            const int c = grp + kGroups * t;
            load_acc(RE, c, x[t]);  // the raw h_E row parked by first_operand
            load_acc(X, c, v);
            const float* b = prm + kP_B3 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              x[t][i] += gate ? fmaf(v[i], sG3, b[i]) : 0.f;
              sum += x[t][i];
            }
          }
          const float mean = row_total(sum, 0) * (1.f / 128.f);
          float var = 0.f;
#pragma unroll
          for (int t = 0; t < kCPT; ++t)
#pragma unroll
            for (int i = 0; i < 32; ++i) { float d = x[t][i] - mean; var += d * d; }
          const float rstd = rsqrtf(row_total(var, 1) * (1.f / 128.f) + 1e-5f);
#pragma unroll
          for (int t = 0; t < kCPT; ++t) {
            const int c = grp + kGroups * t;
            const float* gm = prm + kP_LN2G + c * 32;
            const float* bt = prm + kP_LN2B + c * 32;
            uint32_t eh[16], el[16];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (x[t][i] - mean) * rstd * gm[i] + bt[i];
#pragma unroll
            for (int i = 0; i < 16; ++i) split_f16x2(v[2 * i], v[2 * i + 1], eh[i], el[i], amax);
            store_tmem(RE + c * 32, v);
            tmem_st16(PK + lane_base + c * 16, eh);
            if (PASSES == 3) tmem_st16(PK + 64 + lane_base + c * 16, el);
          }
          tmem_st_wait();
          fence_before_sync();
          mbar_arrive(wk_done);
          stamp(t0);  // 7: e in TMEM
        }
        // ---- FFN: hidden slice j = relu(acc + b_in[j]) -> the four A chunks of FFN-out slice j
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&acc_full[yb], accph[yb]); accph[yb] ^= 1;
          fence_after_sync();
          stamp(t0);  // 8 + 2j: FFN-in slice j complete
          float v2[32];
          load_acc(Y, grp, v);
          if (kCPT == 2) load_acc(Y, grp + kGroups, v2);
          // Y is drained for this thread: the next FFN-in slice / the next tile's head may overwrite it.  Signalled
          // before the chunks are published, which can block on a ring slot that only FFN-out j-1 frees.
          fence_before_sync();
          mbar_arrive(wk_done);
#pragma unroll
          for (int t = 0; t < kCPT; ++t) {
            const int c = grp + kGroups * t;
            const float* b = prm + kP_BIN + j * 128 + c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(fmaf(t ? v2[i] : v[i], sFI, b[i]), 0.f);
            publish(qbase + 14 + 4 * j + c, v, kKC);
          }
          stamp(t0);  // 9 + 2j: hidden slice j published
        }
        // ---- the next tile's first operand is built while FFN-out drains; its head GEMM then overlaps the final
        //      LayerNorm and the stores below.  PK is free: the workers have seen FFN-in 3 complete.
        if (more) first_operand(nx, qbase + kChunksTile, PK, false);
        stamp(t0);  // 16: next first operand published
        // ---- final: y = LN3(e + acc + b_out) * mask   (layers.py:143-146)
        mbar_wait(&acc_full[xb], accph[xb]); accph[xb] ^= 1;
        fence_after_sync();
        stamp(t0);  // 17: FFN-out complete
        float y[kCPT][32];
        float sum = 0.f;
#pragma unroll
        for (int t = 0; t < kCPT; ++t) {
          const int c = grp + kGroups * t;
          load_acc(RE, c, y[t]);
          load_acc(X, c, v);
          const float* b = prm + kP_BOUT + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) { y[t][i] += fmaf(v[i], sFO, b[i]); sum += y[t][i]; }
        }
        const float mean3 = row_total(sum, 2) * (1.f / 128.f);
        float var = 0.f;
#pragma unroll
        for (int t = 0; t < kCPT; ++t)
#pragma unroll
          for (int i = 0; i < 32; ++i) { float d = y[t][i] - mean3; var += d * d; }
        const float rstd3 = rsqrtf(row_total(var, 3) * (1.f / 128.f) + 1e-5f);
        // result rows go into the staging buffer (every worker has read the next tile's rows out of it: the row_total
        // barriers above come after first_operand)
        {
#pragma unroll
          for (int t = 0; t < kCPT; ++t) {
            const int c = grp + kGroups * t;
            const float* gm = prm + kP_LN3G + c * 32;
            const float* bt = prm + kP_LN3B + c * 32;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float4 o;
              const int i = u * 4;
              o.x = on ? (y[t][i + 0] - mean3) * rstd3 * gm[i + 0] + bt[i + 0] : 0.f;
              o.y = on ? (y[t][i + 1] - mean3) * rstd3 * gm[i + 1] + bt[i + 1] : 0.f;
              o.z = on ? (y[t][i + 2] - mean3) * rstd3 * gm[i + 2] + bt[i + 2] : 0.f;
              o.w = on ? (y[t][i + 3] - mean3) * rstd3 * gm[i + 3] + bt[i + 3] : 0.f;
              *reinterpret_cast<float4*>(srow + (c << 14) + ((u ^ swz) << 4)) = o;
            }
          }
        }
        fence_async_smem();  // hand the rows to the tile loader's TMA stores
        mbar_arrive(stage_free);
        stamp(t0);  // 18: outputs written
      }
      // end of tile: this thread has read everything the tile left in TMEM
      fence_before_sync();
      mbar_arrive(tile_done);
      qbase += kChunksTile;
      cx = nx;
      tile = nxt_tile;
    }  // tile loop
    report_overflow(a.overflow, amax);
  }

  // ---- teardown: all tensor-core work of this CTA has been consumed by its workers; in a cluster nobody may exit
  //      while a peer can still multicast into its shared memory or arrive on its barriers
  fence_before_sync();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();
  if (warp == kWarpMMA) tmem_dealloc<512>(tmem);
}

// Tensor map of a [rows][K][128] fp32 tensor with boxes {32 floats, min(K, 32) rows, 1}: one residue's rows of one
// 32-column chunk, 128-byte swizzle (16-byte unit u of row k lands at k * 128 + ((u ^ (k & 7)) << 4))
static int make_row_map(CUtensorMap* tm, const float* base, long long rows, int K) {
  using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
      snprintf(g_last_error, sizeof(g_last_error), "edge_tc_kernel: cuTensorMapEncodeTiled is not available");
      return 1;
    }
    encode = reinterpret_cast<Encode>(fn);
  }
  const cuuint64_t gdim[3] = {128, (cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstr[2] = {512, (cuuint64_t)K * 512};
  const cuuint32_t box[3] = {32, (cuuint32_t)(K < 32 ? K : 32), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error), "edge_tc_kernel: cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return 1;
  }
  return 0;
}

template <int MODE, int PASSES, int CLUSTER>
static int launch(const Args& a, cudaStream_t stream) {
  auto kern = edge_tc_kernel<MODE, PASSES, CLUSTER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTC);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "edge_tc_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return 1;
  }
  const long long R = (long long)a.S * a.G;
  alignas(64) CUtensorMap tm_in, tm_out;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_out, 0, sizeof(tm_out));
  if (make_row_map(&tm_in, a.hE_in, a.he_shared ? a.G : R, a.K)) return 1;
  if (MODE == 1 && make_row_map(&tm_out, a.out, R, a.K)) return 1;
  unsigned tiles = (unsigned)((R + 3) / 4);
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  unsigned grid = tiles < (unsigned)num_sms ? tiles : (unsigned)num_sms;  // persistent: one CTA per SM
  grid = grid / CLUSTER * CLUSTER;
  if (grid == 0) grid = CLUSTER;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsTC);
  cfg.dynamicSmemBytes = kSmemTC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, a, tm_in, tm_out);
  if (e != cudaSuccess) {
    snprintf(g_last_error, sizeof(g_last_error), "edge_tc_kernel: launch: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

}  // namespace tc
}  // namespace pp

using namespace pp;

extern "C" int64_t pp_tc_stream_floats() { return tc::kStreamFloats; }

static unsigned long long* g_tc_trace = nullptr;
static int g_tc_trace_it = 0, g_tc_trace_path = 1;
// Diagnostics: subsequent pp_ipmp_edge_tc launches record clock64() stamps of the first tile of CTA 0 (one worker
// thread, phase boundaries, see the stamp() calls in edge_tc_kernel) into `trace` (device, >= 32 entries); NULL = off.
extern "C" int pp_set_tc_trace_tile(int64_t path, int64_t it) {
  g_tc_trace_path = (int)path;
  g_tc_trace_it = (int)it;
  return 0;
}
extern "C" int pp_set_tc_trace(uint64_t* trace) {
  g_tc_trace = reinterpret_cast<unsigned long long*>(trace);
  return 0;
}

// Tensor-core version of pp_ipmp_edge_node (path = 0) and pp_ipmp_edge_edge (path = 1).
//   wstream: operand images of this layer and path, pp_tc_stream_floats() floats (weights.py: pack_tc_stream)
//   passes : 3 = split fp16 (fp32-grade), 1 = plain fp16 inputs;  cluster: 1 or 2 CTAs sharing the weight stream
//   out    : accsum [S*G][128] (path 0) or hE_out [S*G][K][128] (path 1, may alias hE_in when he_shared == 0)
extern "C" int pp_ipmp_edge_tc(const float* weights, int64_t layer, int64_t path, const float* wstream,
                               const float* geo, const int32_t* nbr, const float* mask_attend, const float* msum,
                               int64_t G, int64_t K, int64_t S, const float* hE_in, int64_t he_shared, const float* wsA, const float* wsN,
                               const float* wsP, float* out, int64_t passes, int64_t cluster, int32_t* overflow,
                               const int32_t* live_tiles, const int32_t* n_live, cudaStream_t stream) {
  PP_REQUIRE(weights && wstream && geo && nbr && mask_attend && msum && hE_in && wsA && wsN && wsP && out,
             "null pointer");
  PP_REQUIRE(layer >= 0 && layer < 3 && (path == 0 || path == 1), "layer / path out of range");
  PP_REQUIRE(G > 0 && S > 0 && K > 0 && K <= PP_KMAX, "bad sizes");
  PP_REQUIRE(passes == 1 || passes == 3, "passes must be 1 or 3");
  PP_REQUIRE(cluster == 1 || cluster == 2, "cluster must be 1 or 2");
  const float* Lb = weights + layer * wl::kLayerStride;
  tc::Args a{};
  a.geo = geo; a.nbr = nbr; a.matt = mask_attend; a.msum = msum;
  a.G = (int)G; a.K = (int)K; a.S = (int)S;
  a.wstream = wstream;
  a.B2 = Lb + (path ? PP_OFF(L0_E_B2) : PP_OFF(L0_N_B2));
  a.B3 = Lb + PP_OFF(L0_E_B3);
  a.LNG = Lb + PP_OFF(L0_LN2_G); a.LNB = Lb + PP_OFF(L0_LN2_B);
  a.BIN = Lb + PP_OFF(L0_EF_BIN); a.BOUT = Lb + PP_OFF(L0_EF_BOUT);
  a.LN3G = Lb + PP_OFF(L0_LN3_G); a.LN3B = Lb + PP_OFF(L0_LN3_B);
  a.hE_in = hE_in; a.he_shared = (int)he_shared;
  a.A = wsA; a.Nn = wsN; a.pglob = wsP;
  a.out = out;
  a.overflow = overflow;
  a.live_list = live_tiles;
  a.n_live = live_tiles ? n_live : nullptr;
  PP_REQUIRE(!live_tiles || n_live, "live_tiles needs n_live");
  a.trace = path == g_tc_trace_path ? g_tc_trace : nullptr;  // the trace follows one of the two kernels
  a.trace_it = g_tc_trace_it;
  int rc;
#define PP_TC_CASE(E, P, C) if (path == (E) && passes == (P) && cluster == (C)) rc = tc::launch<E, P, C>(a, stream); else
  PP_TC_CASE(0, 3, 1) PP_TC_CASE(0, 3, 2) PP_TC_CASE(0, 1, 1) PP_TC_CASE(0, 1, 2)
  PP_TC_CASE(1, 3, 1) PP_TC_CASE(1, 3, 2) PP_TC_CASE(1, 1, 1) PP_TC_CASE(1, 1, 2)
  rc = 2;
#undef PP_TC_CASE
  if (rc) return rc;
  return check_launch("pp_ipmp_edge_tc");
}
