/* libpackppi_b200.so - C ABI of the B200-native PackPPI-MSC sampling / PackPPI-Prox kernels.
 *
 * The reference (Jackz915/PackPPI) is pure Python on PyTorch and has no FFI; each entry point below replaces the
 * Python function cited next to it (paths under the reference's src/).  A maintainer binds them with ctypes
 * (INTEGRATION.md shows the stub): plain device pointers from tensor.data_ptr(), int64 sizes, and the CUDA stream
 * from torch.cuda.current_stream().cuda_stream.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless marked host; fp32 unless the type says otherwise
 *   - the caller owns all memory (inputs, outputs, workspaces); the library never allocates or frees
 *   - calls are asynchronous and stream-ordered; no entry point synchronises with the host
 *   - return value 0 = success; otherwise pp_last_error() (thread-local, host) describes the failure
 *   - G = B*L residues of the padded batch, K = min(32, L) neighbours, S = diffusion samples that share the
 *     backbone; per-sample arrays have S*G rows, row r = s*G + g
 */
#ifndef PACKPPI_B200_H
#define PACKPPI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pp_stream_t; /* cudaStream_t */

int pp_abi_version(void);
const char* pp_last_error(void);
int pp_check_device(void); /* 0 iff the current device is compute capability 10.x */

/* Packed weight blob (packppi_b200/csrc/weights_layout.h).  Entry i has a name, an offset and a size in floats;
 * the host packs the reference state_dict (models/TorsionalDiffusion.py:39-68) by name. */
int64_t pp_layout_count(void);
int64_t pp_layout_total_floats(void);
int pp_layout_entry(int64_t i, const char** name, int64_t* offset, int64_t* size);
int64_t pp_geo_stride(void);   /* floats per residue geometry record */
int64_t pp_table_stride(void); /* floats per residue-type table record */

/* ProteinEncoder._dist (models/components/encoder.py:105-118) + mask_attend (models/components/mpnn.py:49-50).
 * X [B][L][14][3], residue_mask [B][L] -> E_idx int64 [B][L][K] (index inside the complex, ascending distance,
 * lowest index first among equal distances), nbr int32 [G][K] (row of the padded batch), D_neighbors [G][K] or
 * NULL, mask_attend [G][K] or NULL, msum [G] = mean_k mask_attend or NULL. */
int pp_knn_build(const float* X, const float* residue_mask, int64_t B, int64_t L, int64_t K, int64_t* E_idx,
                 int32_t* nbr, float* D_neighbors, float* mask_attend, float* msum, pp_stream_t stream);

/* Cell-list version of pp_knn_build: same arguments and bit-identical outputs, O(L * neighbourhood) instead of O(L^2)
 * (cells of >= 7 A on CA, warp-level top-K, ring expansion until the K-th neighbour is closer than the searched cube).
 * ws_int: B * 2 * (pp_knn_cells_max() + 1) + B * L int32; ws_box: B * 8 floats. */
int64_t pp_knn_cells_max(void);
int pp_knn_build_cells(const float* X, const float* residue_mask, int64_t B, int64_t L, int64_t K, int64_t* E_idx,
                       int32_t* nbr, float* D_neighbors, float* mask_attend, float* msum, int32_t* ws_int,
                       float* ws_box, pp_stream_t stream);

/* Rigid.from_3_points(N,CA,C) (utils/rigid_utils.py:1126-1179, fixed=True) and _impute_CB (encoder.py:137-142).
 * geo [G][pp_geo_stride()]: R(9, row-major) t(3) N CA C O CB(15). */
int pp_geometry_build(const float* X, int64_t G, float* geo, pp_stream_t stream);

/* Edge half of ProteinEncoder.forward (encoder.py:34-47,120-196,231-236,243-244): 468 features per edge,
 * Linear(468,128), LayerNorm -> hE0 [G][K][128].  Step-invariant: call once per batch. */
int pp_edge_embed(const float* weights, const float* geo, const int32_t* nbr, const int64_t* residue_index,
                  const int64_t* chain_indices, int64_t G, int64_t K, float* hE0, pp_stream_t stream);

/* Node half of ProteinEncoder.forward (encoder.py:217-229,239-242; layers.py:248-268) including the sin/cos of the
 * noised chi (models/TorsionalDiffusion.py:91-92).  bb_sincos [G][6], chi [S*G][4], chi_mask [G][4]; or, when
 * sc_sincos [S*G][8] is not NULL, the already masked sin/cos pairs as ProteinEncoder.forward receives them.
 * t: one value (t_stride 0) or one per row (t_stride 1) -> hV [S*G][128]. */
int pp_node_embed(const float* weights, const int64_t* residue_type, const float* bb_sincos, const float* chi,
                  const float* chi_mask, const float* sc_sincos, const float* t, int64_t t_stride, int64_t G,
                  int64_t S, float* hV, pp_stream_t stream);

/* InvariantPointMessagePassing.forward (models/components/layers.py:119-148), layer = 0..2.  hV is updated in
 * place; hE_in has G rows when he_shared != 0 (the step-invariant hE0) else S*G rows; hE_out [S*G][K][128] may
 * alias hE_in when he_shared == 0; edge_update == 0 skips the edge half.
 * Workspaces: wsA, wsN, wsAcc [S*G][128], wsP [S*G][24]. */
int pp_ipmp_layer(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                  const float* mask_attend, const float* msum, const float* residue_mask, int64_t G, int64_t K,
                  int64_t S, float* hV, const float* hE_in, int64_t he_shared, float* hE_out, int64_t edge_update,
                  float* wsA, float* wsN, float* wsP, float* wsAcc, pp_stream_t stream);

/* The four kernels of pp_ipmp_layer as separate entry points (same meaning of the arguments), for callers that
 * time or re-order them: per-residue prologue (IPMP points, hoisted h_V_i / h_V_j products; path 0 = node message,
 * 1 = edge message), per-edge node message + masked sum, per-residue node epilogue, per-edge edge update. */
int pp_ipmp_node_pre(const float* weights, int64_t layer, int64_t path, const float* geo, const int32_t* nbr,
                     const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S,
                     const float* hV, float* wsA, float* wsN, float* wsP, pp_stream_t stream);
int pp_ipmp_edge_node(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                      const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S,
                      const float* hE_in, int64_t he_shared, const float* wsA, const float* wsN, const float* wsP,
                      float* wsAcc, pp_stream_t stream);
int pp_ipmp_node_post(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                      const float* mask_attend, const float* msum, const float* residue_mask, int64_t G, int64_t K,
                      int64_t S, const float* wsAcc, float* hV, pp_stream_t stream);
int pp_ipmp_edge_edge(const float* weights, int64_t layer, const float* geo, const int32_t* nbr,
                      const float* mask_attend, const float* residue_mask, int64_t G, int64_t K, int64_t S,
                      const float* hE_in, int64_t he_shared, const float* wsA, const float* wsN, const float* wsP,
                      float* hE_out, pp_stream_t stream);

/* Tensor-core (tcgen05 / TMEM) version of pp_ipmp_edge_node (path 0) and pp_ipmp_edge_edge (path 1): same inputs,
 * same outputs.  wstream = operand images of this layer and path (pp_tc_stream_floats() floats, built by
 * packppi_b200.weights.pack_tc_stream: fp16 (hi, lo) image pairs, kind::f16 MMAs).  passes 3 = split fp16
 * (fp32-grade), 1 = plain fp16 inputs; cluster = 1 or 2 CTAs that share (multicast) the weight stream.
 * msum [G] = mean of mask_attend over K (pp_knn_build): tiles whose residues all have msum == 0 (padding) are
 * skipped and their output rows left untouched (cluster == 1; zero the buffers once, the kernels of this library
 * never write anything else there).  out = accsum [S*G][128] or hE_out [S*G][K][128]; hE_in / hE_out
 * move through TMA tensor copies and must be 16-byte aligned.
 * overflow (device int32, may be NULL; the same argument of pp_ipmp_node_pre_tc / pp_ipmp_node_post_tc32): the split
 * into fp16 halves has no per-tile scale, so an activation above 65504 becomes inf, the product NaN, and a ReLU would
 * turn that into a plausible 0.  A kernel that splits such a value ORs 1 into *overflow (never clears it): the caller
 * zeroes the flag, reads it after the pass and repeats the work in fp32 if it is set.
 * live_tiles / n_live (device, may be NULL): ids of the tiles (4 consecutive rows of the S*G) that hold at least one
 * residue with msum != 0, in ascending order, and their count; the persistent CTAs then take list positions b, b + grid,
 * ... instead of testing tiles b, b + grid, ... for liveness, which balances ragged, padded batches (+-1 tile).  The
 * same pair of arguments of pp_ipmp_node_pre_tc / pp_ipmp_node_post_tc32 lists live tiles of 128 rows; their output
 * rows in tiles of pure padding are then left untouched. */
int64_t pp_tc_stream_floats(void);
int pp_ipmp_edge_tc(const float* weights, int64_t layer, int64_t path, const float* wstream, const float* geo,
                    const int32_t* nbr, const float* mask_attend, const float* msum, int64_t G, int64_t K,
                    int64_t S, const float* hE_in, int64_t he_shared, const float* wsA, const float* wsN,
                    const float* wsP, float* out, int64_t passes, int64_t cluster, int32_t* overflow,
                    const int32_t* live_tiles, const int32_t* n_live, pp_stream_t stream);

/* pp_ipmp_node_post on the tensor cores with fp32-grade ("promoted") accumulation: every K = 16 step of a GEMM goes
 * into a fresh TMEM accumulator and the row threads sum the steps in fp32 registers, because the tensor core's own
 * fp32 accumulation truncates (csrc/node_post_tc.cu).  wstream = operand images of path 2 of this layer; hV
 * [S*G][128] is updated in place (reference layers.py:127-132). */
int pp_ipmp_node_post_tc32(const float* weights, int64_t layer, const float* wstream, const float* msum,
                           const float* residue_mask, int64_t G, int64_t K, int64_t S, const float* wsAcc, float* hV,
                           int32_t* overflow, const int32_t* live_tiles, const int32_t* n_live, pp_stream_t stream);

/* Tensor-core version of pp_ipmp_node_pre (reference layers.py:72-77,91 and the h_V_i / h_V_j columns of W_in):
 * tile = 128 residue rows, the three weight matrices resident in shared memory as fp16 (hi, lo) images.
 * wstream = operand images of this layer and path (pp_tc_pre_stream_floats() floats, weights.pack_pre_stream). */
int64_t pp_tc_pre_stream_floats(void);
int pp_ipmp_node_pre_tc(const float* weights, int64_t layer, int64_t path, const float* wstream, const float* geo,
                        int64_t G, int64_t S, const float* hV, float* wsA, float* wsN, float* wsP,
                        int32_t* overflow, const int32_t* live_tiles, const int32_t* n_live, pp_stream_t stream);

/* Diagnostics: later pp_ipmp_edge_tc launches write clock64() stamps of the phase boundaries of their first tile
 * (CTA 0, one worker thread) into trace (device memory, >= 32 uint64); NULL switches it off. */
int pp_set_tc_trace(uint64_t* trace);
/* Which kernel (path 0 = node message, 1 = edge update; default 1) and which tile of CTA 0 (default 0 = the first,
 * whose stamps include the prologue; for a later tile the stamps start at "G1 complete") the trace records. */
int pp_set_tc_trace_tile(int64_t path, int64_t it);

/* decoder_score (models/TorsionalDiffusion.py:62-68,106-108) and, if do_step, both SO2VESchedule.step calls in ode
 * mode plus wrap and mask (models/components/schedule.py:198-235, TorsionalDiffusion.py:272-280):
 *   chi <- wrap(chi + [step_mask] c_ode (score * w_anneal)) * chi_mask.
 * score_out [S*G][4] or NULL; step_mask uint8 [G][4]; chi_mask [G][4]; chi [S*G][4] in/out.
 * SDE branch (schedule.py:224-228) when noise_1pi / noise_2pi [S*G][4] (the two torch.normal draws of a step) and
 * mask_1pi uint8 [G][4] are given: chi += c_ode (score * w) + d_sde * noise with c_ode = g^2 dt, d_sde = g sqrt(dt).
 * With d_sde != 0 and the noise pointers NULL the normals are generated in the kernel: Philox4x32-10 keyed by `seed`,
 * counter (row, step_index) - one independent stream per (item, step), nothing stored. */
int pp_decode_step(const float* weights, const float* hV, int64_t G, int64_t S, float* score_out, int64_t do_step,
                   float c_ode, float w_anneal, const uint8_t* step_mask, const float* chi_mask, float* chi,
                   const float* noise_1pi, const float* noise_2pi, const uint8_t* mask_1pi, float d_sde, int64_t seed,
                   int64_t step_index, pp_stream_t stream);

/* get_atom14_coords (models/components/__init__.py:76-120).  tables [21][pp_table_stride()] from
 * packppi_b200.tables.packed_geometry(); chi [S*G][4] -> xyz_out [S*G][14][3]. */
int pp_atom14_fwd(const float* tables, const float* X, const int64_t* residue_type, const float* chi, int64_t G,
                  int64_t S, float* xyz_out, pp_stream_t stream);

/* Residue neighbour list of the clash term (no reference counterpart: clash.py:139-149 is dense).
 * fill = 0: writes reach [G] and counts [G].  The caller builds start [G+1] = exclusive prefix sum of counts.
 * fill = 1: writes list [start[G]] (ascending neighbour index).  cutoff = 2 * max radius - overlap tolerance. */
int pp_clash_neighbours(const float* tables, const float* X, const int64_t* residue_type, const float* atom_exists,
                        const int64_t* residue_index, int64_t B, int64_t L, float cutoff, int64_t fill, float* reach,
                        int32_t* counts, const int64_t* start, int32_t* list, pp_stream_t stream);

/* Spatially hashed version: pp_clash_reach computes reach [G]; pp_clash_neighbours_cells bins the residues into cells
 * of edge h_min >= 2 * max(reach) + cutoff and searches the 27 surrounding cells (same counts and list as
 * pp_clash_neighbours, ascending order).  fill = 0 bins and counts, fill = 1 reuses the bins and writes the list.
 * Workspaces as for pp_knn_build_cells. */
int pp_clash_reach(const float* tables, const float* X, const int64_t* residue_type, const float* atom_exists,
                   int64_t G, float* reach, pp_stream_t stream);
int pp_clash_neighbours_cells(const float* X, const float* reach, const int64_t* residue_index, int64_t B, int64_t L,
                              float cutoff, float h_min, int64_t fill, int32_t* counts, const int64_t* start,
                              int32_t* list, int32_t* ws_int, float* ws_box, pp_stream_t stream);

/* compute_residue_clash (models/components/clash.py:335-365) and its gradient.
 * lower/upper [21][14][14] from make_atom14_dists_bounds (utils/residue_constants.py:809-869).
 * mode 0: per_res [S*G].  mode 1: also grad_chi [S*G][4] = d(sum_r res_w[r] per_res[r]) / d chi (analytic).
 * Workspaces: atoms4 [S*G][14][4], axes [S*G][4][6], bound [S*G]. */
int pp_clash_fwd_bwd(const float* tables, const float* lower, const float* upper, const float* X,
                     const int64_t* residue_type, const float* atom_exists, const int64_t* nbr_start,
                     const int32_t* nbr_list, const float* chi, int64_t G, int64_t S, float tol, float max_cut,
                     int64_t mode, const float* res_w, float* per_res, float* grad_chi, float* atoms4, float* axes,
                     float* bound, pp_stream_t stream);

/* proximal_optimizer (models/components/optimize.py:5-73) for S samples of a padded batch of B complexes.  The
 * reference handles one complex per call (optimize.py:27 asserts num_proteins == 1) and the notebooks loop over
 * decoys; here the S*B (sample, complex) items are independent problems sharing every launch.  Rows: item
 * it = s*B + b owns rows s*B*L + b*L + [0, n_res[b]); n_res int32 [B] (NULL: every complex has L residues) gives the
 * unpadded residue counts, over which the reference's means are taken.
 * pp_prox_init: find_clash_mask + optimiser state per item.  mean_out [S*B][2], slot 1 = mean per-residue loss.
 * pp_prox_step: one Adam step of every item in two launches (atom14 rebuild, pair kernel).  The objective of a step
 *   is left as per-row terms in loss_rows [S*B*L][2]; the NEXT call reduces them into prev_loss_out [S*B][2]
 *   (NULL on the first step) in the tail blocks of its rebuild launch, and pp_prox_loss reduces the last step.
 *   loss[it][0] = objective before the update (the value the reference appends to loss_list), loss[it][1] = mean
 *   per-residue clash; snapshot = where(mask, x_new, SC_D).
 * step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t) as torch.optim.Adam computes them on the host. */
int pp_prox_init(const float* tables, const float* lower, const float* upper, const float* X,
                 const int64_t* residue_type, const float* atom_exists, const int64_t* nbr_start,
                 const int32_t* nbr_list, const float* sc_d, int64_t B, int64_t L, int64_t S, const int32_t* n_res,
                 float tol, float max_cut, uint8_t* mask, float* z, float* x, float* m, float* v, float* per_res,
                 float* mean_out, float* atoms4, float* axes, float* bound, float* loss_rows, pp_stream_t stream);
int pp_prox_step(const float* tables, const float* lower, const float* upper, const float* X,
                 const int64_t* residue_type, const float* atom_exists, const int64_t* nbr_start,
                 const int32_t* nbr_list, const float* sc_d, const uint8_t* mask, const float* z, float* x, float* m,
                 float* v, int64_t B, int64_t L, int64_t S, const int32_t* n_res, float tol, float max_cut, float lamda,
                 float step_size, float bc2_sqrt, float beta1, float beta2, float eps, float* snapshot,
                 float* prev_loss_out, float* per_res, float* atoms4, float* axes, float* bound, float* loss_rows,
                 const uint8_t* owned, int64_t n_total, pp_stream_t stream);
int pp_prox_loss(const float* loss_rows, int64_t B, int64_t L, int64_t S, const int32_t* n_res, float lamda,
                 int64_t n_total, float* loss_out, pp_stream_t stream);
/* Slab-partitioned complex (one rank per slab): B = S = 1, the arrays hold the rank's owned
 * residues plus the halo; owned uint8 [L] (NULL = all) marks the residues this rank optimises and counts in the
 * loss, n_total is the residue count of the whole complex (0 = per-complex counts).  pp_prox_init_from_mean builds
 * mask / z / x from per_res and the all-reduced mean[1]. */
int pp_prox_init_from_mean(const float* per_res, const float* mean, const float* sc_d, const uint8_t* owned, int64_t G,
                           uint8_t* mask, float* z, float* x, float* m, float* v, pp_stream_t stream);

/* Diagnostics: one 128 x 128 tile D = A W^T on the tcgen05 tensor cores (A [128][K], W [128][K], K % 32 == 0)
 * through kind::f16 with fp16 (hi, lo) operand pairs (the format the fused kernels use): passes 1 = plain
 * fp16 inputs, 3 = split fp16 (~fp32).  ts_mode 1 feeds A from tensor memory as packed fp16 pairs; ts_mode 2 starts
 * a fresh accumulator for every K = 16 step and sums the steps in fp32 with round-to-nearest ("promotion", the scheme
 * of pp_ipmp_node_post_tc32), which removes the bias of the tensor core's truncating fp32 accumulation. */
int pp_selftest_umma_f16(const float* A, const float* W, float* D, int64_t K, int64_t passes, int64_t ts_mode,
                         pp_stream_t stream);

/* Device featurisation of a padded batch [B][L]: restates ComplexDataset.prot_to_data
 * (src/datamodules/components/complex_dataset.py:64-148), calc_bb_dihedrals / calc_sc_dihedrals (helper.py:39-101)
 * and the zero padding of collate_fn (complex_datamodule.py:196-226).  Inputs: atom14 coordinates (NaN = missing
 * atom), residue types, atom masks, chain-offset residue indices, 1-based chain numbers, residue count per complex,
 * chi tables [21][7] / [21][4] / [21][4].  Outputs: every tensor field of the batch contract, zero in the padding. */
int pp_featurize(const float* X_in, const int64_t* aatype, const float* atom_mask_in, const int64_t* ridx_in,
                 const int64_t* chain_in, const int32_t* length, int64_t B, int64_t L, const int32_t* chi_atoms,
                 const float* chi_mask, const float* chi_pi, float* X, float* atom_mask, int64_t* residue_type,
                 float* residue_mask, int64_t* residue_index, int64_t* chain_indices, float* bb_d, float* bb_sincos,
                 float* bb_mask, float* sc_d, float* sc_sincos, float* sc_mask, uint8_t* chi_1pi, uint8_t* chi_2pi,
                 pp_stream_t stream);

/* Diagnostics: one TMA gather4 copy (four arbitrary rows of src [rows][128], 32 floats from column col, 128-byte
 * swizzle, tensor-map box {32, box_rows}); out[256] receives the raw 1 KB of shared memory. */
int pp_selftest_gather4(const float* src, int64_t rows, int64_t box_rows, int64_t col, int64_t r0, int64_t r1,
                        int64_t r2, int64_t r3, float* out, pp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
