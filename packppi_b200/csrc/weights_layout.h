// Packed fp32 weight blob: one contiguous device buffer, every matrix stored K-major ("Wt[in][out]",
// the transpose of torch's Linear.weight) so that a GEMM B-operand row is contiguous.  The host side
// (packppi_b200/weights.py: pack_weights) queries this table through pp_layout_entry() and fills the
// blob by name, so the two sides cannot drift apart.
//
// state_dict provenance (reference src/models/TorsionalDiffusion.py:39-68, layers.py:36-63):
//   ENC_NODE_WT  encoder.node_embedding.weight^T           rows padded 51 -> 52
//   ENC_EDGE_WT  encoder.edge_embedding.weight^T, rows reordered to the kernel's feature chunks:
//                [25 atom pairs x 16 RBF | 65 relpos classes + 15 zero rows | chain type, phi, psi + 13 zero rows]
//   per layer l and path p in {N(ode), E(dge)}  (message_fn = node_message_fn / edge_message_fn):
//     WP   points_fn_{node,edge}.weight^T [128][24]
//     WAG  rows 0..127  = message_fn.W_in.weight[:, 0:128]^T     (h_V of the centre residue)
//          rows 128..159 = message_fn.W_in.weight[:, 384:416]^T  (own local points 24 + their norms 8)
//     WN   message_fn.W_in.weight[:, 256:384]^T                  (h_V of the neighbour)
//     WEG  rows 0..127  = message_fn.W_in.weight[:, 128:256]^T   (h_E)
//          rows 128..167 = message_fn.W_in.weight[:, 416:456]^T  (neighbour points in local frame 24, norms 8, global distances 8)
//     W2 / W3  message_fn.W_inter.0 / W_out
//   FFN: {node,edge}_dense.W_in^T [128][512], W_out^T [512][128];  LN0..LN3 = norm.0..3
#pragma once

#define PP_LAYER_ENTRIES(X, L)                                                          \
  X(L##_N_WP, 128 * 24) X(L##_N_BP, 32) X(L##_N_WAG, 160 * 128) X(L##_N_B1, 128)       \
  X(L##_N_WN, 128 * 128) X(L##_N_WEG, 168 * 128) X(L##_N_W2, 128 * 128) X(L##_N_B2, 128) \
  X(L##_N_W3, 128 * 128) X(L##_N_B3, 128)                                               \
  X(L##_E_WP, 128 * 24) X(L##_E_BP, 32) X(L##_E_WAG, 160 * 128) X(L##_E_B1, 128)       \
  X(L##_E_WN, 128 * 128) X(L##_E_WEG, 168 * 128) X(L##_E_W2, 128 * 128) X(L##_E_B2, 128) \
  X(L##_E_W3, 128 * 128) X(L##_E_B3, 128)                                               \
  X(L##_LN0_G, 128) X(L##_LN0_B, 128) X(L##_LN1_G, 128) X(L##_LN1_B, 128)               \
  X(L##_LN2_G, 128) X(L##_LN2_B, 128) X(L##_LN3_G, 128) X(L##_LN3_B, 128)               \
  X(L##_NF_WIN, 128 * 512) X(L##_NF_BIN, 512) X(L##_NF_WOUT, 512 * 128) X(L##_NF_BOUT, 128) \
  X(L##_EF_WIN, 128 * 512) X(L##_EF_BIN, 512) X(L##_EF_WOUT, 512 * 128) X(L##_EF_BOUT, 128)

#define PP_WEIGHT_ENTRIES(X)                                                          \
  X(ENC_NODE_WT, 52 * 128) X(ENC_NODE_B, 128) X(ENC_NODE_LNG, 128) X(ENC_NODE_LNB, 128) \
  X(ENC_EDGE_WT, 496 * 128) X(ENC_EDGE_B, 128) X(ENC_EDGE_LNG, 128) X(ENC_EDGE_LNB, 128) \
  PP_LAYER_ENTRIES(X, L0) PP_LAYER_ENTRIES(X, L1) PP_LAYER_ENTRIES(X, L2)               \
  X(DEC_W0, 128 * 64) X(DEC_B0, 64) X(DEC_W1, 64 * 32) X(DEC_B1, 32)                    \
  X(DEC_W2, 32 * 16) X(DEC_B2, 16) X(DEC_W3, 16 * 4) X(DEC_B3, 4)                       \
  X(RBF_MU, 16) X(TIME_FREQ, 8)

namespace pp {
namespace wl {

enum Entry {
#define X(name, n) name,
  PP_WEIGHT_ENTRIES(X)
#undef X
      NUM_ENTRIES
};

struct Info {
  const char* name;
  long long size;
};

static constexpr Info kInfo[] = {
#define X(name, n) {#name, (long long)(n)},
    PP_WEIGHT_ENTRIES(X)
#undef X
};

// offsets are rounded up to 32 floats (128 B) so every matrix row group is cp.async / float4 friendly
constexpr long long offset_of(int e) {
  long long o = 0;
  for (int i = 0; i < e; ++i) o += (kInfo[i].size + 31) / 32 * 32;
  return o;
}
constexpr long long kTotalFloats = offset_of(NUM_ENTRIES);

// entries of one layer are laid out identically; LAYER_STRIDE lets kernels index by layer number
constexpr long long kLayerStride = offset_of(L1_N_WP) - offset_of(L0_N_WP);
static_assert(offset_of(L2_N_WP) - offset_of(L1_N_WP) == kLayerStride, "layer stride");

}  // namespace wl
}  // namespace pp

namespace pp {
namespace wl {
template <int E>
struct Off {
  static constexpr long long v = offset_of(E);  // forced compile-time evaluation: usable in device code
};
}  // namespace wl
}  // namespace pp
#define PP_OFF(e) (pp::wl::Off<pp::wl::e>::v)
