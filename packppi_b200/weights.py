"""`state_dict` layout of the reference model and a seeded random initialiser.

Key names and shapes are those produced by `TDiffusionModule.__init__`
(reference src/models/TorsionalDiffusion.py:39-68), so a trained reference checkpoint loads verbatim.
The Google-Drive checkpoint is not available offline; benchmarks and parity tests use
`make_state_dict(seed)`, which draws every tensor from its own `torch.Generator` so that the values do
not depend on module construction order (xavier-uniform for matrices as TorsionalDiffusion.py:80-82,
small uniform biases, LayerNorm gains near 1 so that gamma/beta are exercised).
"""
import math
from collections import OrderedDict

import torch

H = 128          # hidden / node / edge feature width (configs/model/model_cfg/MpnnNet.yaml:1)
NODE_IN = 51     # 35 + 16 time-embedding dims (encoder.py:76)
EDGE_IN = 468    # 65 relpos + 400 RBF + 1 chain type + 2 dihedrals (encoder.py:236)
N_POINTS = 8
MSG_IN = 2 * H + H + 9 * N_POINTS  # 456 (layers.py:49)
N_LAYERS = 3
TOP_K = 32


def shapes():
    s = OrderedDict()

    def lin(name, out, inp):
        s[name + ".weight"] = (out, inp)
        s[name + ".bias"] = (out,)

    def ln(name):
        s[name + ".weight"] = (H,)
        s[name + ".bias"] = (H,)

    lin("encoder.node_embedding", H, NODE_IN)
    ln("encoder.norm_nodes")
    lin("encoder.edge_embedding", H, EDGE_IN)
    ln("encoder.norm_edges")
    for l in range(N_LAYERS):
        p = f"mpnn.mpnn_layers.{l}."
        lin(p + "points_fn_node", 3 * N_POINTS, H)
        lin(p + "points_fn_edge", 3 * N_POINTS, H)
        for fn in ("node_message_fn", "edge_message_fn"):
            lin(p + fn + ".W_in", H, MSG_IN)
            lin(p + fn + ".W_inter.0", H, H)
            lin(p + fn + ".W_out", H, H)
        for i in range(4):
            ln(p + f"norm.{i}")
        for fn in ("node_dense", "edge_dense"):
            lin(p + fn + ".W_in", 4 * H, H)
            lin(p + fn + ".W_out", H, 4 * H)
    lin("decoder_score.0.W_in", H // 2, H)
    lin("decoder_score.0.W_out", H // 4, H // 2)
    lin("decoder_score.2.W_in", H // 8, H // 4)
    lin("decoder_score.2.W_out", 4, H // 8)
    return s


def num_parameters():
    n = 0
    for shp in shapes().values():
        k = 1
        for d in shp:
            k *= d
        n += k
    return n


def make_state_dict(seed=0, device="cpu"):
    sd = OrderedDict()
    for i, (name, shp) in enumerate(shapes().items()):
        g = torch.Generator().manual_seed(1000003 * int(seed) + i)
        if len(shp) == 2:
            bound = math.sqrt(6.0 / (shp[0] + shp[1]))
            w = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif ".norm" in name and name.endswith(".weight"):
            w = 1.0 + 0.1 * (torch.rand(shp, generator=g) * 2 - 1)
        else:
            w = 0.1 * (torch.rand(shp, generator=g) * 2 - 1)
        sd[name] = w.to(torch.float32).to(device)
    return sd


def check_state_dict(sd):
    exp = shapes()
    missing = [k for k in exp if k not in sd]
    if missing:
        raise RuntimeError(f"state_dict is missing keys: {missing[:4]}{'...' if len(missing) > 4 else ''}")
    for k, shp in exp.items():
        if tuple(sd[k].shape) != tuple(shp):
            raise RuntimeError(f"state_dict[{k}] has shape {tuple(sd[k].shape)}, expected {shp}")


def pack_weights(sd, layout, total_floats):
    """state_dict -> the flat fp32 blob the kernels read (layout from `_lib.layout()`, documented in
    csrc/weights_layout.h).  Runs on the host; the caller moves the result to the device."""
    check_state_dict(sd)
    f = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    blob = torch.zeros(total_floats, dtype=torch.float32)

    def put(name, t):
        off, size = layout[name]
        t = t.contiguous().reshape(-1)
        assert t.numel() <= size, (name, t.numel(), size)
        blob[off:off + t.numel()] = t

    put("ENC_NODE_WT", f["encoder.node_embedding.weight"].t())
    put("ENC_NODE_B", f["encoder.node_embedding.bias"])
    put("ENC_NODE_LNG", f["encoder.norm_nodes.weight"])
    put("ENC_NODE_LNB", f["encoder.norm_nodes.bias"])
    We = f["encoder.edge_embedding.weight"]  # [128, 468] = [relpos 65 | rbf 400 | type | phi psi]
    Wt = torch.zeros(496, H)
    Wt[0:400] = We[:, 65:465].t()
    Wt[400:465] = We[:, 0:65].t()
    Wt[480:483] = We[:, 465:468].t()
    put("ENC_EDGE_WT", Wt)
    put("ENC_EDGE_B", f["encoder.edge_embedding.bias"])
    put("ENC_EDGE_LNG", f["encoder.norm_edges.weight"])
    put("ENC_EDGE_LNB", f["encoder.norm_edges.bias"])
    for l in range(N_LAYERS):
        p = f"mpnn.mpnn_layers.{l}."
        for tag, pts, fn in (("N", "points_fn_node", "node_message_fn"), ("E", "points_fn_edge", "edge_message_fn")):
            q = f"L{l}_{tag}_"
            Win = f[p + fn + ".W_in.weight"]  # [128, 456]
            put(q + "WP", f[p + pts + ".weight"].t())
            put(q + "BP", f[p + pts + ".bias"])
            put(q + "WAG", torch.cat([Win[:, 0:128].t(), Win[:, 384:416].t()], 0))
            put(q + "B1", f[p + fn + ".W_in.bias"])
            put(q + "WN", Win[:, 256:384].t())
            put(q + "WEG", torch.cat([Win[:, 128:256].t(), Win[:, 416:456].t()], 0))
            put(q + "W2", f[p + fn + ".W_inter.0.weight"].t())
            put(q + "B2", f[p + fn + ".W_inter.0.bias"])
            put(q + "W3", f[p + fn + ".W_out.weight"].t())
            put(q + "B3", f[p + fn + ".W_out.bias"])
        for i in range(4):
            put(f"L{l}_LN{i}_G", f[p + f"norm.{i}.weight"])
            put(f"L{l}_LN{i}_B", f[p + f"norm.{i}.bias"])
        for tag, fn in (("NF", "node_dense"), ("EF", "edge_dense")):
            put(f"L{l}_{tag}_WIN", f[p + fn + ".W_in.weight"].t())
            put(f"L{l}_{tag}_BIN", f[p + fn + ".W_in.bias"])
            put(f"L{l}_{tag}_WOUT", f[p + fn + ".W_out.weight"].t())
            put(f"L{l}_{tag}_BOUT", f[p + fn + ".W_out.bias"])
    put("DEC_W0", f["decoder_score.0.W_in.weight"].t())
    put("DEC_B0", f["decoder_score.0.W_in.bias"])
    put("DEC_W1", f["decoder_score.0.W_out.weight"].t())
    put("DEC_B1", f["decoder_score.0.W_out.bias"])
    put("DEC_W2", f["decoder_score.2.W_in.weight"].t())
    put("DEC_B2", f["decoder_score.2.W_in.bias"])
    put("DEC_W3", f["decoder_score.2.W_out.weight"].t())
    put("DEC_B3", f["decoder_score.2.W_out.bias"])
    put("RBF_MU", torch.linspace(0.0, 20.0, 16))  # encoder.py:122-123
    put("TIME_FREQ", torch.exp(torch.arange(8, dtype=torch.float32) * -(math.log(10000) / 7)))  # layers.py:260-262
    return blob


def _f16_split(w, scale):
    """scale * w ~= hi + lo, both halves rounded to nearest fp16 (same split as split_f16x2 in csrc/umma.cuh); the
    pair resolves 22 mantissa bits.  Returned as int16 bit patterns."""
    ws = w.contiguous().to(torch.float32) * scale
    hi = ws.to(torch.float16)
    lo = (ws - hi.to(torch.float32)).to(torch.float16)
    return hi.view(torch.int16), lo.view(torch.int16)


def _f16_scale(w):
    """Power of two that brings max |w| into [2^13, 2^14): hi stays far below the fp16 maximum and the lo halves of
    all but vanishing weights stay in the normal fp16 range (the product is rescaled exactly in the epilogue)."""
    m = float(w.abs().max())
    return 1.0 if m == 0.0 else 2.0 ** (13 - math.floor(math.log2(m)))


def _umma_image16(w):
    """[rows, kc] K-major fp16 operand (int16 bit patterns) -> flat image in the no-swizzle core-matrix layout of
    csrc/umma.cuh: element offset(row, k) = (k // 8) * rows * 8 + (row // 8) * 64 + (row % 8) * 8 + k % 8."""
    rows, kc = w.shape
    assert rows % 8 == 0 and kc % 8 == 0
    return w.reshape(rows // 8, 8, kc // 8, 8).permute(2, 0, 1, 3).contiguous().reshape(-1)


def pack_tc_stream(sd, stream_floats):
    """Operand images for the tensor-core kernels (csrc/mpnn_tc.cu): [3 layers, 3 paths, stream_floats] float32 words
    holding fp16 (hi, lo) image pairs followed by 8 floats 1 / scale of (G1, G2, G3, FFN-in, FFN-out).

    Per (layer, path) the chunks follow the kernel's consumption order, each chunk = hi image then lo image:
      G1: W_in[:, h_E | pair geometry]  k = 0..167 zero-padded to 176, chunks of 32, 32, 32, 32, 32, 16
      G2: W_inter.0 (4 chunks), G3: W_out (4 chunks)
      FFN slices j = 0..3 (FFN-in = rows 128j..128j+127 of edge_dense.W_in, FFN-out = columns 128j..128j+127 of
      edge_dense.W_out, 4 chunks each), software-pipelined by one slice: in0, in1, out0, in2, out1, in3, out2, out3
    Every matrix is scaled by its own power of two (_f16_scale) before the split.  The node path (path 0) uses G1 and
    G2 of node_message_fn only.  Path 2 is the per-residue node epilogue: node_message_fn.W_out (4 chunks) followed by
    node_dense in the same pipelined order."""
    check_state_dict(sd)
    f = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    out = torch.zeros(N_LAYERS, 3, stream_floats, dtype=torch.float32)
    for l in range(N_LAYERS):
        p = f"mpnn.mpnn_layers.{l}."
        for path, fn in enumerate(("node_message_fn", "edge_message_fn", "node_message_fn")):
            Win = f[p + fn + ".W_in.weight"]
            inv = torch.ones(8)
            mats = []  # (matrix [128, k], scale)
            if path != 2:
                G1 = torch.cat([Win[:, 128:256], Win[:, 416:456], torch.zeros(128, 8)], 1)
                G2 = f[p + fn + ".W_inter.0.weight"]
                s1, s2 = _f16_scale(G1), _f16_scale(G2)
                inv[0], inv[1] = 1.0 / s1, 1.0 / s2
                mats += [(G1, s1), (G2, s2)]
            if path >= 1:
                dense = "edge_dense" if path == 1 else "node_dense"
                G3 = f[p + fn + ".W_out.weight"]
                Fi, Fo = f[p + dense + ".W_in.weight"], f[p + dense + ".W_out.weight"]
                s3, si, so = _f16_scale(G3), _f16_scale(Fi), _f16_scale(Fo)
                inv[2], inv[3], inv[4] = 1.0 / s3, 1.0 / si, 1.0 / so
                mats.append((G3, s3))
                mats.append((Fi[0:128, :], si))
                for j in range(4):
                    if j + 1 < 4:
                        mats.append((Fi[128 * (j + 1):128 * (j + 2), :], si))
                    mats.append((Fo[:, 128 * j:128 * (j + 1)], so))
            pieces = []
            for M, scale in mats:
                for k0 in range(0, M.shape[1], 32):
                    kc = min(32, M.shape[1] - k0)
                    hi, lo = _f16_split(M[:, k0:k0 + kc], scale)
                    pieces += [_umma_image16(hi), _umma_image16(lo)]
            flat = torch.cat(pieces).view(torch.float32)
            assert flat.numel() <= stream_floats - 8
            out[l, path, :flat.numel()] = flat
            out[l, path, stream_floats - 8:] = inv
    return out


def pack_pre_stream(sd, stream_floats):
    """Operand images for the tensor-core residue prologue (csrc/node_pre_tc.cu): [3 layers, 2 paths, stream_floats].

    Per (layer, path: 0 = node message, 1 = edge message), each chunk of 32 k-columns as hi image then lo image:
      W_p   points_fn weight [24 -> 32 rows, 128]                                     4 chunks of 32-row images
      W_ag  W_in[:, h_V_i (0:128) | own geometry (384:416)]  [128, 160]               5 chunks
      W_n   W_in[:, h_V_j (256:384)]                          [128, 128]               4 chunks
    followed by 8 floats: 1 / scale of (W_p, W_ag, W_n)."""
    check_state_dict(sd)
    f = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    out = torch.zeros(N_LAYERS, 2, stream_floats, dtype=torch.float32)
    for l in range(N_LAYERS):
        p = f"mpnn.mpnn_layers.{l}."
        for path, (pts, fn) in enumerate((("points_fn_node", "node_message_fn"), ("points_fn_edge", "edge_message_fn"))):
            Win = f[p + fn + ".W_in.weight"]
            Wp = torch.cat([f[p + pts + ".weight"], torch.zeros(8, H)], 0)
            mats = [Wp, torch.cat([Win[:, 0:128], Win[:, 384:416]], 1), Win[:, 256:384]]
            inv = torch.ones(8)
            pieces = []
            for i, M in enumerate(mats):
                scale = _f16_scale(M)
                inv[i] = 1.0 / scale
                for k0 in range(0, M.shape[1], 32):
                    hi, lo = _f16_split(M[:, k0:k0 + 32], scale)
                    pieces += [_umma_image16(hi), _umma_image16(lo)]
            flat = torch.cat(pieces).view(torch.float32)
            assert flat.numel() == stream_floats - 8, (flat.numel(), stream_floats)
            out[l, path, :flat.numel()] = flat
            out[l, path, stream_floats - 8:] = inv
    return out
