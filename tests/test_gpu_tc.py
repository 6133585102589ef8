"""Tensor-core (tcgen05) execution modes of the message MLPs against the golden vectors and the exact fp32 kernels."""
import pytest
import torch

from util import load_golden, tt, wrapped_diff

pytestmark = pytest.mark.gpu

# split fp16 pairs keep 22 mantissa bits per product: same gates as the exact fp32 path.
# plain fp16 inputs keep 11 bits (the precision of TF32); measured on the fixtures that moves the final chi by up to
# 3.3e-3 rad (mean 3-5e-4; SURVEY.md §0 fact 8: bf16 weights alone move it by 5e-2), so it is a separate, explicitly
# looser, mode.
TOL = {"f16x3": dict(act=2e-4, chi=1e-4), "f16": dict(act=2e-2, chi=2e-2)}


def _model(dev, mode, cluster=1):
    from packppi_b200 import TDiffusionModule, weights
    m = TDiffusionModule()
    m.load_state_dict(weights.make_state_dict(0))
    m.kernel_mode, m.kernel_cluster = mode, cluster
    return m.to(dev).eval()


@pytest.mark.parametrize("case", ["syn5", "syn17", "syn33", "syn64", "synbatch", "syn300", "1brs", "t1124"])
@pytest.mark.parametrize("mode,cluster", [("f16x3", 1), ("f16x3", 2), ("f16", 1), ("f16", 2)])
def test_network_probe_tc(case, mode, cluster):
    dev = torch.device("cuda:0")
    g, b = load_golden(case)
    bd = b.to(dev)
    B, L = b.X.shape[:2]
    model = _model(dev, mode, cluster)
    x = tt(g["in_probe_SC_D"]).to(dev)
    score, hV = model.network(bd, x, torch.full((B * L,), 0.7, device=dev))
    torch.cuda.synchronize()
    tol = TOL[mode]["act"]
    assert (hV.cpu() - tt(g["ref_probe_hV"])).abs().max().item() < tol
    assert (score.cpu() - tt(g["ref_probe_score"])).abs().max().item() < tol


@pytest.mark.parametrize("case", ["syn33", "synbatch", "1brs", "t1124"])
@pytest.mark.parametrize("mode,cluster", [("f16x3", 1), ("f16x3", 2), ("f16", 2)])
def test_sampling_tc(case, mode, cluster):
    dev = torch.device("cuda:0")
    g, b = load_golden(case)
    bd = b.to(dev)
    model = _model(dev, mode, cluster)
    out = model.sampling(bd, init_SC_D=tt(g["in_SC_D_init"]).to(dev))
    torch.cuda.synchronize()
    d = wrapped_diff(out.cpu(), tt(g["ref_SC_D_final"]))
    print(f"{case} {mode} cluster={cluster}: max chi diff {d.max().item():.3e} rad, mean {d.mean().item():.3e}")
    assert d.max().item() < TOL[mode]["chi"]


def test_padding_tiles_are_skipped_and_zeroed():
    """A batch padded to its longest complex: the tensor-core kernels skip the tiles that hold only padding residues.
    The engine zeroes the padding rows of h_E once per graph (the buffers are poisoned first); h_V rows in skipped
    128-row tiles keep the node embedding (nothing unmasked reads them, and `network()` zeroes them for the caller);
    everything else must equal the CUDA-core kernels."""
    from packppi_b200 import TDiffusionModule, weights, synthetic, _lib
    from packppi_b200.batch import collate
    dev = torch.device("cuda:0")
    batch = collate([synthetic.make_complex((20, 21), seed=1), synthetic.make_complex((90, 83), seed=2),
                     synthetic.make_complex((30, 30), seed=3)]).to(dev)
    B, L = batch.X.shape[:2]
    chi = (torch.rand(2, B, L, 4, generator=torch.Generator().manual_seed(5)) * 6.0 - 3.0).to(dev)
    outs = {}
    for mode in ("fp32", "f16x3"):
        m = TDiffusionModule()
        m.load_state_dict(weights.make_state_dict(0))
        m.kernel_mode = mode
        m = m.to(dev).eval()
        eng, graph = m._graph(batch)
        ws = eng.workspace(graph.G, graph.K, 2)
        for buf in (ws.hE, ws.hV, ws.wsAcc):
            buf.fill_(float("nan"))
        ws.clean_for = None
        ni = eng.node_inputs(batch)
        t = torch.full((2 * B * L,), 0.4, device=dev)
        eng.forward_layers(graph, ws, ni, chi.reshape(-1, 4).contiguous(), t, 1)
        torch.cuda.synchronize()
        outs[mode] = (ws.hV.clone(), ws.hE.clone())
    pad = (batch.residue_mask.reshape(-1) == 0).repeat(2)
    assert int(pad.sum()) >= 2 * 200
    for mode, (hV, hE) in outs.items():
        assert torch.isfinite(hV).all() and torch.isfinite(hE).all(), mode
        assert hE.reshape(hV.shape[0], -1)[pad].abs().max().item() == 0.0, mode
        if mode == "fp32":
            assert hV[pad].abs().max().item() == 0.0
    assert (outs["fp32"][0] - outs["f16x3"][0])[~pad].abs().max().item() < TOL["f16x3"]["act"]
    assert (outs["fp32"][1] - outs["f16x3"][1]).abs().max().item() < TOL["f16x3"]["act"]
    # through the public API the padding rows are zero in every mode
    score, hV = m.network(batch, chi, torch.full((2 * B * L,), 0.4, device=dev))
    assert hV.reshape(-1, 128)[pad].abs().max().item() == 0.0 and torch.isfinite(score).all()


@pytest.mark.parametrize("mode", ["f16x3", "f16"])
def test_repeated_evaluation_is_bit_identical(mode):
    """All reductions of the tensor-core kernels have a fixed order, so repeating an evaluation must reproduce every
    bit; a difference means a synchronisation bug (regression: a TMA copy once overtook shared-memory reads that were
    still in flight, corrupting about one residue in 1000 tiles when the copy hit in L2)."""
    from packppi_b200 import TDiffusionModule, weights, synthetic
    from packppi_b200.batch import collate
    dev = torch.device("cuda:0")
    m = TDiffusionModule()
    m.load_state_dict(weights.make_state_dict(0))
    m.kernel_mode = mode
    m = m.to(dev).eval()
    for items, S in (([synthetic.make_complex((500, 500, 500), seed=3)], 2),
                     ([synthetic.make_complex((20, 21), seed=1), synthetic.make_complex((190, 183), seed=2)], 4)):
        b = collate(items).to(dev)
        B, L = b.X.shape[:2]
        x = ((torch.rand(S, B, L, 4, generator=torch.Generator().manual_seed(1)) * 2 - 1) * 3.14).to(dev)
        eng, graph = m._graph(b)
        t = torch.full((S * B * L,), 0.4, device=dev)
        ref = None
        for _ in range(30):
            _, hV = eng.network(graph, b, x.reshape(-1, 4).contiguous(), t)
            if ref is None:
                ref = hV.clone()
            else:
                assert torch.equal(ref, hV)


@pytest.mark.parametrize("case", ["syn17", "synbatch", "t1124"])
def test_node_pre_tc_matches_cuda_core_kernel(case):
    """Residue prologue (IPMP points, A_i, N_j) on the tensor cores against the exact fp32 kernel, same inputs."""
    from packppi_b200 import TDiffusionModule, weights, _lib
    dev = torch.device("cuda:0")
    g, b = load_golden(case)
    bd = b.to(dev)
    B, L = b.X.shape[:2]
    m = _model(dev, "f16x3")
    eng, graph = m._graph(bd)
    G, K, S = graph.G, graph.K, 2
    hV = torch.randn(S * G, 128, generator=torch.Generator().manual_seed(2)).to(dev)
    W = eng.wblob
    for layer in range(3):
        for path in (0, 1):
            ref = [torch.zeros(S * G, n, device=dev) for n in (128, 128, 24)]
            out = [torch.full((S * G, n), float("nan"), device=dev) for n in (128, 128, 24)]
            _lib.call("pp_ipmp_node_pre", W, layer, path, graph.geo, graph.nbr, graph.mask_attend, graph.mask, G, K, S,
                      hV, *ref)
            _lib.call("pp_ipmp_node_pre_tc", W, layer, path, eng.wpre[layer, path], graph.geo, G, S, hV, *out, None, None, None)
            torch.cuda.synchronize()
            for name, r, o in zip(("A", "N", "P"), ref, out):
                scale = max(1.0, r.abs().max().item())
                assert (r - o).abs().max().item() < 2e-5 * scale, (layer, path, name)


@pytest.mark.parametrize("case", ["syn17", "synbatch", "t1124"])
def test_node_post_tc32_matches_cuda_core_kernel(case):
    """Node update with promoted (fp32-grade) tensor-core accumulation against the exact fp32 kernel, same inputs."""
    from packppi_b200 import _lib
    dev = torch.device("cuda:0")
    g, b = load_golden(case)
    bd = b.to(dev)
    m = _model(dev, "f16x3")
    eng, graph = m._graph(bd)
    G, K, S = graph.G, graph.K, 2
    gen = torch.Generator().manual_seed(3)
    hV0 = torch.randn(S * G, 128, generator=gen).to(dev)
    acc = (torch.randn(S * G, 128, generator=gen) * 8.0).to(dev)
    W = eng.wblob
    for layer in range(3):
        ref, out = hV0.clone(), hV0.clone()
        _lib.call("pp_ipmp_node_post", W, layer, graph.geo, graph.nbr, graph.mask_attend, graph.msum, graph.mask, G, K,
                  S, acc, ref)
        _lib.call("pp_ipmp_node_post_tc32", W, layer, eng.wtc[layer, 2], graph.msum, graph.mask, G, K, S, acc, out, None, None, None)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        assert (ref - out).abs().max().item() < 5e-6 * max(1.0, ref.abs().max().item()), layer


@pytest.mark.parametrize("node_epilogue", ["tc32", "ffma"])
def test_node_epilogue_options(node_epilogue):
    """The two homes of the per-residue node update: promoted tensor-core accumulation (default) and the CUDA-core
    kernel, both within the activation gate of one network evaluation.  (A third one, plain TMEM accumulation, missed the
    1e-4 rad gate after the first ODE steps - the tensor core's fp32 accumulation truncates - and was removed.)"""
    from packppi_b200 import TDiffusionModule, weights
    dev = torch.device("cuda:0")
    g, b = load_golden("t1124")
    bd = b.to(dev)
    B, L = b.X.shape[:2]
    m = TDiffusionModule()
    m.load_state_dict(weights.make_state_dict(0))
    m.kernel_mode, m.kernel_node_epilogue = "f16x3", node_epilogue
    m = m.to(dev).eval()
    assert m.engine(dev).node_epilogue == node_epilogue
    score, hV = m.network(bd, tt(g["in_probe_SC_D"]).to(dev), torch.full((B * L,), 0.7, device=dev))
    torch.cuda.synchronize()
    assert (hV.cpu() - tt(g["ref_probe_hV"])).abs().max().item() < TOL["f16x3"]["act"]
    assert (score.cpu() - tt(g["ref_probe_score"])).abs().max().item() < TOL["f16x3"]["act"]


@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_tensor_cores_match_cuda_cores(seed):
    """Shape fuzz: ragged batches of 1-3 complexes of 4-70 residues, 1-4 samples (tiles that straddle samples and
    complexes, K < 32, partial last tiles); one network evaluation, tensor-core mode against the exact fp32 kernels."""
    import random
    from packppi_b200 import TDiffusionModule, synthetic, weights
    from packppi_b200.batch import collate
    rnd = random.Random(1000 + seed)
    dev = torch.device("cuda:0")
    items = []
    for i in range(rnd.randint(1, 3)):
        n = rnd.randint(4, 70)
        a = rnd.randint(2, n - 2)
        items.append(synthetic.make_complex((a, n - a), seed=seed * 10 + i))
    b = collate(items).to(dev)
    B, L = b.X.shape[:2]
    S = rnd.randint(1, 4)
    x = ((torch.rand(S, B, L, 4, generator=torch.Generator().manual_seed(seed)) * 2 - 1) * 3.14).to(dev)
    t = torch.full((S * B * L,), 0.3 + 0.1 * seed, device=dev)
    res = {}
    for mode in ("fp32", "f16x3"):
        m = TDiffusionModule()
        m.load_state_dict(weights.make_state_dict(0))
        m.kernel_mode = mode
        m = m.to(dev).eval()
        eng, graph = m._graph(b)
        score, hV = eng.network(graph, b, x.reshape(-1, 4).contiguous(), t)
        torch.cuda.synchronize()
        res[mode] = (score.clone(), hV.clone())
    assert torch.isfinite(res["f16x3"][1]).all()
    assert (res["fp32"][1] - res["f16x3"][1]).abs().max().item() < TOL["f16x3"]["act"], (B, L, S)
    assert (res["fp32"][0] - res["f16x3"][0]).abs().max().item() < TOL["f16x3"]["act"], (B, L, S)


def test_fp16_overflow_is_detected_and_redone_in_fp32():
    """the split-fp16 operands carry no per-tile scale, so a checkpoint whose hidden activations exceed
    65504 overflows.  Contract: never a silently wrong angle - the result is NaN inside the kernels, `sampling` notices
    and repeats the call with the fp32 kernels."""
    import warnings
    from packppi_b200 import TDiffusionModule, weights
    dev = torch.device("cuda:0")
    g, b = load_golden("syn33")
    sd = weights.make_state_dict(0)
    sd = {k: v.clone() for k, v in sd.items()}
    sd["mpnn.mpnn_layers.0.edge_dense.W_in.weight"] *= 1e6   # FFN hidden of the first edge update ~ 1e6 > 65504
    sd["mpnn.mpnn_layers.0.edge_dense.W_out.weight"] /= 1e6
    init = tt(g["in_SC_D_init"]).to(dev)
    outs = {}
    for mode in ("f16x3", "fp32"):
        m = TDiffusionModule()
        m.load_state_dict(sd)
        m.kernel_mode = mode
        m = m.to(dev).eval()
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            outs[mode] = m.sampling(b.to(dev), init_SC_D=init)
        assert (len([x for x in w if "fp16 range" in str(x.message)]) == 1) == (mode == "f16x3"), mode
    assert torch.isfinite(outs["f16x3"]).all()
    assert torch.equal(outs["f16x3"], outs["fp32"])
    # the same guard protects `network` (the PackPPI-AP feature-extractor call)
    B, L = b.X.shape[:2]
    m.kernel_mode = "f16x3"
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        s16, h16 = m.network(b.to(dev), init, torch.full((B * L,), 0.5, device=dev))
    assert any("fp16 range" in str(x.message) for x in w)
    m.kernel_mode = "fp32"
    s32, h32 = m.network(b.to(dev), init, torch.full((B * L,), 0.5, device=dev))
    assert torch.equal(s16, s32) and torch.equal(h16, h32)
