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
@pytest.mark.parametrize("mode,cluster", [("f16x3", 1), ("f16x3", 2), ("f16x3", 4), ("f16", 1), ("f16", 4)])
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
@pytest.mark.parametrize("mode,cluster", [("f16x3", 1), ("f16x3", 4), ("f16", 2)])
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
