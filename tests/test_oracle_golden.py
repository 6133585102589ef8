"""Pins the CPU oracle (oracle/) against vectors produced by the unmodified reference (tools/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import msc_oracle as mo
from oracle import prox_oracle as po
from packppi_b200 import weights

from util import ALL_CASES, knn_mismatches, load_golden, tt, wrapped_diff

SD = weights.make_state_dict(0)
FAST = ["syn5", "syn17", "syn31", "syn33", "syn64", "synbatch", "1brs"]


@pytest.mark.parametrize("case", ALL_CASES)
def test_knn_graph(case):
    g, b = load_golden(case)
    D, E = mo.knn_graph(b.X[:, :, 1, :], b.residue_mask)
    B, L, K = E.shape
    valid = (b.residue_mask > 0).reshape(-1).numpy()
    bad = knn_mismatches(E.reshape(B * L, K), D.reshape(B * L, K), g["ref_E_idx"].reshape(B * L, K),
                         g["ref_D_neighbors"].reshape(B * L, K), valid)
    assert not bad, f"{case}: rows {bad[:5]}"


@pytest.mark.parametrize("case", FAST + ["syn300"])
def test_network_probe(case):
    g, b = load_golden(case)
    B, L = b.X.shape[:2]
    rows = g["in_rows"]
    with torch.no_grad():
        cache = mo.GraphCache(SD, b)
        # edges to masked / padded neighbours are exact ties in the kNN and carry arbitrary indices; every consumer
        # multiplies them by mask_attend, so they are compared after masking
        att = cache.mask_att[:, rows][..., None]
        assert ((cache.h_E0[:, rows] - tt(g["ref_probe_hE0_rows"])) * att).abs().max() < 2e-5
        score, hV, layers = mo.network(SD, b, tt(g["in_probe_SC_D"]), torch.full((B * L,), 0.7), cache,
                                       return_layers=True)
    assert (layers[0] - tt(g["ref_probe_hV0"])).abs().max() < 2e-5
    for li in range(3):
        assert (layers[li + 1] - tt(g[f"ref_probe_hV_l{li}"])).abs().max() < 5e-5, li
    assert (hV - tt(g["ref_probe_hV"])).abs().max() < 5e-5
    assert (score - tt(g["ref_probe_score"])).abs().max() < 5e-5


@pytest.mark.parametrize("case", FAST)
def test_sampling_trajectory(case):
    g, b = load_golden(case)
    x, traj = mo.sampling(SD, b, tt(g["in_SC_D_init"]), trajectory=True)
    for i, s in enumerate(g["in_traj_steps"]):
        assert wrapped_diff(traj[int(s)], tt(g["ref_traj"][i])).max() < 1e-4, (case, int(s))
    assert wrapped_diff(x, tt(g["ref_SC_D_final"])).max() < 1e-4


def test_initial_noise_formula():
    g, b = load_golden("1brs")
    torch.manual_seed(1)
    x = b.SC_D.reshape(-1, 4)
    e1 = torch.randn_like(x)
    e2 = torch.randn_like(x)
    assert wrapped_diff(mo.initial_noise(b, e1, e2), tt(g["in_SC_D_init"])).max() < 1e-6


@pytest.mark.parametrize("case", ALL_CASES)
def test_atom14(case):
    g, b = load_golden(case)
    for key, chi in (("ref_atom14_final", tt(g["ref_SC_D_final"])), ("ref_atom14_native", b.SC_D)):
        pos = po.atom14_coords(b.X, b.residue_type, b.BB_D, chi)
        assert (pos - tt(g[key])).abs().max() < 1e-4, (case, key)


@pytest.mark.parametrize("case", FAST + ["syn300"])
@pytest.mark.parametrize("sparse", [False, True])
def test_clash_loss_and_grad(case, sparse):
    g, b = load_golden(case)
    chi = tt(g["ref_SC_D_final"])
    for bi in range(b.X.shape[0]):
        sub = {k: (v[bi:bi + 1] if torch.is_tensor(v) else v) for k, v in b.items()}
        pr, gr = po.clash_value_and_grad(sub, chi[bi:bi + 1], sparse=sparse)
        ref_pr, ref_gr = tt(g["ref_clash_per_res"][bi:bi + 1]), tt(g["ref_clash_grad"][bi:bi + 1])
        assert (pr - ref_pr).abs().max() < 1e-5 * max(1.0, float(ref_pr.abs().max()))
        assert (gr - ref_gr).abs().max() < 1e-5 * max(1.0, float(ref_gr.abs().max()))


@pytest.mark.slow
def test_clash_t1124_sparse():
    g, b = load_golden("t1124")
    pr, gr = po.clash_value_and_grad(b, tt(g["ref_SC_D_final"]), sparse=True)
    assert (pr - tt(g["ref_clash_per_res"])).abs().max() < 2e-5
    assert (gr - tt(g["ref_clash_grad"])).abs().max() < 2e-5


@pytest.mark.parametrize("case", ["syn17", "syn64", "1brs"])
def test_proximal(case):
    g, b = load_golden(case)
    n = len(g["ref_prox_losses"])
    snaps, losses, mask = po.proximal(b, tt(g["in_prox_start"]), 12.0, 0.5, 1.0, n, sparse=(case == "1brs"))
    assert torch.equal(mask, tt(g["ref_prox_mask"]))
    np.testing.assert_allclose(np.asarray(losses), g["ref_prox_losses"], rtol=2e-4)
    for i, k in enumerate(g["in_prox_keep"]):
        assert wrapped_diff(snaps[int(k)], tt(g["ref_prox_snaps"][i])).max() < 1e-4


def test_sde_sampling_with_injected_noise():
    """mode "sde" (schedule.py:224-228): the reference's own torch.normal draws are replayed."""
    g, b = load_golden("1brs_sde")
    x = mo.sampling(SD, b, tt(g["in_SC_D_init"]), sde_noise=tt(g["in_sde_noise"]))
    assert wrapped_diff(x, tt(g["ref_SC_D_final"])).max() < 1e-4
