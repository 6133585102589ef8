"""Drop-in boundary test: the reference's OWN callers - `evaluate_model` (src/eval_diffusion.py:52-79) and `main` of
src/proximal_optimize.py (:26-66) - executed unmodified, once with the reference's classes on the CPU and once with
the names INTEGRATION.md tells a maintainer to import from `packppi_b200` instead, in the same process.

The reference runs from baseline/_ref/ (tools/install_reference.sh; git-ignored, shipped to the GPU box) under
tools/ref_shims.py: third-party modules that are absent here (Lightning, Hydra, Biopython ...) are stand-ins, every
line of the reference itself is the original.  MolProbity is not available, so `--molprobity_clash_loc` points at a
script that prints a fixed clashscore; everything else of `get_metric` (protein_analysis.py:36-91) runs for real.
"""
import argparse
import os
import stat
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shims.available(),
                                                  reason="reference not installed (tools/install_reference.sh)")]


@pytest.fixture(scope="module")
def ref():
    return ref_shims.import_reference()


@pytest.fixture(scope="module")
def fake_molprobity(tmp_path_factory):
    p = tmp_path_factory.mktemp("bin") / "molprobity.clashscore"
    p.write_text("#!/bin/sh\necho 'clashscore = 12.50'\n")
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(p)


def _pdb_coords(path):
    rows = [(line[12:16], line[17:20], line[21], int(line[22:26]), float(line[30:38]), float(line[38:46]),
             float(line[46:54])) for line in open(path) if line.startswith("ATOM")]
    return [r[:4] for r in rows], np.array([r[4:] for r in rows])


def test_evaluate_model_with_packppi_b200_swapped_in(ref, fake_molprobity, tmp_path, capsys):
    import src.eval_diffusion as ed
    import packppi_b200
    from packppi_b200 import weights
    sd = weights.make_state_dict(0)
    pdb = os.path.join(ref_shims.REFERENCE_ROOT, "data", "1BRS.pdb")

    # --- the reference, CPU ---------------------------------------------------------------------------
    ref_model = ref_shims.build_reference_model(ref)
    ref_model.load_state_dict(sd)
    drawn = {}
    orig_noise = ref_model.add_sc_noise

    def recording_noise(batch, t):
        out = orig_noise(batch, t)
        drawn["SC_D"] = out[0].clone()
        return out

    ref_model.add_sc_noise = recording_noise
    out_ref = tmp_path / "ref"
    args = argparse.Namespace(input=pdb, outdir=str(out_ref), molprobity_clash_loc=fake_molprobity,
                              use_proximal=True, device="cpu")
    torch.manual_seed(1)
    ed.evaluate_model(ref_model, args)
    text_ref = capsys.readouterr().out
    assert "Metric" in text_ref

    # --- packppi_b200 behind the same caller (the two imported names change) ----------
    model = packppi_b200.TDiffusionModule(encoder_cfg=ref_shims.REF_CFG["encoder_cfg"],
                                          model_cfg=ref_shims.REF_CFG["model_cfg"],
                                          sample_cfg=ref_shims.REF_CFG["sample_cfg"])
    model.load_state_dict(ref_model.state_dict())  # the reference's own state_dict, key for key
    model = model.to("cuda:0").eval()
    plain_sampling = model.sampling
    # parity hook: the reference's own noise draw is injected (north_star: "reference's noise tensors injected")
    model.sampling = lambda batch, use_proximal=False: plain_sampling(batch, use_proximal=use_proximal,
                                                                      init_SC_D=drawn["SC_D"].to("cuda:0"))
    saved = (ed.TDiffusionModule, ed.get_atom14_coords)
    ed.TDiffusionModule, ed.get_atom14_coords = packppi_b200.TDiffusionModule, packppi_b200.get_atom14_coords
    try:
        out_new = tmp_path / "new"
        args2 = argparse.Namespace(input=pdb, outdir=str(out_new), molprobity_clash_loc=fake_molprobity,
                                   use_proximal=True, device="cuda:0")
        ed.evaluate_model(model, args2)
    finally:
        ed.TDiffusionModule, ed.get_atom14_coords = saved
    text_new = capsys.readouterr().out

    ids_r, xyz_r = _pdb_coords(out_ref / "structure.pdb")
    ids_n, xyz_n = _pdb_coords(out_new / "structure.pdb")
    assert ids_r == ids_n and len(ids_r) > 1400
    # PDB files carry 3 decimals; the proximal step moves a few noise-driven angles by up to ~1e-2 rad
    d = np.abs(xyz_r - xyz_n).max(axis=1)
    assert np.quantile(d, 0.99) <= 2.5e-3 and d.max() < 0.1, (np.quantile(d, 0.99), d.max())  # 3-decimal PDB fields

    def metrics(text):
        line = [ln for ln in text.splitlines() if "Metric" in ln][0]
        import re
        return {k: float(v) for k, v in re.findall(r"'(\w+)': (?:tensor\()?([-0-9.e]+)", line)}

    m_r, m_n = metrics(text_ref), metrics(text_new)
    assert set(m_r) == set(m_n) and "atom_rmsd" in m_r and "interface_acc" in m_r
    for k in m_r:  # end metrics within 1 % (BASELINE.json north_star)
        assert abs(m_r[k] - m_n[k]) <= 0.01 * max(abs(m_r[k]), 1e-3), (k, m_r[k], m_n[k])


def test_proximal_optimize_main_with_packppi_b200_swapped_in(ref, fake_molprobity, tmp_path, capsys):
    """src/proximal_optimize.py main(), unmodified.  The script keeps its batch on the host (it has no --device flag),
    so the swapped-in functions run under `packppi_b200.host_staging`."""
    import src.proximal_optimize as pm
    import packppi_b200
    pdb = os.path.join(ref_shims.REFERENCE_ROOT, "data", "1BRS.pdb")
    mk = lambda d: argparse.Namespace(input=pdb, outdir=str(tmp_path / d), molprobity_clash_loc=fake_molprobity,  # noqa: E731
                                      violation_tolerance_factor=12.0, clash_overlap_tolerance=0.5, lamda=1.0,
                                      num_steps=50)
    torch.manual_seed(0)
    pm.main(mk("ref"))
    saved = (pm.proximal_optimizer, pm.get_atom14_coords)
    pm.proximal_optimizer, pm.get_atom14_coords = packppi_b200.proximal_optimizer, packppi_b200.get_atom14_coords
    try:
        with packppi_b200.host_staging("cuda:0"):
            pm.main(mk("new"))
    finally:
        pm.proximal_optimizer, pm.get_atom14_coords = saved
    capsys.readouterr()
    ids_r, xyz_r = _pdb_coords(tmp_path / "ref" / "structure.pdb")
    ids_n, xyz_n = _pdb_coords(tmp_path / "new" / "structure.pdb")
    assert ids_r == ids_n
    d = np.abs(xyz_r - xyz_n).max(axis=1)
    assert np.quantile(d, 0.99) <= 2.5e-3 and d.max() < 0.1, (np.quantile(d, 0.99), d.max())  # 3-decimal PDB fields
    # the optimisation did something: the written structure differs from the input's side chains
    _, xyz_in = _pdb_coords(pdb)
    assert len(xyz_in) != len(xyz_r) or np.abs(xyz_in - xyz_r).max() > 0.05
