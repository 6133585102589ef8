"""tests/golden/pdb_text.json + metric_block.npz: outputs of the UNMODIFIED reference for the file format and the
metric block either side of the path, generated in the build container (needs /root/reference):
  * `to_pdb` (src/utils/protein.py:207-314) on the parsed 1BRS and T1124 records: sha256 of the text, first / last lines
  * `get_interface_mask` (helper.py:104-129) of both files
  * `ProteinAnalysis.get_metric` (protein_analysis.py:36-91) for 1BRS against a copy whose side chains were rebuilt
    from perturbed angles (MolProbity replaced by a script that prints a fixed clashscore)
Biopython is absent: the reference's PDBParser / NeighborSearch calls run on tools/mini_biopdb.py."""
import hashlib
import json
import os
import stat
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402


def main():
    ref = ref_shims.import_reference()
    from pathlib import Path

    import src.utils.protein as rp
    import src.utils.protein_analysis as rpa
    out_json, out_npz = {}, {}
    tmp = tempfile.mkdtemp()
    fake = os.path.join(tmp, "molprobity.clashscore")
    with open(fake, "w") as f:
        f.write("#!/bin/sh\necho 'clashscore = 7.25'\n")
    os.chmod(fake, os.stat(fake).st_mode | stat.S_IEXEC)
    pa = rpa.ProteinAnalysis(fake, tmp)
    for name, fn in (("1brs", "1BRS.pdb"), ("t1124", "T1124_lig.pdb")):
        path = os.path.join(ref_shims.REFERENCE_ROOT, "data", fn)
        prot = vars(rp.from_pdb_file(Path(path), mse_to_met=True))
        text = rp.to_pdb(prot)
        lines = text.split("\n")
        out_json[name] = {"sha256": hashlib.sha256(text.encode()).hexdigest(), "n_lines": len(lines),
                          "head": lines[:4], "tail": lines[-5:],
                          "ter": [ln for ln in lines if ln.startswith("TER")]}
        data = pa.get_prot(path, get_interface=True)
        out_npz[f"{name}_interface_mask"] = data.interface_mask[0].numpy()
        if name == "1brs":
            # predicted structure: same backbone, side chains rebuilt from perturbed angles
            g = torch.Generator().manual_seed(3)
            chi = (data.SC_D + 0.4 * torch.randn(data.SC_D.shape, generator=g)) * data.SC_D_mask
            xyz = ref.comp.get_atom14_coords(data.X, data.residue_type, data.BB_D, chi)
            pred = dict(prot)
            pred["atom_positions"] = xyz[0].numpy()
            pred_path = os.path.join(tmp, "pred.pdb")
            with open(pred_path, "w") as f:
                f.write(rp.to_pdb(pred))
            metric = pa.get_metric(true_pdb=path, pred_pdb=pred_path)
            out_npz["1brs_pred_atom_positions"] = pred["atom_positions"]
            for k, v in metric.items():
                out_npz[f"1brs_metric_{k}"] = np.float64(float(v))
    with open(os.path.join(ROOT, "tests", "golden", "pdb_text.json"), "w") as f:
        json.dump(out_json, f, indent=1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metric_block.npz"), **out_npz)
    print({k: v["sha256"][:12] for k, v in out_json.items()}, sorted(out_npz)[:4])


if __name__ == "__main__":
    main()
