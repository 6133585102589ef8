"""tests/golden/score_target.npz: the reference's `TDiffusionModule.add_sc_noise` (TorsionalDiffusion.py:111-124) on the
1BRS inputs at three diffusion times, with the two randn draws it made recorded as inputs.  Run in the build container
(needs /root/reference): python tools/make_golden_score.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402
from packppi_b200.batch import TENSOR_FIELDS  # noqa: E402


def main():
    ref = ref_shims.import_reference()
    model = ref_shims.build_reference_model(ref)
    with np.load(os.path.join(ROOT, "tests", "golden", "1brs.npz")) as z:
        fields = {k: torch.from_numpy(z["in_" + k]) for k in TENSOR_FIELDS}
    L = fields["X"].shape[1]
    batch = ref.Data(**fields, num_proteins=1, max_size=L)
    out = {}
    for i, tv in enumerate((1.0, 0.37, 0.02)):
        t = torch.full((L,), tv)
        torch.manual_seed(40 + i)
        noised, score = model.add_sc_noise(batch, t)
        torch.manual_seed(40 + i)  # the same two draws, in the order add_noise makes them (schedule.py:186)
        e1 = torch.randn(L, 4)
        e2 = torch.randn(L, 4)
        out[f"in_t_{i}"], out[f"in_eps1_{i}"], out[f"in_eps2_{i}"] = t.numpy(), e1.numpy(), e2.numpy()
        out[f"ref_noised_{i}"], out[f"ref_score_{i}"] = noised.numpy(), score.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "score_target.npz"), **out)
    print("wrote score_target.npz", {k: v.shape for k, v in out.items() if k.endswith("_0")})


if __name__ == "__main__":
    main()
