"""Clash loss + analytic gradient at 5000 residues, a few repetitions (workload for an ncu capture of the clash kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from packppi_b200 import compute_residue_clash, synthetic  # noqa: E402

dev = torch.device("cuda:0")
b = synthetic.make_complex((500,) * 10, seed=5000).to(dev)
for _ in range(6):
    x = b.SC_D.clone().requires_grad_(True)
    compute_residue_clash(b, x).sum().backward()
torch.cuda.synchronize()
print("ok")
