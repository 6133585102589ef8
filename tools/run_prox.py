"""PackPPI-Prox workloads for ncu launch lists / captures: a 195-residue and a 5000-residue complex (single item,
eager launches: run with PACKPPI_B200_GRAPH_ROWS=0) and one batched micro-batch (8 complexes x 8 decoys)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from packppi_b200 import collate, get_atom14_coords, proximal_optimizer, synthetic  # noqa: E402

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for chains in ((110, 85), (500,) * 10):
    b = synthetic.make_complex(chains, seed=sum(chains)).to(dev)
    b["X"] = (get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D) * b.atom_mask[..., None]).contiguous()
    for _ in range(2):
        proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, steps)
batch = collate([synthetic.make_complex((200 + 30 * i, 180 + 20 * i), seed=300 + i) for i in range(8)]).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
chi = ((torch.rand(8, *batch.SC_D.shape, device=dev, generator=gen) * 2 - 1) * math.pi) * batch.SC_D_mask
for _ in range(2):
    proximal_optimizer(batch, chi, 12.0, 0.5, 1.0, steps)
torch.cuda.synchronize()
print("ok")
