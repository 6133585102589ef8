"""Every kernel of libpackppi_b200.so launched once or twice on a medium input (1100 residues: the cell-list paths;
a ragged batch of 4 complexes x 2 samples): the workload of the per-kernel
`ncu --set full` capture in profiles/ (one reverse-ODE step instead of 30, so that the capture stays short)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from packppi_b200 import (TDiffusionModule, collate, compute_residue_clash, featurize, get_atom14_coords,  # noqa: E402
                          proximal_optimizer, synthetic, weights)

dev = torch.device("cuda:0")
sd = weights.make_state_dict(0)
big = synthetic.make_complex((600, 500), seed=1100)
ragged = collate([synthetic.make_complex((120 + 30 * i, 100 + 20 * i), seed=40 + i) for i in range(4)])
for mode, work in (("f16x3", ((big, 1), (ragged, 2))), ("fp32", ((ragged, 2),))):
    m = TDiffusionModule()
    m.load_state_dict(sd)
    m.kernel_mode = mode
    m = m.to(dev).eval()
    for b, S in work:
        bd = b.to(dev)
        eng, graph = m._graph(bd)                       # kNN (scan or cells), geometry, edge embedding
        chi = ((torch.rand(S * graph.G, 4, device=dev) * 2 - 1) * math.pi) * bd.SC_D_mask.reshape(-1, 4).repeat(S, 1)
        eng.sample(graph, bd, chi, n_steps=1)           # node embed, 3 IPMP layers, decoder + ODE step
bd = big.to(dev)
bd["X"] = (get_atom14_coords(bd.X, bd.residue_type, bd.BB_D, bd.SC_D) * bd.atom_mask[..., None]).contiguous()
x = bd.SC_D.clone().requires_grad_(True)
compute_residue_clash(bd, x).sum().backward()           # atom14, neighbour list (cells), pair kernel <0> and <1>
proximal_optimizer(bd, bd.SC_D, 12.0, 0.5, 1.0, 1)      # prox init / step / loss
rb = ragged.to(dev)
chi = ((torch.rand(2, *rb.SC_D.shape, device=dev) * 2 - 1) * math.pi) * rb.SC_D_mask
proximal_optimizer(rb, chi, 12.0, 0.5, 1.0, 1)          # batched items (scan neighbour list)
prots = []
for b in (big, synthetic.make_complex((300, 200), seed=3)):
    X = b.X[0].clone()
    X[b.atom_mask[0] == 0] = float("nan")
    prots.append(dict(atom_positions=X.numpy(), aaindex=b.residue_type[0].numpy(), atom_mask=b.atom_mask[0].numpy(),
                      residue_index=b.residue_index[0].numpy(), chain_id=[str(int(c)) for c in b.chain_indices[0]]))
featurize.proteins_to_batch_device(prots, dev)
torch.cuda.synchronize()
print("ok")
