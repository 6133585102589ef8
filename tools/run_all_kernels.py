"""Launch every kernel family of libpackppi_b200.so a few times on small-to-medium inputs: the workload for the
per-kernel `ncu --set full` evidence (profiles/) and for compute-sanitizer.  Argument: "small" (sanitizer-sized) or
"medium" (default; one 1100-residue complex through the cell-list paths + a ragged batch x 4 samples)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from packppi_b200 import (TDiffusionModule, collate, compute_residue_clash, featurize, get_atom14_coords,  # noqa: E402
                          proximal_optimizer, synthetic, weights)

size = sys.argv[1] if len(sys.argv) > 1 else "medium"
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["f16x3"]
dev = torch.device("cuda:0")
sd = weights.make_state_dict(0)
if size == "small":
    big = synthetic.make_complex((20, 13), seed=33)               # 33 residues (one residue over a tile boundary)
    ragged = collate([synthetic.make_complex((6, 5), seed=1), synthetic.make_complex((25, 15), seed=5),
                      synthetic.make_complex((4, 3), seed=2)])
    os.environ.setdefault("PACKPPI_B200_CELL_LIST_MIN_L", "16")  # exercise the cell-list kernels at this size too
else:
    big = synthetic.make_complex((600, 500), seed=1100)          # >= 1024 residues: cell-list kNN and clash lists
    ragged = collate([synthetic.make_complex((200 + 40 * i, 150 + 30 * i), seed=40 + i) for i in range(4)])
import packppi_b200.engine as engine  # noqa: E402
engine.CELL_LIST_MIN_L = int(os.environ.get("PACKPPI_B200_CELL_LIST_MIN_L", engine.CELL_LIST_MIN_L))

for mode in modes:
    m = TDiffusionModule()
    m.load_state_dict(sd)
    m.kernel_mode = mode
    m = m.to(dev).eval()
    for b, S in ((big, 1), (ragged, 4)):
        bd = b.to(dev)
        out = m.sampling(bd, n_samples=S)                        # graph, edge embed, 30 x (node embed, 3 layers, decode)
        assert torch.isfinite(out).all()
    bd = big.to(dev)
    bd["X"] = (get_atom14_coords(bd.X, bd.residue_type, bd.BB_D, bd.SC_D) * bd.atom_mask[..., None]).contiguous()
    x = bd.SC_D.clone().requires_grad_(True)
    compute_residue_clash(bd, x).sum().backward()               # atom14, neighbour list, pair kernel <0> and <1>
    proximal_optimizer(bd, bd.SC_D, 12.0, 0.5, 1.0, 5)           # prox init / step / loss
    rb = ragged.to(dev)
    chi = ((torch.rand(2, *rb.SC_D.shape, device=dev) * 2 - 1) * math.pi) * rb.SC_D_mask
    proximal_optimizer(rb, chi, 12.0, 0.5, 1.0, 5)               # batched items
# device featurisation from raw records
prots = []
for b in (big, synthetic.make_complex((30, 20), seed=3)):
    X = b.X[0].clone()
    X[b.atom_mask[0] == 0] = float("nan")
    prots.append(dict(atom_positions=X.numpy(), aaindex=b.residue_type[0].numpy(), atom_mask=b.atom_mask[0].numpy(),
                      residue_index=b.residue_index[0].numpy(), chain_id=[str(int(c)) for c in b.chain_indices[0]]))
featurize.proteins_to_batch_device(prots, dev)
torch.cuda.synchronize()
print("ok", size, modes)
