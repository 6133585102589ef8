"""Accuracy of the execution modes against the CPU oracle on fresh random inputs (run on the GPU box):
max chi difference after 2 and after 30 reverse-ODE steps, several seeds."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def main():
    from oracle import msc_oracle as mo
    from oracle import prox_oracle as po
    from packppi_b200 import TDiffusionModule, synthetic, weights
    dev = torch.device("cuda:0")
    sd = weights.make_state_dict(0)
    configs = (("fp32", "ffma"), ("f16x3", "ffma"), ("f16x3", "tc32"))
    models = {}
    for mode, ne in configs:
        m = TDiffusionModule()
        m.load_state_dict(sd)
        m.kernel_mode, m.kernel_node_epilogue = mode, ne
        models[(mode, ne)] = m.to(dev).eval()
    worst = {c: [0.0, 0.0] for c in configs}
    total_degenerate = 0
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        scale = int(os.environ.get("PP_DIAG_SCALE", "1"))  # PP_DIAG_SCALE=4: complexes of 256-340 residues
        b = synthetic.make_complex((scale * (40 + 3 * seed), scale * (24 + seed)), seed=seed,
                                   place_side_chains=po.atom14_coords)
        L = b.X.shape[1]
        g = torch.Generator().manual_seed(100 + seed)
        x0 = ((torch.rand(1, L, 4, generator=g) * 2 - 1) * 3.14159) * b.SC_D_mask
        refs = {n: mo.sampling(sd, b, x0, n_steps=n) for n in (2, 30)}
        # The inter-residue dihedral features are raw signed angles: at near-planar geometry (|cos| rounding past 1
        # -> NaN -> 0 in the reference, or an angle at +-pi) one ulp moves the feature by pi or 2 pi.  The kernel
        # repeats the reference's rounding sequence, so the count of such edges is expected to be 0; it is COUNTED
        # here, and every complex is scored.
        E_idx = mo.knn_graph(b.X[:, :, 1, :], b.residue_mask)[1]
        hE_ref = mo.edge_embedding(sd, b, E_idx)
        _, graph0 = models[configs[0]]._graph(b.to(dev))
        dE = (graph0.hE0.cpu().reshape(hE_ref.shape) - hE_ref).abs().amax(-1)[0]
        self_edge = E_idx[0] == torch.arange(L).view(-1, 1)
        degenerate = int(((dE > 1e-2) & ~self_edge).sum())
        bd = b.to(dev)
        line = [f"seed {seed} L {L:3d}"]
        for c in configs:
            eng, graph = models[c]._graph(bd)
            for i, n in enumerate((2, 30)):
                chi = eng.sample(graph, bd, x0.reshape(-1, 4).to(dev), n_steps=n).cpu().reshape(1, L, 4)
                d = (chi - refs[n]).abs()
                d = torch.minimum(d, 2 * 3.141592653589793 - d).max().item()
                worst[c][i] = max(worst[c][i], d)
                line.append(f"{c[0]}/{c[1]} n={n}: {d:.1e}")
        if degenerate:
            line.append(f"[{degenerate} edge(s) with |h_E0 - oracle| > 1e-2]")
        total_degenerate += degenerate
        print("  ".join(line), flush=True)
    for c in configs:
        print(f"worst {c[0]}/{c[1]}: 2 steps {worst[c][0]:.2e}, 30 steps {worst[c][1]:.2e}")
    print(f"edges off by more than 1e-2 in h_E0 (degenerate dihedral features): {total_degenerate}; unscored complexes: 0")


if __name__ == "__main__":
    main()
