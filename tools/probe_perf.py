"""Performance probes (run on the GPU box): cluster size of the tensor-core kernels on an unpadded micro-batch, and the
proximal loop on small / large complexes.  Prints one JSON line per probe."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def ev_ms(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from packppi_b200 import TDiffusionModule, _lib, collate, get_atom14_coords, proximal_optimizer, synthetic, weights
    dev = torch.device("cuda:0")
    sd = weights.make_state_dict(0)
    what = sys.argv[1:] or ["cluster", "prox"]
    if "cluster" in what:
        L = int(os.environ.get("PROBE_L", "512"))
        items = [synthetic.make_complex((L // 2, L - L // 2), seed=100 + i) for i in range(8)]
        b = collate(items).to(dev)
        for cluster in (1, 2):
            m = TDiffusionModule()
            m.load_state_dict(sd)
            m.kernel_cluster = cluster
            m = m.to(dev).eval()
            eng, graph = m._graph(b)
            S = 8
            chi = torch.zeros(S * graph.G, 4, device=dev)
            ws = eng.workspace(graph.G, graph.K, S)
            ni = eng.node_inputs(b)
            t = torch.full((1,), 0.5, device=dev)
            _lib.PROFILE = {"pp_ipmp_edge_tc:edge": [], "pp_ipmp_edge_tc:node": []}
            for _ in range(6):
                eng.forward_layers(graph, ws, ni, chi, t, 0)
            torch.cuda.synchronize()
            prof = {k: [a.elapsed_time(c) for a, c, _ in v][-8:] for k, v in _lib.PROFILE.items()}
            _lib.PROFILE = None
            ms = ev_ms(lambda: eng.forward_layers(graph, ws, ni, chi, t, 0), 10)
            print(json.dumps({"probe": "cluster", "cluster": cluster, "rows": S * graph.G, "L": L,
                              "forward_layers_ms": ms, "edge_ms": sum(prof["pp_ipmp_edge_tc:edge"]) / 8,
                              "node_ms": sum(prof["pp_ipmp_edge_tc:node"]) / 8,
                              "residue_steps_per_s": S * graph.G / (ms * 1e-3)}), flush=True)
    if "prox" in what:
        for name, chains in (("195", (110, 85)), ("1500", (500,) * 3), ("5000", (500,) * 10)):
            b = synthetic.make_complex(chains, seed=sum(chains)).to(dev)
            b["X"] = (get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D) * b.atom_mask[..., None]).contiguous()
            ms = ev_ms(lambda: proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 50), 5)
            from packppi_b200 import compute_residue_clash
            x = b.SC_D.clone().requires_grad_(True)

            def fb():
                x.grad = None
                compute_residue_clash(b, x).sum().backward()

            print(json.dumps({"probe": "prox", "residues": name, "proximal_50_steps_ms": ms,
                              "clash_grad_ms": ev_ms(fb, 20),
                              "clash_fwd_ms": ev_ms(lambda: compute_residue_clash(b, b.SC_D), 20)}), flush=True)


if __name__ == "__main__":
    main()
