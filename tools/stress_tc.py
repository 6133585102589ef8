"""Determinism stress test of the tensor-core kernels (run on the GPU box): the same network evaluation repeated N
times must give bit-identical h_V (all reductions have a fixed order); a mismatch exposes a synchronisation bug."""
import os
import sys

import torch

sys.path.insert(0, os.environ.get("PP_STRESS_ROOT") or os.path.join(os.path.dirname(__file__), ".."))


def main():
    from packppi_b200 import TDiffusionModule, weights, synthetic
    from packppi_b200.batch import collate
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    cases = {"1x1500": [synthetic.make_complex((500, 500, 500), seed=3)],
             "1x17": [synthetic.make_complex((9, 8), seed=4)],
             "3 padded": [synthetic.make_complex((20, 21), seed=1), synthetic.make_complex((90, 83), seed=2),
                          synthetic.make_complex((30, 30), seed=3)],
             "8x~500": [synthetic.make_complex(c, seed=10 + i) for i, c in enumerate(synthetic.sweep_lengths(8, seed=64))]}
    for mode in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["f16x3", "f16"]):
        m = TDiffusionModule()
        m.load_state_dict(weights.make_state_dict(0))
        m.kernel_mode = mode
        m = m.to(dev).eval()
        for name, items in cases.items():
            b = collate(items).to(dev)
            B, L = b.X.shape[:2]
            S = 2 if name != "8x~500" else 8
            x = ((torch.rand(S, B, L, 4, generator=torch.Generator().manual_seed(1)) * 2 - 1) * 3.14).to(dev)
            eng, graph = m._graph(b)
            t = torch.full((S * B * L,), 0.4, device=dev)
            ref = None
            bad = 0
            worst = 0.0
            for i in range(n):
                score, hV = eng.network(graph, b, x.reshape(-1, 4).contiguous(), t)
                hV = hV.clone()
                if ref is None:
                    ref = hV
                elif not torch.equal(ref, hV):
                    bad += 1
                    worst = max(worst, (ref - hV).abs().max().item())
            torch.cuda.synchronize()
            print(f"{mode:6s} {name:9s} rows {S * B * L:6d}: {bad} of {n - 1} repeats differ (max |diff| {worst:.2e})", flush=True)


if __name__ == "__main__" and not os.environ.get("PP_STRESS_KERNELS"):
    main()


def kernels(n=100, passes=3):
    """Each tensor-core kernel alone, repeated on fixed inputs (layer 0 of a 2 x 1500-residue problem)."""
    from packppi_b200 import TDiffusionModule, weights, synthetic, _lib
    from packppi_b200.batch import collate
    dev = torch.device("cuda:0")
    m = TDiffusionModule()
    m.load_state_dict(weights.make_state_dict(0))
    m.kernel_mode = "fp32"
    m = m.to(dev).eval()
    b = collate([synthetic.make_complex((500, 500, 500), seed=3)]).to(dev)
    B, L = b.X.shape[:2]
    S = 2
    x = ((torch.rand(S, B, L, 4, generator=torch.Generator().manual_seed(1)) * 2 - 1) * 3.14).to(dev)
    eng, graph = m._graph(b)
    G, K = graph.G, graph.K
    t = torch.full((S * G,), 0.4, device=dev)
    eng.network(graph, b, x.reshape(-1, 4).contiguous(), t)   # fills wsA / wsN / wsP / hV / hE with layer-2 state
    ws = eng.workspace(G, K, S)
    W, wtc = eng.wblob, eng.wtc
    common = (graph.geo, graph.nbr, graph.mask_attend, graph.msum)
    torch.cuda.synchronize()

    def rep(name, fn, out):
        ref, bad, worst = None, 0, 0.0
        for i in range(n):
            fn()
            o = out().clone()
            if ref is None:
                ref = o
            elif not torch.equal(ref, o):
                bad += 1
                worst = max(worst, (ref - o).abs().max().item())
        torch.cuda.synchronize()
        print(f"kernel {name:28s}: {bad} of {n - 1} repeats differ (max |diff| {worst:.2e})", flush=True)

    acc = torch.zeros_like(ws.wsAcc)
    rep("node message, shared h_E0", lambda: _lib.call("pp_ipmp_edge_tc", W, 0, 0, wtc[0, 0], *common, G, K, S, graph.hE0, 1,
                                                       ws.wsA, ws.wsN, ws.wsP, acc, passes, 1, None, None, None), lambda: acc)
    hin = ws.hE.clone()
    rep("node message, per-sample h_E", lambda: _lib.call("pp_ipmp_edge_tc", W, 1, 0, wtc[1, 0], *common, G, K, S, hin, 0,
                                                          ws.wsA, ws.wsN, ws.wsP, acc, passes, 1, None, None, None), lambda: acc)
    hout = torch.zeros_like(ws.hE)
    rep("edge update, shared h_E0", lambda: _lib.call("pp_ipmp_edge_tc", W, 0, 1, wtc[0, 1], *common, G, K, S, graph.hE0, 1,
                                                      ws.wsA, ws.wsN, ws.wsP, hout, passes, 1, None, None, None), lambda: hout)
    rep("edge update, separate out", lambda: _lib.call("pp_ipmp_edge_tc", W, 1, 1, wtc[1, 1], *common, G, K, S, hin, 0,
                                                       ws.wsA, ws.wsN, ws.wsP, hout, passes, 1, None, None, None), lambda: hout)
    work = hin.clone()

    def inplace():
        work.copy_(hin)
        _lib.call("pp_ipmp_edge_tc", W, 1, 1, wtc[1, 1], *common, G, K, S, work, 0, ws.wsA, ws.wsN, ws.wsP, work, passes, 1, None, None, None)
    rep("edge update, in place", inplace, lambda: work)
    hv0 = ws.hV.clone()
    hv = hv0.clone()

    def post():
        hv.copy_(hv0)
        _lib.call("pp_ipmp_node_post_tc32", W, 0, wtc[0, 2], graph.msum, graph.mask, G, K, S, ws.wsAcc, hv, None, None, None)
    rep("node epilogue", post, lambda: hv)


if __name__ == "__main__" and os.environ.get("PP_STRESS_KERNELS"):
    kernels()
