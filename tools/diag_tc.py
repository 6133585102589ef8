"""Diagnostics for the tensor-core kernels (run on the GPU box): chi / activation error per execution mode against the
golden fixtures, and the clock64 stage trace of one tile (pp_set_tc_trace)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from util import load_golden, tt, wrapped_diff  # noqa: E402


def model(dev, mode):
    from packppi_b200 import TDiffusionModule, weights
    m = TDiffusionModule()
    m.load_state_dict(weights.make_state_dict(0))
    m.kernel_mode = mode
    return m.to(dev).eval()


def accuracy(dev, modes):
    for case in ("syn64", "syn300", "1brs", "t1124"):
        g, b = load_golden(case)
        bd = b.to(dev)
        B, L = b.X.shape[:2]
        for mode in modes:
            m = model(dev, mode)
            x = tt(g["in_probe_SC_D"]).to(dev)
            score, hV = m.network(bd, x, torch.full((B * L,), 0.7, device=dev))
            e_hv = (hV.cpu() - tt(g["ref_probe_hV"])).abs().max().item()
            out = m.sampling(bd, init_SC_D=tt(g["in_SC_D_init"]).to(dev))
            d = wrapped_diff(out.cpu(), tt(g["ref_SC_D_final"]))
            print(f"{case:8s} {mode:7s} hV err {e_hv:.2e}  chi max {d.max().item():.2e} mean {d.mean().item():.2e}")


def trace(dev, mode):
    from packppi_b200 import _lib
    g, b = load_golden("t1124")
    bd = b.to(dev)
    B, L = b.X.shape[:2]
    m = model(dev, mode)
    x = tt(g["in_probe_SC_D"]).to(dev)
    buf = torch.zeros(64, dtype=torch.int64, device=dev)
    m.network(bd, x, torch.full((B * L,), 0.7, device=dev))
    _lib.load().pp_set_tc_trace(ctypes.c_void_p(buf.data_ptr()))
    m.network(bd, x, torch.full((B * L,), 0.7, device=dev))
    torch.cuda.synchronize()
    _lib.load().pp_set_tc_trace(None)
    t = buf.cpu().tolist()
    names = ["start", "first operand", "G1 done", "x1 published", "G2 done", "x2 published", "G3 done", "e in TMEM"]
    for j in range(4):
        names += [f"FFN-in {j} done", f"hidden {j} published"]
    names += ["next first operand", "FFN-out done", "stored"]
    prev = t[0]
    for i, n in enumerate(names):
        if t[i] == 0:
            break
        print(f"  {i:2d} {n:22s} +{t[i] - prev:7d}  (total {t[i] - t[0]:7d})")
        prev = t[i]


def trace_big(dev, mode, path, it):
    """Trace of tile `it` of CTA 0 on a bench-like micro-batch (8 complexes of ~500 residues, 8 samples)."""
    from packppi_b200 import _lib, synthetic
    from packppi_b200.batch import collate
    items = [synthetic.make_complex(c, seed=10 + i) for i, c in enumerate(synthetic.sweep_lengths(8, seed=64))]
    b = collate(items).to(dev)
    B, L = b.X.shape[:2]
    S = 8
    m = model(dev, mode)
    eng, graph = m._graph(b)
    x = ((torch.rand(S, B, L, 4, generator=torch.Generator().manual_seed(1)) * 2 - 1) * 3.14).to(dev).reshape(-1, 4)
    t = torch.full((S * B * L,), 0.4, device=dev)
    buf = torch.zeros(64, dtype=torch.int64, device=dev)
    eng.network(graph, b, x, t)
    lib = _lib.load()
    lib.pp_set_tc_trace_tile(ctypes.c_int64(path), ctypes.c_int64(it))
    layer_sel = int(os.environ.get("PP_DIAG_LAYER", "-1"))  # trace only this layer's launch (-1: the last one wins)
    orig_call = _lib.call

    def traced_call(name, *args, **kw):
        if name == "pp_ipmp_edge_tc":
            on = layer_sel < 0 or args[1] == layer_sel
            lib.pp_set_tc_trace(ctypes.c_void_p(buf.data_ptr()) if on else None)
        return orig_call(name, *args, **kw)

    _lib.call = traced_call
    import packppi_b200.engine as _eng
    _eng._lib.call = traced_call
    eng.network(graph, b, x, t)
    torch.cuda.synchronize()
    _lib.call = orig_call
    lib.pp_set_tc_trace(None)
    lib.pp_set_tc_trace_tile(ctypes.c_int64(1), ctypes.c_int64(0))
    ts = [v for v in buf.cpu().tolist() if v]
    if path == 0:
        names = ["tile start", "G1 done", "x1 published", "next first operand", "G2 done", "sums written"]
    else:
        names = ["tile start", "G1 done", "x1 published", "G2 done", "x2 published", "G3 done", "e in TMEM"]
        for j in range(4):
            names += [f"FFN-in {j} done", f"hidden {j} published"]
        names += ["next first operand", "FFN-out done", "stored"]
    print(f"bench-like trace, mode {mode}, {'node message' if path == 0 else 'edge update'}, tile {it} of CTA 0")
    for i in range(1, min(len(ts), len(names))):
        print(f"  {names[i]:22s} +{ts[i] - ts[i - 1]:7d}  (total {ts[i] - ts[0]:7d})")


if __name__ == "__main__" and os.environ.get("PP_DIAG_BIG"):
    for p_, it_ in ((1, 5), (1, 20), (0, 5), (0, 20)):
        trace_big(torch.device("cuda:0"), "f16x3", p_, it_)
    sys.exit(0)

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "f16x3", "f16"]
    if not os.environ.get("PP_DIAG_TRACE_ONLY"):
        accuracy(dev, modes)
    for mode in modes:
        if mode != "fp32":
            print("trace", mode)
            trace(dev, mode)
