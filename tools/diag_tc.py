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


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "f16x3", "f16"]
    if not os.environ.get("PP_DIAG_TRACE_ONLY"):
        accuracy(dev, modes)
    for mode in modes:
        if mode != "fp32":
            print("trace", mode)
            trace(dev, mode)
