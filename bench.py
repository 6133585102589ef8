"""Benchmark of the PackPPI-MSC reverse-diffusion sampling path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl packppi_b200|reference] [--complexes C]

Workload (BASELINE.json configs[4]): a sweep of C = 64 synthetic two-chain complexes with
L ~ U{200..800} residues (seed 64), 8 diffusion samples each, random-init weights (seed 0).  One bench "step" = one
full pass over the sweep: for every complex the kNN graph, the edge embedding and 30 reverse-ODE steps of all 8
samples.  The sweep is processed in micro-batches of 8 complexes (sorted by length, padded to the longest of the
micro-batch).  Metric: residue.denoise-steps/s = sum(valid residues) x 8 samples x 30 / seconds.
With N > 1 every rank runs its own sweep (same lengths, different coordinates): weak scaling, no data-path
collective; the only NCCL call is the all_gather of the sampled angles at the end of each step.

`value`  : inputs resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`    : same pass through the public API (TDiffusionModule.sampling) from pinned HOST batches, including the
           host->device copy of every micro-batch and the device->host copy of the sampled angles.
`roofline`: dominant kernel (edge_edge_kernel, the per-edge message MLP + FFN), timed with CUDA events around each
           of its launches in one instrumented pass.
`cpu_baseline` / `--impl reference`: the CPU oracle port (oracle/msc_oracle.py, dense like the reference, graph
           rebuilt every step like the reference) on a bounded sample of the same sweep, all host threads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SAMPLES = 8
N_ODE = 30
MICRO = 8
# executed FLOPs of one edge_edge_kernel launch per residue row (K = 32 edges): first Linear 168 wide (h_E + pair
# geometry; the h_V parts are hoisted per residue), two 128x128 Linears, FFN 128-512-128   
EDGE_KERNEL_FLOP_PER_RES = 32 * 2 * (168 * 128 + 128 * 128 + 128 * 128 + 2 * 128 * 512)
# algorithmic HBM bytes of the same launch per residue row: read h_E (K*128*4), write h_E (K*128*4), A, N gathers
EDGE_KERNEL_BYTES_PER_RES = 2 * 32 * 128 * 4 + 2 * 128 * 4 + 24 * 4
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per valid residue row of the same kernel, from the
# `ncu --set full` capture in profiles/r01_tc_f16x3_ncu_raw.csv: 966.6 MB for a launch over 29 616 valid rows
NCU_EDGE_TRAFFIC_PER_RES = {"f16x3": 966.555e6 / 29616}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,enforced.power.limit")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def wait_first_sample(self, timeout=10.0):
        """nvidia-smi takes a moment to attach to the driver and that can stall kernel launches: the timed regions
        start only after it is polling."""
        t0 = time.monotonic()
        while self.proc is not None and not self.rows and time.monotonic() - t0 < timeout:
            time.sleep(0.05)

    def stop(self):
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()

    def window(self, t0, t1):
        """Summary of the samples taken between two time.monotonic() stamps."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, pw, plim = [], None, set(), [], None
        for ts, r in list(self.rows):
            if ts < t0 or ts > t1:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            try:
                pw.append(float(r[3]))
                plim = float(r[8])
            except (ValueError, IndexError):
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None, "power_limit_w": plim}


def build_sweep(n_complexes, seed_base, rank):
    from packppi_b200 import synthetic
    from packppi_b200.batch import collate
    lengths = synthetic.sweep_lengths(n_complexes, 200, 800, seed=64)
    items = [synthetic.make_complex(ch, seed=seed_base + 1000 * rank + i) for i, ch in enumerate(lengths)]
    items.sort(key=lambda b: b.max_size)
    micro = [collate(items[i:i + MICRO]) for i in range(0, len(items), MICRO)]
    residues = sum(int(b.residue_mask.sum()) for b in micro)
    return micro, residues


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_rate(batch, n_ode, threads):
    """residue.steps/s of the CPU oracle (dense, graph rebuilt every step like the reference) on one complex."""
    from oracle import msc_oracle as mo
    from packppi_b200 import weights
    torch.set_num_threads(threads)
    sd = weights.make_state_dict(0)
    g = torch.Generator().manual_seed(1)
    x0 = ((torch.rand(batch.SC_D.shape, generator=g) * 2 - 1) * math.pi) * batch.SC_D_mask
    t0 = time.perf_counter()
    mo.sampling(sd, batch, x0, n_steps=n_ode, hoist=False)
    dt = time.perf_counter() - t0
    return float(batch.residue_mask.sum()) * n_ode / dt, dt


def run_reference(args):
    """`--impl reference`: the reference algorithm on the host cores (oracle port; the reference itself is Python on
    torch and cannot travel to the GPU box).  Bounded sample: one ~300-residue complex of the sweep,
    n_ode reverse-ODE steps per bench step, calibrated so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from packppi_b200 import synthetic
    threads = os.cpu_count() or 1
    lengths = synthetic.sweep_lengths(args.complexes, 200, 800, seed=64)
    ch = min(lengths, key=lambda c: abs(sum(c) - 300))
    b = synthetic.make_complex(ch, seed=7)
    L = sum(ch)
    _, dt1 = oracle_rate(b, 1, threads)  # calibration (also warms the thread pool)
    budget = 150.0 / max(1, args.steps + args.warmup)
    n_ode = int(max(1, min(N_ODE, budget / max(dt1, 1e-3))))
    for _ in range(args.warmup):
        oracle_rate(b, n_ode, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_rate(b, n_ode, threads)
    dt = time.perf_counter() - t0
    value = L * n_ode * args.steps / dt
    sample = f"1 complex of {L} residues (sweep member), 1 sample, {n_ode} reverse-ODE steps per bench step"
    line = {"impl": "reference", "metric": "residue.denoise-steps/s", "value": value, "unit": "residue.steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "residue.steps/s", "cores": threads, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "residue.steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"kernel_mode": os.environ.get("PACKPPI_B200_MODE", "f16x3"),
            "kernel_node_epilogue": os.environ.get("PACKPPI_B200_NODE_EPILOGUE") or "tc32",
            "kernel_cluster": int(os.environ.get("PACKPPI_B200_CLUSTER", "1")),
            "workload": f"sweep of {args.complexes} synthetic 2-chain complexes, L~U{{200..800}} (seed 64) x "
                        f"{N_SAMPLES} diffusion samples x {N_ODE} reverse-ODE steps, micro-batches of {MICRO} complexes; "
                        "BASELINE.json configs[4]",
            "weights": "random init, packppi_b200.weights.make_state_dict(0)", "samples_per_complex": N_SAMPLES,
            "denoise_steps": N_ODE, "l2": "per-micro-batch h_E working set 0.3-1.6 GB > 126 MB L2 (inputs larger than L2)",
            "arithmetic": {"f16x3": "fp32-grade: every product is 3 tcgen05 fp16 MMAs on rounded (hi, lo) operand pairs "
                                    "with fp32 accumulation, the node update with promoted fp32 sums; meets the fp32 "
                                    "parity gates (chi <= 1e-4 rad at every step against the reference)",
                           "fp32": "fp32 FFMA on CUDA cores",
                           "f16": "fast mode: plain fp16 tensor-core inputs, fp32 accumulation (own tolerance 2e-2 rad)"
                           }.get(os.environ.get("PACKPPI_B200_MODE", "f16x3"), "see kernel_mode"),
            "parallelism": "one sweep per GPU, no data-path collective"}


def _events_ms(fn, reps):
    for _ in range(3):  # first call eager, second call captures the CUDA graph (small problems), third replays
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_secondary(model, dev, with_cpu):
    """Second headline metric of BASELINE.json ("clash-grad ms @5k residues", configs[3]) and the single-complex
    configurations (configs[0..2]) that are latency- rather than throughput-bound.  Device-timed, rank 0 only."""
    import numpy as np

    from packppi_b200 import compute_residue_clash, get_atom14_coords, proximal_optimizer, synthetic
    from packppi_b200.batch import ComplexBatch, TENSOR_FIELDS
    out = {}
    # -- clash loss + analytic gradient and the full proximal loop on one 5000-residue complex
    b = synthetic.make_complex((500,) * 10, seed=5000).to(dev)
    b["X"] = (get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D) * b.atom_mask[..., None]).contiguous()
    x = b.SC_D.clone().requires_grad_(True)

    def fwd_bwd():
        x.grad = None
        compute_residue_clash(b, x).sum().backward()

    out["clash_grad_ms_5k"] = _events_ms(fwd_bwd, 20)
    out["clash_fwd_ms_5k"] = _events_ms(lambda: compute_residue_clash(b, b.SC_D), 20)
    t0 = time.perf_counter()
    snaps, losses = proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 50)
    torch.cuda.synchronize()
    out["proximal_50_steps_first_call_ms_5k"] = 1e3 * (time.perf_counter() - t0)  # cold: neighbour list, allocations
    out["proximal_50_steps_ms_5k"] = _events_ms(lambda: proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 50), 5)
    out["proximal_loss_first_last_5k"] = [losses[0], losses[-1]]
    out["peak_mem_GB"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    if with_cpu:
        from oracle import prox_oracle as po
        bc = b.to("cpu")
        t0 = time.perf_counter()
        po.clash_value_and_grad(bc, bc.SC_D, sparse=True)
        out["cpu_port_clash_grad_ms_5k"] = 1e3 * (time.perf_counter() - t0)
        out["cpu_note"] = ("sparse (KD-tree) CPU oracle; the reference's dense [N,N,14,14] formulation needs ~257 GB at "
                           "5000 residues and cannot run")
    # -- single complexes: 1BRS (195) and T1124 (739) from the golden inputs, synthetic 1500; one sample, 30 steps
    gold = os.path.join(ROOT, "tests", "golden")
    cases = []
    for name in ("1brs", "t1124"):
        with np.load(os.path.join(gold, name + ".npz")) as z:
            cb = ComplexBatch(**{k: torch.from_numpy(z["in_" + k]) for k in TENSOR_FIELDS})
        cb["num_proteins"], cb["max_size"] = 1, int(cb.X.shape[1])
        cases.append((name, cb.to(dev)))
    cases.append(("synthetic1500", synthetic.make_complex((500,) * 3, seed=1500).to(dev)))
    for name, cb in cases:
        L = int(cb.residue_mask.sum())

        def run():
            model._graph_cache = (None, model._graph_cache[1])  # rebuild graph + edge embedding, keep the buffers
            return model.sampling(cb)

        ms = _events_ms(run, 5)
        out[f"sampling_{name}"] = {"residues": L, "ms": ms, "residue_steps_per_s": L * N_ODE / (ms * 1e-3)}
    cb = cases[0][1]
    chi = model.sampling(cb)
    t0 = time.perf_counter()
    proximal_optimizer(cb, chi, 12.0, 0.5, 1.0, 50)
    torch.cuda.synchronize()
    out["proximal_50_steps_first_call_ms_1brs"] = 1e3 * (time.perf_counter() - t0)
    out["proximal_50_steps_ms_1brs"] = _events_ms(lambda: proximal_optimizer(cb, chi, 12.0, 0.5, 1.0, 50), 5)
    return out


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="packppi_b200", choices=["packppi_b200", "reference"])
    ap.add_argument("--complexes", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the clash-grad @5k and single-complex timings")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from packppi_b200 import TDiffusionModule, _lib, weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    model = TDiffusionModule()
    model.load_state_dict(weights.make_state_dict(0))
    model = model.to(dev).eval()
    micro_host, residues = build_sweep(args.complexes, 10_000, rank)
    micro_pinned = [b.pin_memory() for b in micro_host]
    micro_dev = [b.to(dev) for b in micro_host]
    units = residues * N_SAMPLES * N_ODE
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    gather_buf = None

    host_out = [torch.empty(N_SAMPLES, b.X.shape[0], b.X.shape[1], 4, dtype=torch.float32).pin_memory()
                for b in micro_host]

    def one_pass(batches, from_host):
        """One pass over the sweep.  from_host: every micro-batch is copied from pinned host memory and its sampled
        angles are copied back to pinned host memory (both asynchronous on the compute stream; the caller's
        synchronize at the end of the timed region waits for the last byte)."""
        outs = []
        for i, b in enumerate(batches):
            bd = b.to(dev, non_blocking=True) if from_host else b
            model._graph_cache = (None, model._graph_cache[1])  # new complex: graph + edge embedding are rebuilt every call
            chi = model.sampling(bd, n_samples=N_SAMPLES, generator=gen)
            if from_host:
                host_out[i].copy_(chi, non_blocking=True)
                outs.append(host_out[i])
            else:
                outs.append(chi)
        return outs

    def gather(outs):
        if world == 1:
            return
        flat = torch.cat([o.reshape(-1) for o in outs])
        bufs = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(bufs, flat)

    def timed(batches, from_host, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            outs = one_pass(batches, from_host)
            if not from_host:
                gather(outs)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    clocks = ClockSampler(local)  # one nvidia-smi poller for the whole run, started before the warm-up
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        one_pass(micro_dev, False)
    if rank == 0:
        clocks.wait_first_sample()
    _lib.LAUNCHES = 0
    t_a = time.monotonic()
    ms = timed(micro_dev, False, args.steps)
    t_b = time.monotonic()
    launches = _lib.LAUNCHES
    clk = clocks.window(t_a, t_b) if rank == 0 else None
    value = world * units * args.steps / (ms / 1e3)

    one_pass(micro_pinned, True)  # warm-up of the host path: its allocations differ from the device-resident pass
    t_a = time.monotonic()
    if os.environ.get("PP_BENCH_DEBUG"):
        for i in range(4):
            print(f"[debug] e2e pass {i}: {timed(micro_pinned, True, 1):.1f} ms; device pass: "
                  f"{timed(micro_dev, False, 1):.1f} ms", file=sys.stderr, flush=True)
    ms_e2e = timed(micro_pinned, True, max(1, min(args.steps, 2)))
    clk_e2e = clocks.window(t_a, time.monotonic()) if rank == 0 else None
    clocks.stop()
    e2e_steps = max(1, min(args.steps, 2))
    e2e_value = world * units * e2e_steps / (ms_e2e / 1e3)
    h2d = sum(b.nbytes() for b in micro_host)
    d2h = sum(int(b.X.shape[0] * b.X.shape[1]) for b in micro_host) * N_SAMPLES * 4 * 4

    # instrumented pass: CUDA events around every launch of the dominant kernel (the per-edge edge update)
    mode, cluster = model.engine(dev).mode, model.kernel_cluster
    pkey = "pp_ipmp_edge_edge" if mode == "fp32" else "pp_ipmp_edge_tc:edge"
    _lib.PROFILE = {pkey: []}
    t_pass = timed(micro_dev, False, 1)
    ev = _lib.PROFILE[pkey]
    _lib.PROFILE = None
    torch.cuda.synchronize()
    k_ms = [a.elapsed_time(b) for a, b, _ in ev]
    # algorithmic work = the valid residues (padding rows of a ragged micro-batch are not work; the tensor-core
    # kernels skip them).  Every micro-batch launches the kernel for 2 layers x 30 steps.
    valid = sum(int(b.residue_mask.sum()) for b in micro_host) * N_SAMPLES
    peaks = measured_peaks()
    if k_ms:
        tot_ms, tot_rows = sum(k_ms), valid * 2 * N_ODE
        assert len(k_ms) == len(micro_host) * 2 * N_ODE
        tflops = EDGE_KERNEL_FLOP_PER_RES * tot_rows / (tot_ms * 1e-3) / 1e12
        gbs = EDGE_KERNEL_BYTES_PER_RES * tot_rows / (tot_ms * 1e-3) / 1e9
        kname = {"fp32": "edge_edge_kernel (per-edge message MLP + FFN, fp32 FFMA on CUDA cores)",
                 "f16x3": "edge_tc_kernel<edge> (per-edge message MLP + FFN, tcgen05 kind::f16, split fp16 operand "
                          "pairs: 3 MMAs per product, fp32 accumulation)",
                 "f16": "edge_tc_kernel<edge> (per-edge message MLP + FFN, tcgen05 kind::f16, plain fp16 inputs)"}[mode]
        roof = {"kernel": kname, "bound": "tensor",
                "achieved": tflops, "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": tflops / peaks["tf_sust"],
                "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside a long step)",
                "traffic": (NCU_EDGE_TRAFFIC_PER_RES[mode] * tot_rows / len(k_ms)
                            if mode in NCU_EDGE_TRAFFIC_PER_RES else None),
                "traffic_note": "bytes per launch = ncu DRAM bytes per valid residue row (profiles/"
                                "r01_tc_f16x3_ncu_raw.csv) x the average valid rows per launch of this run",
                "algorithmic_bytes_per_launch": EDGE_KERNEL_BYTES_PER_RES * tot_rows / len(k_ms),
                "avg_launch_ms": tot_ms / len(k_ms), "launches": len(k_ms),
                "share_of_step": tot_ms / t_pass, "hbm_view": {"achieved_GBps": gbs, "peak_GBps": peaks["hbm"],
                                                              "frac": gbs / peaks["hbm"]},
                "flop_per_residue_row": EDGE_KERNEL_FLOP_PER_RES, "bytes_per_residue_row": EDGE_KERNEL_BYTES_PER_RES}
    else:
        roof = None

    secondary = None
    if rank == 0 and not args.no_secondary:
        secondary = run_secondary(model, dev, not args.no_cpu_baseline)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        small = min(micro_host[0:1], key=lambda b: b.max_size)
        from packppi_b200.batch import ComplexBatch
        one = ComplexBatch(**{k: (v[:1] if torch.is_tensor(v) else v) for k, v in small.items()})
        Lc = int(one.residue_mask.sum())
        for k, v in list(one.items()):  # strip the padding of the micro-batch
            if torch.is_tensor(v):
                one[k] = v[:, :Lc].contiguous()
        one["num_proteins"], one["max_size"] = 1, Lc
        threads = os.cpu_count() or 1
        rate, dt = oracle_rate(one, 10, threads)
        cpu = {"value": rate, "unit": "residue.steps/s", "cores": threads, "kind": "port",
               "sample": f"1 complex of {Lc} residues (shortest of the sweep), 1 sample, 10 reverse-ODE steps, {dt:.1f} s"}

    if rank == 0:
        line = {"metric": "residue.denoise-steps/s", "value": value, "unit": "residue.steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args), "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "residue.steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / e2e_steps, "clocks": clk_e2e},
                "gpu_launches": launches, "residues_per_gpu": residues, "secondary": secondary}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
