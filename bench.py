"""Benchmark of the PackPPI-MSC reverse-diffusion sampling path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl packppi_b200|reference] [--complexes C]

Workload (BASELINE.json configs[4]): ONE fixed sweep of C = 64 synthetic two-chain complexes
with L ~ U{200..800} residues (seed 64), 8 diffusion samples each, random-init weights (seed 0).  One bench "step" =
one full pass over the sweep: for every complex the kNN graph, the edge embedding and 30 reverse-ODE steps of all 8
samples.  Metric: residue.denoise-steps/s = sum(valid residues) x 8 samples x 30 / seconds.
With N > 1 the SAME sweep is partitioned over the ranks (`packppi_b200.shard.partition`: longest-first bin packing on
the residue counts; the 8 samples of a complex share graph and edge embedding and stay together): strong scaling, no
data-path collective; one pre-allocated `all_gather_into_tensor` returns every rank's sampled angles at the end of
each step (`"scaling": "strong"`).  The line also carries `weak_scaling` (every rank runs a whole sweep of its own, as
round 1 measured it) so that both curves can be read from the same run.  Each rank processes its complexes in
micro-batches of 8 (sorted by length, padded to the longest of the micro-batch).

`value`  : inputs resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`    : same pass through the public API (TDiffusionModule.sampling) from pinned HOST batches, including the
           host->device copy of every micro-batch and the device->host copy of the sampled angles.
`roofline`: dominant kernel (the per-edge edge update), timed with CUDA events around each of its launches in one
           instrumented pass; `roofline_clash`: the batched clash loss + gradient kernel on the same sweep's 512 items.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (`TDiffusionModule.sampling`,
           /root/reference/src/models/TorsionalDiffusion.py:254-298, copied to baseline/_ref/ by
           tools/install_reference.sh and imported under tools/ref_shims.py) on the box's host cores, all threads, on a
           bounded, length-stratified sample of the same sweep (kind "reference"); the CPU oracle port is the fallback
           only if baseline/_ref/ is absent (kind "port").
`parity_check`: the GPU path against that reference run on one complex of the sample, reference noise injected.
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
MICRO = int(os.environ.get("PP_BENCH_MICRO", "8"))  # complexes per micro-batch
# arithmetic of the message-passing GEMMs per execution mode (the `dtype` key of the line)
DTYPE = {"f16x3": "f32 via 3x split-f16 tensor-core products (fp32 accumulate, fp32-grade)", "fp32": "f32",
         "f16": "f16 inputs, f32 accumulate"}
# executed FLOPs of one edge_edge_kernel launch per residue row (K = 32 edges): first Linear 168 wide (h_E + pair
# geometry; the h_V parts are hoisted per residue), two 128x128 Linears, FFN 128-512-128   
EDGE_KERNEL_FLOP_PER_RES = 32 * 2 * (168 * 128 + 128 * 128 + 128 * 128 + 2 * 128 * 512)
# algorithmic HBM bytes of the same launch per residue row: read h_E (K*128*4), write h_E (K*128*4), A, N gathers
EDGE_KERNEL_BYTES_PER_RES = 2 * 32 * 128 * 4 + 2 * 128 * 4 + 24 * 4
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per valid residue row of the same kernel, from the
# `ncu --set full` capture in profiles/r02_edge_tc_ncu_raw.csv: 520.9 MB read + 444.5 MB written for a launch over
# 29 616 valid rows (round 1: 966.6 MB - unchanged, the traffic is the algorithmic h_E read + write)
NCU_EDGE_TRAFFIC_PER_RES = {"f16x3": (520.866304e6 + 444.490752e6) / 29616}


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


def sweep_items(n_complexes, seed_base=10_000):
    """The fixed sweep: list of single-complex host batches (identical on every rank)."""
    from packppi_b200 import synthetic
    lengths = synthetic.sweep_lengths(n_complexes, 200, 800, seed=64)
    return [synthetic.make_complex(ch, seed=seed_base + i) for i, ch in enumerate(lengths)]


def micro_batches(items):
    from packppi_b200.batch import collate
    items = sorted(items, key=lambda b: b.max_size)
    micro = [collate(items[i:i + MICRO]) for i in range(0, len(items), MICRO)]
    residues = sum(int(b.residue_mask.sum()) for b in micro)
    return micro, residues


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------ reference arm
def stratified(items, n):
    """n complexes spread evenly over the length-sorted sweep (mid-points of n equal quantile bins)."""
    order = sorted(range(len(items)), key=lambda i: (items[i].max_size, i))
    n = max(1, min(n, len(order)))
    return [items[order[min(len(order) - 1, (2 * k + 1) * len(order) // (2 * n))]] for k in range(n)]


class ReferenceCPU:
    """The unmodified reference on the host cores: `TDiffusionModule.sampling(batch)` exactly as
    src/eval_diffusion.py:62 calls it (fp32, autograd graph and per-step graph rebuild included), with the bench
    weights loaded through its own `load_state_dict`.  None of packppi_b200's kernels are on this path."""

    def __init__(self, threads):
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        os.environ.setdefault("TQDM_DISABLE", "1")  # the reference builds its score tables under tqdm progress bars
        import ref_shims
        from packppi_b200 import weights
        if not ref_shims.available():
            raise FileNotFoundError("baseline/_ref is absent: run tools/install_reference.sh in the build container")
        torch.set_num_threads(threads)
        self.threads = threads
        self.ref = ref_shims.import_reference()
        self.model = ref_shims.build_reference_model(self.ref)
        self.model.load_state_dict(weights.make_state_dict(0))
        self.kind = "reference"
        self.drawn = None
        orig = self.model.add_sc_noise

        def recording(batch, t):  # keeps the reference's own initial noise draw for the parity check
            out = orig(batch, t)
            self.drawn = out[0].detach().clone()
            return out

        self.model.add_sc_noise = recording

    def sample(self, b, seed=1):
        """-> (sampled angles [1,L,4], seconds)"""
        d = self.ref.Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()})
        torch.manual_seed(seed)
        t0 = time.perf_counter()
        out = self.model.sampling(d, use_proximal=False)
        return out.detach(), time.perf_counter() - t0


class OraclePortCPU:
    """Fallback when baseline/_ref is absent: the CPU oracle port (dense like the reference, graph rebuilt every step)."""

    def __init__(self, threads):
        from packppi_b200 import weights
        torch.set_num_threads(threads)
        self.threads, self.kind, self.drawn = threads, "port", None
        self.sd = weights.make_state_dict(0)

    def sample(self, b, seed=1):
        from oracle import msc_oracle as mo
        g = torch.Generator().manual_seed(seed)
        x0 = ((torch.rand(b.SC_D.shape, generator=g) * 2 - 1) * math.pi) * b.SC_D_mask
        self.drawn = x0
        t0 = time.perf_counter()
        out = mo.sampling(self.sd, b, x0, n_steps=N_ODE, hoist=False)
        return out, time.perf_counter() - t0


def cpu_arm(threads):
    try:
        return ReferenceCPU(threads)
    except (FileNotFoundError, ImportError) as e:
        print(f"[bench] reference unavailable ({e}); timing the oracle port instead", file=sys.stderr)
        return OraclePortCPU(threads)


def run_reference(args):
    """`--impl reference`: every bench step = one call of the reference's `sampling` (30 reverse-ODE steps, 1 diffusion
    sample) on ONE complex of the sweep; the K timed steps walk K complexes spread evenly over the length-sorted sweep,
    so the rate is that of a length-stratified sample of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "WORLD_SIZE" in os.environ and os.environ.get("PP_BENCH_REFERENCE_CHILD") != "1":
        # Under torchrun every rank inherits OMP_NUM_THREADS=1 (read when the OpenMP runtime starts, i.e. at `import
        # torch`), which would pin the reference to one core.  Rank 0 re-runs this arm in a fresh interpreter without the
        # launcher's environment and relays its line; the other ranks have already left.
        env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "WORLD_SIZE",
                                                                "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK",
                                                                "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID")}
        env["PP_BENCH_REFERENCE_CHILD"] = "1"
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", str(args.gpus),
                              "--steps", str(args.steps), "--warmup", str(args.warmup), "--complexes",
                              str(args.complexes)], env=env, capture_output=True, text=True)
        sys.stderr.write(out.stderr[-2000:])
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        print(lines[-1] if lines else json.dumps({"impl": "reference", "unavailable": "child run failed"}))
        return
    threads = os.cpu_count() or 1
    items = sweep_items(args.complexes)
    arm = cpu_arm(threads)
    timed = stratified(items, args.steps)
    timed = [timed[k % len(timed)] for k in range(args.steps)]
    # bound the run: ~1.5 s per 10 k residue.steps on 16 cores; shrink to the shorter half of the sweep if K is large
    _, dt0 = arm.sample(min(items, key=lambda b: b.max_size))
    rate0 = min(items, key=lambda b: b.max_size).max_size * N_ODE / dt0
    est = sum(b.max_size for b in timed) * N_ODE / rate0
    bounded = ""
    if est > 240.0:
        keep = sorted(items, key=lambda b: b.max_size)[:max(1, int(len(items) * 240.0 / est))]
        timed = stratified(keep, args.steps)
        timed = [timed[k % len(timed)] for k in range(args.steps)]
        bounded = f" (restricted to the {len(keep)} shortest complexes to bound the run)"
    warm = stratified(items, max(1, args.warmup))[:1] * max(0, args.warmup - 1)  # the calibration call was warm-up 1
    for b in warm:
        arm.sample(b)
    res, secs = 0, 0.0
    for b in timed:
        _, dt = arm.sample(b)
        res += int(b.residue_mask.sum())
        secs += dt
    value = res * N_ODE / secs
    sample = (f"{args.steps} calls of TDiffusionModule.sampling (30 reverse-ODE steps, 1 diffusion sample), one "
              f"complex of the sweep per bench step, {min(b.max_size for b in timed)}-{max(b.max_size for b in timed)} "
              f"residues spread evenly over the length-sorted sweep{bounded}; {res} residues, {secs:.1f} s; "
              f"CPU: {cpu_model_name()}")
    cfg = workload_config(args)  # the same workload description as our arm; what was actually timed: `sample`
    line = {"impl": "reference", "sample": sample, "metric": "residue.denoise-steps/s", "value": value, "unit": "residue.steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "residue.steps/s", "cores": threads, "kind": arm.kind,
                             "sample": sample, "cpu_model": cpu_model_name()},
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
            "parallelism": (f"the fixed sweep partitioned over {args.gpus} rank(s) by residue count "
                            "(packppi_b200.shard.partition), no data-path collective, one all_gather_into_tensor of "
                            "the sampled angles per step")}


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


CLASH_BYTES_PER_RES = 330  # backbone 36 + chi 16 + type/index 8 in, atom records, 16 + 4 out


def run_batched_clash(items, dev, peaks):
    """`roofline_clash`: the clash loss + analytic gradient (atom14 rebuild included) and the full 50-step PackPPI-Prox
    over the sweep's 512 (complex, sample) items, micro-batch by micro-batch (B = 8 complexes x S = 8 decoys per
    launch) - the batched variant for which SURVEY §8d asks an HBM fraction (~330 B per residue and evaluation)."""
    from packppi_b200 import _lib, proximal_optimizer
    from packppi_b200.components import clash_context
    micro, residues = micro_batches(items)
    rows = residues * N_SAMPLES
    gen = torch.Generator(device=dev).manual_seed(77)
    work = []
    for b in micro:
        bd = b.to(dev)
        chi = ((torch.rand(N_SAMPLES, *bd.SC_D.shape, device=dev, generator=gen) * 2 - 1) * math.pi) * bd.SC_D_mask
        cc = clash_context(bd)
        w = torch.ones(chi.numel() // 4, device=dev)
        work.append((bd, chi.contiguous(), cc, w))
    def evals():
        for bd, chi, cc, w in work:
            cc.evaluate(chi.reshape(-1, 4), res_w=w)

    for _ in range(3):
        evals()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        evals()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = CLASH_BYTES_PER_RES * rows / (ms * 1e-3) / 1e9
    # the 50-step proximal loop on the same items (first call eager, second captures the CUDA graph, then replays)
    t_prox = 0.0
    accepted = 0
    for bd, chi, cc, w in work:
        run = lambda: cc.proximal(chi.reshape(-1, 4), 1.0, 50)  # noqa: E731
        run(); run()
        torch.cuda.synchronize()
        e0.record()
        _, losses, _ = run()
        e1.record()
        torch.cuda.synchronize()
        t_prox += e0.elapsed_time(e1)
        accepted += int((losses[-1] < losses[0]).sum())
    n_items = len(items) * N_SAMPLES
    return {"kernel": "atom14_kernel + clash_pair_kernel<1> (loss + analytic dL/dchi), batched over B x S items",
            "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
            "peak_source": f"{peaks['src']} HBM copy bandwidth", "traffic": None,
            "algorithmic_bytes_per_launch": CLASH_BYTES_PER_RES * rows / len(work), "bytes_per_residue": CLASH_BYTES_PER_RES,
            "avg_launch_ms": ms / len(work), "launches": len(work), "residue_rows": rows, "items": n_items,
            "evaluations_per_s": n_items / (ms * 1e-3), "residue_evals_per_s": rows / (ms * 1e-3),
            "note": "latency / L2 bound: the pair kernel walks neighbour lists with dependent loads; the HBM fraction is "
                    "reported because SURVEY §8d asks for it on the batched variant, not because HBM binds",
            "proximal_50_steps_all_items_ms": t_prox, "proximal_items_per_s": n_items / (t_prox * 1e-3),
            "proximal_items_accepted": accepted}


def run_slab_proximal(dev, world, rank):
    """BASELINE configs[3]: PackPPI-Prox of ONE large complex cut into `world` spatial slabs (packppi_b200.shard.
    SlabProximal; one NCCL all-gather of the owned angles per step, the loop replayed as one CUDA graph).  Collective:
    every rank takes part; the numbers are the max over ranks."""
    import torch.distributed as dist

    from packppi_b200 import get_atom14_coords, shard, synthetic
    out = {}
    for tag, chains in (("5k", 10), ("50k", 100)):
        b = synthetic.make_complex((500,) * chains, seed=5000 if chains == 10 else 50000).to(dev)
        b["X"] = (get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D) * b.atom_mask[..., None]).contiguous()
        t0 = time.perf_counter()
        sp = shard.SlabProximal(b, 12.0, 0.5)
        torch.cuda.synchronize()
        setup_ms = 1e3 * (time.perf_counter() - t0)
        sp.run(b.SC_D, 1.0, 50)
        sp.run(b.SC_D, 1.0, 50)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            _, losses = sp.run(b.SC_D, 1.0, 50)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps, float(len(sp.local))], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[f"slab_proximal_50_steps_ms_{tag}"] = float(ms[0])
        out[f"slab_proximal_local_residues_max_{tag}"] = int(ms[1])
        out[f"slab_proximal_setup_ms_{tag}"] = setup_ms
        out[f"slab_proximal_loss_first_last_{tag}"] = [float(losses[0]), float(losses[-1])]
        del sp, b
        torch.cuda.empty_cache()
    out["slab_proximal_world"] = world
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
    ap.add_argument("--no-weak", action="store_true", help="skip the extra weak-scaling pass at N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    from packppi_b200 import TDiffusionModule, _lib, shard, weights

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
    items = sweep_items(args.complexes)                      # the fixed sweep, identical on every rank
    lengths = [int(b.max_size) for b in items]
    plan = shard.partition(lengths, world)                   # strong scaling: this rank's share of the sweep
    total_residues = sum(int(b.residue_mask.sum()) for b in items)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    class Pass:
        """Device / pinned-host copies of a list of complexes as micro-batches + pre-allocated output buffers."""

        def __init__(self, its, gather_cap=None):
            self.micro_host, self.residues = micro_batches(its) if its else ([], 0)
            self.pinned = [b.pin_memory() for b in self.micro_host]
            self.dev = [b.to(dev) for b in self.micro_host]
            self.host_out = [torch.empty(N_SAMPLES, b.X.shape[0], b.X.shape[1], 4, dtype=torch.float32).pin_memory()
                             for b in self.micro_host]
            self.sizes = [N_SAMPLES * b.X.shape[0] * b.X.shape[1] * 4 for b in self.micro_host]
            self.flat = self.all = None
            if gather_cap is not None:  # pre-allocated all_gather buffers: [cap] per rank -> [world, cap]
                self.flat = torch.zeros(gather_cap, dtype=torch.float32, device=dev)
                self.all = torch.empty(world, gather_cap, dtype=torch.float32, device=dev)

        def run(self, from_host, gather):
            """One pass.  from_host: every micro-batch is copied from pinned host memory and its sampled angles are
            copied back to pinned host memory (asynchronous on the compute stream; the synchronize that ends the timed
            region waits for the last byte)."""
            o = 0
            src = self.pinned if from_host else self.dev
            for i, b in enumerate(src):
                bd = b.to(dev, non_blocking=True) if from_host else b
                model._graph_cache = (None, model._graph_cache[1])  # new complexes: graph + edge embedding rebuilt
                chi = model.sampling(bd, n_samples=N_SAMPLES, generator=gen)
                if from_host:
                    self.host_out[i].copy_(chi, non_blocking=True)
                if gather and self.flat is not None:
                    self.flat[o:o + self.sizes[i]].copy_(chi.reshape(-1))
                    o += self.sizes[i]
            if gather and self.flat is not None and world > 1:
                dist.all_gather_into_tensor(self.all, self.flat)

    def padded_floats(its):
        mb, _ = micro_batches(its) if its else ([], 0)
        return sum(N_SAMPLES * b.X.shape[0] * b.X.shape[1] * 4 for b in mb)

    cap = max(padded_floats([items[i] for i in plan[r]]) for r in range(world))
    mine = Pass([items[i] for i in plan[rank]], gather_cap=cap)
    units = total_residues * N_SAMPLES * N_ODE               # whole job, all ranks together

    def timed(ps, from_host, steps, gather=True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ps.run(from_host, gather)
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
        mine.run(False, True)
    if rank == 0:
        clocks.wait_first_sample()
    _lib.LAUNCHES = 0
    t_a = time.monotonic()
    ms = timed(mine, False, args.steps)
    t_b = time.monotonic()
    launches = _lib.LAUNCHES
    clk = clocks.window(t_a, t_b) if rank == 0 else None
    value = units * args.steps / (ms / 1e3)

    mine.run(True, True)  # warm-up of the host path: its allocations differ from the device-resident pass
    t_a = time.monotonic()
    e2e_steps = max(1, min(args.steps, 2))
    ms_e2e = timed(mine, True, e2e_steps)
    clk_e2e = clocks.window(t_a, time.monotonic()) if rank == 0 else None
    e2e_value = units * e2e_steps / (ms_e2e / 1e3)
    h2d = torch.tensor([float(sum(b.nbytes() for b in mine.micro_host)),
                        float(sum(int(b.X.shape[0] * b.X.shape[1]) for b in mine.micro_host) * N_SAMPLES * 4 * 4)],
                       device=dev)
    if world > 1:
        dist.all_reduce(h2d)  # bytes of the whole job
    h2d, d2h = int(h2d[0].item()), int(h2d[1].item())

    # weak-scaling curve (round 1's definition): every rank runs a WHOLE sweep of its own, no gather
    weak = None
    if world > 1 and not args.no_weak:
        from packppi_b200 import synthetic
        wl = synthetic.sweep_lengths(args.complexes, 200, 800, seed=64)
        whole = Pass([synthetic.make_complex(ch, seed=10_000 + 1000 * rank + i) for i, ch in enumerate(wl)])
        whole.run(False, False)
        wsteps = max(1, min(args.steps, 3))
        wms = timed(whole, False, wsteps, gather=False)
        weak = {"value": world * whole.residues * N_SAMPLES * N_ODE * wsteps / (wms / 1e3), "unit": "residue.steps/s",
                "ms_per_step": wms / wsteps, "steps": wsteps,
                "note": "one whole sweep per rank (same lengths, different coordinates), no gather"}
        del whole
    clocks.stop()

    # instrumented pass: CUDA events around every launch of the dominant kernel (the per-edge edge update)
    mode, cluster = model.engine(dev).mode, model.kernel_cluster
    pkey = "pp_ipmp_edge_edge" if mode == "fp32" else "pp_ipmp_edge_tc:edge"
    _lib.PROFILE = {pkey: []}
    t_pass = timed(mine, False, 1)
    ev = _lib.PROFILE[pkey]
    _lib.PROFILE = None
    torch.cuda.synchronize()
    k_ms = [a.elapsed_time(b) for a, b, _ in ev]
    # algorithmic work = the valid residues (padding rows of a ragged micro-batch are not work; the tensor-core
    # kernels skip them).  Every micro-batch launches the kernel for 2 layers x 30 steps.
    valid = mine.residues * N_SAMPLES
    peaks = measured_peaks()
    if k_ms:
        tot_ms, tot_rows = sum(k_ms), valid * 2 * N_ODE
        assert len(k_ms) == len(mine.micro_host) * 2 * N_ODE
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
                                "r02_edge_tc_ncu_raw.csv) x the average valid rows per launch of this run",
                "algorithmic_bytes_per_launch": EDGE_KERNEL_BYTES_PER_RES * tot_rows / len(k_ms),
                "avg_launch_ms": tot_ms / len(k_ms), "launches": len(k_ms),
                "share_of_step": tot_ms / t_pass, "hbm_view": {"achieved_GBps": gbs, "peak_GBps": peaks["hbm"],
                                                              "frac": gbs / peaks["hbm"]},
                "flop_per_residue_row": EDGE_KERNEL_FLOP_PER_RES, "bytes_per_residue_row": EDGE_KERNEL_BYTES_PER_RES}
    else:
        roof = None

    slab = run_slab_proximal(dev, world, rank) if not args.no_secondary else {}  # collective: every rank takes part
    roof_clash = run_batched_clash(items, dev, peaks) if rank == 0 and not args.no_secondary else None
    secondary = None
    if rank == 0 and not args.no_secondary:
        secondary = run_secondary(model, dev, not args.no_cpu_baseline)
        secondary.update(slab)

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the CPU arm is a single-process measurement (N = 1)
        arm = cpu_arm(os.cpu_count() or 1)
        picks = stratified(items, 4)
        res, secs, check = 0, 0.0, None
        for j, b in enumerate(picks):
            out, dt = arm.sample(b)
            res += int(b.residue_mask.sum())
            secs += dt
            if j == 1:  # parity check on the second pick (~350 residues), the reference's own noise injected
                got = model.sampling(b.to(dev), init_SC_D=arm.drawn.to(dev)).cpu()
                d = (got - out).abs()
                d = torch.minimum(d, 2 * math.pi - d)
                check = {"against": arm.kind, "residues": int(b.max_size), "denoise_steps": N_ODE,
                         "max_chi_diff_rad": float(d.max()), "mean_chi_diff_rad": float(d.mean()),
                         "tolerance_rad": 1e-4, "pass": bool(d.max() < 1e-4)}
        parity = check
        cpu = {"value": res * N_ODE / secs, "unit": "residue.steps/s", "cores": arm.threads, "kind": arm.kind,
               "cpu_model": cpu_model_name(),
               "sample": f"{len(picks)} complexes of the sweep at the 1/8, 3/8, 5/8, 7/8 length quantiles "
                         f"({', '.join(str(int(b.max_size)) for b in picks)} residues), 1 diffusion sample, full 30 "
                         f"reverse-ODE steps each through TDiffusionModule.sampling, {secs:.1f} s"}

    if rank == 0:
        cfg = workload_config(args)
        line = {"metric": "residue.denoise-steps/s", "value": value, "unit": "residue.steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": DTYPE.get(mode, mode), "data": "synthetic",
                "config": cfg, "roofline": roof, "roofline_clash": roof_clash, "cpu_baseline": cpu,
                "parity_check": parity, "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "residue.steps/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / e2e_steps, "clocks": clk_e2e},
                "gpu_launches": launches, "launches_note": "kernels launched by rank 0 inside the timed region",
                "residues_total": total_residues, "residues_this_rank": mine.residues,
                "rank_residue_share": [sum(lengths[i] for i in plan[r]) for r in range(world)],
                "weak_scaling": weak, "secondary": secondary}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()  # rank 0's CPU baseline runs last: the others wait here instead of tearing the group down
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
