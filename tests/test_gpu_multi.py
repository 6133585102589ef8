"""Multi-GPU tests (skipped below two GPUs): one process per GPU under torchrun, NCCL over NVLink."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _torchrun(n, script, *args, timeout=600):
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_slab_partitioned_proximal_on_two_gpus_equals_one_gpu():
    """BASELINE config 4: 5000 residues cut into two slabs, halo angles exchanged by one NCCL
    all-gather per step inside a captured CUDA graph; snapshots equal the single-GPU run (tools/dist_slab_check.py
    asserts 1e-4 rad and that the graph replay equals the eager loop bit for bit)."""
    r = _torchrun(2, "tools/dist_slab_check.py", "10")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["world"] == 2 and line["max_chi_diff_vs_one_gpu"] < 1e-4
    print(line)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_bench_strong_scaling_line_on_two_gpus():
    r = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "1", "--warmup", "3", "--complexes", "8",
                  "--no-cpu-baseline", "--no-secondary")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0
    assert sum(line["rank_residue_share"]) == line["residues_total"]
    assert line["weak_scaling"]["value"] > 0
