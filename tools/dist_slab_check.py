"""Run under torchrun with >= 2 GPUs: slab-partitioned PackPPI-Prox (halo exchange over NCCL, the whole loop one CUDA
graph) against the single-GPU run of the same complex; prints one JSON line on rank 0.
`torchrun --nproc-per-node 2 tools/dist_slab_check.py [n_chains x 500 residues]`.  tests/test_gpu_multi.py runs it
when the box has two GPUs; bench.py times the same thing at every world size (`secondary.slab_proximal_*`)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))


def main():
    from packppi_b200 import get_atom14_coords, proximal_optimizer, shard, synthetic
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    b = synthetic.make_complex((500,) * chains, seed=5000).to(dev)
    b["X"] = (get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D) * b.atom_mask[..., None]).contiguous()
    sp = shard.SlabProximal(b, 12.0, 0.5)
    eager, _ = sp.run(b.SC_D, 1.0, 50)   # first call: eager launches
    sp.run(b.SC_D, 1.0, 50)              # second call: captures the loop (NCCL included) in a CUDA graph
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    snaps, losses = sp.run(b.SC_D, 1.0, 50)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert torch.equal(eager, snaps), "graph replay differs from the eager loop"
    ref_snaps, ref_losses = proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 50)
    proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 50)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref_snaps, ref_losses = proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 50)
    torch.cuda.synchronize()
    dt1 = time.perf_counter() - t0
    d = (torch.stack(ref_snaps)[:, 0] - snaps).abs().max().item()
    dl = max(abs(a - c) / max(abs(c), 1e-9) for a, c in zip(losses.cpu().tolist(), ref_losses))
    n_local = len(sp.local)
    if rank == 0:
        import json
        print(json.dumps({"world": world, "residues": chains * 500, "local_rank0": n_local,
                          "max_chi_diff_vs_one_gpu": d, "bit_identical": bool(d == 0.0), "max_rel_loss_diff": dl,
                          "slab_ms": 1e3 * dt, "one_gpu_ms": 1e3 * dt1}))
    assert d < 1e-4 and dl < 1e-4, (d, dl)
    # the captured graph holds NCCL work: release it (and everything queued) before the group is torn down
    sp._state.clear()
    del sp
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
