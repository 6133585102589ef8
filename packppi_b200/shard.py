"""Sharding of independent complexes over the GPUs of one box (SURVEY.md §8e, row "batched sampling").

(complex, sample) items are independent and the samples of one complex share graph and edge embedding, so a
complex is never split: complexes are assigned to ranks by greedy longest-first bin packing on their residue
count, every rank samples its own share with no data-path collective, and one `all_gather` returns the sampled
angles.  Works with any `torch.distributed` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def partition(lengths, world):
    """-> list over ranks of lists of complex indices; deterministic, balanced by total residues (ties -> lower rank)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    for r in range(world):
        out[r].sort()
    return out


def gather_angles(local, lengths, plan, n_samples):
    """local: {complex index: tensor [S, L_c, 4]} of this rank -> list over ALL complexes of [S, L_c, 4] tensors,
    identical on every rank.  One all_gather of a flat, equally padded buffer per rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [local[i] for i in range(len(lengths))]
    sizes = [sum(int(lengths[i]) for i in plan[r]) * n_samples * 4 for r in range(world)]
    cap = max(sizes)
    some = next(iter(local.values())) if local else None
    dev = some.device if some is not None else torch.device("cpu")
    rank = dist.get_rank()
    flat = torch.zeros(cap, dtype=torch.float32, device=dev)
    if local:
        mine = torch.cat([local[i].reshape(-1) for i in plan[rank]])
        flat[:mine.numel()] = mine
    bufs = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(bufs, flat)
    out = [None] * len(lengths)
    for r in range(world):
        o = 0
        for i in plan[r]:
            n = int(lengths[i]) * n_samples * 4
            out[i] = bufs[r][o:o + n].reshape(n_samples, int(lengths[i]), 4)
            o += n
    return out


def sample_sharded(sample_fn, batches, n_samples):
    """batches: list of single-complex batches (same list on every rank).  `sample_fn(batch, n_samples)` ->
    [S, 1, L, 4].  Returns the per-complex angles of ALL complexes on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lengths = [int(b.max_size) for b in batches]
    plan = partition(lengths, world)
    local = {i: sample_fn(batches[i], n_samples).reshape(n_samples, lengths[i], 4) for i in plan[rank]}
    return gather_angles(local, lengths, plan, n_samples)
