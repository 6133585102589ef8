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


# ----------------------------------------------------------------------------------------------------------------
# One very large complex: PackPPI-Prox partitioned into spatial slabs (BASELINE.json configs[3])
# ----------------------------------------------------------------------------------------------------------------
def slab_partition(ca, valid, world):
    """Residue ownership by slabs along the longest principal axis of the CA cloud, equal residue counts.

    ca [L,3] numpy, valid [L] bool -> owner [L] int (rank of every residue; invalid residues go to rank 0)."""
    import numpy as np
    ca = np.asarray(ca, np.float64)
    idx = np.nonzero(valid)[0]
    owner = np.zeros(len(ca), np.int64)
    if len(idx) == 0 or world == 1:
        return owner
    c = ca[idx] - ca[idx].mean(0)
    _, _, vt = np.linalg.svd(c, full_matrices=False)
    proj = c @ vt[0]
    order = idx[np.argsort(proj, kind="stable")]
    for r, chunk in enumerate(np.array_split(order, world)):
        owner[chunk] = r
    return owner


def halo_of(ca, reach, owner, rank, cutoff):
    """Residues not owned by `rank` that can interact with an owned one: CA distance < reach_i + reach_j + cutoff
    (the criterion of the clash neighbour list).  Returns the sorted local id list owned + halo."""
    import numpy as np
    from scipy.spatial import cKDTree
    ca = np.asarray(ca, np.float64)
    reach = np.asarray(reach, np.float64)
    own = np.nonzero((owner == rank) & (reach >= 0))[0]
    others = np.nonzero((owner != rank) & (reach >= 0))[0]
    if len(own) == 0 or len(others) == 0:
        return np.sort(np.nonzero(owner == rank)[0])
    rmax = float(reach[reach >= 0].max())
    tree = cKDTree(ca[others])
    halo = set()
    for i in own:
        for t in tree.query_ball_point(ca[i], reach[i] + rmax + cutoff):
            j = others[t]
            if np.linalg.norm(ca[i] - ca[j]) < reach[i] + reach[j] + cutoff:
                halo.add(int(j))
    return np.sort(np.concatenate([np.nonzero(owner == rank)[0], np.fromiter(halo, np.int64, len(halo))]))


def gather_owned_rows(mine, ids_of, counts, L):
    """mine [n_owned, C] of this rank -> full [L, C] with every row taken from its owner (one all_gather of equally
    padded buffers).  ids_of[r] = row ids owned by rank r, counts[r] = their number."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    full = torch.zeros(L, mine.shape[1], dtype=mine.dtype, device=mine.device)
    if world == 1:
        full[ids_of[0]] = mine
        return full
    buf = torch.zeros(max(counts), mine.shape[1], dtype=mine.dtype, device=mine.device)
    buf[:mine.shape[0]] = mine
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    for r in range(world):
        full[ids_of[r]] = bufs[r][:counts[r]]
    return full


class SlabProximal:
    """proximal_optimizer (reference optimize.py:21-73) for ONE complex spread over the ranks of a process group.

    Every rank owns one slab of residues, keeps copies of the halo residues it interacts with, evaluates loss and
    gradient of its owned residues with the same kernels as the single-GPU path (every cross-boundary pair is
    evaluated on both sides, so no force has to travel back), and updates their angles.  Per step the ranks exchange
    only the current angles of the owned residues (one all_gather, 16 B per residue); the loss values are summed
    once at the end.  All ranks hold the full input batch (replicated); outputs are identical on every rank."""

    def __init__(self, batch, violation_tolerance_factor=12.0, clash_overlap_tolerance=0.5):
        import numpy as np

        from . import tables
        from .engine import ClashContext
        assert batch.num_proteins == 1
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        dev = batch.X.device
        L = int(batch.X.shape[1])
        self.L, self.dev = L, dev
        ca = batch.X[0, :, 1, :].detach().cpu().numpy()
        has_atoms = (batch.atom_mask[0].sum(-1) > 0).cpu().numpy()
        rtype = batch.residue_type[0].cpu().numpy()
        bb = (batch.X[0, :, :4, :] - batch.X[0, :, 1:2, :]).norm(dim=-1).max(-1)[0].cpu().numpy()
        reach = np.maximum(tables.max_reach()[np.clip(rtype, 0, 20)], bb + 1e-3).astype(np.float64)
        reach[~has_atoms] = -1.0
        cutoff = max(2.0 * float(tables.raw()["clash_radius"].max()) - float(clash_overlap_tolerance), 0.0)
        self.owner = slab_partition(ca, has_atoms, self.world)
        self.local = halo_of(ca, reach, self.owner, self.rank, cutoff)
        self.counts = [int((self.owner == r).sum()) for r in range(self.world)]
        self.ids_of = [torch.from_numpy(np.nonzero(self.owner == r)[0]).to(dev) for r in range(self.world)]
        loc = torch.from_numpy(self.local).to(dev)
        self.loc = loc
        self.owned_local = torch.from_numpy((self.owner[self.local] == self.rank).astype(np.uint8)).to(dev)
        self.own_pos = torch.nonzero(self.owned_local)[:, 0]          # positions of owned residues inside the local set
        self.halo_pos = torch.nonzero(self.owned_local == 0)[:, 0]
        self.halo_ids = loc[self.halo_pos]
        sel = lambda t: t[:, loc].contiguous()  # noqa: E731
        self.cc = ClashContext(dev, sel(batch.X), sel(batch.residue_type), sel(batch.atom_mask),
                               sel(batch.residue_index), violation_tolerance_factor, clash_overlap_tolerance)

    def _gather_owned(self, local_vals):
        return gather_owned_rows(local_vals[self.own_pos], self.ids_of, self.counts, self.L)

    def run(self, SC_D, lamda, num_steps=50, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8):
        """-> (snapshots [num_steps, L, 4], losses [num_steps]) like ClashContext.proximal, for the whole complex."""
        import math

        from . import _lib
        cc, dev, n = self.cc, self.dev, len(self.local)
        f = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
        sc_full = SC_D.reshape(self.L, 4).to(torch.float32)
        sc_d = sc_full[self.loc].contiguous()        # starting angles of the local set; halo rows are refreshed per step
        ws = cc.scratch(1)
        per_res, _ = cc.evaluate(sc_d)
        tot = per_res[self.own_pos].sum().reshape(1)
        if self.world > 1:
            dist.all_reduce(tot)
        mean = torch.stack([tot[0] * 0, tot[0] / self.L])
        mask = torch.zeros(n, 4, dtype=torch.uint8, device=dev)
        z, x, m, v = f(n, 4), f(n, 4), f(n, 4), f(n, 4)
        _lib.call("pp_prox_init_from_mean", per_res, mean, sc_d, self.owned_local, n, mask, z, x, m, v)
        partial = f(int(_lib.load().pp_prox_partial_floats(n)))
        snap, losses = f(n, 4), f(num_steps, 2)
        snaps = f(num_steps, self.L, 4)
        static = (cc.tables.geo, cc.lower, cc.upper, cc.X, cc.rtype, cc.exists, cc.start, cc.list, sc_d)
        for k in range(num_steps):
            t = k + 1
            _lib.call("pp_prox_step", *static, mask, z, x, m, v, n, cc.tol, cc.max_cut, float(lamda),
                      lr / (1 - beta1 ** t), math.sqrt(1 - beta2 ** t), beta1, beta2, eps, snap, losses[k], per_res,
                      ws["atoms4"], ws["axes"], ws["bound"], partial, self.owned_local, self.L)
            full = self._gather_owned(snap)          # angles after the update, every residue from its owner
            snaps[k] = full
            if len(self.halo_pos):
                sc_d[self.halo_pos] = full[self.halo_ids]   # halo copies follow their owners
        if self.world > 1:
            dist.all_reduce(losses)
        return snaps, losses[:, 0]
