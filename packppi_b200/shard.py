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


def _collective_device():
    """Where collective buffers must live: the current CUDA device under NCCL, the host under gloo."""
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_angles(local, lengths, plan, n_samples, device=None):
    """local: {complex index: tensor [S, L_c, 4]} of this rank -> list over ALL complexes of [S, L_c, 4] tensors,
    identical on every rank.  One all_gather of a flat, equally padded buffer per rank.  A rank may own nothing
    (more ranks than complexes): its buffer then lives on `device` (default: the backend's device), not on
    whatever its empty result list would suggest."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [local[i] for i in range(len(lengths))]
    sizes = [sum(int(lengths[i]) for i in plan[r]) * n_samples * 4 for r in range(world)]
    cap = max(max(sizes), 1)
    some = next(iter(local.values())) if local else None
    dev = torch.device(device) if device is not None else (some.device if some is not None else _collective_device())
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
    only the current angles of the owned residues (one `all_gather_into_tensor` of equally padded buffers, 16 B per
    residue); the loss values are summed once at the end.  All ranks hold the full input batch (replicated); outputs
    are identical on every rank.

    Everything per step is stream-ordered device work with pre-allocated buffers - two kernels, one packing gather,
    the NCCL all-gather, one halo scatter - so the whole `num_steps` loop is captured in ONE CUDA graph (NCCL included)
    on the second call with the same schedule and replayed from then on.  The halo is found on the GPU from the
    clash neighbour list of the whole complex (`clash_nbr_cells_kernel`), not by a host-side tree search."""

    def __init__(self, batch, violation_tolerance_factor=12.0, clash_overlap_tolerance=0.5):
        import numpy as np

        from .engine import ClashContext
        assert batch.num_proteins == 1
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        dev = batch.X.device
        L = int(batch.X.shape[1])
        self.L, self.dev = L, dev
        ca = batch.X[0, :, 1, :].detach().cpu().numpy()
        has_atoms = (batch.atom_mask[0].sum(-1) > 0).cpu().numpy()
        self.owner = slab_partition(ca, has_atoms, self.world)
        owner = torch.from_numpy(self.owner).to(dev)
        mine = owner == self.rank
        if self.world > 1:
            # halo = residues of other ranks on the neighbour list of an owned residue (the list's own criterion:
            # CA distance < reach_i + reach_j + cutoff), read off the whole complex's list on the device
            whole = ClashContext(dev, batch.X, batch.residue_type, batch.atom_mask, batch.residue_index,
                                 violation_tolerance_factor, clash_overlap_tolerance)
            counts = (whole.start[1:] - whole.start[:-1])
            row_of = torch.repeat_interleave(torch.arange(L, device=dev), counts)
            nb = whole.list[:int(whole.start[-1].item())].long()
            sel = mine[row_of] & ~mine[nb]
            keep = mine.clone()
            keep[nb[sel]] = True
            del whole
        else:
            keep = mine
        loc = torch.nonzero(keep)[:, 0]                      # ascending global ids: same summation order as one GPU
        self.local = loc.cpu().numpy()
        self.loc = loc
        self.counts = [int((self.owner == r).sum()) for r in range(self.world)]
        self.cap = max(self.counts)
        self.owned_local = mine[loc].to(torch.uint8).contiguous()
        self.own_pos = torch.nonzero(self.owned_local)[:, 0]  # positions of owned residues inside the local set
        self.halo_pos = torch.nonzero(self.owned_local == 0)[:, 0]
        # position of every residue of the complex inside the gathered [world, cap] buffer
        slot = torch.zeros(L, dtype=torch.long, device=dev)
        for r in range(self.world):
            ids = torch.nonzero(owner == r)[:, 0]
            slot[ids] = r * self.cap + torch.arange(len(ids), device=dev)
        self.slot = slot
        self.halo_src = slot[loc[self.halo_pos]]
        sel_ = lambda t: t[:, loc].contiguous()  # noqa: E731
        self.cc = ClashContext(dev, sel_(batch.X), sel_(batch.residue_type), sel_(batch.atom_mask),
                               sel_(batch.residue_index), violation_tolerance_factor, clash_overlap_tolerance)
        self._state = {}

    def _loop(self, st, lamda, num_steps, lr, beta1, beta2, eps):
        import math

        from . import _lib
        cc, n = self.cc, len(self.local)
        ws = cc.scratch(1)
        static = (cc.tables.geo, cc.lower, cc.upper, cc.X, cc.rtype, cc.exists, cc.start, cc.list, st["sc_d"])
        for k in range(num_steps):
            t = k + 1
            _lib.call("pp_prox_step", *static, st["mask"], st["z"], st["x"], st["m"], st["v"], 1, n, 1, None, cc.tol,
                      cc.max_cut, float(lamda), lr / (1 - beta1 ** t), math.sqrt(1 - beta2 ** t), beta1, beta2, eps,
                      st["snap"], st["losses"][k - 1] if k else None, st["per_res"], ws["atoms4"], ws["axes"],
                      ws["bound"], st["loss_rows"], self.owned_local, self.L)
            torch.index_select(st["snap"], 0, self.own_pos, out=st["send"][:len(self.own_pos)])
            if self.world > 1:
                dist.all_gather_into_tensor(st["all"][k], st["send"])
                if len(self.halo_pos):  # halo copies follow their owners
                    st["sc_d"].index_copy_(0, self.halo_pos, st["all"][k].index_select(0, self.halo_src))
            else:
                st["all"][k].copy_(st["send"])
        _lib.call("pp_prox_loss", st["loss_rows"], 1, n, 1, None, float(lamda), self.L, st["losses"][num_steps - 1])

    def run(self, SC_D, lamda, num_steps=50, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8, graph=True):
        """-> (snapshots [num_steps, L, 4], losses [num_steps]) like ClashContext.proximal, for the whole complex."""
        from . import _lib
        cc, dev, n = self.cc, self.dev, len(self.local)
        key = (int(num_steps), float(lamda), lr, beta1, beta2, eps)
        st = self._state.get(key)
        if st is None:
            f = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
            st = dict(sc_d=f(n, 4), mask=torch.zeros(n, 4, dtype=torch.uint8, device=dev), z=f(n, 4), x=f(n, 4),
                      m=f(n, 4), v=f(n, 4), per_res=f(n), snap=f(n, 4), loss_rows=f(n, 2), losses=f(num_steps, 1, 2),
                      send=f(self.cap, 4), all=f(num_steps, self.world * self.cap, 4), calls=0, graph=None)
            self._state = {key: st}
        sc_full = SC_D.reshape(self.L, 4).to(torch.float32)
        st["sc_d"].copy_(sc_full[self.loc])     # starting angles of the local set; halo rows are refreshed per step
        per_res, _ = cc.evaluate(st["sc_d"])
        tot = per_res[self.own_pos].sum().reshape(1)
        if self.world > 1:
            dist.all_reduce(tot)
        mean = torch.stack([tot[0] * 0, tot[0] / self.L])
        _lib.call("pp_prox_init_from_mean", per_res, mean, st["sc_d"], self.owned_local, n, st["mask"], st["z"], st["x"],
                  st["m"], st["v"])
        st["calls"] += 1
        args = (st, lamda, num_steps, lr, beta1, beta2, eps)
        if st["calls"] == 1 or not graph:
            self._loop(*args)
        else:
            if st["graph"] is None:
                torch.cuda.synchronize(dev)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, capture_error_mode="thread_local"):
                    self._loop(*args)
                st["graph"] = cg
                # the capture did not execute anything: the state set up above is still the starting state
            st["graph"].replay()
        losses = st["losses"].clone()
        if self.world > 1:
            dist.all_reduce(losses)
        snaps = st["all"].index_select(1, self.slot)   # every residue from its owner, all steps at once
        return snaps, losses[:, 0, 0]
