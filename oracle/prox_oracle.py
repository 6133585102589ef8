"""CPU ORACLE (test infrastructure, not product code) for chi -> atom14 rebuild, the PackPPI-Prox
structural-violation loss and the proximal Adam loop.

torch-CPU fp32 restatement of the reference (paths under /root/reference/src), differentiable through
torch autograd so that the analytic CUDA gradient can be checked against it.  The between-residue term
exists in two forms: `dense_between` follows the reference's [N,N,14,14] tensor formulation (usable up to
~1500 residues) and `sparse_between` evaluates the same sum over the atom pairs inside the interaction
cutoff only (KD-tree), validated against the dense one in tests/ and trusted at 5000 residues where the
dense tensor would need ~257 GB.  Pinned against the unmodified reference through
tests/golden/*.npz (tools/make_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from packppi_b200 import tables
from .msc_oracle import backbone_frames


def _t(name, dtype=torch.float32):
    return torch.from_numpy(np.asarray(tables.raw()[name])).to(dtype)


# ---------------------------------------------------------------------------------------------- atom14
def atom14_coords(X, S, BB_D, SC_D, return_frames=False):
    """models/components/__init__.py:76-120, utils/features.py:95-194.

    Only chi frames matter for the output: slots 0-3 are overwritten with the input backbone and CB sits in
    the backbone frame, so BB_D (omega/phi/psi frames) has no effect.
    F_g = Default_g * Rx(chi_g); chi2..4 are chained; atom = bb o F_group(atom) (lit_pos), times ideal mask.
    """
    R_bb, t_bb = backbone_frames(X)  # [..,3,3], [..,3]
    sc = torch.stack((torch.sin(SC_D), torch.cos(SC_D)), -1)  # [..,4,2]
    sc = sc / torch.sqrt(torch.clamp((sc ** 2).sum(-1, keepdim=True), min=1e-12))
    dflt = _t("default_frames")[S][..., 4:8, :, :]  # [..,4,4,4]
    Rd, td = dflt[..., :3, :3], dflt[..., :3, 3]
    s, c = sc[..., 0], sc[..., 1]
    zero, one = torch.zeros_like(s), torch.ones_like(s)
    Rx = torch.stack([one, zero, zero, zero, c, -s, zero, s, c], -1).reshape(*s.shape, 3, 3)
    Rl = Rd @ Rx  # local chi frames: rotation Default*Rx, translation Default.t
    Rs, ts = [Rl[..., 0, :, :]], [td[..., 0, :]]
    for k in range(1, 4):  # chi_k frame to backbone = chi_{k-1} frame o local_k
        ts.append((Rs[-1] @ td[..., k, :, None])[..., 0] + ts[-1])
        Rs.append(Rs[-1] @ Rl[..., k, :, :])
    Rg = [R_bb] + [R_bb @ r for r in Rs]  # group 0 (backbone), groups 4..7 -> index 1..4
    tg = [t_bb] + [(R_bb @ v[..., None])[..., 0] + t_bb for v in ts]
    Rg, tg = torch.stack(Rg, -3), torch.stack(tg, -2)  # [..,5,3,3], [..,5,3]

    grp = _t("group_idx", torch.int64)[S]  # [..,14]
    gsel = torch.clamp(grp - 3, min=0)  # 0 -> 0, 4..7 -> 1..4 (groups 1..3 only own overwritten slots)
    lit = _t("lit_positions")[S]
    Ra = torch.gather(Rg, -3, gsel[..., None, None].expand(*gsel.shape, 3, 3))
    ta = torch.gather(tg, -2, gsel[..., None].expand(*gsel.shape, 3))
    pos = (Ra @ lit[..., None])[..., 0] + ta
    pos = pos * _t("atom14_ideal_mask")[S][..., None]
    pos = torch.cat([X[..., :4, :], pos[..., 4:, :]], -2)
    if return_frames:
        return pos, Rg[..., 1:, :, :], tg[..., 1:, :]
    return pos


# ---------------------------------------------------------------------------------------------- clash
def clash_radius(S, atom_exists):
    """models/components/clash.py:263-289."""
    return atom_exists * _t("clash_radius")[S]


def within_residue(pos, exists, S, cot, vtf, eps=1e-10):
    """clash.py:7-99 (`within_residue_violations`), per-atom loss sum only."""
    lo, hi = tables.dist_bounds(cot, vtf)
    lo, hi = torch.from_numpy(lo)[S], torch.from_numpy(hi)[S]
    m = exists[..., :, None] * exists[..., None, :] * (1.0 - torch.eye(14))
    bb = torch.zeros(14, 14)
    bb[:4, :4] = 1.0
    m = m * (1.0 - bb)
    d = torch.sqrt(eps + ((pos[..., :, None, :] - pos[..., None, :, :]) ** 2).sum(-1))
    loss = m * (F.relu(lo - d) + F.relu(d - hi))
    return loss.sum(-2) + loss.sum(-1)


def dense_between(pos, exists, radius, ridx, cot, eps=1e-10):
    """clash.py:102-254 (`between_residue_clash_loss`), per-atom loss sum only; pos [N,14,3]."""
    d = torch.sqrt(eps + ((pos[:, None, :, None, :] - pos[None, :, None, :, :]) ** 2).sum(-1))
    m = exists[:, None, :, None] * exists[None, :, None, :]
    bb = torch.zeros(14, 14)
    bb[:4, :4] = 1.0
    m = m * (1.0 - bb)
    m = m * (ridx[:, None, None, None] < ridx[None, :, None, None])
    # C(i)-N(i+1) exclusion (clash.py:171-196): slots 2 and 0 are both backbone, already removed above
    ss = torch.zeros(14, 14)
    ss[5, 5] = 1.0  # the "disulfide" one-hot removes every slot-5/slot-5 pair (clash.py:198-210)
    m = m * (1.0 - ss)
    lb = m * (radius[:, None, :, None] + radius[None, :, None, :])
    e = m * F.relu(lb - cot - d)
    return e.sum(dim=(0, 2)) + e.sum(dim=(1, 3))


def sparse_between(pos, exists, radius, ridx, cot, eps=1e-10):
    """Same sum as `dense_between`, restricted to pairs that can be non-zero: d < r_a + r_b - cot <= 2*r_max - cot."""
    from scipy.spatial import cKDTree

    N = pos.shape[0]
    flat = pos.reshape(-1, 3)
    ex = exists.reshape(-1) > 0
    ids = torch.nonzero(ex)[:, 0]
    cutoff = 2.0 * float(radius.max()) - cot
    out = torch.zeros(N * 14, dtype=pos.dtype)
    if cutoff <= 0 or len(ids) == 0:
        return out.reshape(N, 14)
    tree = cKDTree(flat[ids].detach().numpy().astype(np.float64))
    pairs = tree.query_pairs(cutoff + 1e-3, output_type="ndarray")
    if len(pairs) == 0:
        return out.reshape(N, 14)
    p = ids[torch.from_numpy(pairs[:, 0].astype(np.int64))]
    q = ids[torch.from_numpy(pairs[:, 1].astype(np.int64))]
    ri, rj, a, b = p // 14, q // 14, p % 14, q % 14
    keep = (ridx[ri] != ridx[rj]) & ~((a < 4) & (b < 4)) & ~((a == 5) & (b == 5))
    p, q = p[keep], q[keep]
    d = torch.sqrt(eps + ((flat[p] - flat[q]) ** 2).sum(-1))
    rr = radius.reshape(-1)
    e = F.relu((rr[p] + rr[q]) - cot - d)
    out = out.index_add(0, p, e).index_add(0, q, e)
    return out.reshape(N, 14)


def residue_clash(batch, SC_D, vtf=12.0, cot=0.5, eps=1e-10, sparse=False):
    """clash.py:335-365 (`compute_residue_clash`) -> [B,L]."""
    am = batch["atom_mask"]
    n_sc = am[..., 4:].sum(-1)
    pos = atom14_coords(batch["X"], batch["residue_type"], batch["BB_D"], SC_D)
    rad = clash_radius(batch["residue_type"], am)
    fn = sparse_between if sparse else dense_between
    between = torch.stack([fn(pos[b], am[b], rad[b], batch["residue_index"][b], cot) for b in range(pos.shape[0])])
    per_atom = between + within_residue(pos, am, batch["residue_type"], cot, vtf)
    per_atom = torch.cat([torch.zeros_like(per_atom[..., :4]), per_atom[..., 4:]], -1)
    return per_atom.sum(-1) / (eps + n_sc)


def clash_value_and_grad(batch, SC_D, vtf=12.0, cot=0.5, sparse=False):
    """per-residue loss [B,L] and d(sum of per-residue loss)/d(SC_D) by autograd."""
    x = SC_D.clone().requires_grad_(True)
    per_res = residue_clash(batch, x, vtf, cot, sparse=sparse)
    (g,) = torch.autograd.grad(per_res.sum(), x)
    return per_res.detach(), g


def proximal(batch, SC_D, vtf, cot, lamda, num_steps=50, sparse=False):
    """models/components/optimize.py:5-73: clash mask = per-residue loss above its mean; 50 Adam steps
    (lr 1e-2, betas 0.9/0.999, eps 1e-8) on f(x) = mean_res |x' - z|^2 + lamda * mean_res clash(x')."""
    assert batch["num_proteins"] == 1
    with torch.no_grad():
        pr = residue_clash(batch, SC_D, vtf, cot, sparse=sparse)
        mask = (pr > pr.mean())[..., None].expand(-1, -1, 4)
    z = SC_D * mask

    def f(x):
        x = torch.where(mask, x * mask, SC_D)
        return ((x - z).abs() ** 2).sum(-1).mean() + lamda * residue_clash(batch, x, vtf, cot, sparse=sparse).mean()

    x = z.clone().requires_grad_(True)
    opt = torch.optim.Adam([x], lr=1e-2)
    snaps, losses = [], []
    for _ in range(num_steps):
        opt.zero_grad()
        loss = f(x)
        loss.backward()
        opt.step()
        snaps.append(torch.where(mask, x.detach().clone(), SC_D))
        losses.append(loss.item())
    return snaps, losses, mask
