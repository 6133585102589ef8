"""Protein arrays -> per-residue batch tensors (host side, torch CPU).

Restates `ComplexDataset.prot_to_data` (reference src/datamodules/components/complex_dataset.py:64-148),
`calc_dihedrals / calc_bb_dihedrals / calc_sc_dihedrals` (src/datamodules/components/helper.py:20-101) and
`ProteinAnalysis.get_prot` (src/utils/protein_analysis.py:103-122, minus the interface mask).  It defines the
batch contract the kernels consume; `proteins_to_batch_device` is the device version.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import tables
from .batch import ComplexBatch


def _normalize(t):
    return torch.nan_to_num(t / torch.norm(t, dim=-1, keepdim=True))


def calc_dihedrals(p, eps=1e-8):
    """helper.py:20-36."""
    u = _normalize(p[..., 1:, :] - p[..., :-1, :])
    u2, u1, u0 = u[..., :-2, :], u[..., 1:-1, :], u[..., 2:, :]
    n2 = _normalize(torch.cross(u2, u1, dim=-1))
    n1 = _normalize(torch.cross(u1, u0, dim=-1))
    c = torch.clamp((n2 * n1).sum(-1), -1 + eps, 1 - eps)
    return torch.sign((u2 * n1).sum(-1)) * torch.acos(c)


def calc_bb_dihedrals(X, residue_index):
    """helper.py:39-74 with use_pre_omega=True: columns (pre-omega, phi, psi)."""
    n = X.shape[0]
    bb = X[:, :3].reshape(3 * n, 3)
    d = calc_dihedrals(bb)
    d = F.pad(d, [1, 2], value=float("nan")).reshape(n, 3)
    pre = torch.cat((torch.tensor([0.0]), (residue_index[1:] - 1 == residue_index[:-1]).float()))
    post = torch.cat(((residue_index[:-1] + 1 == residue_index[1:]).float(), torch.tensor([0.0])))
    m = torch.stack((pre, post, post), dim=-1)
    d[:, 2] = torch.cat((torch.tensor([float("nan")]), d[:-1, 2]))
    d[:, [0, 1, 2]] = d[:, [2, 0, 1]]
    m[:, 1] = m[:, 0]
    m = m * torch.isfinite(d).float()
    return d, m


def calc_sc_dihedrals(X, aatype):
    """helper.py:77-101."""
    t = tables.raw()
    idx = torch.from_numpy(t["chi_atom_indices_atom14"].astype(np.int64))[aatype]
    cm = torch.from_numpy(t["chi_mask_atom14"])[aatype]
    pos = torch.gather(X, -2, idx[..., None].expand(*idx.shape, 3))
    d = torch.nan_to_num(calc_dihedrals(pos)) * cm
    return d, (d != 0.0).float()


def chain_codes(chain_id):
    """1-based chain number in order of first appearance (complex_dataset.py:82-85)."""
    seen, out = {}, []
    for c in chain_id:
        c = str(c)
        if c not in seen:
            seen[c] = len(seen) + 1
        out.append(seen[c])
    return np.asarray(out, np.int64)


def offset_residue_index(ridx, chain):
    """complex_dataset.py:86-92: running offset = max index of the previous chains + 100 (ridx is modified)."""
    uniq = torch.unique(chain)
    if len(uniq) > 1:
        off = 0
        for c in uniq[:-1]:
            off += int(ridx[chain == c].max())
            off += 100
            ridx[chain == c + 1] += off
    return ridx


_DEV_TABLES = {}


def proteins_to_batch_device(proteins, device):
    """The same batch as collate([protein_to_batch(p) for p in proteins]).to(device), featurised ON the device
    (csrc/featurize.cu): the raw atom records are padded on the host, copied once, and one kernel
    writes every field.  Angles agree with the host version to fp32 rounding of the dihedral arithmetic (the acos near
    a planar angle amplifies it: up to ~1e-4 rad there), masks and integer fields exactly."""
    from . import _lib
    device = torch.device(device)
    t = tables.raw()
    B = len(proteins)
    lens = [int(np.asarray(p["aaindex"]).shape[0]) for p in proteins]
    L = max(lens)
    X = np.zeros((B, L, 14, 3), np.float32)
    aa = np.zeros((B, L), np.int64)
    am = np.zeros((B, L, 14), np.float32)
    ri = np.zeros((B, L), np.int64)
    ch = np.zeros((B, L), np.int64)
    for b, p in enumerate(proteins):
        n = lens[b]
        X[b, :n] = np.asarray(p["atom_positions"], np.float32)
        aa[b, :n] = np.asarray(p["aaindex"], np.int64)
        am[b, :n] = np.asarray(p["atom_mask"], np.float32)
        chain = torch.from_numpy(chain_codes(p["chain_id"]))
        ridx = offset_residue_index(torch.from_numpy(np.asarray(p["residue_index"])).to(torch.int64).clone(), chain)
        ri[b, :n] = ridx.numpy()
        ch[b, :n] = chain.numpy()
    if device not in _DEV_TABLES:
        _DEV_TABLES[device] = (torch.from_numpy(t["chi_atom_indices_atom14"].astype(np.int32)).to(device).contiguous(),
                               torch.from_numpy(t["chi_mask_atom14"].astype(np.float32)).to(device).contiguous(),
                               torch.from_numpy(t["chi_pi_periodic"].astype(np.float32)).to(device).contiguous())
    up = lambda a: torch.from_numpy(a).to(device)  # noqa: E731
    f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)  # noqa: E731
    i64 = lambda *s: torch.empty(*s, dtype=torch.int64, device=device)  # noqa: E731
    u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=device)  # noqa: E731
    out = ComplexBatch(X=f32(B, L, 14, 3), atom_mask=f32(B, L, 14), residue_type=i64(B, L), residue_mask=f32(B, L),
                       residue_index=i64(B, L), chain_indices=i64(B, L), BB_D=f32(B, L, 3), BB_D_sincos=f32(B, L, 3, 2),
                       BB_D_mask=f32(B, L, 3), SC_D=f32(B, L, 4), SC_D_sincos=f32(B, L, 4, 2), SC_D_mask=f32(B, L, 4),
                       chi_1pi_periodic_mask=u8(B, L, 4), chi_2pi_periodic_mask=u8(B, L, 4))
    _lib.call("pp_featurize", up(X), up(aa), up(am), up(ri), up(ch), torch.tensor(lens, dtype=torch.int32, device=device),
              B, L, *_DEV_TABLES[device], *[out[k] for k in ("X", "atom_mask", "residue_type", "residue_mask",
                                                             "residue_index", "chain_indices", "BB_D", "BB_D_sincos",
                                                             "BB_D_mask", "SC_D", "SC_D_sincos", "SC_D_mask",
                                                             "chi_1pi_periodic_mask", "chi_2pi_periodic_mask")])
    out["chi_1pi_periodic_mask"] = out["chi_1pi_periodic_mask"].bool()
    out["chi_2pi_periodic_mask"] = out["chi_2pi_periodic_mask"].bool()
    out["num_nodes"], out["num_proteins"], out["max_size"] = L, B, L
    return out


def protein_to_batch(protein):
    """dict(atom_positions[L,14,3], aaindex[L], atom_mask[L,14], residue_index[L], chain_id[L]) -> batch of one.

    Every per-residue tensor gets a leading batch dimension of 1, `num_proteins = 1`, `max_size = L`.
    """
    t = tables.raw()
    X = torch.from_numpy(np.asarray(protein["atom_positions"])).to(torch.float32)
    L = X.shape[0]
    S = torch.from_numpy(np.asarray(protein["aaindex"])).to(torch.int64)
    atom_mask = torch.from_numpy(np.asarray(protein["atom_mask"])).to(torch.float32)
    ridx = torch.from_numpy(np.asarray(protein["residue_index"])).to(torch.int64).clone()
    chain = torch.from_numpy(chain_codes(protein["chain_id"]))

    ridx = offset_residue_index(ridx, chain)

    rmask = torch.isfinite(X[:, :4].sum(dim=(-1, -2))).float()
    BB_D, BB_m = calc_bb_dihedrals(X, ridx)
    SC_D, SC_m = calc_sc_dihedrals(X, S)
    BB_sc = torch.stack((torch.sin(BB_D), torch.cos(BB_D)), -1) * BB_m[..., None]
    SC_sc = torch.stack((torch.sin(SC_D), torch.cos(SC_D)), -1) * SC_m[..., None]
    p1 = torch.from_numpy(t["chi_pi_periodic"])[S].bool()
    p2 = ~p1

    rm1, rm2 = rmask[..., None], rmask[..., None, None]
    out = ComplexBatch(
        num_nodes=L,
        X=X * rm2,
        atom_mask=atom_mask * rm1,
        residue_type=(S * rmask).to(torch.int64),
        residue_mask=rmask,
        residue_index=(ridx * rmask).to(torch.int64),
        chain_indices=(chain * rmask).to(torch.int64),
        BB_D=BB_D * rm1,
        BB_D_sincos=BB_sc * rm2,
        BB_D_mask=BB_m * rm1,
        SC_D=SC_D * rm1,
        SC_D_sincos=SC_sc * rm2,
        SC_D_mask=SC_m * rm1,
    )
    scm = out["SC_D_mask"]
    out["chi_1pi_periodic_mask"] = torch.logical_and(scm, p1 * rm1)
    out["chi_2pi_periodic_mask"] = torch.logical_and(scm, p2 * rm1)
    for k, v in list(out.items()):
        if torch.is_tensor(v):
            out[k] = torch.nan_to_num(v).unsqueeze(0)
    out["num_proteins"] = 1
    out["max_size"] = L
    return out
