"""Synthetic protein complexes of a named size (BASELINE.json configs 3-5).

Backbone: per chain a self-avoiding CA walk (3.8 A steps, >= 4.2 A between non-bonded CA, virtual bond angle
85-150 deg) confined to a sphere of radius 4.29 * N^(1/3) A, which puts the 32nd-neighbour CA distance at
10-13 A like the real fixtures; N / C / O are placed in the plane of consecutive CA triplets with ideal
peptide bond lengths.  Residue types are uniform over the 20 standard ones, chi ~ U(-pi, pi),
`atom_mask` = ideal atom14 mask, `SC_D_mask` = chi_angles_mask[type], residue numbers 1..n per chain with
the reference's inter-chain offset rule (complex_dataset.py:86-92).  Side-chain slots of X are left at
zero unless `place_side_chains` is given (callable X,S,BB_D,SC_D -> atom14), because the sampling path only
reads the four backbone slots.
"""
import math

import numpy as np
import torch

from . import tables
from .batch import ComplexBatch
from .featurize import calc_bb_dihedrals


def _unit(v):
    return v / np.linalg.norm(v)


def _ca_walk(n, rng, occupied, grid, radius, cell=4.2, min_sep=4.2):
    def key(p):
        return tuple(np.floor(p / cell).astype(np.int64))

    def free(p, skip):
        k = key(p)
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for idx in grid.get((k[0] + dx, k[1] + dy, k[2] + dz), ()):
                        if idx in skip:
                            continue
                        if np.sum((occupied[idx] - p) ** 2) < min_sep * min_sep:
                            return False
        return True

    def push(p):
        occupied.append(p)
        grid.setdefault(key(p), []).append(len(occupied) - 1)

    def pop():
        p = occupied.pop()
        grid[key(p)].remove(len(occupied))

    for _ in range(2000):  # start point
        p0 = rng.normal(size=3)
        p0 = _unit(p0) * radius * rng.random() ** (1 / 3)
        if free(p0, ()):
            break
    start = len(occupied)
    push(p0)
    stuck = 0
    while len(occupied) - start < n:
        m = len(occupied) - start
        prev = occupied[-1]
        placed = False
        for _ in range(60):
            d = _unit(rng.normal(size=3))
            if m >= 2:
                back = _unit(occupied[-2] - prev)
                cosang = float(np.dot(d, back))  # virtual bond angle CA(i-1)-CA(i)-CA(i+1)
                if not (math.cos(math.radians(150)) <= cosang <= math.cos(math.radians(85))):
                    continue
            p = prev + 3.8 * d
            if np.linalg.norm(p) > radius:
                continue
            if free(p, (len(occupied) - 1,)):
                push(p)
                placed = True
                break
        if not placed:
            stuck += 1
            for _ in range(min(m - 1, 5 + stuck % 20)):  # back-track
                pop()
            if stuck > 20000:
                raise RuntimeError("synthetic CA walk did not converge")
    return np.asarray(occupied[start:])


def _backbone_from_ca(ca, rng):
    """N, C, O in the plane of (CA_i, CA_i+1, bisector normal); bond lengths N-CA 1.46, CA-C 1.52, C-O 1.23."""
    n = len(ca)
    N = np.zeros((n, 3))
    C = np.zeros((n, 3))
    O = np.zeros((n, 3))
    for i in range(n):
        nxt = ca[i + 1] - ca[i] if i + 1 < n else ca[i] - ca[i - 1]
        prv = ca[i - 1] - ca[i] if i > 0 else ca[i] - ca[i + 1]
        ex = _unit(nxt)
        nrm = np.cross(ex, prv)
        if np.linalg.norm(nrm) < 1e-3:
            nrm = np.cross(ex, rng.normal(size=3))
        ez = _unit(nrm)
        ey = np.cross(ez, ex)
        # C_i leaves CA_i 20.7 deg off the CA-CA axis; N_i arrives 14.6 deg off the previous CA-CA axis
        C[i] = ca[i] + 1.52 * (math.cos(math.radians(20.7)) * ex + math.sin(math.radians(20.7)) * ey)
        # N_i: 1.46 A from CA_i at the ideal N-CA-C angle (111 deg), in the plane, on the side of CA_(i-1)
        uc = _unit(C[i] - ca[i])
        perp = prv - np.dot(prv, uc) * uc
        if np.linalg.norm(perp) < 1e-3:
            perp = ez
        N[i] = ca[i] + 1.46 * (math.cos(math.radians(111.0)) * uc + math.sin(math.radians(111.0)) * _unit(perp))
        O[i] = C[i] + 1.23 * _unit(math.cos(math.radians(60)) * ex + math.sin(math.radians(60)) * ey + 0.2 * ez)
    return N, C, O


def make_complex(chain_lengths, seed=0, place_side_chains=None):
    """-> ComplexBatch of one complex ([1,L,...], L = sum(chain_lengths))."""
    rng = np.random.default_rng(seed)
    t = tables.raw()
    L = int(sum(chain_lengths))
    radius = 4.29 * L ** (1.0 / 3.0)
    occupied, grid = [], {}
    X = np.zeros((L, 14, 3), np.float32)
    ridx = np.zeros(L, np.int64)
    chain = np.zeros(L, np.int64)
    o = 0
    offset = 0
    for c, n in enumerate(chain_lengths):
        ca = _ca_walk(n, rng, occupied, grid, radius)
        N, C, O = _backbone_from_ca(ca, rng)
        X[o:o + n, 0], X[o:o + n, 1], X[o:o + n, 2], X[o:o + n, 3] = N, ca, C, O
        ridx[o:o + n] = np.arange(1, n + 1) + offset
        offset += n + 100  # running max of the previous chain + 100
        chain[o:o + n] = c + 1
        o += n
    S = rng.integers(0, 20, size=L).astype(np.int64)
    chi_mask = t["chi_angles_mask"][S].astype(np.float32)
    SC_D = (rng.uniform(-math.pi, math.pi, size=(L, 4)).astype(np.float32)) * chi_mask
    atom_mask = t["atom14_ideal_mask"][S].astype(np.float32)

    Xt = torch.from_numpy(X)
    ridx_t = torch.from_numpy(ridx)
    BB_D, BB_m = calc_bb_dihedrals(Xt, ridx_t)
    BB_D = torch.nan_to_num(BB_D)
    SC_D_t = torch.from_numpy(SC_D)
    SC_m = torch.from_numpy(chi_mask) * (SC_D_t != 0).float()
    p1 = torch.from_numpy(t["chi_pi_periodic"])[torch.from_numpy(S)].bool()
    if place_side_chains is not None:
        Xt = place_side_chains(Xt[None], torch.from_numpy(S)[None], BB_D[None], SC_D_t[None])[0]
        Xt = Xt * torch.from_numpy(atom_mask)[..., None]
    b = ComplexBatch(
        num_nodes=L,
        X=Xt,
        atom_mask=torch.from_numpy(atom_mask),
        residue_type=torch.from_numpy(S),
        residue_mask=torch.ones(L),
        residue_index=ridx_t,
        chain_indices=torch.from_numpy(chain),
        BB_D=BB_D * BB_m,
        BB_D_sincos=torch.stack((torch.sin(BB_D), torch.cos(BB_D)), -1) * BB_m[..., None],
        BB_D_mask=BB_m,
        SC_D=SC_D_t,
        SC_D_sincos=torch.stack((torch.sin(SC_D_t), torch.cos(SC_D_t)), -1) * SC_m[..., None],
        SC_D_mask=SC_m,
        chi_1pi_periodic_mask=torch.logical_and(SC_m.bool(), p1),
        chi_2pi_periodic_mask=torch.logical_and(SC_m.bool(), ~p1),
    )
    for k, v in list(b.items()):
        if torch.is_tensor(v):
            b[k] = v.unsqueeze(0).contiguous()
    b["num_proteins"] = 1
    b["max_size"] = L
    return b


def sweep_lengths(n_complexes=64, lo=200, hi=800, seed=64):
    """Config 5: L_c ~ U{lo..hi}, two chains each."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_complexes):
        L = int(rng.integers(lo, hi + 1))
        a = L // 2
        out.append((a, L - a))
    return out
