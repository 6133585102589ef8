"""Lookup tables of the hot path (ideal side-chain geometry, radii, bond statistics).

The numbers come from `packppi_b200/data/tables.npz`, exported by tools/gen_tables.py from the
reference's `src/utils/residue_constants.py` (see that script for the line-by-line provenance).
"""
import functools
import json
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tables.npz")


@functools.lru_cache(maxsize=None)
def raw():
    with np.load(_PATH) as z:
        return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=None)
def names():
    return json.loads(bytes(raw()["names_json"]).decode())


def restypes():
    return names()["restypes"]


def atom14_names(resname3):
    return names()["atom14_names"][resname3]


@functools.lru_cache(maxsize=None)
def restype_3to1():
    return {v: k for k, v in names()["restype_1to3"].items()}


def dist_bounds(overlap_tolerance=0.5, bond_length_tolerance_factor=12.0):
    """Within-residue distance bounds [21,14,14] (lower, upper), float32.

    Same float64 arithmetic, then float32 store, as `make_atom14_dists_bounds`
    (reference src/utils/residue_constants.py:809-869).
    """
    t = raw()
    cot = float(overlap_tolerance)
    vtf = float(bond_length_tolerance_factor)
    lower = np.where(t["pair_named"], t["pair_rsum"] - cot, 0.0)
    upper = np.where(t["pair_named"], 1e10, 0.0)
    lower = np.where(t["bonded"], t["bond_len"] - vtf * t["bond_std"], lower)
    upper = np.where(t["bonded"], t["bond_len"] + vtf * t["bond_std"], upper)
    return lower.astype(np.float32), upper.astype(np.float32)


@functools.lru_cache(maxsize=None)
def max_reach():
    """Upper bound [21] on the distance CA -> any atom14 slot of the residue type, over all chi.

    |translation| of every rigid group on the chain chi1..chik plus |literature position| is invariant
    under the chi rotations, so the sum bounds the reach rigorously.  Used to size the residue-level
    clash neighbour list; not part of the reference (its clash term is dense).
    """
    t = raw()
    frames, grp, lit, msk = t["default_frames"], t["group_idx"], t["lit_positions"], t["atom14_ideal_mask"]
    out = np.zeros(21, np.float64)
    for r in range(21):
        tn = np.linalg.norm(frames[r, :, :3, 3].astype(np.float64), axis=-1)
        for a in range(14):
            if msk[r, a] == 0:
                continue
            g = int(grp[r, a])
            d = float(np.linalg.norm(lit[r, a].astype(np.float64)))
            if g >= 4:
                d += float(tn[4:g + 1].sum())
            elif g > 0:
                d += float(tn[g])
            out[r] = max(out[r], d)
    return (out + 1e-3).astype(np.float32)


def packed_geometry():
    """One contiguous float32 blob per residue type for the atom14 kernel.

    layout per type (stride GEO_STRIDE floats):
      [0:48)   chi1..chi4 default frames, each 3x4 row-major (rotation | translation)
      [48:90)  literature positions 14 x 3
      [90:104) rigid group id per slot (as float)
      [104:118) ideal atom mask per slot
      [118:132) clash radius per slot
      [132]    max reach
    """
    t = raw()
    G = np.zeros((21, GEO_STRIDE), np.float32)
    G[:, 0:48] = t["default_frames"][:, 4:8, :3, :].reshape(21, 48)
    G[:, 48:90] = t["lit_positions"].reshape(21, 42)
    G[:, 90:104] = t["group_idx"].astype(np.float32)
    G[:, 104:118] = t["atom14_ideal_mask"]
    G[:, 118:132] = t["clash_radius"].astype(np.float32)
    G[:, 132] = max_reach()
    return G


GEO_STRIDE = 136
