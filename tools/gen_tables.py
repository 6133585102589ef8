"""Export the reference's numeric lookup tables to packppi_b200/data/tables.npz.

The tables are chemistry DATA (ideal geometry, van-der-Waals radii, bond statistics,
atom14 naming), SURVEY.md §2 row 10: `src/utils/residue_constants.py` and
`src/utils/stereo_chemical_props.py`.  They are exported numerically instead of being
re-typed, so the kernels consume byte-identical values.  Run here (needs /root/reference):

    python tools/gen_tables.py

`tests/test_tables.py` re-derives the same arrays from the reference when it is present
and asserts equality with the committed file.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "packppi_b200", "data", "tables.npz")


def build_tables():
    ref_shims.install()
    import src.utils.residue_constants as rc

    t = {}
    # residue_constants.py:585-677 (_make_rigid_group_constants)
    t["default_frames"] = np.asarray(rc.restype_rigid_group_default_frame, np.float32)  # [21,8,4,4]
    t["group_idx"] = np.asarray(rc.restype_atom14_to_rigid_group, np.int32)  # [21,14]
    t["atom14_ideal_mask"] = np.asarray(rc.restype_atom14_mask, np.float32)  # [21,14]
    t["lit_positions"] = np.asarray(rc.restype_atom14_rigid_group_positions, np.float32)  # [21,14,3]
    # residue_constants.py:507-555
    cam = np.zeros((21, 4), np.float32)
    cam[:20] = np.asarray(rc.chi_angles_mask, np.float32)
    t["chi_angles_mask"] = cam
    t["chi_pi_periodic"] = np.asarray(rc.chi_pi_periodic, np.float32)  # [21,4]
    # residue_constants.py:872-905
    t["chi_atom_indices_atom14"] = np.asarray(rc.chi_atom_indices_atom14, np.int32)  # [21,7]
    t["chi_mask_atom14"] = np.asarray(rc.chi_mask_atom14, np.float32)  # [21,4]

    # clash.py:263-289: per (restype, slot) van-der-Waals radius by first letter of the atom37 name the
    # slot maps to; empty slots map to atom37 index 0 ('N').
    radius = np.zeros((21, 14), np.float64)
    for r, letter in enumerate(rc.restypes):
        names = rc.restype_name_to_atom14_names[rc.restype_1to3[letter]]
        for a, name in enumerate(names):
            a37 = rc.atom_order[name] if name else 0
            radius[r, a] = rc.van_der_waals_radius[rc.atom_types[a37][0]]
    radius[20, :] = rc.van_der_waals_radius[rc.atom_types[0][0]]
    t["clash_radius"] = radius  # float64; cast to f32 exactly as new_tensor() does

    # residue_constants.py:809-869 (make_atom14_dists_bounds) split into its ingredients so that the
    # bounds can be rebuilt for any (overlap_tolerance, bond_length_tolerance_factor) in float64.
    pair_rsum = np.zeros((21, 14, 14), np.float64)
    pair_named = np.zeros((21, 14, 14), np.bool_)
    bond_len = np.zeros((21, 14, 14), np.float64)
    bond_std = np.zeros((21, 14, 14), np.float64)
    bonded = np.zeros((21, 14, 14), np.bool_)
    residue_bonds, residue_virtual_bonds, _ = rc.load_stereo_chemical_props()
    for r, letter in enumerate(rc.restypes):
        resname = rc.restype_1to3[letter]
        names = rc.restype_name_to_atom14_names[resname]
        for a, n1 in enumerate(names):
            if not n1:
                continue
            for b, n2 in enumerate(names):
                if (not n2) or a == b:
                    continue
                # the reference writes [a,b] and [b,a] with the same sum, evaluated as r1 + r2
                s = rc.van_der_waals_radius[n1[0]] + rc.van_der_waals_radius[n2[0]]
                pair_rsum[r, a, b] = s
                pair_rsum[r, b, a] = s
                pair_named[r, a, b] = pair_named[r, b, a] = True
        for bnd in residue_bonds[resname] + residue_virtual_bonds[resname]:
            a = names.index(bnd.atom1_name)
            b = names.index(bnd.atom2_name)
            bond_len[r, a, b] = bond_len[r, b, a] = bnd.length
            bond_std[r, a, b] = bond_std[r, b, a] = bnd.stddev
            bonded[r, a, b] = bonded[r, b, a] = True
    t.update(pair_rsum=pair_rsum, pair_named=pair_named, bond_len=bond_len, bond_std=bond_std, bonded=bonded)

    names = {
        "restypes": list(rc.restypes),
        "restype_1to3": dict(rc.restype_1to3),
        "atom14_names": {k: list(v) for k, v in rc.restype_name_to_atom14_names.items()},
        "atom_types": list(rc.atom_types),
    }
    t["names_json"] = np.frombuffer(json.dumps(names, sort_keys=True).encode(), dtype=np.uint8)
    return t, rc


def rebuild_bounds(t, cot, vtf):
    """Same arithmetic as packppi_b200.tables.dist_bounds (float64, cast to f32 at the end)."""
    lower = np.where(t["pair_named"], t["pair_rsum"] - cot, 0.0)
    upper = np.where(t["pair_named"], 1e10, 0.0)
    lower = np.where(t["bonded"], t["bond_len"] - vtf * t["bond_std"], lower)
    upper = np.where(t["bonded"], t["bond_len"] + vtf * t["bond_std"], upper)
    return lower.astype(np.float32), upper.astype(np.float32)


def main():
    t, rc = build_tables()
    for cot, vtf in [(0.5, 12.0), (1.5, 15.0), (0.3, 7.5)]:
        lo, hi = rebuild_bounds(t, cot, vtf)
        refb = rc.make_atom14_dists_bounds(overlap_tolerance=cot, bond_length_tolerance_factor=vtf)
        assert np.array_equal(lo, refb["lower_bound"]), (cot, vtf)
        assert np.array_equal(hi, refb["upper_bound"]), (cot, vtf)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **t)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
