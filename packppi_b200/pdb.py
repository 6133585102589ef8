"""Minimal PDB reader / writer for the hot path's on-disk format.

`read_pdb` follows the ordering rules of the reference `from_pdb_file / from_pdb_string`
(src/utils/protein.py:55-199), whose parsing is delegated to Biopython's PDBParser:
only lines starting with "ATOM" are read; chains are visited sorted by id and residues stably sorted by
resseq; HOH and non-standard residues are skipped; MSE -> MET; every residue with an insertion code bumps
a running offset that is added to the residue number (never reset between chains); duplicate numbers
inside a chain are moved to the next free number; atoms are placed in atom14 slots by name, missing atoms
are NaN; of alternate locations the highest occupancy wins, the first one on ties.
"""
import numpy as np

from . import tables


def read_pdb(path, mse_to_met=True):
    names3 = tables.names()["atom14_names"]
    three2one = tables.restype_3to1()
    order = {r: i for i, r in enumerate(tables.restypes())}
    chains = {}
    with open(path) as f:
        for line in f:
            if not line.startswith("ATOM"):
                continue
            line = line.strip()
            chain = line[21]
            key = (int(line[22:26]), line[26] if len(line) > 26 else " ")
            resname = line[17:20].strip()
            res = chains.setdefault(chain, {}).setdefault(key, {"resname": resname, "atoms": {}})
            name = line[12:16].strip()
            alt = line[16]
            occ = float(line[54:60]) if len(line) >= 60 and line[54:60].strip() else 0.0
            xyz = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
            b = float(line[60:66]) if len(line) >= 66 and line[60:66].strip() else 0.0
            prev = res["atoms"].get(name)
            if prev is None:
                res["atoms"][name] = (xyz, occ, b, alt)
            elif alt != " " and prev[3] != " " and occ > prev[1]:
                res["atoms"][name] = (xyz, occ, b, alt)

    pos, aa, mask, ridx, cids, bfs = [], [], [], [], [], []
    offset = 0
    for chain in sorted(chains):
        for key in sorted(chains[chain], key=lambda k: k[0]):
            res = chains[chain][key]
            resname = res["resname"]
            atoms = res["atoms"]
            if resname == "HOH":
                continue
            if mse_to_met and resname == "MSE":
                resname = "MET"
                if "SE" in atoms:
                    atoms["SD"] = atoms.pop("SE")
            short = three2one.get(resname, "X")
            if short == "X":
                continue
            if key[1] != " ":
                offset += 1
            slots = names3[resname]
            p = np.full((14, 3), np.nan)
            m = np.zeros(14)
            bf = np.zeros(14)
            for name, (xyz, _, b, _) in atoms.items():
                if name not in slots:
                    continue
                i = slots.index(name)
                p[i] = np.asarray(xyz, np.float32)
                m[i] = 1.0
                bf[i] = b
            if m.sum() < 0.5:
                continue
            aa.append(order[short])
            pos.append(p)
            mask.append(m)
            ridx.append(key[0] + offset)
            cids.append(chain)
            bfs.append(bf)

    used, new_idx = {}, []
    for c, i in zip(cids, ridx):
        s = used.setdefault(c, set())
        while i in s:
            i += 1
        s.add(i)
        new_idx.append(i)
    return dict(atom_positions=np.array(pos), atom_mask=np.array(mask), aaindex=np.array(aa),
                residue_index=np.array(new_idx), chain_id=np.array(cids), b_factors=np.array(bfs))


def to_pdb(prot, keep_chains=None):
    """PDB text of a protein record, byte for byte what the reference's `to_pdb` writes (src/utils/protein.py:207-314;
    pinned by tests/golden/pdb_text.json): `MODEL     1`, one 80-column ATOM line per atom whose mask is >= 0.5 in
    atom14 slot order, a TER record whenever the chain id changes and after the last residue (it takes a serial
    number), `ENDMDL`, `END`.  Serial numbers start at 1, occupancy is 1.00, the element is the first letter of the
    atom name.  `prot`: dict (or object) with atom_positions [L,14,3], atom_mask, aaindex, residue_index, chain_id,
    b_factors; `keep_chains` restricts the output to those chain ids."""
    get = (lambda k: prot[k]) if isinstance(prot, dict) else (lambda k: getattr(prot, k))
    fields = {k: np.asarray(get(k)) for k in ("atom_mask", "aaindex", "atom_positions", "residue_index", "chain_id",
                                              "b_factors")}
    rt = tables.restypes()
    if np.any(fields["aaindex"] > len(rt)):
        raise ValueError("Invalid aaindexs.")
    if keep_chains is not None:
        sel = np.isin(fields["chain_id"], keep_chains)
        fields = {k: v[sel] for k, v in fields.items()}
    if fields["atom_positions"].shape[-2] != 14:
        raise ValueError("Invalid number of atoms per residue.")
    one2three = tables.names()["restype_1to3"]
    names3 = tables.names()["atom14_names"]
    res3 = [one2three.get(rt[int(a)], "UNK") if int(a) < len(rt) else "UNK" for a in fields["aaindex"]]
    chain, rnum = fields["chain_id"], fields["residue_index"]

    def ter(serial, i):
        return f"{'TER':<6}{serial:>5}      {res3[i]:>3} {chain[i]:>1}{rnum[i]:>4}"

    lines, serial = ["MODEL     1"], 1
    for i in range(len(res3)):
        if i > 0 and chain[i] != chain[i - 1]:
            lines.append(ter(serial, i - 1))
            serial += 1
        for name, xyz, m, b in zip(names3[res3[i]], fields["atom_positions"][i], fields["atom_mask"][i],
                                   fields["b_factors"][i]):
            if m < 0.5:
                continue
            shown = name if len(name) == 4 else f" {name}"
            lines.append(f"{'ATOM':<6}{serial:>5} {shown:<4}{'':>1}{res3[i]:>3} {chain[i]:>1}{rnum[i]:>4}{'':>1}   "
                         f"{xyz[0]:>8.3f}{xyz[1]:>8.3f}{xyz[2]:>8.3f}{1.0:>6.2f}{b:>6.2f}          {name[0]:>2}{'':>2}")
            serial += 1
    lines.append(ter(serial, len(res3) - 1))
    lines += ["ENDMDL", "END"]
    return "\n".join(line.ljust(80) for line in lines) + "\n"


def write_pdb(protein, path=None):
    """ATOM/TER/END records for atom14 coordinates (same record layout as protein.py:207-314)."""
    rt = tables.restypes()
    one2three = tables.names()["restype_1to3"]
    names3 = tables.names()["atom14_names"]
    pos = np.asarray(protein["atom_positions"])
    mask = np.asarray(protein["atom_mask"])
    lines, serial = [], 1
    L = pos.shape[0]
    for i in range(L):
        res3 = one2three[rt[int(protein["aaindex"][i])]]
        chain = str(protein["chain_id"][i])
        rnum = int(protein["residue_index"][i])
        for a, name in enumerate(names3[res3]):
            if not name or mask[i, a] < 0.5 or not np.all(np.isfinite(pos[i, a])):
                continue
            nm = name if len(name) == 4 else f" {name}"
            b = float(protein["b_factors"][i, a]) if "b_factors" in protein else 0.0
            x, y, z = pos[i, a]
            lines.append(f"ATOM  {serial:>5} {nm:<4} {res3:>3} {chain:>1}{rnum:>4}    "
                         f"{x:>8.3f}{y:>8.3f}{z:>8.3f}{1.0:>6.2f}{b:>6.2f}          {name[0]:>2}  ")
            serial += 1
        last = i == L - 1 or str(protein["chain_id"][i + 1]) != chain
        if last:
            lines.append(f"TER   {serial:>5}      {res3:>3} {chain:>1}{rnum:>4}")
            serial += 1
    lines.append("END")
    text = "\n".join(lines) + "\n"
    if path is not None:
        with open(path, "w") as f:
            f.write(text)
    return text
