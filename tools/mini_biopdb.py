"""A few dozen lines of Bio.PDB, enough to run the UNMODIFIED reference callers (test / baseline infrastructure).

Biopython is not installed in this image, and the reference's file readers (`src/utils/protein.py:78-81`,
`src/utils/interface.py:20-32`) delegate to it.  This module provides the pieces they touch - `PDBParser`
(ATOM / HETATM records -> structure / model / chain / residue / atom objects, alternate locations resolved to the
highest occupancy, first one on ties, like Biopython's DisorderedAtom), `NeighborSearch.search_all(radius, "R")` and
`Selection` - with Biopython's attribute names.  tools/ref_shims.py registers it as `Bio.PDB`.  Not product code.
"""
import numpy as np


class Atom:
    def __init__(self, name, coord, bfactor, occupancy, altloc, parent):
        self.name, self.coord, self.bfactor, self.occupancy, self.altloc, self.parent = \
            name, coord, bfactor, occupancy, altloc, parent

    def get_parent(self):
        return self.parent


class _Entity:
    def __init__(self, id_, parent=None):
        self.id, self.parent, self.child_list, self.child_dict = id_, parent, [], {}

    def __iter__(self):
        return iter(self.child_list)

    def __len__(self):
        return len(self.child_list)

    def __getitem__(self, key):
        return self.child_dict[key]

    def add(self, child, key):
        self.child_list.append(child)
        self.child_dict[key] = child

    def get_parent(self):
        return self.parent


class Residue(_Entity):
    def __init__(self, id_, resname, parent):
        super().__init__(id_, parent)
        self.resname = resname

    def get_atoms(self):
        return iter(self.child_list)


class Chain(_Entity):
    def get_residues(self):
        return iter(self.child_list)

    def get_atoms(self):
        for r in self.child_list:
            yield from r.child_list


class Model(_Entity):
    def get_chains(self):
        return iter(self.child_list)


class Structure(_Entity):
    def get_models(self):
        return iter(self.child_list)

    def get_chains(self):
        for m in self.child_list:
            yield from m.child_list


class PDBParser:
    def __init__(self, QUIET=True, **kw):
        pass

    def get_structure(self, name, file):
        lines = open(file).read().splitlines() if isinstance(file, str) else file.read().splitlines()
        s = Structure(name)
        model = Model(0, s)
        s.add(model, 0)
        for line in lines:
            rec = line[:6]
            if rec.startswith("ENDMDL"):
                break  # first model only is ever used by the reference
            if not (rec.startswith("ATOM") or rec.startswith("HETATM")):
                continue
            line = line.ljust(80)
            cid = line[21]
            het = " " if rec.startswith("ATOM") else ("W" if line[17:20] == "HOH" else "H_" + line[17:20].strip())
            rid = (het, int(line[22:26]), line[26])
            if cid not in model.child_dict:
                model.add(Chain(cid, model), cid)
            chain = model[cid]
            if rid not in chain.child_dict:
                chain.add(Residue(rid, line[17:20].strip(), chain), rid)
            res = chain[rid]
            aname = line[12:16].strip()
            occ = float(line[54:60]) if line[54:60].strip() else 0.0
            bf = float(line[60:66]) if line[60:66].strip() else 0.0
            atom = Atom(aname, np.array([float(line[30:38]), float(line[38:46]), float(line[46:54])], np.float32), bf,
                        occ, line[16], res)
            prev = res.child_dict.get(aname)
            if prev is None:
                res.add(atom, aname)
            elif atom.altloc != " " and prev.altloc != " " and occ > prev.occupancy:
                res.child_list[res.child_list.index(prev)] = atom
                res.child_dict[aname] = atom
        return s


class NeighborSearch:
    """search_all(radius, "R"): unordered pairs of distinct residues with any two atoms within `radius`."""

    def __init__(self, atoms):
        self.atoms = list(atoms)

    def search_all(self, radius, level="A"):
        from scipy.spatial import cKDTree
        xyz = np.array([a.coord for a in self.atoms], np.float64)
        pairs = cKDTree(xyz).query_pairs(radius, output_type="ndarray")
        if level == "A":
            return [(self.atoms[i], self.atoms[j]) for i, j in pairs]
        seen, out = set(), []
        for i, j in pairs:
            ri, rj = self.atoms[i].parent, self.atoms[j].parent
            if ri is rj:
                continue
            key = (id(ri), id(rj)) if id(ri) < id(rj) else (id(rj), id(ri))
            if key not in seen:
                seen.add(key)
                out.append((ri, rj))
        return out


class Selection:
    pass
