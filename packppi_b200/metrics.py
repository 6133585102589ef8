"""Evaluation metrics on the two sides of the sampling path: the arithmetic of
`ProteinAnalysis.get_metric` (reference src/utils/protein_analysis.py:53-88) and the interface mask it weights with
(`get_interface_mask`, src/datamodules/components/helper.py:104-129 -> `get_interface_residues`,
src/utils/interface.py:11-55).  Host code: it runs once per structure, next to file I/O.

The reference needs Biopython (`NeighborSearch`) for the interface and MolProbity for the clashscore; here the interface
search is a k-d tree over the file's atoms and the clashscore is an argument (None = not measured; the reference returns
no metrics at all in that case, protein_analysis.py:48-51).
"""
import numpy as np
import torch

from .components import get_atom14_coords


def interface_residues(pdb_file, radius=10.0):
    """{chain id: sorted residue numbers} of residues with ANY atom within `radius` of an atom of another protein chain,
    or None with fewer than two protein chains (interface.py:11-55).  Like the reference's Biopython structure this
    reads ATOM and HETATM records of the first model, hydrogens included, and keeps the chains that hold at least one
    non-hetero residue."""
    from scipy.spatial import cKDTree
    xyz, chain, resseq, hetero = [], [], [], []
    with open(pdb_file) as f:
        for line in f:
            rec = line[:6]
            if rec.startswith("ENDMDL"):
                break
            if not (rec.startswith("ATOM") or rec.startswith("HETATM")):
                continue
            xyz.append((float(line[30:38]), float(line[38:46]), float(line[46:54])))
            chain.append(line[21])
            resseq.append((int(line[22:26]), line[26:27], rec.startswith("HETATM")))
            hetero.append(rec.startswith("HETATM"))
    chain = np.array(chain)
    hetero = np.array(hetero)
    protein_chains = [c for c in dict.fromkeys(chain.tolist()) if np.any(~hetero[chain == c])]
    if len(protein_chains) < 2:
        return None
    keep = np.isin(chain, protein_chains)
    xyz = np.asarray(xyz, np.float64)[keep]
    chain = chain[keep]
    resid = [r for r, k in zip(resseq, keep) if k]
    out = {c: set() for c in protein_chains}
    for i, j in cKDTree(xyz).query_pairs(radius, output_type="ndarray"):
        if chain[i] != chain[j]:
            out[chain[i]].add(resid[i][0])
            out[chain[j]].add(resid[j][0])
    return {c: sorted(v) for c, v in out.items()}


def interface_mask(protein, pdb_file, radius=10.0, as_reference=True):
    """float32 [L] mask of interface residues in the order of `protein` (helper.py:104-129); None for one chain.

    as_reference: in the reference `get_prot` calls `prot_to_data` first, and that adds the inter-chain offset
    (running max + 100, complex_dataset.py:86-92) to `protein["residue_index"]` IN PLACE (the tensor shares the
    array's memory); `get_interface_mask` then matches those offset numbers of every chain after the first against
    the file's own numbering, so such chains keep only accidental matches.  True (default) reproduces that - it is what
    the reference's `interface_acc` is computed with; False compares like with like."""
    from .featurize import chain_codes, offset_residue_index
    chain_id = np.asarray(protein["chain_id"])
    if len(np.unique(chain_id)) == 1:
        return None
    inter = interface_residues(pdb_file, radius)
    ridx = np.asarray(protein["residue_index"]).astype(np.int64)
    if as_reference:
        ridx = offset_residue_index(torch.from_numpy(ridx.copy()), torch.from_numpy(chain_codes(chain_id))).numpy()
    parts = []
    for c in np.unique(chain_id):
        sub = ridx[chain_id == c]
        parts.append(np.isin(sub, inter[c]) if inter is not None and c in inter else np.zeros(len(sub), bool))
    return torch.from_numpy(np.concatenate(parts)).to(torch.float32)


def compute_rmsd(true_coords, pred_coords, atom_mask, residue_mask, eps=1e-6):
    """protein_analysis.py:93-101 (a mean squared deviation: the reference never takes the root)."""
    err = torch.sum((true_coords - pred_coords) ** 2, dim=-1) * atom_mask * residue_mask[..., None]
    count = torch.sum(atom_mask * residue_mask[..., None] + eps, dim=-1)
    return torch.sum(torch.sum(err, dim=-1)) / torch.sum(count)


def get_metric(true_data, pred_data, clashscore=None, atom14_fn=None):
    """protein_analysis.py:53-88 on two featurised structures (batches of one complex): chi MAE / accuracy per chi,
    atom RMSD of the side chains rebuilt from the predicted angles on the TRUE backbone, total and interface accuracy.
    `true_data` carries `interface_mask` (zeros if absent).  `atom14_fn` defaults to the CUDA `get_atom14_coords`
    (tensors must then live on the GPU); pass another callable to evaluate host tensors."""
    interface = true_data["interface_mask"] if "interface_mask" in true_data else torch.zeros_like(true_data["residue_mask"])
    chis_true, chis_pred = true_data["SC_D"], pred_data["SC_D"]
    chi_mask, p1 = true_data["SC_D_mask"], true_data["chi_1pi_periodic_mask"]
    metric, total_acc, interface_acc = {}, 0, 0
    for i in range(4):
        n = chi_mask[..., i].sum()
        n = 1 if n == 0 else n
        ni = (chi_mask[..., i] * interface).sum()
        ni = 1 if ni == 0 else ni
        diff = (chis_pred[..., i] - chis_true[..., i]).abs()
        acc = torch.where(torch.logical_and(diff * 180 / np.pi < 20, diff > 0), 1., 0.)
        ae = torch.minimum(diff, 2 * np.pi - diff)
        ae = torch.where(p1[..., i], torch.minimum(ae, np.pi - ae), ae)
        metric[f"chi_{i}_ae_rad"] = ae.sum() / n
        metric[f"chi_{i}_ae_deg"] = (ae * 180 / np.pi).sum() / n
        metric[f"chi_{i}_acc"] = acc.sum() / n
        total_acc = total_acc + acc.sum() / n
        interface_acc = interface_acc + (acc * interface).sum() / ni
    fn = atom14_fn or get_atom14_coords
    xyz = fn(true_data["X"], true_data["residue_type"], true_data["BB_D"], pred_data["SC_D"])
    metric["atom_rmsd"] = compute_rmsd(true_data["X"], xyz, true_data["atom_mask"], true_data["residue_mask"])
    metric["total_acc"] = total_acc / 4
    metric["interface_acc"] = interface_acc / 4
    metric["clashscore"] = clashscore
    return metric
