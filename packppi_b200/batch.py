"""Attribute bag with the batch contract of the hot path.

The reference uses `torch_geometric.data.Data` purely as an attribute container
(src/datamodules/components/complex_dataset.py:123-139, src/utils/protein_analysis.py:115-120);
this class offers the same access patterns (`batch.X`, `batch['X']`, `.to(device)`, `.keys()`)
without the dependency.  Any object exposing the same attributes is accepted by the model.
"""
import torch

TENSOR_FIELDS = ("X", "atom_mask", "residue_type", "residue_mask", "residue_index", "chain_indices", "BB_D",
                 "BB_D_sincos", "BB_D_mask", "SC_D", "SC_D_sincos", "SC_D_mask", "chi_1pi_periodic_mask",
                 "chi_2pi_periodic_mask")


class ComplexBatch(dict):
    def __init__(self, **kw):
        super().__init__(**kw)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to(self, device, non_blocking=False):
        out = ComplexBatch()
        for k, v in self.items():
            out[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        return out

    def pin_memory(self):
        out = ComplexBatch()
        for k, v in self.items():
            out[k] = v.pin_memory() if torch.is_tensor(v) else v
        return out

    def apply(self, fn):
        for k, v in list(self.items()):
            self[k] = fn(v)
        return self

    def clone(self):
        return ComplexBatch(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.items()})

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in self.values() if torch.is_tensor(v))


def collate(items):
    """Pad-and-stack a list of single-complex batches ([1,L_c,...]) to [B,L_max,...].

    Same result as the reference `collate_fn` (src/datamodules/complex_datamodule.py:196-226):
    zero padding at the end of every per-residue tensor, `num_proteins = B`, `max_size = L_max`.
    """
    L = max(int(it["X"].shape[-3]) for it in items)
    out = ComplexBatch()
    for k in items[0].keys():
        v0 = items[0][k]
        if not torch.is_tensor(v0):
            continue
        rows = []
        for it in items:
            v = it[k]
            v = v[0] if v.shape[0] == 1 and v.dim() == v0.dim() else v
            pad = L - v.shape[0]
            if pad:
                z = torch.zeros((pad,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                v = torch.cat([v, z], 0)
            rows.append(v)
        out[k] = torch.stack(rows, 0)
    out["num_proteins"] = len(items)
    out["max_size"] = L
    out["num_nodes"] = L
    return out
