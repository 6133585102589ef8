"""packppi_b200: B200-native (sm_100a) kernels for PackPPI-MSC reverse-diffusion sampling and PackPPI-Prox.

Public surface = the reference's own names on this path:
    TDiffusionModule, ProteinEncoder, MpnnNet          (model.py)
    get_atom14_coords, compute_residue_clash, find_clash_mask, proximal_optimizer   (components.py)
    protein_to_batch (host) / proteins_to_batch_device (csrc/featurize.cu): ComplexDataset.prot_to_data + collate_fn
Importing the package needs no GPU; calling a kernel without the built extension or off a sm_100 device raises.
"""
from .batch import ComplexBatch, collate  # noqa: F401
from .featurize import protein_to_batch, proteins_to_batch_device  # noqa: F401
from .components import (compute_residue_clash, find_clash_mask, get_atom14_coords,  # noqa: F401
                         host_staging, proximal_optimizer)
from .model import MpnnNet, ProteinEncoder, TDiffusionModule  # noqa: F401

__version__ = "0.1.0"
