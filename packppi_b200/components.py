"""Drop-in mirrors of the reference's free functions on the hot path, backed by libpackppi_b200.so.

  get_atom14_coords(X, S, BB_D, SC_D)                       reference src/models/components/__init__.py:76-120
  compute_residue_clash(batch, SC_D, vtf, cot, eps)         src/models/components/clash.py:335-365
  find_clash_mask / proximal_optimizer                      src/models/components/optimize.py:5-73
Same names, positional order and return types.  CUDA tensors only (RuntimeError otherwise: the reference
remains the `--device cpu` path).  `compute_residue_clash` is differentiable in SC_D through an analytic
backward kernel.

`src/proximal_optimize.py:38-55` never moves its batch off the host.  To run that script unchanged,
`with host_staging("cuda:0"):` makes `proximal_optimizer` and `get_atom14_coords` accept HOST tensors: they are copied
to the named device, the kernels run there, and the results come back as host tensors (a host<->device staging of
the buffers, not a CPU implementation).
"""
import contextlib

import torch

from . import _lib
from .batch import ComplexBatch
from .engine import ClashContext, DeviceTables

_STAGE_DEVICE = None


@contextlib.contextmanager
def host_staging(device="cuda"):
    """Inside this context host tensors passed to proximal_optimizer / get_atom14_coords are staged through `device`."""
    global _STAGE_DEVICE
    prev, _STAGE_DEVICE = _STAGE_DEVICE, torch.device(device)
    if _STAGE_DEVICE.type != "cuda":
        _STAGE_DEVICE = prev
        raise RuntimeError("host_staging: the staging device must be a CUDA device")
    try:
        yield
    finally:
        _STAGE_DEVICE = prev


def _staged(t):
    return _STAGE_DEVICE is not None and torch.is_tensor(t) and not t.is_cuda


def _cuda_only(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"packppi_b200.{what}: CUDA tensors only, there is no CPU fallback "
                           "(use the reference implementation for --device cpu)")


def get_atom14_coords(X, S, BB_D, SC_D):
    """chi angles -> atom14 coordinates [..., L, 14, 3].  BB_D is accepted for signature parity; the omega/phi/psi
    frames only own slots that are overwritten with the input backbone, so it cannot influence the result.
    Leading dimensions of SC_D beyond those of X are treated as samples of the same backbone."""
    if _staged(X):
        d = _STAGE_DEVICE
        return get_atom14_coords(X.to(d), S.to(d), BB_D.to(d), SC_D.to(d)).cpu()
    _cuda_only(X, "get_atom14_coords")
    L = X.shape[-3]
    G = X.numel() // 42
    Xf = X.reshape(G, 14, 3).to(torch.float32).contiguous()
    Sf = S.reshape(-1).to(torch.int64).contiguous()
    chi = SC_D.reshape(-1, 4).to(torch.float32).contiguous()
    if chi.shape[0] % G != 0:
        raise RuntimeError("get_atom14_coords: SC_D does not match X")
    n = chi.shape[0] // G
    out = torch.empty(n * G, 14, 3, dtype=torch.float32, device=X.device)
    _lib.call("pp_atom14_fwd", DeviceTables.get(X.device).geo, Xf, Sf, chi, G, n, out)
    return out.reshape(*SC_D.shape[:-2], L, 14, 3)


_ctx_cache = {}


def clash_context(batch, violation_tolerance_factor=12., clash_overlap_tolerance=0.5):
    """Static neighbour list / tables for `batch`, cached while the same tensors are passed again.

    The key holds the tensors themselves (identity + version counter): data pointers are recycled by the allocator as
    soon as a batch is freed, so a pointer key would serve the neighbour list of the previous complex to the next one
    of the same size."""
    tensors = (batch.X, batch.residue_type, batch.atom_mask, batch.residue_index)
    key = (tensors, tuple(t._version for t in tensors), float(violation_tolerance_factor),
           float(clash_overlap_tolerance))
    hit = _ctx_cache.get("ctx")
    same = (hit is not None and hit[0][1:] == key[1:] and all(a is b for a, b in zip(hit[0][0], key[0])))
    if not same:
        ctx = ClashContext(batch.X.device, batch.X, batch.residue_type, batch.atom_mask, batch.residue_index,
                           violation_tolerance_factor, clash_overlap_tolerance)
        _ctx_cache["ctx"] = (key, ctx)
    return _ctx_cache["ctx"][1]


class _ResidueClash(torch.autograd.Function):
    @staticmethod
    def forward(ctx, SC_D, cc):
        chi = SC_D.detach().reshape(-1, 4).to(torch.float32).contiguous()
        per_res, _ = cc.evaluate(chi)
        ctx.cc = cc
        ctx.save_for_backward(chi)
        ctx.in_shape = SC_D.shape
        return per_res.reshape(SC_D.shape[:-1])

    @staticmethod
    def backward(ctx, grad_out):
        (chi,) = ctx.saved_tensors
        w = grad_out.reshape(-1).to(torch.float32).contiguous()
        _, g = ctx.cc.evaluate(chi, res_w=w)
        return g.reshape(ctx.in_shape), None


def compute_residue_clash(batch, SC_D, violation_tolerance_factor=12., clash_overlap_tolerance=0.5, eps=1e-10):
    """Per-residue structural-violation loss [B, L] of the side chains rebuilt from SC_D (clash.py:335-365)."""
    _cuda_only(SC_D, "compute_residue_clash")
    if eps != 1e-10:
        raise NotImplementedError("compute_residue_clash: eps is fixed to the reference default 1e-10")
    cc = clash_context(batch, violation_tolerance_factor, clash_overlap_tolerance)
    return _ResidueClash.apply(SC_D, cc)


def find_clash_mask(batch, SC_D, violation_tolerance_factor, clash_overlap_tolerance):
    """optimize.py:5-18: residues whose loss exceeds the mean of their own complex, expanded to the four chi ->
    bool [B, L, 4] (the reference handles B = 1; with B > 1 the mean runs over the unpadded residues of each item)."""
    per = compute_residue_clash(batch, SC_D.detach(), violation_tolerance_factor, clash_overlap_tolerance)
    if per.numel() == per.shape[-1]:  # one complex, one sample: the reference's case
        return (per > per.mean()).unsqueeze(-1).expand(*per.shape, 4)
    cc = clash_context(batch, violation_tolerance_factor, clash_overlap_tolerance)
    n_res = cc.n_res()
    n = n_res.to(per.dtype) if n_res is not None else float(per.shape[-1])
    mean = per.sum(-1) / n  # padding rows carry no loss
    return (per > mean.unsqueeze(-1)).unsqueeze(-1).expand(*per.shape, 4)


def proximal_optimizer(batch, SC_D, violation_tolerance_factor, clash_overlap_tolerance, lamda, num_steps=50):
    """optimize.py:21-73 -> (list of num_steps tensors [1, L, 4], list of num_steps floats).

    The whole loop (rebuild, loss, analytic gradient, Adam, snapshot) runs on the device; the losses are copied
    to the host once at the end instead of one `.item()` per step.

    Beyond the reference (which asserts one complex, optimize.py:27, and is looped per decoy by its notebooks): a padded
    batch of B complexes and / or SC_D with a leading sample dimension [S, B, L, 4] optimises all S*B items in the same
    launches, each against the mean of its own residues; the snapshots then have SC_D's shape and every entry of the
    loss list is a tensor [S, B] (or [B]) of per-item objectives, equal to what a per-item call returns."""
    B, L = int(batch.X.shape[0]), int(SC_D.shape[-2])
    if _staged(SC_D):
        d = _STAGE_DEVICE
        on_dev = ComplexBatch(**{k: (batch[k].to(d) if torch.is_tensor(batch[k]) else batch[k]) for k in
                                 ("X", "residue_type", "atom_mask", "residue_index", "BB_D", "num_proteins")})
        snaps, loss_list = proximal_optimizer(on_dev, SC_D.to(d), violation_tolerance_factor,
                                              clash_overlap_tolerance, lamda, num_steps)
        return [s.cpu() for s in snaps], ([x.cpu() for x in loss_list] if torch.is_tensor(loss_list[0]) else loss_list)
    _cuda_only(SC_D, "proximal_optimizer")
    cc = clash_context(batch, violation_tolerance_factor, clash_overlap_tolerance)
    snaps, losses, _ = cc.proximal(SC_D.detach().reshape(-1, 4).to(torch.float32), float(lamda), int(num_steps))
    single = B == 1 and SC_D.numel() == L * 4
    if single:
        loss_list = losses[:, 0].cpu().tolist()
        return [snaps[k].reshape(1, L, 4) for k in range(num_steps)], loss_list
    lead = SC_D.shape[:-2]  # (B,) or (S, B)
    return [snaps[k].reshape(*lead, L, 4) for k in range(num_steps)], [losses[k].reshape(*lead) for k in range(num_steps)]
