"""The rounding sequence of the inter-residue dihedral feature (encoder.py:155-174): the plain-C restatement
(oracle/dihedral_rounding.c), which the CUDA kernel repeats operation by operation, against torch CPU itself.
The cosine that enters arccos and the sign of the triple product must be BIT-identical: at near-planar geometry one
ulp decides between NaN -> 0 and ~pi, or between +pi and -pi."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import msc_oracle as mo

ORACLE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle")


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(ORACLE_DIR, "build", "liboracle_c.so"))
    lib.pp_oracle_dihedral.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_long] + [ctypes.c_void_p] * 3
    return lib


def c_dihedral(lib, p):
    p = [x.contiguous().float() for x in p]
    n = p[0].shape[0]
    out = [torch.empty(n) for _ in range(3)]
    lib.pp_oracle_dihedral(*[x.data_ptr() for x in p], n, *[o.data_ptr() for o in out])
    return out


def torch_parts(p0, p1, p2, p3):
    """the reference's own expression, split so that the cosine and the sign are visible"""
    u0, u1, u2 = p2 - p1, p0 - p1, p3 - p2
    unit = lambda v: torch.nan_to_num(v / torch.norm(v, dim=-1, keepdim=True))  # noqa: E731
    n1, n2 = unit(torch.cross(u0, u1, dim=-1)), unit(torch.cross(u0, u2, dim=-1))
    sgn = torch.sign((torch.cross(u1, u2, dim=-1) * u0).sum(-1))
    return (n1 * n2).sum(-1), sgn


def near_planar(n, seed):
    """four points whose dihedral is within ~1e-4 rad of 0 or pi (or exactly planar), random pose and scale"""
    g = torch.Generator().manual_seed(seed)
    p1 = torch.randn(n, 3, generator=g) * 10
    ax = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    a = torch.nn.functional.normalize(torch.linalg.cross(ax, torch.randn(n, 3, generator=g)), dim=-1)
    b = torch.linalg.cross(ax, a)
    p2 = p1 + ax * (1.2 + torch.rand(n, 1, generator=g))
    delta = (torch.randn(n, 1, generator=g) * 1e-4) * (torch.rand(n, 1, generator=g) > 0.3)
    flip = torch.where(torch.rand(n, 1, generator=g) > 0.5, 1.0, -1.0)
    p0 = p1 + a * 1.4 + ax * torch.randn(n, 1, generator=g)
    p3 = p2 + flip * (a * torch.cos(delta) + b * torch.sin(delta)) * 1.3 + ax * torch.randn(n, 1, generator=g)
    return p0, p1, p2, p3


@pytest.mark.parametrize("kind", ["random", "planar"])
def test_c_restatement_is_bit_exact_against_torch(clib, kind):
    n = 120000
    if kind == "random":
        g = torch.Generator().manual_seed(5)
        p = [torch.randn(n, 3, generator=g) * 6 for _ in range(4)]
    else:
        p = near_planar(n, 7)
    cos_t, sgn_t = torch_parts(*p)
    cos_c, sgn_c, ang_c = c_dihedral(clib, p)
    assert np.array_equal(cos_t.numpy().view(np.int32), cos_c.numpy().view(np.int32))
    assert torch.equal(sgn_t, sgn_c)
    ang_t = mo._dihedral(*p)
    # arccos itself: Sleef (torch) vs libm, a few ulp of a value <= pi
    assert (ang_t - ang_c).abs().max().item() < 2e-6
    if kind == "planar":  # the cases this is about do occur in the sample
        assert int((cos_t.abs() > 1).sum()) > 100 and int((ang_t.abs() > 3.14).sum()) > 100
