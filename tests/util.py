"""Shared helpers for the parity tests: golden fixtures -> batches, wrapped angle differences."""
import math
import os

import numpy as np
import torch

from packppi_b200.batch import ComplexBatch, TENSOR_FIELDS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL_CASES = ["syn5", "syn17", "syn31", "syn33", "syn64", "synbatch", "syn300", "1brs", "t1124"]


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        g = {k: z[k] for k in z.files}
    b = ComplexBatch()
    for k in TENSOR_FIELDS:
        b[k] = torch.from_numpy(g["in_" + k])
    b["num_proteins"] = int(b["X"].shape[0])
    b["max_size"] = int(b["X"].shape[1])
    b["num_nodes"] = int(b["X"].shape[1])
    return g, b


def tt(a):
    return torch.from_numpy(np.asarray(a))


def wrapped_diff(a, b):
    d = (a - b).abs()
    return torch.minimum(d, 2 * math.pi - d)


def _ulps(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def knn_mismatches(E_a, D_a, E_b, D_b, valid_rows, ulp=0):
    """Rows whose neighbour lists differ beyond the order of (near-)tied distances.

    torch.topk leaves the order of equal keys unspecified, so: the distance lists must agree within
    `ulp` units in the last place, and inside every run of reference distances whose neighbours are within 2*ulp of
    each other the index SETS must agree; only a run that touches slot K-1 may differ in membership (the tie then
    extends past the K-th neighbour).  ulp=0 demands bit-identical distances (oracle vs reference, both torch CPU);
    ulp=1 is used for the CUDA kernel, whose sqrt is IEEE-correctly rounded while torch-CPU's vectorised sqrt is
    off by one ulp in about 1 % of the entries."""
    E_a, E_b = np.asarray(E_a), np.asarray(E_b)
    D_a, D_b = np.asarray(D_a, np.float32), np.asarray(D_b, np.float32)
    bad = []
    for r in np.nonzero(np.asarray(valid_rows))[0]:
        if _ulps(D_a[r], D_b[r]).max() > ulp:
            bad.append(int(r))
            continue
        if np.array_equal(E_a[r], E_b[r]):
            continue
        d = D_b[r]
        K = len(d)
        start = 0
        while start < K:
            end = start
            while end + 1 < K and _ulps(d[end + 1:end + 2], d[end:end + 1])[0] <= 2 * ulp:
                end += 1
            if end < K - 1 and set(E_a[r, start:end + 1].tolist()) != set(E_b[r, start:end + 1].tolist()):
                bad.append(int(r))
                break
            start = end + 1
    return bad
