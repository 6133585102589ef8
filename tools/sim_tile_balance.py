"""Host simulation behind DESIGN.md §8 "Live-tile lists": how evenly the live 4-row tiles of the benchmark's ragged,
padded micro-batches fall on the 148 persistent CTAs, (a) when CTA b tests tiles b, b + 148, ... for liveness (the
scheme up to round 2) and (b) when it takes positions b, b + 148, ... of the compacted live-tile list.  Prints the
fullest-CTA / mean ratio summed over the micro-batches of a rank's share (the kernel ends with its fullest CTA), for
the single-GPU sweep and for the per-rank shares at 2 / 4 / 8 GPUs.  No GPU needed:  python tools/sim_tile_balance.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from packppi_b200 import shard, synthetic  # noqa: E402


def ratio(groups, S=8, ncta=148, compact=False):
    tot_max = tot_mean = 0.0
    for grp in groups:
        Lp, B = max(grp), len(grp)
        rows = np.zeros(B * Lp, bool)
        for b, n in enumerate(grp):
            rows[b * Lp:b * Lp + n] = True
        rows = np.tile(rows, S)
        nt = (len(rows) + 3) // 4
        pad = np.zeros(nt * 4, bool)
        pad[:len(rows)] = rows
        live = np.nonzero(pad.reshape(nt, 4).any(1))[0]
        owner = (np.arange(len(live)) if compact else live) % ncta
        per = np.bincount(owner, minlength=ncta)
        tot_max += per.max()
        tot_mean += per.mean()
    return tot_max / tot_mean


def micro(lengths, size=8):
    lengths = sorted(lengths)
    return [lengths[i:i + size] for i in range(0, len(lengths), size)]


def main():
    L = [sum(c) for c in synthetic.sweep_lengths()]
    print("GPUs  test-each-tile  compacted-list   (fullest CTA / mean, worst rank)")
    for world in (1, 2, 4, 8):
        plan = shard.partition(L, world)
        a = max(ratio(micro([L[i] for i in p])) for p in plan)
        c = max(ratio(micro([L[i] for i in p]), compact=True) for p in plan)
        print(f"{world:4d}  {a:14.3f}  {c:14.3f}")


if __name__ == "__main__":
    main()
