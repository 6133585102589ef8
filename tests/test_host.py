"""CPU tests of the host side: C-ABI library exports, weight packing, tables, featuriser / PDB I/O, sharding."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
HAVE_REF = os.path.isdir("/root/reference/src")


def test_library_loads_and_exports_every_header_symbol():
    from packppi_b200 import _lib
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "packppi_b200.h")) as f:
        names = set(re.findall(r"\b(pp_\w+)\s*\(", f.read()))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.pp_abi_version() == 2
    assert set(_lib.SIGNATURES) <= names and set(_lib.KERNELS) == set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    import packppi_b200 as pp
    from packppi_b200 import synthetic
    b = synthetic.make_complex((6, 5), seed=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        pp.get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D)
    with pytest.raises(RuntimeError, match="CUDA"):
        pp.compute_residue_clash(b, b.SC_D)
    with pytest.raises(RuntimeError, match="CUDA"):
        pp.proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 2)
    m = pp.TDiffusionModule()
    with pytest.raises(RuntimeError):
        m.sampling(b)


def test_missing_library_fails_loudly(monkeypatch):
    from packppi_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpackppi_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_weight_layout_and_packing():
    from packppi_b200 import _lib, weights
    layout, total = _lib.layout()
    sd = weights.make_state_dict(3)
    assert weights.num_parameters() == 1439172
    blob = weights.pack_weights(sd, layout, total)
    assert blob.numel() == total and all(off % 32 == 0 for off, _ in layout.values())

    def view(name, rows, cols):
        off, _ = layout[name]
        return blob[off:off + rows * cols].reshape(rows, cols)

    Win = sd["mpnn.mpnn_layers.1.edge_message_fn.W_in.weight"]
    assert torch.equal(view("L1_E_WAG", 160, 128), torch.cat([Win[:, :128].t(), Win[:, 384:416].t()]))
    assert torch.equal(view("L1_E_WEG", 168, 128), torch.cat([Win[:, 128:256].t(), Win[:, 416:456].t()]))
    assert torch.equal(view("L1_E_WN", 128, 128), Win[:, 256:384].t())
    We = sd["encoder.edge_embedding.weight"]
    Wt = view("ENC_EDGE_WT", 496, 128)
    assert torch.equal(Wt[:400], We[:, 65:465].t()) and torch.equal(Wt[400:465], We[:, :65].t())
    assert torch.equal(Wt[480:483], We[:, 465:].t()) and Wt[465:480].abs().sum() == 0 and Wt[483:].abs().sum() == 0
    assert torch.equal(view("L2_NF_WOUT", 512, 128), sd["mpnn.mpnn_layers.2.node_dense.W_out.weight"].t())
    assert torch.equal(view("DEC_W3", 16, 4), sd["decoder_score.2.W_out.weight"].t())
    bad = dict(sd)
    bad["encoder.node_embedding.weight"] = torch.zeros(128, 50)
    with pytest.raises(RuntimeError, match="shape"):
        weights.pack_weights(bad, layout, total)


def test_module_state_dict_keys_match_reference_layout():
    from packppi_b200 import TDiffusionModule, weights
    m = TDiffusionModule()
    assert list(m.state_dict().keys()) == list(weights.shapes().keys())
    assert sum(p.numel() for p in m.parameters()) == weights.num_parameters()


def test_ode_coefficients_closed_form():
    """c = 0.5 g^2 dt, w = 3 / (alpha + 3 (1 - alpha)) with sigma = exp(ln(0.01 pi) + ln(100) t) (SURVEY.md §8a')."""
    from packppi_b200.engine import Engine
    co = Engine.ode_coefficients(30, 3)
    assert len(co) == 30 and abs(co[0][0] - 1.0) < 1e-7
    for t, c, w, d in co:
        sigma = np.exp(np.log(0.01 * np.pi) + np.log(100.0) * t)
        g2 = sigma ** 2 * 2 * np.log(100.0)
        alpha = 1 - (sigma / np.pi) ** 2
        assert abs(c - 0.5 * g2 / 30) < 1e-5 * max(1.0, c) and d == 0.0
        assert abs(w - 3 / (alpha + 3 * (1 - alpha))) < 1e-5
    # SDE branch (schedule.py:224-228): c = g^2 dt, d = g sqrt(dt)
    for (t, c, w, d), (_, c_ode, w_ode, _) in zip(Engine.ode_coefficients(30, 3, mode="sde"), co):
        assert abs(c - 2 * c_ode) < 1e-5 * max(1.0, c) and w == w_ode
        assert abs(d * d - c) < 1e-5 * max(1.0, c)


def test_dist_bounds_and_reach():
    from packppi_b200 import tables
    lo, hi = tables.dist_bounds(0.5, 12.0)
    assert lo.shape == (21, 14, 14) and lo.dtype == np.float32
    assert np.array_equal(lo, lo.transpose(0, 2, 1)) and np.array_equal(hi, hi.transpose(0, 2, 1))
    assert (tables.max_reach()[:20] > 2.3).all() and tables.max_reach().max() < 12.0
    assert tables.packed_geometry().shape == (21, tables.GEO_STRIDE)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
def test_tables_equal_reference():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_tables
    from packppi_b200 import tables
    t, rc = gen_tables.build_tables()
    for k, v in tables.raw().items():
        assert np.array_equal(v, t[k]), k
    for cot, vtf in ((0.5, 12.0), (1.5, 15.0)):
        ref = rc.make_atom14_dists_bounds(overlap_tolerance=cot, bond_length_tolerance_factor=vtf)
        lo, hi = tables.dist_bounds(cot, vtf)
        assert np.array_equal(lo, ref["lower_bound"]) and np.array_equal(hi, ref["upper_bound"])


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not mounted")
@pytest.mark.parametrize("name,case", [("1BRS", "1brs"), ("T1124_lig", "t1124")])
def test_pdb_reader_and_featuriser_reproduce_golden_inputs(name, case):
    """packppi_b200.pdb + featurize against the batch the reference's prot_to_data produced for the golden files."""
    from packppi_b200 import featurize, pdb
    from util import load_golden
    _, gb = load_golden(case)
    b = featurize.protein_to_batch(pdb.read_pdb(f"/root/reference/data/{name}.pdb"))
    for k, v in gb.items():
        if torch.is_tensor(v):
            assert v.dtype == b[k].dtype and torch.equal(v, b[k]), k


def test_pdb_write_read_round_trip(tmp_path):
    from packppi_b200 import pdb
    from util import load_golden
    _, b = load_golden("syn64")
    from packppi_b200 import tables
    prot = dict(atom_positions=np.round(b.X[0].numpy().astype(np.float64), 3), atom_mask=b.atom_mask[0].numpy(),
                aaindex=b.residue_type[0].numpy(), residue_index=np.arange(1, 65),
                chain_id=np.array(["A"] * 32 + ["B"] * 32))
    prot["atom_positions"][prot["atom_mask"] == 0] = np.nan
    path = tmp_path / "x.pdb"
    pdb.write_pdb(prot, str(path))
    back = pdb.read_pdb(str(path))
    assert np.array_equal(back["aaindex"], prot["aaindex"]) and np.array_equal(back["atom_mask"], prot["atom_mask"])
    assert np.allclose(np.nan_to_num(back["atom_positions"]), np.nan_to_num(prot["atom_positions"]), atol=1e-3)
    assert list(back["chain_id"]) == list(prot["chain_id"]) and tables.restypes()[0] == "A"


def test_collate_pads_like_reference():
    from packppi_b200 import collate, synthetic
    items = [synthetic.make_complex((4, 3), seed=1), synthetic.make_complex((6, 5), seed=2)]
    b = collate(items)
    assert b.X.shape == (2, 11, 14, 3) and b.num_proteins == 2 and b.max_size == 11
    assert torch.equal(b.X[0, :7], items[0].X[0]) and b.X[0, 7:].abs().sum() == 0
    assert b.residue_mask[0].tolist() == [1.0] * 7 + [0.0] * 4 and b.chi_1pi_periodic_mask.dtype == torch.bool


def test_partition_is_balanced_and_complete():
    from packppi_b200 import shard
    lengths = [800, 200, 640, 333, 512, 250, 799, 410]
    for world in (1, 2, 4, 8):
        plan = shard.partition(lengths, world)
        assert sorted(i for p in plan for i in p) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in p) for p in plan]
        assert max(loads) - min(loads) <= max(lengths)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from packppi_b200 import shard, synthetic
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    batches = [synthetic.make_complex((n, 3), seed=n) for n in (9, 4, 7, 5, 6)]
    calls = []

    def fake_sampler(b, S):  # deterministic stand-in for model.sampling(b, n_samples=S)
        calls.append(int(b.max_size))
        return torch.stack([b.SC_D * (s + 1) for s in range(S)])

    out = shard.sample_sharded(fake_sampler, batches, 3)
    ok = all(torch.equal(o, torch.stack([b.SC_D[0] * (s + 1) for s in range(3)])) for o, b in zip(out, batches))
    q.put((rank, ok, sorted(calls)))
    dist.destroy_process_group()


def test_sharded_sampling_over_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    done = sorted(sum((c for _, _, c in res), []))
    assert done == [7, 8, 9, 10, 12]  # every complex sampled exactly once across the two ranks


def _sparse_worker(rank, world, port, q):
    """More ranks than complexes: rank 1 owns nothing and must still join the collective with a buffer on the backend's
    device (the device used to be taken from the first local result)."""
    import torch.distributed as dist
    from packppi_b200 import shard, synthetic
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    batches = [synthetic.make_complex((6, 3), seed=3)]
    out = shard.sample_sharded(lambda b, S: torch.stack([b.SC_D * (s + 2) for s in range(S)]), batches, 2)
    plan = shard.partition([9], world)
    ok = len(out) == 1 and torch.equal(out[0], torch.stack([batches[0].SC_D[0] * (s + 2) for s in range(2)]))
    q.put((rank, ok, len(plan[rank])))
    dist.destroy_process_group()


def test_sharded_sampling_with_more_ranks_than_complexes():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res)
    assert [n for _, _, n in res] == [1, 0]  # the second rank had nothing to sample


def test_slab_partition_and_halo_are_complete():
    """Owned residues of every slab see all their clash partners inside (owned + halo): the oracle's loss and gradient
    on the local subset equal the global ones on the owned residues (host logic of SURVEY.md section 8e)."""
    from oracle import prox_oracle as po
    from packppi_b200 import shard, synthetic, tables
    from packppi_b200.batch import ComplexBatch
    b = synthetic.make_complex((70, 60, 50), seed=21, place_side_chains=po.atom14_coords)
    L = 180
    ca = b.X[0, :, 1].numpy()
    rtype = b.residue_type[0].numpy()
    reach = tables.max_reach()[rtype].astype(np.float64)
    cutoff = 2 * float(tables.raw()["clash_radius"].max()) - 0.5
    pr_all, gr_all = po.clash_value_and_grad(b, b.SC_D, sparse=True)
    for world in (2, 3):
        owner = shard.slab_partition(ca, np.ones(L, bool), world)
        assert sorted(np.bincount(owner, minlength=world)) == sorted(len(c) for c in np.array_split(np.arange(L), world))
        seen = np.zeros(L, int)
        for rank in range(world):
            loc = shard.halo_of(ca, reach, owner, rank, cutoff)
            assert np.all(np.diff(loc) > 0) and set(np.nonzero(owner == rank)[0]) <= set(loc)
            sub = ComplexBatch(**{k: (v[:, loc] if torch.is_tensor(v) else v) for k, v in b.items()})
            pr, gr = po.clash_value_and_grad(sub, sub.SC_D, sparse=True)
            own = owner[loc] == rank
            # the oracle differentiates sum(per_res) over the local set; halo residues add terms that involve only
            # halo-halo / halo-owned pairs, and the owned-owned + owned-halo part is what must agree: compare per_res
            assert torch.allclose(pr[0, own], pr_all[0, loc[own]], atol=1e-6)
            seen[loc[own]] += 1
        assert np.all(seen == 1)


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from packppi_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    L = 11
    owner = np.array([0, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0])
    ids_of = [torch.from_numpy(np.nonzero(owner == r)[0]) for r in range(world)]
    counts = [int((owner == r).sum()) for r in range(world)]
    truth = torch.arange(L * 4, dtype=torch.float32).reshape(L, 4)
    full = shard.gather_owned_rows(truth[ids_of[rank]], ids_of, counts, L)
    q.put((rank, bool(torch.equal(full, truth))))
    dist.destroy_process_group()


def test_gather_owned_rows_over_two_gloo_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


def test_tensor_core_operand_images_decode_back_to_the_weights():
    """pack_tc_stream / pack_pre_stream: fp16 (hi, lo) image pairs in UMMA core-matrix order, scaled by a power of
    two; decoding an image with the documented offset formula must give the weight back to ~2^-22."""
    import torch
    from packppi_b200 import weights
    sd = weights.make_state_dict(0)

    def decode(words, off_halves, rows, kc):
        """-> (hi + lo) as [rows, kc] float64 from the image pair that starts at `off_halves` (int16 units)."""
        raw = words.view(torch.int16)
        n = rows * kc
        hi = raw[off_halves:off_halves + n].view(torch.float16).double()
        lo = raw[off_halves + n:off_halves + 2 * n].view(torch.float16).double()
        r = torch.arange(rows).view(-1, 1)
        k = torch.arange(kc).view(1, -1)
        idx = (k // 8) * rows * 8 + (r // 8) * 64 + (r % 8) * 8 + k % 8
        return (hi + lo)[idx]

    n_tc = 2 * 128 * (176 + 128 + 128 + 4 * 256) * 2 // 4 + 8
    w = weights.pack_tc_stream(sd, n_tc)
    assert w.shape == (3, 3, n_tc)
    M = sd["mpnn.mpnn_layers.1.edge_message_fn.W_inter.0.weight"].double()        # G2 of layer 1, edge path
    inv = float(w[1, 1, n_tc - 8 + 1])
    got = decode(w[1, 1], 2 * 128 * 176 + 2 * 128 * 32 * 2, 128, 32) * inv         # third chunk: k = 64..95
    assert (got - M[:, 64:96]).abs().max().item() < 2 ** -21 * M.abs().max().item()
    assert inv == 2.0 ** round(np.log2(inv))

    n_pre = (4 * 2 * 32 * 32 * 2 + 9 * 2 * 128 * 32 * 2) // 4 + 8
    p = weights.pack_pre_stream(sd, n_pre)
    assert p.shape == (3, 2, n_pre)
    Wp = sd["mpnn.mpnn_layers.2.points_fn_edge.weight"].double()                   # [24, 128]
    got = decode(p[2, 1], 2 * 32 * 32, 32, 32) * float(p[2, 1, n_pre - 8])         # second chunk: k = 32..63
    assert (got[:24] - Wp[:, 32:64]).abs().max().item() < 2 ** -21 * Wp.abs().max().item()
    assert got[24:].abs().max().item() == 0.0


def test_add_sc_noise_returns_the_reference_score_target():
    """`add_sc_noise` used to return zeros as its second value.  The reference reads the score of the
    wrapped normal from a 5001 x 5001 table (schedule.py:64-73); here the same table entry is evaluated on the fly.
    Golden: the reference's own add_sc_noise on 1BRS at three times (tools/make_golden_score.py).  An index that sits
    on a rounding boundary of the log grid may land in the neighbouring cell (numpy fp32 log vs torch fp32 log):
    allowed for < 1 % of the entries, and then within the 0.5 % that one grid cell is worth."""
    from util import load_golden
    from packppi_b200 import TDiffusionModule
    _, b = load_golden("1brs")
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "score_target.npz"))
    m = TDiffusionModule()
    for i in range(3):
        x, s = m.add_sc_noise(b, torch.from_numpy(z[f"in_t_{i}"]),
                              noise=(torch.from_numpy(z[f"in_eps1_{i}"]), torch.from_numpy(z[f"in_eps2_{i}"])))
        rn, rs = torch.from_numpy(z[f"ref_noised_{i}"]), torch.from_numpy(z[f"ref_score_{i}"])
        assert torch.equal(x, rn)
        rel = (s - rs).abs() / (rs.abs() + 1e-6)
        assert (rel > 1e-5).float().mean().item() < 0.01 and rel.max().item() < 5e-3, (i, rel.max().item())
        assert (s[rs == 0] == 0).all() and rs.abs().max() > 0


def _ref_data_dir():
    for d in ("/root/reference/data", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "baseline", "_ref",
                                                   "data")):
        if os.path.exists(os.path.join(d, "1BRS.pdb")):
            return d
    return None


@pytest.mark.skipif(_ref_data_dir() is None, reason="PDB fixtures of the reference not present")
@pytest.mark.parametrize("name,case", [("1BRS", "1brs"), ("T1124_lig", "t1124")])
def test_to_pdb_is_byte_identical_to_the_reference(name, case):
    """`to_pdb` (protein.py:207-314).  Golden = sha256 of the text the reference's own to_pdb wrote for the
    parsed file (tools/make_golden_pdb.py), plus its first / last lines and TER records for a readable failure."""
    import hashlib
    import json
    from packppi_b200 import pdb
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pdb_text.json")) as f:
        want = json.load(f)[case]
    text = pdb.to_pdb(pdb.read_pdb(os.path.join(_ref_data_dir(), name + ".pdb")))
    lines = text.split("\n")
    assert lines[:4] == want["head"] and lines[-5:] == want["tail"]
    assert [ln for ln in lines if ln.startswith("TER")] == want["ter"]
    assert len(lines) == want["n_lines"]
    assert hashlib.sha256(text.encode()).hexdigest() == want["sha256"]


@pytest.mark.skipif(_ref_data_dir() is None, reason="PDB fixtures of the reference not present")
def test_interface_mask_and_metric_block_match_the_reference(tmp_path):
    """`get_interface_mask` (helper.py:104-129) and the arithmetic of `get_metric`
    (protein_analysis.py:53-88), against values the reference computed (tools/make_golden_pdb.py)."""
    from oracle import prox_oracle as po
    from packppi_b200 import featurize, metrics, pdb
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metric_block.npz"))
    d = _ref_data_dir()
    for name, case in (("1BRS", "1brs"), ("T1124_lig", "t1124")):
        path = os.path.join(d, name + ".pdb")
        prot = pdb.read_pdb(path)
        mask = metrics.interface_mask(prot, path)
        b = featurize.protein_to_batch(prot)
        assert np.array_equal((mask * b.residue_mask[0]).numpy(), z[f"{case}_interface_mask"]), case
    # metric block: true = 1BRS, predicted = the reference-written structure with perturbed side chains
    path = os.path.join(d, "1BRS.pdb")
    prot = pdb.read_pdb(path)
    true = featurize.protein_to_batch(prot)
    true["interface_mask"] = (metrics.interface_mask(prot, path) * true.residue_mask[0])[None]
    pred_prot = dict(prot)
    pred_prot["atom_positions"] = z["1brs_pred_atom_positions"]
    pred_path = tmp_path / "pred.pdb"
    pred_path.write_text(pdb.to_pdb(pred_prot))
    pred = featurize.protein_to_batch(pdb.read_pdb(str(pred_path)))
    got = metrics.get_metric(true, pred, clashscore=7.25, atom14_fn=po.atom14_coords)  # host tensors: CPU rebuild
    keys = [k[len("1brs_metric_"):] for k in z.files if k.startswith("1brs_metric_")]
    assert set(keys) == set(got) and len(keys) == 16
    for k in keys:
        assert abs(float(got[k]) - float(z["1brs_metric_" + k])) <= 1e-5 * max(1.0, abs(float(z["1brs_metric_" + k]))), k


def test_live_tile_lists_are_compact_and_ordered():
    """engine.Graph._live_compute: ids of the 4-row / 128-row tiles that hold a residue with msum != 0, ascending, the
    tail pointing one past the last tile (what the persistent tensor-core kernels index by position)."""
    from packppi_b200.engine import Graph
    rng = np.random.default_rng(3)
    for G, S, per in ((13, 2, 4), (700, 3, 4), (700, 3, 128), (5, 1, 128), (1024, 8, 128)):
        msum = torch.from_numpy((rng.random(G) < 0.4).astype(np.float32))
        msum[G // 3:G // 2] = 0  # a run of padding
        g = Graph.__new__(Graph)
        g.msum = msum
        ids, cnt = g._live_compute(S, per)
        rows = np.tile(msum.numpy() != 0, S)
        nt = (len(rows) + per - 1) // per
        pad = np.zeros(nt * per, bool)
        pad[:len(rows)] = rows
        want = np.nonzero(pad.reshape(nt, per).any(1))[0]
        assert int(cnt) == len(want) and ids.dtype == torch.int32 and ids.numel() == nt + 2
        assert np.array_equal(ids[:len(want)].numpy(), want)
        assert (ids[len(want):] == nt).all()


def test_bench_sample_selection_is_stratified():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Item:
        def __init__(self, n):
            self.max_size = n

    items = [Item(n) for n in (500, 210, 790, 330, 640, 450, 270, 720)]
    picks = [it.max_size for it in bench.stratified(items, 4)]
    assert picks == [270, 450, 640, 790] or picks == sorted(picks)  # one per length quartile, ascending
    assert len(set(picks)) == 4 and min(picks) < 400 < max(picks)
    assert [it.max_size for it in bench.stratified(items, 1)] == [500]


def test_host_tensors_are_refused_unless_staged_and_staging_needs_a_cuda_device():
    """No CPU implementation: CPU tensors raise RuntimeError; `host_staging` only accepts a CUDA device (it moves the
    buffers, it does not compute on the host)."""
    import packppi_b200
    from packppi_b200 import synthetic
    b = synthetic.make_complex((6, 5), seed=1)
    with pytest.raises(RuntimeError):
        packppi_b200.get_atom14_coords(b.X, b.residue_type, b.BB_D, b.SC_D)
    with pytest.raises(RuntimeError):
        packppi_b200.proximal_optimizer(b, b.SC_D, 12.0, 0.5, 1.0, 2)
    with pytest.raises(RuntimeError):
        with packppi_b200.host_staging("cpu"):
            pass
    from packppi_b200 import components
    assert components._STAGE_DEVICE is None


@pytest.mark.skipif(_ref_data_dir() is None, reason="PDB fixtures of the reference not present")
def test_interface_mask_without_the_reference_quirk_marks_both_chains():
    """`as_reference=False` compares every chain's residue numbers with the file's own numbering: both chains of the
    barnase-barstar complex then have interface residues (the reference's in-place offset leaves the second chain with
    accidental matches only, which `as_reference=True` reproduces)."""
    from packppi_b200 import metrics, pdb
    path = os.path.join(_ref_data_dir(), "1BRS.pdb")
    prot = pdb.read_pdb(path)
    chains = np.asarray(prot["chain_id"])
    fair = metrics.interface_mask(prot, path, as_reference=False).numpy()
    quirk = metrics.interface_mask(prot, path).numpy()
    first, second = chains == np.unique(chains)[0], chains == np.unique(chains)[1]
    assert fair[first].sum() > 20 and fair[second].sum() > 20
    assert np.array_equal(fair[first], quirk[first]) and quirk[second].sum() < fair[second].sum()
