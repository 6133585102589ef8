"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under tools/ref_shims.py.

    python tools/make_golden.py            # all cases (a few minutes on 8 cores)

Protocol: weights = packppi_b200.weights.make_state_dict(0) loaded into the model built by the
reference constructor; initial noise = the reference's own `add_sc_noise` under torch.manual_seed(1); the
reference `sampling` is then re-run under the same seed, so the noise it draws is the recorded one.  Every
array saved as `ref_*` is an output of reference code; `in_*` arrays are the inputs that produced it.
The PDB cases are parsed with packppi_b200.pdb (Biopython is absent) and featurised by the reference's
`ComplexDataset.prot_to_data`.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shims  # noqa: E402
from packppi_b200 import pdb, synthetic, weights  # noqa: E402
from packppi_b200.batch import TENSOR_FIELDS  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_DATA = os.path.join(ref_shims.REFERENCE_ROOT, "data")


def ref_batch_from_protein(ref, prot):
    d = ref.cds.ComplexDataset.prot_to_data({k: (np.array(v, copy=True) if hasattr(v, "shape") else v)
                                             for k, v in prot.items()}, cache_processed_data=False)
    for k in list(d.keys()):
        if not isinstance(d[k], int):
            d[k] = d[k].unsqueeze(0)
    d.num_proteins = 1
    d.max_size = d.num_nodes
    return d


def to_ref_data(ref, b):
    d = ref.Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()})
    return d


def pad_stack(ref, items):
    """reference collate semantics (complex_datamodule.py:196-226) on already-batched ([1,L,..]) items."""
    import torch.nn.functional as F
    L = max(int(it.X.shape[1]) for it in items)
    out = ref.Data(num_proteins=len(items), max_size=L, num_nodes=L)
    for k in TENSOR_FIELDS:
        rows = []
        for it in items:
            v = it[k][0]
            rows.append(F.pad(v, [0, 0] * (v.dim() - 1) + [0, L - v.shape[0]]))
        out[k] = torch.stack(rows)
    return out


def save_inputs(store, batch):
    for k in TENSOR_FIELDS:
        store["in_" + k] = batch[k].numpy()


def run_case(ref, model, name, batch, do_prox=0, traj_steps=None, clash=True, prox_on="sample", rows=(0,),
             full_layers=True):
    t0 = time.time()
    g = {}
    save_inputs(g, batch)
    B, L = batch.X.shape[:2]

    with torch.no_grad():
        # graph
        D_nb, E_idx, _ = model.encoder._dist(batch.X[:, :, 1, :], batch.residue_mask)
        g["ref_E_idx"] = E_idx.numpy().astype(np.int16)
        g["ref_D_neighbors"] = D_nb.numpy()

        # one network call at t = 0.7 on a fixed noised input, with per-layer taps
        gen = torch.Generator().manual_seed(7)
        x_probe = ((torch.rand(B, L, 4, generator=gen) * 2 - 1) * np.pi) * batch.SC_D_mask
        taps = {}
        hooks = []
        hooks.append(model.encoder.register_forward_hook(lambda m, i, o: taps.__setitem__("enc", o)))
        for li, layer in enumerate(model.mpnn.mpnn_layers):
            hooks.append(layer.register_forward_hook(lambda m, i, o, li=li: taps.__setitem__(f"l{li}", o)))
        t = torch.full((B * L,), 0.7)
        score, h_V = model.network(batch, x_probe, t)
        for h in hooks:
            h.remove()
        g["in_probe_SC_D"] = x_probe.numpy()
        g["ref_probe_score"] = score.numpy()
        g["ref_probe_hV"] = h_V.numpy()
        rows = np.asarray([r for r in rows if r < L], np.int64)
        g["in_rows"] = rows
        g["in_full_layers"] = np.asarray(int(full_layers))
        g["ref_probe_hV0"] = taps["enc"][0].numpy() if full_layers else taps["enc"][0][:, rows].numpy()
        g["ref_probe_hE0_rows"] = taps["enc"][1][:, rows].numpy()
        for li in range(3):
            hv = taps[f"l{li}"][0]
            g[f"ref_probe_hV_l{li}"] = hv.numpy() if full_layers else hv[:, rows].numpy()
        for li in range(2):
            g[f"ref_probe_hE_l{li}_rows"] = taps[f"l{li}"][1][:, rows].numpy()

        # sampling with recorded noise
        torch.manual_seed(1)
        t1 = torch.tensor([1.]).repeat_interleave(L * B)
        x_init, _ = model.add_sc_noise(batch, t1)
        g["in_SC_D_init"] = x_init.numpy()
        trace = []
        orig_network = model.network

        def tapped(b, x, t):
            trace.append(x.detach().clone())
            return orig_network(b, x, t)

        model.network = tapped
        torch.manual_seed(1)
        x_final = model.sampling(batch, use_proximal=False)
        model.network = orig_network
        assert torch.equal(trace[0], x_init)
        traj = torch.stack(trace[1:] + [x_final])  # state after step j, j = 0..29
        keep = np.arange(30) if traj_steps is None else np.asarray(traj_steps)
        g["in_traj_steps"] = keep
        g["ref_traj"] = traj[keep].numpy()
        g["ref_SC_D_final"] = x_final.numpy()
        g["ref_atom14_final"] = ref.comp.get_atom14_coords(batch.X, batch.residue_type, batch.BB_D, x_final).numpy()
        g["ref_atom14_native"] = ref.comp.get_atom14_coords(batch.X, batch.residue_type, batch.BB_D,
                                                            batch.SC_D).numpy()

    if clash:
        outs, grads = [], []
        for bi in range(B):  # compute_residue_clash broadcasts over a leading batch dim of one complex at a time
            sub = ref.Data(**{k: (v[bi:bi + 1] if torch.is_tensor(v) else v) for k, v in batch.items()})
            x = x_final[bi:bi + 1].clone().requires_grad_(True)
            pr = ref.clash.compute_residue_clash(sub, x, 12., 0.5)
            (gr,) = torch.autograd.grad(pr.sum(), x)
            outs.append(pr.detach())
            grads.append(gr)
        g["ref_clash_per_res"] = torch.cat(outs).numpy()
        g["ref_clash_grad"] = torch.cat(grads).numpy()

    if do_prox:
        assert B == 1
        start = x_final if prox_on == "sample" else batch.SC_D
        g["in_prox_start"] = start.numpy()
        snaps, losses = ref.optimize.proximal_optimizer(batch, start.clone(), 12., 0.5, 1., do_prox)
        g["ref_prox_losses"] = np.asarray(losses, np.float64)
        keep = sorted(set([0, do_prox // 2, do_prox - 1]))
        g["in_prox_keep"] = np.asarray(keep)
        g["ref_prox_snaps"] = torch.stack([snaps[k] for k in keep]).numpy()
        g["ref_prox_mask"] = ref.optimize.find_clash_mask(batch, start, 12., 0.5).numpy()

    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **g)
    print(f"{name}: B={B} L={L} -> {os.path.getsize(path) / 1024:.0f} KiB in {time.time() - t0:.1f}s", flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = ref_shims.import_reference()
    model = ref_shims.build_reference_model(ref)
    sd = weights.make_state_dict(0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not [m for m in missing if "schedule" not in m and "loss" not in m], (missing, unexpected)
    assert sum(p.numel() for p in model.parameters()) == weights.num_parameters()
    only = sys.argv[1:]

    def want(n):
        return not only or n in only

    place = lambda X, S, BB, SC: ref.comp.get_atom14_coords(X, S, BB, SC)  # noqa: E731

    if want("1brs"):
        b = ref_batch_from_protein(ref, pdb.read_pdb(os.path.join(REF_DATA, "1BRS.pdb")))
        run_case(ref, model, "1brs", b, do_prox=50, rows=(0, 57, 120, 194))
    if want("1brs_sde"):
        # mode "sde" (schedule.py:224-228): record the torch.normal draws of the reference run so they can be injected
        b = ref_batch_from_protein(ref, pdb.read_pdb(os.path.join(REF_DATA, "1BRS.pdb")))
        model.schedule_1pi_periodic.mode = model.schedule_2pi_periodic.mode = "sde"
        draws, orig_normal = [], torch.normal

        def recording_normal(*a, **k):
            out = orig_normal(*a, **k)
            draws.append(out.clone())
            return out

        torch.manual_seed(1)
        x_init, _ = model.add_sc_noise(b, torch.ones(b.X.shape[1]))
        torch.manual_seed(1)
        torch.normal = recording_normal
        try:
            with torch.no_grad():
                x_final = model.sampling(b, use_proximal=False)
        finally:
            torch.normal = orig_normal
            model.schedule_1pi_periodic.mode = model.schedule_2pi_periodic.mode = "ode"
        assert len(draws) == 60
        gd = {}
        save_inputs(gd, b)
        gd["in_SC_D_init"] = x_init.numpy()
        gd["in_sde_noise"] = torch.stack(draws).reshape(30, 2, -1, 4).numpy()
        gd["ref_SC_D_final"] = x_final.numpy()
        np.savez_compressed(os.path.join(OUT, "1brs_sde.npz"), **gd)
        print("1brs_sde written", flush=True)
    if want("t1124"):
        b = ref_batch_from_protein(ref, pdb.read_pdb(os.path.join(REF_DATA, "T1124_lig.pdb")))
        run_case(ref, model, "t1124", b, do_prox=0, traj_steps=[0, 1, 9, 19, 29], rows=(100, 619, 620),
                 full_layers=False)
    if want("small"):
        # ragged batch of tiny complexes: K = min(32, L) edge cases, padding, an interior masked residue
        items = []
        for L, seed in ((5, 11), (17, 12), (31, 13), (33, 14), (64, 15)):
            it = to_ref_data(ref, synthetic.make_complex((L - L // 2, L // 2), seed=seed, place_side_chains=place))
            items.append(it)
        it = items[3]
        for k in TENSOR_FIELDS:  # mask residue 7 of the 33-residue complex the way prot_to_data would
            it[k][0, 7] = 0
        for L, it in zip((5, 17, 31, 33, 64), items):
            run_case(ref, model, f"syn{L}", it, do_prox=5, traj_steps=[0, 29], prox_on="native")
        run_case(ref, model, "synbatch", pad_stack(ref, items), do_prox=0, traj_steps=[0, 29], rows=(0, 4, 40))
    if want("syn300"):
        b = to_ref_data(ref, synthetic.make_complex((150, 150), seed=300, place_side_chains=place))
        run_case(ref, model, "syn300", b, do_prox=50, traj_steps=[0, 14, 29], prox_on="native", rows=(0, 299))


if __name__ == "__main__":
    main()
