"""Host-side orchestration of the kernels: buffers, launch order, per-step scalars.  PyTorch is used for device
memory and streams only; every arithmetic step of the hot path is a kernel of libpackppi_b200.so.
"""
import math
import os

import numpy as np
import torch

from . import _lib, tables
from .weights import pack_pre_stream, pack_tc_stream, pack_weights

SIGMA_MIN, SIGMA_MAX = 0.01 * np.pi, np.pi  # schedule.py:148-149 defaults
TOP_K = 32
# Below this many residue rows (S*G) a denoising step is launch-latency bound (17 launches, ~0.1 ms of device work),
# so repeated sampling on buffers of the same shape replays one captured CUDA graph of the whole 30-step loop.
# From this many residues per complex the kNN graph is built with the cell-list kernel instead of the O(L^2) scan.
CELL_LIST_MIN_L = int(os.environ.get("PACKPPI_B200_CELL_LIST_MIN_L", "1024"))
GRAPH_ROWS_MAX = int(os.environ.get("PACKPPI_B200_GRAPH_ROWS", "16384"))
USE_LIVE_LIST = os.environ.get("PACKPPI_B200_LIVE_LIST", "1") != "0"  # compacted live-tile list for the tc kernels
NODE_CLUSTER = int(os.environ.get("PACKPPI_B200_NODE_CLUSTER", "1"))  # CTAs per cluster of the node-message kernel


def _f32(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _i64(t, dev):
    return t.to(device=dev, dtype=torch.int64).contiguous()


class DeviceTables:
    """Per-device chemistry tables (uploaded once, cached)."""
    _cache = {}

    @classmethod
    def get(cls, dev):
        key = (dev.type, dev.index)
        if key not in cls._cache:
            cls._cache[key] = cls(dev)
        return cls._cache[key]

    def __init__(self, dev):
        self.dev = dev
        self.geo = torch.from_numpy(tables.packed_geometry()).to(dev).contiguous()
        assert self.geo.shape[1] == _lib.load().pp_table_stride()
        self._bounds = {}
        self.max_radius = float(tables.raw()["clash_radius"].max())

    def bounds(self, cot, vtf):
        key = (float(cot), float(vtf))
        if key not in self._bounds:
            lo, hi = tables.dist_bounds(cot, vtf)
            self._bounds[key] = (torch.from_numpy(lo).to(self.dev).contiguous(),
                                 torch.from_numpy(hi).to(self.dev).contiguous())
        return self._bounds[key]


class Graph:
    """Step-invariant state of one padded batch: neighbour lists, geometry records,
    attention mask and, once `edge_embed` ran, the embedded edge features h_E0."""

    ni = step_mask = mask_1pi = None  # filled by Engine.build_graph

    def __init__(self, X, residue_mask, top_k=TOP_K):
        dev = X.device
        self.B, self.L = int(X.shape[0]), int(X.shape[1])
        self.G = self.B * self.L
        self.K = min(top_k, self.L)
        self.X = _f32(X, dev).clone()
        self.mask = _f32(residue_mask, dev).clone()
        self._cells = None
        self._live = {}
        self._replay = {}  # captured CUDA graphs of the sampling loop, keyed by (S, steps, ...)
        self._seen = set()
        G, K = self.G, self.K
        self.E_idx = torch.empty(self.B, self.L, K, dtype=torch.int64, device=dev)
        self.nbr = torch.empty(G, K, dtype=torch.int32, device=dev)
        self.D_neighbors = torch.empty(self.B, self.L, K, dtype=torch.float32, device=dev)
        self.mask_attend = torch.empty(G, K, dtype=torch.float32, device=dev)
        self.msum = torch.empty(G, dtype=torch.float32, device=dev)
        self.geo = torch.empty(G, _lib.load().pp_geo_stride(), dtype=torch.float32, device=dev)
        self.hE0 = None
        self._build()

    _serial = 0

    def _build(self):
        Graph._serial += 1
        self.serial = Graph._serial  # identifies this set of masks (Workspace.clean_for)
        outs = (self.E_idx, self.nbr, self.D_neighbors, self.mask_attend, self.msum)
        if self.L >= CELL_LIST_MIN_L:  # cell list: O(L * neighbourhood); identical output
            if self._cells is None:
                nc = int(_lib.load().pp_knn_cells_max()) + 1
                self._cells = (torch.empty(self.B * 2 * nc + self.G, dtype=torch.int32, device=self.X.device),
                               torch.empty(self.B * 8, dtype=torch.float32, device=self.X.device))
            _lib.call("pp_knn_build_cells", self.X, self.mask, self.B, self.L, self.K, *outs, *self._cells)
        else:
            _lib.call("pp_knn_build", self.X, self.mask, self.B, self.L, self.K, *outs)
        _lib.call("pp_geometry_build", self.X, self.G, self.geo)

    def _live_compute(self, S, per=4):
        dev = self.msum.device
        rows = (self.msum != 0).repeat(S)
        nt = (rows.numel() + per - 1) // per
        padded = torch.zeros(nt * per, dtype=torch.bool, device=dev)
        padded[:rows.numel()] = rows
        live = padded.view(nt, per).any(1)
        pos = torch.cumsum(live, 0) - 1
        ids = torch.full((nt + 2,), nt, dtype=torch.int32, device=dev)
        ids.scatter_(0, torch.where(live, pos, torch.full_like(pos, nt + 1)), torch.arange(nt, dtype=torch.int32, device=dev))
        ids[nt + 1] = nt  # the slot the dead tiles were scattered to
        return ids, live.sum().to(torch.int32).reshape(1)

    def live_tiles(self, S, per=4):
        """(ids int32 [ntiles + 2], count int32 [1]) of the `per`-row tiles (4: per-edge kernels, 128: per-residue
        kernels) of the S*G rows that hold a residue with msum != 0 (anything else is padding the tensor-core kernels
        skip), ascending; entries past the count point one past the last tile.  Built with a handful of stream-ordered torch ops, no host synchronisation; cached per S.
        The tensors keep their addresses for the life of the Graph (captured CUDA graphs hold them): `rebuild`
        refills them in place."""
        if getattr(self, "_live", None) is None:
            self._live = {}
        hit = self._live.get((S, per))
        if hit is None:
            hit = self._live[(S, per)] = self._live_compute(S, per)
        return hit

    def rebuild(self, X, residue_mask):
        """Same shape, new complex: refill the existing buffers in place (pointers stay valid for captured graphs)."""
        self.X.copy_(X)
        self.mask.copy_(residue_mask)
        self._build()

    def edge_embed(self, wblob, residue_index, chain_indices):
        dev = self.X.device
        if self.hE0 is None:
            self.hE0 = torch.empty(self.G, self.K, 128, dtype=torch.float32, device=dev)
        _lib.call("pp_edge_embed", wblob, self.geo, self.nbr, _i64(residue_index, dev), _i64(chain_indices, dev),
                  self.G, self.K, self.hE0)
        return self.hE0


class Workspace:
    """Per-sample activations and scratch for S*G residue rows: views into the capacity buffers of a `WorkspacePool`
    (or freshly allocated when no pool is given, e.g. for a captured CUDA graph that must own its memory)."""

    def __init__(self, G, K, S, dev, pool=None):
        R = S * G
        self.G, self.K, self.S = G, K, S
        if pool is None:
            z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
            self.hV, self.hE = z(R, 128), z(R, K, 128)
            self.wsA, self.wsN, self.wsAcc = z(R, 128), z(R, 128), z(R, 128)
            self.wsP, self.score = z(R, 24), z(R, 4)
        else:
            v = lambda name, *shape: pool.view(name, shape)  # noqa: E731
            self.hV, self.hE = v("hV", R, 128), v("hE", R, K, 128)
            self.wsA, self.wsN, self.wsAcc = v("wsA", R, 128), v("wsN", R, 128), v("wsAcc", R, 128)
            self.wsP, self.score = v("wsP", R, 24), v("score", R, 4)
        self.clean_for = None  # Graph.serial whose padding rows of hE / wsAcc are known to be zero


class WorkspacePool:
    """One set of flat buffers per engine, grown to the largest micro-batch seen and then reused for every shape
    (a sweep of ragged micro-batches used to free and re-allocate 1-2 GB whenever the padded shape changed)."""

    def __init__(self, dev):
        self.dev, self.buf = dev, {}

    def view(self, name, shape):
        n = 1
        for d in shape:
            n *= int(d)
        cur = self.buf.get(name)
        if cur is None or cur.numel() < n:
            self.buf[name] = None  # release the old block before the larger one is requested
            cur = self.buf[name] = torch.zeros(n, dtype=torch.float32, device=self.dev)
        return cur[:n].view(*shape)


class Engine:
    """Packed weights + kernels for one device."""

    # precision / execution modes of the per-edge message MLPs
    #   fp32   CUDA-core FFMA kernels (csrc/mpnn.cu), exact fp32
    #   f16x3  tcgen05 tensor cores (kind::f16), every operand split into a rounded fp16 (hi, lo) pair, 3 MMAs per
    #          product, fp32 accumulation: fp32-grade, the parity mode of the tensor path (default)
    #   f16    tensor cores, hi halves only (11 mantissa bits, the precision of TF32): fast mode, looser stated tolerance
    # node_epilogue: where the per-residue node update (W_out, LayerNorm, FFN 128-512-128, LayerNorm) runs:
    #   "tc32" tensor cores with promoted accumulation (csrc/node_post_tc.cu; default of the tensor-core modes)
    #   "ffma" exact-fp32 CUDA-core kernel
    # The tensor core accumulates in fp32 with truncation (measured bias -7e-7 relative at K = 128, growing linearly
    # with K), which the h_V path amplifies (1.2e-4 rad after two ODE steps with plain TMEM accumulation, over the 1e-4
    # gate; that variant was removed).  Against the CPU oracle on fresh inputs (tests/diag_accuracy.py, 64 complexes)
    # the max chi error after 2 / 30 steps is 6.5e-5 / 9.5e-6 rad with "ffma" and 5.7e-5 / 1.05e-5 with "tc32" (fp32
    # mode: 4.4e-5 / 7.6e-6).  The residue prologue (points, A_i, N_j) is accuracy-neutral and always runs on the tensor
    # cores in these modes.
    MODES = ("fp32", "f16x3", "f16")
    _serial = 0

    def __init__(self, state_dict, device, mode="f16x3", cluster=1, node_epilogue=None):
        if node_epilogue is None:
            node_epilogue = "tc32"
        if node_epilogue not in ("tc32", "ffma"):
            raise ValueError("node_epilogue must be 'tc32' or 'ffma'")
        if int(cluster) not in (1, 2):
            raise ValueError("cluster must be 1 or 2")
        self.node_epilogue = node_epilogue
        if mode not in self.MODES:
            raise ValueError(f"mode must be one of {self.MODES}")
        if mode != "fp32":
            # LayerNorm outputs are the one operand class the kernels split without tracking (|LN(x)| <= |gain| *
            # sqrt(127) + |bias|): make sure no checkpoint can push them out of the fp16 range
            worst = 0.0
            for k, v in state_dict.items():
                if "norm" in k and k.endswith(".weight"):
                    b = state_dict.get(k[:-len("weight")] + "bias")
                    worst = max(worst, float(v.abs().max()) * math.sqrt(127.0) + (float(b.abs().max()) if b is not None else 0.0))
            if worst > 32768.0:
                import warnings
                warnings.warn(f"packppi_b200: LayerNorm outputs of this checkpoint can reach {worst:.3g}, outside the "
                              "fp16 range of the split-fp16 tensor-core mode; using the fp32 CUDA-core kernels")
                mode = "fp32"
        self.mode, self.cluster = mode, int(cluster)
        Engine._serial += 1
        self.serial = Engine._serial  # identifies this set of packed weights (keys of captured CUDA graphs); not id():
        #                               ids are recycled when an engine is rebuilt after a weight update
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("packppi_b200 runs on CUDA devices only; there is no CPU fallback "
                               "(use the reference implementation for --device cpu)")
        layout, total = _lib.layout()
        self.wblob = pack_weights(state_dict, layout, total).to(self.dev)
        self.wtc = pack_tc_stream(state_dict, _lib.load().pp_tc_stream_floats()).to(self.dev)
        self.wpre = pack_pre_stream(state_dict, _lib.load().pp_tc_pre_stream_floats()).to(self.dev)
        self.tables = DeviceTables.get(self.dev)
        self._ws = {}
        self._pool = WorkspacePool(self.dev)
        # sticky flag the tensor-core kernels set when an activation leaves the fp16 range (include/packppi_b200.h)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._sched = {}

    # ------------------------------------------------------------------ graph
    def build_graph(self, batch, with_edges=True, reuse=None):
        """kNN graph, geometry records and edge embedding of `batch`.  `reuse`: a Graph of the same shape whose
        buffers (and captured CUDA graphs) are refilled in place instead of allocating new ones."""
        X, mask = batch.X.to(self.dev), batch.residue_mask.to(self.dev)
        if reuse is not None and (reuse.B, reuse.L) == tuple(X.shape[:2]) and reuse.X.device == X.device:
            g = reuse
            g.rebuild(X, mask)
        else:
            g = Graph(X, mask)
        if with_edges:
            g.edge_embed(self.wblob, batch.residue_index, batch.chain_indices)
        # per-batch constants of the sampling loop, built once per graph instead of once per sampling call
        g.ni = self.node_inputs(batch)
        m1 = batch.chi_1pi_periodic_mask.to(self.dev).reshape(-1, 4).bool()
        g.step_mask = (m1 | batch.chi_2pi_periodic_mask.to(self.dev).reshape(-1, 4).bool()).to(torch.uint8).contiguous()
        g.mask_1pi = m1.to(torch.uint8).contiguous()
        return g

    def workspace(self, G, K, S):
        """Activations for S*G rows: views into this engine's capacity buffers (one live shape at a time: asking for
        another shape re-labels the same memory, so its padding rows count as dirty again)."""
        key = (G, K, S)
        if key not in self._ws:
            self._ws.clear()
            self._ws[key] = Workspace(G, K, S, self.dev, pool=self._pool)
        return self._ws[key]

    # ------------------------------------------------------------------ network
    def node_inputs(self, batch):
        dev = self.dev
        return dict(rtype=_i64(batch.residue_type.reshape(-1), dev),
                    bb=_f32(batch.BB_D_sincos.reshape(-1, 6), dev),
                    chi_mask=_f32(batch.SC_D_mask.reshape(-1, 4), dev))

    def forward_layers(self, graph, ws, ni, chi, t, t_stride, hE0=None, sc_sincos=None):
        """node embedding + 3 IPMP layers on ws (rows S*G); leaves h_V in ws.hV."""
        G, K, S = graph.G, graph.K, ws.S
        W = self.wblob
        if self.mode != "fp32" and ws.clean_for != graph.serial:
            # the tensor-core kernels skip tiles that hold only padding residues: zero those rows once per graph
            # (the CUDA-core kernels write the zeros themselves)
            ws.hE.zero_()
            ws.wsAcc.zero_()
            ws.clean_for = graph.serial
        _lib.call("pp_node_embed", W, ni["rtype"], ni["bb"], chi, ni["chi_mask"], sc_sincos, t, t_stride, G, S, ws.hV)
        hE0 = graph.hE0 if hE0 is None else hE0
        for layer in range(3):
            first = layer == 0
            hE_in, shared, edge = (hE0 if first else ws.hE), (1 if first else 0), layer < 2
            common = (graph.geo, graph.nbr, graph.mask_attend)
            size = (graph.mask, G, K, S)
            if self.mode == "fp32" and _lib.PROFILE is None:
                _lib.call("pp_ipmp_layer", W, layer, *common, graph.msum, graph.mask, G, K, S, ws.hV, hE_in, shared,
                          ws.hE, 1 if edge else 0, ws.wsA, ws.wsN, ws.wsP, ws.wsAcc, kernels=5 if edge else 3)
                continue
            tcp = (3 if self.mode == "f16x3" else 1, self.cluster)
            use_list = USE_LIVE_LIST and getattr(graph, "_live", None) is not None
            live = graph.live_tiles(S) if use_list else (None, None)
            live128 = graph.live_tiles(S, 128) if use_list else (None, None)
            # node-message kernel as a cluster of 2 (each CTA fetches half of every weight image and multicasts it):
            # 0.549 -> 0.490 ms on an unpadded micro-batch (tools/probe_perf.py), but on the benchmark's ragged sweep
            # the lockstep of the pair costs more than the halved weight fetch saves (6.17 -> 6.10 M residue.steps/s
            # with pair-granular padding skip), so the default stays 1
            tcp_node = (tcp[0], max(self.cluster, NODE_CLUSTER))
            # the five kernels of a layer through their own entry points (instrumented fp32 mode, tensor-core modes)
            node_tc = self.mode != "fp32"  # residue prologue on the tensor cores
            pre = lambda path: (  # noqa: E731
                _lib.call("pp_ipmp_node_pre_tc", W, layer, path, self.wpre[layer, path], graph.geo, G, S, ws.hV, ws.wsA,
                          ws.wsN, ws.wsP, self.overflow, *live128, rows=S * G) if node_tc else
                _lib.call("pp_ipmp_node_pre", W, layer, path, *common, *size, ws.hV, ws.wsA, ws.wsN, ws.wsP, rows=S * G))
            pre(0)
            if self.mode == "fp32":
                _lib.call("pp_ipmp_edge_node", W, layer, *common, *size, hE_in, shared, ws.wsA, ws.wsN, ws.wsP,
                          ws.wsAcc, rows=S * G)
            else:
                _lib.call("pp_ipmp_edge_tc", W, layer, 0, self.wtc[layer, 0], *common, graph.msum, G, K, S, hE_in, shared, ws.wsA,
                          ws.wsN, ws.wsP, ws.wsAcc, *tcp_node, self.overflow, *live, rows=S * G, tag="node")
            if self.mode == "fp32" or self.node_epilogue == "ffma":
                _lib.call("pp_ipmp_node_post", W, layer, *common, graph.msum, *size, ws.wsAcc, ws.hV, rows=S * G)
            else:
                _lib.call("pp_ipmp_node_post_tc32", W, layer, self.wtc[layer, 2], graph.msum, graph.mask, G, K, S,
                          ws.wsAcc, ws.hV, self.overflow, *live128, rows=S * G)
            if edge:
                pre(1)
                if self.mode == "fp32":
                    _lib.call("pp_ipmp_edge_edge", W, layer, *common, *size, hE_in, shared, ws.wsA, ws.wsN, ws.wsP,
                              ws.hE, rows=S * G)
                else:
                    _lib.call("pp_ipmp_edge_tc", W, layer, 1, self.wtc[layer, 1], *common, graph.msum, G, K, S, hE_in, shared,
                              ws.wsA, ws.wsN, ws.wsP, ws.hE, *tcp, self.overflow, *live, rows=S * G, tag="edge")
        return ws.hV

    def network(self, graph, batch, chi, t):
        """chi [S*G,4] (device), t [S*G] -> (score [S*G,4], h_V [S*G,128])  (TorsionalDiffusion.py:90-109)."""
        S = chi.shape[0] // graph.G
        ws = self.workspace(graph.G, graph.K, S)
        ni = graph.ni if getattr(graph, "ni", None) is not None else self.node_inputs(batch)
        self.forward_layers(graph, ws, ni, chi, t, 1)
        if self.mode != "fp32":  # tiles of pure padding are skipped by the tensor-core kernels: the API returns zeros there
            ws.hV.view(S, graph.G, 128).mul_(graph.mask.view(1, graph.G, 1))
        _lib.call("pp_decode_step", self.wblob, ws.hV, graph.G, S, ws.score, 0, 0.0, 0.0, None, None, None, None, None,
                  None, 0.0, 0, 0)
        return ws.score, ws.hV

    # ------------------------------------------------------------------ sampling
    @staticmethod
    def ode_coefficients(n_steps=30, annealed_temp=3.0, mode="ode"):
        """Per-step (t, c, w, d) with the reference's fp32 tensor arithmetic (schedule.py:165-235,286-288,
        TorsionalDiffusion.py:259-263): the schedule entries are 0-dim fp32 tensors, the numpy scalars fold in.
        ode: c = 0.5 g^2 dt, d = 0;  sde: c = g^2 dt, d = g sqrt(dt)."""
        sched = torch.linspace(1, 0, n_steps + 1)
        lo, hi = np.log(SIGMA_MIN), np.log(SIGMA_MAX)
        out = []
        for j in range(n_steps):
            time, dt = sched[j], sched[j] - sched[j + 1]
            sigma = torch.exp(lo + (hi - lo) * time)
            g = sigma * np.sqrt(2 * np.log(SIGMA_MAX / SIGMA_MIN))
            if annealed_temp:
                alpha = 1 - (sigma / np.exp(hi)) ** 2
                w = annealed_temp / (alpha + (1 - alpha) * annealed_temp)
            else:
                w = torch.tensor(1.0)
            if mode == "sde":
                c, d = g ** 2 * dt, g * torch.sqrt(dt)
            else:
                c, d = 0.5 * g ** 2 * dt, torch.tensor(0.0)
            out.append((float(time), float(c), float(w), float(d)))
        return out

    def _schedule(self, n_steps, annealed_temp, mode):
        """(coefficients, device tensor of the step times), cached: building them costs a host->device copy."""
        key = (n_steps, float(annealed_temp), mode)
        if key not in self._sched:
            coefs = self.ode_coefficients(n_steps, annealed_temp, mode)
            self._sched[key] = (coefs, torch.tensor([c[0] for c in coefs], dtype=torch.float32, device=self.dev))
        return self._sched[key]

    def _run_steps(self, graph, ws, ni, step_mask, chi, tvals, coefs, trajectory=None, sde=None, seed=0):
        """sde = (mask_1pi uint8 [G,4], noise [steps, 2, S*G, 4]) switches the update to the SDE branch with injected
        draws; sde = None with non-zero diffusion coefficients uses the in-kernel Philox stream of `seed`."""
        G, S = graph.G, ws.S
        for j, (_, c, w, d) in enumerate(coefs):
            self.forward_layers(graph, ws, ni, chi, tvals[j:j + 1], 0)
            n1, n2, m1 = (sde[1][j, 0], sde[1][j, 1], sde[0]) if sde is not None else (None, None, None)
            _lib.call("pp_decode_step", self.wblob, ws.hV, G, S, None, 1, c, w, step_mask, ni["chi_mask"], chi, n1, n2,
                      m1, d, int(seed), j)
            if trajectory is not None:
                trajectory.append(chi.clone())

    def sample(self, graph, batch, chi_init, n_steps=30, annealed_temp=3.0, trajectory=None, mode="ode",
               sde_noise=None, generator=None):
        """Reverse-ODE loop (TorsionalDiffusion.py:259-280) on S samples that share `graph`.

        chi_init [S*G,4] on the device; returns the final chi [S*G,4] (a new tensor).  Small problems that come back
        with the same buffers (decoy loops on one complex, same-shape batches through `build_graph(reuse=)`) replay a
        captured CUDA graph of the whole loop from the second call on."""
        G, K = graph.G, graph.K
        S = chi_init.shape[0] // G
        if getattr(graph, "ni", None) is not None:
            ni, step_mask = graph.ni, graph.step_mask
        else:  # a Graph built outside build_graph
            ni = self.node_inputs(batch)
            step_mask = (batch.chi_1pi_periodic_mask.to(self.dev).reshape(-1, 4).bool() |
                         batch.chi_2pi_periodic_mask.to(self.dev).reshape(-1, 4).bool()).to(torch.uint8).contiguous()
        coefs, tvals = self._schedule(n_steps, annealed_temp, mode)
        if mode == "sde":  # fresh noise every step: no graph replay
            chi = chi_init.clone().contiguous()
            if sde_noise is None:
                # one Philox stream per (row, step) inside the decode kernel, keyed by a seed drawn from `generator`
                # (or torch's default CPU generator): nothing of size [steps, 2, rows, 4] is materialised
                gdev = generator.device if generator is not None else torch.device("cpu")
                seed = int(torch.randint(0, 2 ** 62, (1,), device=gdev, generator=generator).item())
                self._run_steps(graph, self.workspace(G, K, S), ni, step_mask, chi, tvals, coefs, trajectory, seed=seed)
                return chi
            m1 = graph.mask_1pi if getattr(graph, "mask_1pi", None) is not None else \
                batch.chi_1pi_periodic_mask.to(self.dev).reshape(-1, 4).to(torch.uint8).contiguous()
            self._run_steps(graph, self.workspace(G, K, S), ni, step_mask, chi, tvals, coefs, trajectory,
                            sde=(m1, sde_noise.to(self.dev, torch.float32).contiguous()))
            return chi
        key = (S, n_steps, float(annealed_temp), self.mode, self.cluster, self.serial)
        small = S * G <= GRAPH_ROWS_MAX and trajectory is None and _lib.PROFILE is None and GRAPH_ROWS_MAX > 0
        if not small or (key not in graph._seen and key not in graph._replay):
            graph._seen.add(key)
            chi = chi_init.clone().contiguous()
            self._run_steps(graph, self.workspace(G, K, S), ni, step_mask, chi, tvals, coefs, trajectory)
            return chi
        st = graph._replay.get(key)
        if st is None:  # second call with these buffers: capture
            st = dict(ws=Workspace(G, K, S, self.dev), chi=chi_init.clone().contiguous(),
                      ni={k: v.clone() for k, v in ni.items()}, step_mask=step_mask.clone(), tvals=tvals)
            torch.cuda.synchronize(self.dev)
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                self._run_steps(graph, st["ws"], st["ni"], st["step_mask"], st["chi"], st["tvals"], coefs)
            st["graph"] = cg
            graph._replay[key] = st
        st["chi"].copy_(chi_init)
        for k, v in ni.items():
            st["ni"][k].copy_(v)
        st["step_mask"].copy_(step_mask)
        st["graph"].replay()
        return st["chi"].clone()

    # ------------------------------------------------------------------ atom14 / clash / proximal
    def atom14(self, X, residue_type, chi):
        """X [G,14,3], residue_type [G], chi [S*G,4] -> [S*G,14,3]."""
        G = X.shape[0]
        S = chi.shape[0] // G
        out = torch.empty(S * G, 14, 3, dtype=torch.float32, device=self.dev)
        _lib.call("pp_atom14_fwd", self.tables.geo, X, residue_type, chi, G, S, out)
        return out


class ClashContext:
    """Static part of the clash term for one padded batch: tables, bounds, residue neighbour list."""

    def __init__(self, dev, X, residue_type, atom_mask, residue_index, vtf=12.0, cot=0.5):
        self.dev = dev
        self.tables = DeviceTables.get(dev)
        self.B, self.L = int(X.shape[0]), int(X.shape[1])
        self.G = self.B * self.L
        self.X = _f32(X.reshape(self.G, 14, 3), dev)
        self.rtype = _i64(residue_type.reshape(-1), dev)
        self.exists = _f32(atom_mask.reshape(self.G, 14), dev)
        self.ridx = _i64(residue_index.reshape(-1), dev)
        self.tol = float(cot)
        self.lower, self.upper = self.tables.bounds(cot, vtf)
        self.max_cut = max(2.0 * self.tables.max_radius - float(cot), 0.0)
        G = self.G
        self.reach = torch.empty(G, dtype=torch.float32, device=dev)
        counts = torch.empty(G, dtype=torch.int32, device=dev)
        self.start = torch.zeros(G + 1, dtype=torch.int64, device=dev)
        if self.L >= CELL_LIST_MIN_L:
            # spatially hashed: bin the residues on CA, search the 27 surrounding cells
            _lib.call("pp_clash_reach", self.tables.geo, self.X, self.rtype, self.exists, G, self.reach)
            nc = int(_lib.load().pp_knn_cells_max()) + 1
            wi = torch.empty(self.B * 2 * nc + G, dtype=torch.int32, device=dev)
            wb = torch.empty(self.B * 8, dtype=torch.float32, device=dev)
            h_min = 2.0 * float(self.reach.max().item()) + self.max_cut + 1e-3
            args = (self.X, self.reach, self.ridx, self.B, self.L, self.max_cut, h_min)
            _lib.call("pp_clash_neighbours_cells", *args, 0, counts, None, None, wi, wb)
            self.start[1:] = torch.cumsum(counts.to(torch.int64), 0)
            total = int(self.start[-1].item())  # one host sync per complex, not per step
            self.list = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            _lib.call("pp_clash_neighbours_cells", *args, 1, None, self.start, self.list, wi, wb, kernels=1)
        else:
            args = (self.tables.geo, self.X, self.rtype, self.exists, self.ridx, self.B, self.L, self.max_cut)
            _lib.call("pp_clash_neighbours", *args, 0, self.reach, counts, None, None)
            self.start[1:] = torch.cumsum(counts.to(torch.int64), 0)
            total = int(self.start[-1].item())  # one host sync per complex, not per step
            self.list = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
            _lib.call("pp_clash_neighbours", *args, 1, self.reach, None, self.start, self.list)
        self._ws = {}
        self._prox = {}

    def scratch(self, S):
        if S not in self._ws:
            R = S * self.G
            z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=self.dev)  # noqa: E731
            self._ws[S] = dict(atoms4=z(R, 14, 4), axes=z(R, 4, 6), bound=z(R))
        return self._ws[S]

    def evaluate(self, chi, res_w=None):
        """chi [S*G,4] -> per_res [S*G] (and grad [S*G,4] of sum_r res_w[r]*per_res[r] when res_w is given)."""
        S = chi.shape[0] // self.G
        ws = self.scratch(S)
        per_res = torch.empty(S * self.G, dtype=torch.float32, device=self.dev)
        grad = torch.empty(S * self.G, 4, dtype=torch.float32, device=self.dev) if res_w is not None else None
        _lib.call("pp_clash_fwd_bwd", self.tables.geo, self.lower, self.upper, self.X, self.rtype, self.exists,
                  self.start, self.list, chi, self.G, S, self.tol, self.max_cut, 0 if res_w is None else 1, res_w,
                  per_res, grad, ws["atoms4"], ws["axes"], ws["bound"])
        return per_res, grad

    def n_res(self):
        """int32 [B]: residues of every complex without the padding (index of the last residue that has an atom + 1);
        the reference's means run over exactly these rows (it never sees a padded batch, optimize.py:27)."""
        if self.B == 1:
            return None  # one unpadded complex: every row counts, as in the reference
        if getattr(self, "_n_res", None) is None:
            has = self.exists.reshape(self.B, self.L, 14).sum(-1) > 0
            idx = torch.arange(1, self.L + 1, device=self.dev).unsqueeze(0)
            self._n_res = (has * idx).amax(1).clamp(min=1).to(torch.int32).contiguous()
        return self._n_res

    def _prox_run(self, st, S, lamda, num_steps, lr, beta1, beta2, eps):
        B, L = self.B, self.L
        ws = self.scratch(S)
        n_res = self.n_res()
        static = (self.tables.geo, self.lower, self.upper, self.X, self.rtype, self.exists, self.start, self.list,
                  st["sc_d"])
        _lib.call("pp_prox_init", *static, B, L, S, n_res, self.tol, self.max_cut, st["mask"], st["z"], st["x"], st["m"],
                  st["v"], st["per_res"], st["mean"], ws["atoms4"], ws["axes"], ws["bound"], st["loss_rows"])
        for k in range(num_steps):
            t = k + 1
            step_size = lr / (1 - beta1 ** t)  # torch.optim.Adam, single-tensor path
            bc2_sqrt = math.sqrt(1 - beta2 ** t)
            _lib.call("pp_prox_step", *static, st["mask"], st["z"], st["x"], st["m"], st["v"], B, L, S, n_res, self.tol,
                      self.max_cut, float(lamda), step_size, bc2_sqrt, beta1, beta2, eps, st["snaps"][k],
                      st["losses"][k - 1] if k else None, st["per_res"], ws["atoms4"], ws["axes"], ws["bound"],
                      st["loss_rows"], None, 0)
        _lib.call("pp_prox_loss", st["loss_rows"], B, L, S, n_res, float(lamda), 0, st["losses"][num_steps - 1])

    def proximal(self, sc_d, lamda, num_steps, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-8):
        """optimize.py:21-73 for every (sample, complex) item of the padded batch at once: sc_d [S*G, 4].
        Returns (snapshots [num_steps, S*G, 4], losses [num_steps, S*B], mask [S*G, 4]); everything stays on the
        device, the caller decides when to synchronise.  The 5 + 2*num_steps launches are captured in a CUDA graph
        the second time the same context runs the same schedule."""
        G, dev = self.G, self.dev
        sc_d = sc_d.reshape(-1, 4)
        S = sc_d.shape[0] // G
        if S * G != sc_d.shape[0]:
            raise RuntimeError("proximal: SC_D does not match the batch")
        R, items = S * G, S * self.B
        key = (S, int(num_steps), float(lamda), lr, beta1, beta2, eps)
        st = self._prox.get(key)
        if st is None:
            f = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
            st = dict(sc_d=f(R, 4), mask=torch.zeros(R, 4, dtype=torch.uint8, device=dev), z=f(R, 4), x=f(R, 4),
                      m=f(R, 4), v=f(R, 4), per_res=f(R), mean=f(items, 2), snaps=f(num_steps, R, 4),
                      losses=f(num_steps, items, 2), loss_rows=f(R, 2), calls=0, graph=None)
            self._prox = {key: st}  # one schedule at a time keeps the memory bounded
        st["sc_d"].copy_(sc_d)
        st["calls"] += 1
        if st["calls"] == 1 or GRAPH_ROWS_MAX <= 0:
            self._prox_run(st, S, lamda, num_steps, lr, beta1, beta2, eps)
        else:
            if st["graph"] is None:
                torch.cuda.synchronize(dev)
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    self._prox_run(st, S, lamda, num_steps, lr, beta1, beta2, eps)
                st["graph"] = cg
            st["graph"].replay()
        return st["snaps"].clone(), st["losses"][:, :, 0].clone(), st["mask"].clone()
