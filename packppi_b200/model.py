"""Drop-in mirrors of the reference's model classes for the sampling path, backed by libpackppi_b200.so.

Same class names, constructor arguments, method names, argument order, return shapes and `state_dict` keys as
  TDiffusionModule   (reference src/models/TorsionalDiffusion.py:21-298)
  ProteinEncoder     (src/models/components/encoder.py:59-246)
  MpnnNet            (src/models/components/mpnn.py:7-62)
so that `src/eval_diffusion.py` and the notebooks can switch by changing one import.  The
nn.Linear / nn.LayerNorm members only hold parameters; `forward` never runs them - it packs the parameters
into the kernel layout (re-packed when a parameter changes) and calls the CUDA kernels.  Inference only:
no autograd through the network, CUDA tensors only, `RuntimeError` otherwise.

Extensions, all trailing keyword arguments with None defaults: `sampling(..., init_SC_D=, noise=, n_samples=,
generator=)` to inject the initial noise (parity) and to draw several decoys that share the graph.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .components import proximal_optimizer
from .engine import Engine, Graph, SIGMA_MAX, SIGMA_MIN
from .weights import H, MSG_IN, N_POINTS

DEFAULT_ENCODER_CFG = dict(node_in=35, edge_in=468, node_features=128, edge_features=128,
                           time_embedding_type="sinusoidal", time_embedding_dim=16, num_positional_embeddings=16,
                           num_rbf=16, top_k=32, af2_relpos=True)  # configs/model/encoder_cfg/ProteinEncoder.yaml
DEFAULT_MODEL_CFG = dict(hidden_dim=128, num_mpnn_layers=3, n_points=8, dropout=0.1, act="relu", position_scale=1.0,
                         use_ipmp=True, k_neighbors=32)  # configs/model/model_cfg/MpnnNet.yaml
DEFAULT_SAMPLE_CFG = dict(eval_epochs=1, sample_during_training=True, annealed_temp=3, mode="ode", use_proximal=True,
                          violation_tolerance_factor=12., clash_overlap_tolerance=0.5, lamda=1.,
                          num_steps=50)  # configs/model/sample_cfg/Sampling.yaml


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _cfg(given, default):
    out = _Cfg(default)
    if given is not None:
        items = given.items() if hasattr(given, "items") else vars(given).items()
        out.update(items)
    return out


class MLP(nn.Module):
    """Parameter holder with the reference layout (layers.py:10-33)."""

    def __init__(self, num_in, num_inter, num_out, num_layers, act="relu", bias=True):
        super().__init__()
        self.W_in = nn.Linear(num_in, num_inter, bias=bias)
        self.W_inter = nn.ModuleList([nn.Linear(num_inter, num_inter, bias=bias) for _ in range(num_layers - 2)])
        self.W_out = nn.Linear(num_inter, num_out, bias=bias)


class InvariantPointMessagePassing(nn.Module):
    """Parameter holder with the reference layout (layers.py:36-63, edge_update=True)."""

    def __init__(self, node_dim=H, edge_dim=H, hidden_dim=H, n_points=N_POINTS, dropout=0.1, act="relu",
                 edge_update=True, position_scale=1.0):
        super().__init__()
        self.points_fn_node = nn.Linear(node_dim, n_points * 3)
        self.points_fn_edge = nn.Linear(node_dim, n_points * 3)
        self.node_message_fn = MLP(2 * node_dim + edge_dim + 9 * n_points, hidden_dim, hidden_dim, 3)
        self.edge_message_fn = MLP(2 * node_dim + edge_dim + 9 * n_points, hidden_dim, hidden_dim, 3)
        self.norm = nn.ModuleList([nn.LayerNorm(hidden_dim) for _ in range(4)])
        self.node_dense = MLP(hidden_dim, hidden_dim * 4, hidden_dim, 2)
        self.edge_dense = MLP(hidden_dim, hidden_dim * 4, hidden_dim, 2)


def _require_supported(cond, what):
    if not cond:
        raise NotImplementedError(f"packppi_b200 implements the shipped PackPPI-MSC configuration only: {what}")


class _PackedModule(nn.Module):
    """Caches an Engine (packed weights on one device) and rebuilds it when a parameter changes.

    `kernel_mode` selects how the per-edge message MLPs run (engine.Engine.MODES): "fp32" = CUDA-core FFMA (exact),
    "f16x3" = tcgen05 tensor cores with split fp16 operand pairs (fp32-grade, default), "f16" = tensor cores, plain
    fp16 inputs (fast, looser tolerance).  `kernel_cluster` = CTAs per cluster sharing the weight stream in the
    tensor-core modes (1 or 2).  `kernel_node_epilogue` = "tc32" (default) | "ffma": where the per-residue node update runs.
    Defaults come from the environment (PACKPPI_B200_MODE, PACKPPI_B200_CLUSTER, PACKPPI_B200_NODE_EPILOGUE)."""
    kernel_mode = os.environ.get("PACKPPI_B200_MODE", "f16x3")
    kernel_cluster = int(os.environ.get("PACKPPI_B200_CLUSTER", "1"))
    kernel_node_epilogue = os.environ.get("PACKPPI_B200_NODE_EPILOGUE") or None
    check_finite = os.environ.get("PACKPPI_B200_CHECK_FINITE", "1") != "0"  # overflow guard of the split-fp16 modes

    def _full_state_dict(self):
        raise NotImplementedError

    def engine(self, device):
        device = torch.device(device)
        sig = (str(device), self.kernel_mode, self.kernel_cluster, self.kernel_node_epilogue) + tuple((p.data_ptr(), p._version)
                                                                            for p in self.parameters())
        if getattr(self, "_engine_sig", None) != sig:
            object.__setattr__(self, "_engine_obj", Engine(self._full_state_dict(), device, self.kernel_mode,
                                                           self.kernel_cluster, self.kernel_node_epilogue))
            object.__setattr__(self, "_engine_sig", sig)
        return self._engine_obj


def _zeros_like_shapes(prefix_filter):
    from .weights import shapes
    return {k: torch.zeros(s) for k, s in shapes().items() if not prefix_filter(k)}


class ProteinEncoder(_PackedModule):
    def __init__(self, node_in, edge_in, node_features, edge_features, time_embedding_type="sinusoidal",
                 time_embedding_dim=16, num_positional_embeddings=16, num_rbf=16, top_k=32, af2_relpos=True):
        super().__init__()
        _require_supported(node_in + time_embedding_dim == 51 and edge_in == 468 and node_features == H and
                           edge_features == H and time_embedding_type == "sinusoidal" and num_rbf == 16 and
                           top_k == 32 and af2_relpos, "encoder_cfg must equal configs/model/encoder_cfg/ProteinEncoder.yaml")
        self.node_embedding = nn.Linear(node_in + time_embedding_dim, node_features, bias=True)
        self.norm_nodes = nn.LayerNorm(node_features)
        self.edge_embedding = nn.Linear(edge_in, edge_features, bias=True)
        self.norm_edges = nn.LayerNorm(edge_features)
        self.top_k = top_k
        self.num_rbf = num_rbf

    def _full_state_dict(self):
        sd = _zeros_like_shapes(lambda k: k.startswith("encoder."))
        sd.update({"encoder." + k: v for k, v in self.state_dict().items()})
        return sd

    def _dist(self, X, mask, eps=1E-6):
        """encoder.py:105-118: (D_neighbors [B,L,K], E_idx [B,L,K] int64, mask_neighbors [B,L,K,1]).  `X` is X_ca."""
        B, L = X.shape[:2]
        X14 = torch.zeros(B, L, 14, 3, dtype=torch.float32, device=X.device)
        X14[:, :, 1] = X
        g = Graph(X14, mask, self.top_k)
        return g.D_neighbors, g.E_idx, g.mask_attend.reshape(B, L, g.K, 1)

    def forward(self, X, S, BB_D_sincos, SC_D_sincos, chain_indices, mask, residue_index=None, t=None):
        """encoder.py:198-246 -> (h_V [B,L,128], h_E [B,L,K,128], E_idx [B,L,K], X)."""
        if residue_index is None or t is None:
            raise NotImplementedError("packppi_b200.ProteinEncoder needs residue_index and t (the MSC configuration)")
        eng = self.engine(X.device)
        B, L = X.shape[:2]
        g = Graph(X, mask, self.top_k)
        g.edge_embed(eng.wblob, residue_index, chain_indices)
        hV = torch.empty(B * L, H, dtype=torch.float32, device=X.device)
        _lib.call("pp_node_embed", eng.wblob, S.reshape(-1).to(torch.int64).contiguous(),
                  BB_D_sincos.reshape(-1, 6).float().contiguous(), None, None,
                  SC_D_sincos.reshape(-1, 8).float().contiguous(), t.reshape(-1).float().contiguous(), 1, B * L, 1, hV)
        return hV.reshape(B, L, H), g.hE0.reshape(B, L, g.K, H), g.E_idx, X


class MpnnNet(_PackedModule):
    def __init__(self, node_features=128, edge_features=128, hidden_dim=128, num_mpnn_layers=3, n_points=8,
                 dropout=0.1, act="relu", position_scale=1.0, use_ipmp=True, k_neighbors=32):
        super().__init__()
        _require_supported(node_features == H and edge_features == H and hidden_dim == H and num_mpnn_layers == 3 and
                           n_points == N_POINTS and act == "relu" and float(position_scale) == 1.0 and use_ipmp,
                           "model_cfg must equal configs/model/model_cfg/MpnnNet.yaml")
        self.use_ipmp = use_ipmp
        self.mpnn_layers = nn.ModuleList([InvariantPointMessagePassing() for _ in range(num_mpnn_layers)])

    def _full_state_dict(self):
        sd = _zeros_like_shapes(lambda k: k.startswith("mpnn."))
        sd.update({"mpnn." + k: v for k, v in self.state_dict().items()})
        return sd

    def forward(self, h_V, h_E, E_idx, X, S, mask):
        """mpnn.py:47-62 -> h_V [B,L,128] (the edge update of the last layer is dead work and skipped)."""
        eng = self.engine(h_V.device)
        B, L, K = E_idx.shape
        g = Graph.__new__(Graph)
        g.B, g.L, g.G, g.K = B, L, B * L, K
        dev = h_V.device
        g.X = X.float().contiguous()
        g.mask = mask.float().reshape(-1).contiguous()
        off = (torch.arange(B, device=dev) * L).view(B, 1, 1)
        g.nbr = (E_idx + off).to(torch.int32).reshape(B * L, K).contiguous()
        mj = torch.gather(g.mask.reshape(B, L), 1, E_idx.reshape(B, L * K)).reshape(B, L, K)
        g.mask_attend = (g.mask.reshape(B, L, 1) * mj).reshape(B * L, K).contiguous()
        g.msum = (g.mask_attend.sum(-1) / K).contiguous()
        g.geo = torch.empty(B * L, _lib.load().pp_geo_stride(), dtype=torch.float32, device=dev)
        _lib.call("pp_geometry_build", g.X, B * L, g.geo)
        ws = eng.workspace(g.G, K, 1)
        ws.hV.copy_(h_V.reshape(B * L, H))
        hE0 = h_E.reshape(B * L, K, H).float().contiguous()
        for layer in range(3):
            first = layer == 0
            _lib.call("pp_ipmp_layer", eng.wblob, layer, g.geo, g.nbr, g.mask_attend, g.msum, g.mask, g.G, K, 1, ws.hV,
                      hE0 if first else ws.hE, 0, ws.hE, 1 if layer < 2 else 0, ws.wsA, ws.wsN, ws.wsP, ws.wsAcc)
        return ws.hV.reshape(B, L, H).clone()


class TDiffusionModule(_PackedModule):
    def __init__(self, optimizer=None, scheduler=None, encoder_cfg=None, model_cfg=None, sample_cfg=None, **kwargs):
        super().__init__()
        self.NUM_CHI_ANGLES = 4
        self.eps = 1e-6
        enc, mdl, smp = _cfg(encoder_cfg, DEFAULT_ENCODER_CFG), _cfg(model_cfg, DEFAULT_MODEL_CFG), \
            _cfg(sample_cfg, DEFAULT_SAMPLE_CFG)
        self.hparams = _Cfg(optimizer=optimizer, scheduler=scheduler, encoder_cfg=enc, model_cfg=mdl, sample_cfg=smp,
                            **kwargs)
        self.encoder = ProteinEncoder(enc.node_in, enc.edge_in, enc.node_features, enc.edge_features,
                                      enc.time_embedding_type, enc.time_embedding_dim, enc.num_positional_embeddings,
                                      enc.num_rbf, enc.top_k, enc.af2_relpos)
        self.mpnn = MpnnNet(enc.node_features, enc.edge_features, mdl.hidden_dim, mdl.num_mpnn_layers, mdl.n_points,
                            mdl.dropout, mdl.act, mdl.position_scale, mdl.use_ipmp, mdl.k_neighbors)
        self.decoder_score = nn.ModuleList([MLP(H, H // 2, H // 4, 2), nn.ReLU(), MLP(H // 4, H // 8, 4, 2)])
        self.schedule = torch.linspace(1, 0, 31)  # SO2VESchedule.reverse_t_schedule (schedule.py:286-288)
        for p in self.parameters():  # TorsionalDiffusion.py:80-82
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self._graph_cache = (None, None)

    # -- plumbing ---------------------------------------------------------------------------------------
    @property
    def device(self):
        return next(self.parameters()).device

    def _full_state_dict(self):
        return self.state_dict()

    def _fp32_engine(self, device):
        """Exact-fp32 engine with the same weights (the overflow fallback of the tensor-core modes)."""
        sig = getattr(self, "_engine_sig", None)
        if getattr(self, "_fp32_sig", None) != sig:
            object.__setattr__(self, "_fp32_obj", Engine(self._full_state_dict(), device, "fp32"))
            object.__setattr__(self, "_fp32_sig", sig)
        return self._fp32_obj

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=False, **kwargs):
        """Lightning-style checkpoint ({'state_dict': ...}) or a bare state_dict (eval_diffusion.py:33-40)."""
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        sd = ckpt.get("state_dict", ckpt)
        model = cls(**kwargs)
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=strict)
        return model

    def _graph(self, batch):
        """Graph + edge embedding of `batch`, reused while the same tensors (and weights) are passed again.

        The cache key holds the tensors that define the graph themselves (identity + version counter), not `id()`s or
        data pointers: both are recycled as soon as a batch is freed, and a recycled key once served the graph of a
        5-residue complex to a 17-residue one."""
        eng = self.engine(batch.X.device)
        tensors = (batch.X, batch.residue_mask, batch.residue_index, batch.chain_indices, batch.residue_type,
                   batch.BB_D_sincos, batch.SC_D_mask, batch.chi_1pi_periodic_mask, batch.chi_2pi_periodic_mask)
        key = (tensors, tuple(t._version for t in tensors), getattr(self, "_engine_sig", None))
        old = self._graph_cache[0]
        same = (old is not None and old[2] == key[2] and old[1] == key[1]
                and all(a is b for a, b in zip(old[0], key[0])))
        if not same:
            prev = self._graph_cache[1]
            if prev is not None and (prev.B, prev.L) != tuple(batch.X.shape[:2]):
                prev = None
                self._graph_cache = (None, None)  # release the old buffers before the new ones are allocated
            self._graph_cache = (key, eng.build_graph(batch, reuse=prev))
        return eng, self._graph_cache[1]

    # -- reference API ----------------------------------------------------------------------------------
    def network(self, batch, SC_D_noised, t):
        """TorsionalDiffusion.py:90-109 -> (pred_score [B,L,4], h_V [B,L,128]).  Unlike the reference's
        SinusoidalEmbedding (layers.py:258) the caller's `t` is not scaled in place."""
        eng, g = self._graph(batch)
        B, L = g.B, g.L
        chi = SC_D_noised.reshape(-1, 4).to(torch.float32).contiguous()
        if chi.shape[0] == 0 or chi.shape[0] % (B * L) != 0:
            raise ValueError(f"network: SC_D_noised has {chi.shape[0]} rows, not a multiple of B*L = {B * L}")
        lead = chi.shape[0] // (B * L)
        t = t.reshape(-1).to(device=chi.device, dtype=torch.float32)
        if t.numel() == B * L and lead > 1:
            t = t.repeat(lead)  # one time per residue, shared by the samples
        elif t.numel() != chi.shape[0]:
            raise ValueError(f"network: t has {t.numel()} entries; expected B*L = {B * L} or S*B*L = {chi.shape[0]}")
        t = t.contiguous()
        check = eng.mode != "fp32" and self.check_finite
        if check:
            eng.overflow.zero_()
        score, hV = eng.network(g, batch, chi, t)
        if check and bool(eng.overflow.item()):  # same guard as in `sampling`: an activation left the fp16 range
            import warnings
            warnings.warn("packppi_b200: an activation left the fp16 range in the split-fp16 tensor-core mode; repeating "
                          "this call with the fp32 CUDA-core kernels")
            score, hV = self._fp32_engine(chi.device).network(g, batch, chi, t)
        shape = (B, L) if lead == 1 else (lead, B, L)
        return score.reshape(*shape, 4).clone(), hV.reshape(*shape, H).clone()

    @staticmethod
    def _wrapped_normal_score(noise, sigma, PI):
        """SO2Schedule.score (schedule.py:64-73): the reference looks the score of the wrapped normal up in a
        5001 x 5001 table over log-spaced (x, sigma) grids (schedule.py:39-54, 200 MB per schedule, 20 s to build).
        Here the table ENTRY the reference would read - same index rounding, same 201-image sums of `p` and `grad`
        (schedule.py:10-22) in float64 - is evaluated on the fly for the values that are needed."""
        X_MIN, X_N, S_MIN, S_MAX, S_N, N = 1e-5, 5000, 3e-3, 2.0, 5000, 100
        # the index arithmetic runs in fp32 like the reference's (numpy on the fp32 arrays it is handed)
        x = (noise.float() + PI) % (2 * PI) - PI
        sign = torch.sign(x).double()
        xi = torch.round(torch.clip((torch.log(x.abs() / PI + 1e-10) - np.log(X_MIN)) / (0 - np.log(X_MIN)) * X_N, 0, X_N))
        si = torch.round(torch.clip((torch.log(sigma.float() / PI) - np.log(S_MIN)) / (np.log(S_MAX) - np.log(S_MIN)) * S_N,
                                    0, S_N))
        xi, si = xi.double(), si.double()
        xg = 10 ** (np.log10(X_MIN) + xi * ((0 - np.log10(X_MIN)) / X_N)) * PI          # SO2Schedule.x[xi]
        sg = 10 ** (np.log10(S_MIN) + si * ((np.log10(S_MAX) - np.log10(S_MIN)) / S_N)) * PI  # SO2Schedule.sigma[si]
        xg = torch.where(xi == X_N, torch.full_like(xg, PI), xg)      # np.linspace ends exactly on its stop value
        sg = torch.where(si == S_N, torch.full_like(sg, S_MAX * PI), sg)
        k = torch.arange(-N, N + 1, device=x.device, dtype=torch.float64)
        y = xg.unsqueeze(-1) + 2 * PI * k
        sg = sg.expand_as(xg).unsqueeze(-1)
        e = torch.exp(-y ** 2 / 2 / sg ** 2)
        p, g = e.sum(-1), (y / sg ** 2 * e).sum(-1)
        return (-sign * (g / torch.where(p == 0, torch.full_like(p, 1e-10), p))).to(noise.dtype)

    @torch.no_grad()
    def add_sc_noise(self, batch, t, noise=None, generator=None):
        """TorsionalDiffusion.py:111-124 / schedule.py:176-196 -> (noised angles, score target), both [B, L, 4].
        `noise` = (eps_1pi, eps_2pi), each [B*L,4] standard normal, injects the two randn draws.  The score target
        (training only; `sampling` drops it) is the reference's table value, evaluated instead of looked up."""
        x = batch.SC_D.reshape(-1, 4)
        dev = x.device
        sigma = torch.exp(np.log(SIGMA_MIN) + (np.log(SIGMA_MAX) - np.log(SIGMA_MIN)) * t.to(dev)).unsqueeze(-1)
        if noise is None:
            noise = (torch.randn(x.shape, device=dev, dtype=x.dtype, generator=generator),
                     torch.randn(x.shape, device=dev, dtype=x.dtype, generator=generator))
        m1, m2 = batch.chi_1pi_periodic_mask.reshape(-1, 4), batch.chi_2pi_periodic_mask.reshape(-1, 4)
        n1, n2 = noise[0].to(dev).reshape(-1, 4) * sigma, noise[1].to(dev).reshape(-1, 4) * sigma
        score = torch.where(m1, self._wrapped_normal_score(n1, sigma, 0.5 * np.pi) * m1,
                            self._wrapped_normal_score(n2, sigma, np.pi) * m2)
        x = x + n1 * m1
        x = x + n2 * m2
        x = (x + np.pi) % (2 * np.pi) - np.pi
        return x.reshape(batch.num_proteins, -1, 4), score.reshape(batch.num_proteins, -1, 4)

    @torch.no_grad()
    def _initial_samples(self, batch, S, noise, generator):
        """add_sc_noise at t = 1 for S decoys in one shot: [S, B, L, 4].  noise = (eps_1pi, eps_2pi), each [S, B*L, 4]."""
        x = batch.SC_D.reshape(1, -1, 4)
        dev = x.device
        # t_to_sigma(t = 1) with the fp32 tensor arithmetic of add_sc_noise (schedule.py:165-174)
        sigma = torch.exp(np.log(SIGMA_MIN) + (np.log(SIGMA_MAX) - np.log(SIGMA_MIN)) * torch.ones(1, device=dev))
        if noise is None:
            shape = (S, x.shape[1], 4)
            noise = (torch.randn(shape, device=dev, dtype=x.dtype, generator=generator),
                     torch.randn(shape, device=dev, dtype=x.dtype, generator=generator))
        n1 = noise[0].to(dev).reshape(S, -1, 4)
        n2 = noise[1].to(dev).reshape(S, -1, 4)
        x = x + (n1 * sigma) * batch.chi_1pi_periodic_mask.reshape(1, -1, 4)
        x = x + (n2 * sigma) * batch.chi_2pi_periodic_mask.reshape(1, -1, 4)
        x = (x + np.pi) % (2 * np.pi) - np.pi
        return x.reshape(S, batch.num_proteins, -1, 4)

    def forward(self, batch):
        raise NotImplementedError("training (score-matching loss) is outside the hot path this package replaces; "
                                  "train with the reference and load its state_dict here")

    def sampling(self, batch, use_proximal=False, return_list=False, init_SC_D=None, noise=None, n_samples=None,
                 generator=None, sde_noise=None):
        """TorsionalDiffusion.py:254-298.  Returns SC_D_sample [B,L,4]; with use_proximal the accepted proximal
        result; with return_list (SC_D_sample, list of 50 [1,L,4] tensors, list of 50 floats).  With B > 1 or
        n_samples the proximal stage runs all (sample, complex) items in one batch (components.proximal_optimizer) and
        the accept rule is applied per item.

        n_samples = S draws S decoys that share graph and edge embedding and returns [S,B,L,4]
        (init_SC_D / noise then carry a leading S).  With sample_cfg.mode = "sde" (schedule.py:224-228) every step adds
        g sqrt(dt) * noise; `sde_noise` [steps, 2, S*B*L, 4] injects the two torch.normal draws per step."""
        eng, g = self._graph(batch)
        B, L = g.B, g.L
        S = 1 if n_samples is None else int(n_samples)
        dev = batch.X.device
        if init_SC_D is None:
            init_SC_D = self._initial_samples(batch, S, noise if n_samples is not None or noise is None
                                              else (noise[0][None], noise[1][None]), generator)
        chi0 = init_SC_D.to(device=dev, dtype=torch.float32).reshape(S * B * L, 4).contiguous()
        smp = self.hparams.sample_cfg
        if smp.mode not in ("ode", "sde"):
            raise NotImplementedError(f"sample_cfg.mode = {smp.mode!r}: the reference knows 'ode' and 'sde'")
        run = lambda e: e.sample(g, batch, chi0, n_steps=len(self.schedule) - 1, annealed_temp=smp.annealed_temp,  # noqa: E731
                                 mode=smp.mode, sde_noise=sde_noise, generator=generator)
        check = eng.mode != "fp32" and self.check_finite
        if check:
            eng.overflow.zero_()
        chi = run(eng)
        if check and bool(eng.overflow.item()):
            # The tensor-core modes split fp32 activations into fp16 (hi, lo) pairs without a per-tile scale: a hidden
            # activation above 65504 (possible with an unusually scaled checkpoint; never seen with xavier-initialised or
            # LayerNorm-bounded values) becomes inf, the product NaN, and the next ReLU would turn that into a plausible
            # zero.  The kernels raise a flag when they split such a value; redo the call in exact fp32.
            import warnings
            warnings.warn("packppi_b200: an activation left the fp16 range in the split-fp16 tensor-core mode; repeating "
                          "this call with the fp32 CUDA-core kernels")
            chi = run(self._fp32_engine(dev))
        SC_D_sample = chi.reshape(S, B, L, 4) if n_samples is not None else chi.reshape(B, L, 4)
        if not use_proximal:
            return SC_D_sample
        SC_D_resample_list, loss_list = proximal_optimizer(batch, SC_D_sample, smp.violation_tolerance_factor,
                                                           smp.clash_overlap_tolerance, smp.lamda, smp.num_steps)
        if return_list:
            return SC_D_sample, SC_D_resample_list, loss_list
        if not torch.is_tensor(loss_list[0]):  # one complex, one sample: the reference's accept rule (:295-298)
            if loss_list[-1] < loss_list[0]:
                return SC_D_resample_list[-1]
            return SC_D_sample
        # batched (B > 1 and / or n_samples): the same accept rule per (sample, complex) item
        better = (loss_list[-1] < loss_list[0])[..., None, None]
        return torch.where(better, SC_D_resample_list[-1], SC_D_sample)

    def compute_rmsd(self, true_coords, pred_coords, atom_mask, residue_mask):
        """TorsionalDiffusion.py:300-309 (mean squared deviation; the reference never takes the root)."""
        err = torch.sum((true_coords - pred_coords) ** 2, dim=-1) * atom_mask * residue_mask[..., None]
        count = torch.sum(atom_mask * residue_mask[..., None] + self.eps, dim=-1)
        return torch.sum(err) / torch.sum(count)

    def analyze_samples(self, batch, SC_D_sample=None):
        """TorsionalDiffusion.py:311-341: chi MAE / accuracy per chi and atom RMSD (host-side metric arithmetic)."""
        from .components import get_atom14_coords
        true, pred = batch["SC_D"].clone(), SC_D_sample.clone()
        m, p1 = batch["SC_D_mask"], batch["chi_1pi_periodic_mask"]
        metric = {}
        for i in range(self.NUM_CHI_ANGLES):
            n = m[..., i].sum()
            n = 1 if n == 0 else n
            diff = (pred[..., i] - true[..., i]).abs()
            acc = torch.where(torch.logical_and(diff * 180 / np.pi < 20, diff > 0), 1., 0.)
            ae = torch.minimum(diff, 2 * np.pi - diff)
            ae = torch.where(p1[..., i], torch.minimum(ae, np.pi - ae), ae)
            metric[f"chi_{i}_ae_rad"] = ae.sum() / n
            metric[f"chi_{i}_ae_deg"] = (ae * 180 / np.pi).sum() / n
            metric[f"chi_{i}_acc"] = acc.sum() / n
        xyz = get_atom14_coords(batch.X, batch.residue_type, batch.BB_D, SC_D_sample)
        metric["atom_rmsd"] = self.compute_rmsd(batch.X, xyz, batch.atom_mask, batch.residue_mask)
        return metric
