"""CPU ORACLE (test infrastructure, not product code) for PackPPI-MSC's reverse-diffusion sampling step.

A torch-CPU fp32 restatement of the reference algorithm, written from the formulas and kept dense like the
reference so that it is also a fair stand-in for the reference's `--device cpu` path when timed
(bench.py `cpu_baseline`, kind "port").  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it.  Each function cites the reference lines it follows (paths under /root/reference/src).

Pinning: tests/test_oracle_golden.py checks every function here against tests/golden/*.npz, vectors produced
by running the UNMODIFIED reference under tools/ref_shims.py (generator: tools/make_golden.py).  The reference
itself ships no tests or golden vectors for this path.
"""
import math

import torch
import torch.nn.functional as F

TWO_PI = 2 * math.pi


# ---------------------------------------------------------------------------------------------- graph
def knn_graph(X_ca, mask, top_k=32, eps=1e-6):
    """models/components/encoder.py:105-118 (`ProteinEncoder._dist`).

    D = mask2D * sqrt(|xi - xj|^2 + eps); masked candidates sit at 2*rowmax(D); K smallest, ascending.
    torch.topk leaves the order of equal keys unspecified; the oracle fixes it to "lowest index first"
    (stable sort), which is the contract of the CUDA kernel.
    """
    m2 = mask[:, None, :] * mask[:, :, None]
    dX = X_ca[:, None, :, :] - X_ca[:, :, None, :]
    D = m2 * torch.sqrt((dX ** 2).sum(3) + eps)
    Dmax = D.max(-1, keepdim=True)[0]
    Dadj = D + 2 * (1.0 - m2) * Dmax
    K = min(top_k, X_ca.shape[1])
    vals, idx = torch.sort(Dadj, dim=-1, stable=True)
    return vals[..., :K].contiguous(), idx[..., :K].contiguous()


def _gather_pairs(M, E_idx):
    """[B,L,L,...] at neighbour indices [B,L,K] -> [B,L,K,...] (components/__init__.py:9-13)."""
    idx = E_idx.reshape(*E_idx.shape, *([1] * (M.dim() - 3))).expand(*E_idx.shape, *M.shape[3:])
    return torch.gather(M, 2, idx)


def _gather_rows(V, E_idx):
    """[B,L,C] at [B,L,K] -> [B,L,K,C] (components/__init__.py:16-30)."""
    B, L, K = E_idx.shape
    flat = E_idx.reshape(B, L * K, 1).expand(-1, -1, V.shape[-1])
    return torch.gather(V, 1, flat).reshape(B, L, K, V.shape[-1])


def _dihedral(p0, p1, p2, p3):
    """encoder.py:156-174: sign * arccos(n1.n2), NaN -> 0 (normalisation and the final value)."""
    u0, u1, u2 = p2 - p1, p0 - p1, p3 - p2

    def unit(v):
        return torch.nan_to_num(v / torch.norm(v, dim=-1, keepdim=True))

    n1 = unit(torch.cross(u0, u1, dim=-1))
    n2 = unit(torch.cross(u0, u2, dim=-1))
    sgn = torch.sign((torch.cross(u1, u2, dim=-1) * u0).sum(-1))
    return torch.nan_to_num(sgn * torch.arccos((n1 * n2).sum(-1)))


def edge_features(X, E_idx, residue_index, chain_indices, num_rbf=16):
    """468 raw edge features, encoder.py:34-47,120-153,176-196,231-236.

    [one_hot(clip(ridx_i - ridx_j + 32, 0, 64), 65) | 25 atom pairs (N,CA,C,O,CB)^2 x 16 RBF | 1 + same_chain |
     dih(C_i,N_j,CA_j,C_j) | dih(N_i,CA_i,C_i,N_j)]
    """
    N, CA, C, O = X[:, :, 0], X[:, :, 1], X[:, :, 2], X[:, :, 3]
    b, c = CA - N, C - CA
    CB = -0.58273431 * torch.cross(b, c, dim=-1) + 0.56802827 * b - 0.54067466 * c + CA  # encoder.py:137-142
    atoms = (N, CA, C, O, CB)

    off = residue_index[:, :, None] - residue_index[:, None, :]
    rel = torch.clip(_gather_pairs(off, E_idx) + 32, 0, 64)
    feats = [F.one_hot(rel, 65).float()]

    mu = torch.linspace(0.0, 20.0, num_rbf).view(1, 1, 1, -1)
    sigma = 20.0 / num_rbf
    for A in atoms:
        for Bm in atoms:
            D = torch.sqrt(((A[:, :, None, :] - Bm[:, None, :, :]) ** 2).sum(-1) + 1e-6)
            Dn = _gather_pairs(D, E_idx)
            feats.append(torch.exp(-(((Dn[..., None] - mu) / sigma) ** 2)))

    same = (chain_indices[:, :, None] == chain_indices[:, None, :]).float()
    feats.append((_gather_pairs(same, E_idx) + 1)[..., None])

    L = X.shape[1]
    ex_i = lambda A: A[:, :, None, :].expand(-1, -1, L, -1)  # noqa: E731
    ex_j = lambda A: A[:, None, :, :].expand(-1, L, -1, -1)  # noqa: E731
    phi = _dihedral(ex_i(C), ex_j(N), ex_j(CA), ex_j(C))
    psi = _dihedral(ex_i(N), ex_i(CA), ex_i(C), ex_j(N))
    feats.append(torch.stack((_gather_pairs(phi, E_idx), _gather_pairs(psi, E_idx)), -1))
    return torch.cat(feats, -1)


def time_embedding(t, dim=16, scale=10000.0, max_positions=10000):
    """layers.py:248-268 (`SinusoidalEmbedding`); the caller's tensor is NOT scaled in place here."""
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
    a = (t * scale).float()[:, None] * f[None, :]
    return torch.cat([torch.sin(a), torch.cos(a)], 1)


def node_features(S, BB_D_sincos, SC_D_sincos, t):
    """51 raw node features, encoder.py:217-229,239."""
    B, L = S.shape
    return torch.cat([F.one_hot(S, 21).float(), BB_D_sincos.reshape(B, L, 6), SC_D_sincos.reshape(B, L, 8),
                      time_embedding(t).reshape(B, L, 16)], -1)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def _mlp(sd, name, x, n_inter):
    """layers.py:10-33 with relu."""
    x = F.relu(_lin(sd, name + ".W_in", x))
    for i in range(n_inter):
        x = F.relu(_lin(sd, f"{name}.W_inter.{i}", x))
    return _lin(sd, name + ".W_out", x)


def edge_embedding(sd, batch, E_idx):
    """Step-invariant h_E0 = LN(Linear(468 features)), encoder.py:243-244."""
    E = edge_features(batch["X"], E_idx, batch["residue_index"], batch["chain_indices"])
    return _ln(sd, "encoder.norm_edges", _lin(sd, "encoder.edge_embedding", E))


# ---------------------------------------------------------------------------------------------- IPMP
def backbone_frames(X, eps=1e-8):
    """utils/features.py:90 -> utils/rigid_utils.py:1126-1179 with fixed=True.

    R = [e0 e1 e0xe1] (columns), e0 ~ C - CA, e1 ~ (N - CA) orthogonalised against e0, origin CA.
    """
    N, CA, C = X[..., 0, :], X[..., 1, :], X[..., 2, :]
    e0 = C - CA
    e1 = N - CA
    e0 = e0 / torch.sqrt((e0 * e0).sum(-1, keepdim=True) + eps)
    e1 = e1 - e0 * (e0 * e1).sum(-1, keepdim=True)
    e1 = e1 / torch.sqrt((e1 * e1).sum(-1, keepdim=True) + eps)
    e2 = torch.cross(e0, e1, dim=-1)
    return torch.stack((e0, e1, e2), -1), CA


def message_input(sd, pts_name, h_V, h_E, E_idx, R, t):
    """456-wide message rows, layers.py:65-117 (position_scale = 1)."""
    B, L, K = E_idx.shape
    p_loc = _lin(sd, pts_name, h_V).reshape(B, L, 8, 3)
    p_glob = torch.einsum("blij,blnj->blni", R, p_loc) + t[:, :, None, :]
    pg_j = _gather_rows(p_glob.reshape(B, L, 24), E_idx).reshape(B, L, K, 8, 3)
    p_loc_k = p_loc[:, :, None].expand(-1, -1, K, -1, -1)
    n_loc = torch.sqrt((p_loc_k ** 2).sum(-1) + 1e-8)
    q = torch.einsum("blji,blknj->blkni", R, pg_j - t[:, :, None, None, :])  # R^T (p_glob_j - t_i)
    n_q = torch.sqrt((q ** 2).sum(-1) + 1e-8)
    n_g = torch.sqrt(((p_glob[:, :, None] - pg_j) ** 2).sum(-1) + 1e-8)
    return torch.cat([h_V[:, :, None, :].expand(-1, -1, K, -1), h_E, _gather_rows(h_V, E_idx),
                      p_loc_k.reshape(B, L, K, 24), n_loc, q.reshape(B, L, K, 24), n_q, n_g], -1)


def ipmp_layer(sd, l, h_V, h_E, E_idx, R, t, mask_V, mask_att, edge_update=True):
    """layers.py:119-148.  mean over K divides by K, not by the number of valid neighbours."""
    p = f"mpnn.mpnn_layers.{l}."
    m = _mlp(sd, p + "node_message_fn", message_input(sd, p + "points_fn_node", h_V, h_E, E_idx, R, t), 1)
    m = (m * mask_att[..., None]).mean(-2)
    h_V = _ln(sd, p + "norm.0", h_V + m)
    h_V = _ln(sd, p + "norm.1", h_V + _mlp(sd, p + "node_dense", h_V, 0))
    h_V = h_V * mask_V[..., None]
    if edge_update:
        m = _mlp(sd, p + "edge_message_fn", message_input(sd, p + "points_fn_edge", h_V, h_E, E_idx, R, t), 1)
        h_E = _ln(sd, p + "norm.2", h_E + m * mask_att[..., None])
        h_E = _ln(sd, p + "norm.3", h_E + _mlp(sd, p + "edge_dense", h_E, 0))
        h_E = h_E * mask_att[..., None]
    return h_V, h_E


def decoder(sd, h_V):
    """models/TorsionalDiffusion.py:62-68,106-108: MLP(128,64,32) - ReLU - MLP(32,16,4)."""
    x = _mlp(sd, "decoder_score.0", h_V, 0)
    return _mlp(sd, "decoder_score.2", F.relu(x), 0)


class GraphCache:
    """Step-invariant part of `network`: E_idx, h_E0, frames, attention mask."""

    def __init__(self, sd, batch, top_k=32):
        self.E_idx = knn_graph(batch["X"][:, :, 1, :], batch["residue_mask"], top_k)[1]
        self.h_E0 = edge_embedding(sd, batch, self.E_idx)
        self.R, self.t = backbone_frames(batch["X"])
        m = batch["residue_mask"]
        self.mask_att = m[..., None] * _gather_rows(m[..., None], self.E_idx)[..., 0]  # mpnn.py:49-50


def network(sd, batch, SC_D_noised, t, cache=None, return_layers=False):
    """models/TorsionalDiffusion.py:90-109.  `cache=None` recomputes the graph like the reference does."""
    if cache is None:
        cache = GraphCache(sd, batch)
    sc = torch.stack((torch.sin(SC_D_noised), torch.cos(SC_D_noised)), -1) * batch["SC_D_mask"][..., None]
    V = node_features(batch["residue_type"], batch["BB_D_sincos"], sc, t)
    h_V = _ln(sd, "encoder.norm_nodes", _lin(sd, "encoder.node_embedding", V))
    h_E = cache.h_E0
    layers = [h_V]
    for l in range(3):
        # the layer-3 edge update is computed and dropped by the reference (mpnn.py:53-62); skipping it
        # changes nothing that is returned
        h_V, h_E = ipmp_layer(sd, l, h_V, h_E, cache.E_idx, cache.R, cache.t, batch["residue_mask"],
                              cache.mask_att, edge_update=(l < 2))
        layers.append(h_V)
    score = decoder(sd, h_V)
    if return_layers:
        return score, h_V, layers
    return score, h_V


# ---------------------------------------------------------------------------------------------- schedule
SIGMA_MIN, SIGMA_MAX = 0.01 * math.pi, math.pi


def t_to_sigma(t):
    """schedule.py:165-174 with float64 scalars folded as numpy does."""
    lo, hi = math.log(SIGMA_MIN), math.log(SIGMA_MAX)
    return torch.exp(lo + (hi - lo) * t)


def ode_step(x, score, time, dt, mask, annealed_temp=3):
    """schedule.py:198-235, ode branch.  x, score [..,4]; time, dt 0-dim f32 tensors; mask bool."""
    sigma = t_to_sigma(time)
    g = sigma * math.sqrt(2 * math.log(SIGMA_MAX / SIGMA_MIN))
    alpha = 1 - (sigma / math.exp(math.log(SIGMA_MAX))) ** 2
    w = annealed_temp / (alpha + (1 - alpha) * annealed_temp)
    new = x + 0.5 * g ** 2 * dt * (score * w)
    return torch.where(mask, new, x)


def sde_step(x, score, time, dt, mask, noise, annealed_temp=3):
    """schedule.py:198-235, sde branch, with the torch.normal draw injected."""
    sigma = t_to_sigma(time)
    g = sigma * math.sqrt(2 * math.log(SIGMA_MAX / SIGMA_MIN))
    alpha = 1 - (sigma / math.exp(math.log(SIGMA_MAX))) ** 2
    w = annealed_temp / (alpha + (1 - alpha) * annealed_temp)
    new = x + (g ** 2 * dt * (score * w) + g * torch.sqrt(dt) * noise)
    return torch.where(mask, new, x)


def wrap(x):
    return (x + math.pi) % TWO_PI - math.pi


def initial_noise(batch, eps1, eps2):
    """add_sc_noise at t = 1 (TorsionalDiffusion.py:111-124, schedule.py:176-196) with the two
    randn draws injected: x = SC_D + eps1*sigma(1)*mask_1pi + eps2*sigma(1)*mask_2pi, wrapped."""
    B, L = batch["SC_D"].shape[:2]
    sig = t_to_sigma(torch.ones(B * L))[:, None]
    x = batch["SC_D"].reshape(-1, 4)
    x = x + (eps1.reshape(-1, 4) * sig) * batch["chi_1pi_periodic_mask"].reshape(-1, 4)
    x = x + (eps2.reshape(-1, 4) * sig) * batch["chi_2pi_periodic_mask"].reshape(-1, 4)
    return wrap(x).reshape(B, L, 4)


def sampling(sd, batch, SC_D_init, n_steps=30, hoist=True, trajectory=False, sde_noise=None):
    """TorsionalDiffusion.py:254-283 from an injected initial sample (ODE mode has no other randomness).

    hoist=False rebuilds graph and edge embedding every step, as the reference does (CPU-baseline timing).
    """
    sched = torch.linspace(1, 0, n_steps + 1)
    B, L = batch["SC_D"].shape[:2]
    cache = GraphCache(sd, batch) if hoist else None
    x = SC_D_init.clone()
    traj = []
    m1 = batch["chi_1pi_periodic_mask"].reshape(-1, 4)
    m2 = batch["chi_2pi_periodic_mask"].reshape(-1, 4)
    with torch.no_grad():
        for j in range(n_steps):
            time, dt = sched[j], sched[j] - sched[j + 1]
            t = time.repeat_interleave(B * L)
            score, _ = network(sd, batch, x, t, cache)
            s = score.reshape(-1, 4)
            if sde_noise is None:
                y = ode_step(x.reshape(-1, 4), s, time, dt, m1)
                y = ode_step(y, s, time, dt, m2)
            else:  # mode "sde": one normal draw per schedule.step call, [n_steps, 2, B*L, 4]
                y = sde_step(x.reshape(-1, 4), s, time, dt, m1, sde_noise[j, 0])
                y = sde_step(y, s, time, dt, m2, sde_noise[j, 1])
            x = wrap(y).reshape(B, L, 4) * batch["SC_D_mask"]
            if trajectory:
                traj.append(x.clone())
    return (x, traj) if trajectory else x
