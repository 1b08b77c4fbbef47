"""CPU oracle for the VaDE / recurrent (GRU + CensNet) hot path of mlfpm/deepof.

TEST INFRASTRUCTURE — NOT THE PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / the timed CPU baseline.  The product path
(``deepof_b200``) never routes through it.

What it is: a *restatement* (own code, functional style, plain torch CPU tensor ops,
fp32 or fp64) of the arithmetic the reference performs for one VaDE training step
and for eval-mode embedding, each function citing the reference ``file:line`` it
follows (paths relative to /root/reference).  Gradients come from torch autograd on
this restatement (the reference does the same with its own graph).

Parity pinning: ``tests/golden/*.npz`` hold outputs of the UNMODIFIED reference run
in the build container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py``
checks this oracle against every one of them (eval outputs, all loss terms, the
flat gradient, post-Adam parameters).  Parity is therefore pinned.

Parameter naming: the dict keys are exactly the reference ``state_dict()`` keys
(SURVEY.md Appendix A.6), e.g. ``encoder.node_recurrent_block.gru1.weight_ih_l0``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor
LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# configuration (mirrors the reference dataclasses, model_utils_new.py:37-169)
# ----------------------------------------------------------------------------
@dataclass
class LossCfg:
    """Active VadeLoss hyper-parameters for ONE phase (losses.py:383-457)."""

    pretrain_mode: bool = False
    kl_weight: float = 1.0            # Dynamic_weight_manager.get_weight() at this step
    l1_activity_weight: float = 0.1   # losses.py:389
    kmeans_loss_weight: float = 0.0   # mode_params[...]["kmeans_loss"]
    model_kmeans_weight: float = 1.0  # GaussianMixtureLatentPT.kmeans_weight (training.py:1556)
    repel_weight: float = 0.0
    repel_length_scale: float = 1.0
    nonempty_weight: float = 2e-2
    nonempty_floor: float = 0.05 / 8  # max(1e-4, floor_percent / K)
    nonempty_p: int = 2
    tf_cluster_weight: float = 0.0
    reg_cat_clusters_weight: float = 0.0
    temporal_cohesion_weight: float = 0.0
    reg_scatter_weight: float = 0.0
    reg_scatter_beta: float = 1.0
    gmm_logvar_clamp: Tuple[float, float] = (-8.0, 8.0)
    mc_samples: int = 32              # losses.py:526
    # distillation (losses.py:730-760); tau_star rows are gathered by batch idx
    lambda_distill: float = 0.0
    distill_sharpen_T: float = 0.5
    distill_conf_weight: bool = False
    distill_conf_thresh: float = 0.3

    @staticmethod
    def pretrain_defaults(n_components: int, kl_weight: float = 0.0) -> "LossCfg":
        # VaDECfg defaults, model_utils_new.py:152-157
        return LossCfg(pretrain_mode=True, kl_weight=kl_weight, kmeans_loss_weight=1.0,
                       repel_weight=0.5, repel_length_scale=0.5, nonempty_weight=2e-2,
                       nonempty_floor=max(1e-4, 0.05 / n_components), nonempty_p=2)

    @staticmethod
    def main_defaults(n_components: int, kl_weight: float = 1.0) -> "LossCfg":
        # VaDECfg / CommonFitCfg defaults, model_utils_new.py:66,135-150
        return LossCfg(pretrain_mode=False, kl_weight=kl_weight, kmeans_loss_weight=0.0,
                       repel_weight=0.0, repel_length_scale=1.0, nonempty_weight=2e-2,
                       nonempty_floor=max(1e-4, 0.05 / n_components), nonempty_p=2)


# ----------------------------------------------------------------------------
# graph operators  (censNetConv_pt.py:160-370)
# ----------------------------------------------------------------------------
def gcn_filter(A: np.ndarray) -> np.ndarray:
    """D^-1/2 (A + I) D^-1/2 with zero-degree -> 1  (censNetConv_pt.py:182-243)."""
    A_hat = A.astype(np.float64) + np.eye(A.shape[0])
    deg = A_hat.sum(axis=1)
    deg[deg == 0] = 1.0
    d = deg ** -0.5
    return (d[:, None] * A_hat) * d[None, :]


def incidence_matrix(A: np.ndarray) -> np.ndarray:
    """[N, E] incidence, edges = nonzeros of triu(A) in row-major order
    (censNetConv_pt.py:296-370)."""
    tri = np.triu(A)
    rows, cols = np.nonzero(tri)  # row-major order, same as torch.nonzero
    inc = np.zeros((A.shape[0], len(rows)), dtype=np.float64)
    for e, (i, j) in enumerate(zip(rows, cols)):
        inc[i, e] = 1.0
        inc[j, e] = 1.0
    return inc


def graph_operators(adjacency: np.ndarray):
    """(laplacian[N,N], edge_laplacian[E,E], incidence[N,E]) as float32 tensors
    (censNetConv_pt.py:160-175; buffers at models_new.py:102-105).

    The reference feeds a float64 adjacency (np default) through float32 eye/ops
    via type promotion and finally ``.float()``; we compute in float64 and round
    once, which tests confirm is bit-identical for 0/1 adjacencies.
    """
    A = np.asarray(adjacency)
    lap = gcn_filter(A)
    inc = incidence_matrix(A)
    line = inc.T @ inc - 2.0 * np.eye(inc.shape[1])
    elap = gcn_filter(line)
    f = lambda m: torch.from_numpy(np.ascontiguousarray(m)).float()
    return f(lap), f(elap), f(inc)


# ----------------------------------------------------------------------------
# A.1 group reshape (models_new.py:120-138)
# ----------------------------------------------------------------------------
def group_gather_index(T: int, G: int, F: int) -> Tensor:
    """Index table idx[g, t', f] into one window flattened as [T*G*F] (row-major
    [T, G, F]) such that out[b, g, t', f] = window[b].flatten()[idx[g, t', f]].

    The reference does reshape(B,T,G*F) -> permute(2,1,0) -> reshape(F,T,G,B) ->
    permute(3,2,1,0).  Element law: lin = (f*T + t')*G + g; j = lin // T; t = lin % T;
    source = flat[b, t, j].  It is a fixed permutation of the window's T*G*F
    elements (NOT a transpose).
    """
    g = torch.arange(G).view(G, 1, 1)
    t2 = torch.arange(T).view(1, T, 1)
    f = torch.arange(F).view(1, 1, F)
    lin = (f * T + t2) * G + g
    j = lin // T
    t = lin % T
    return (t * (G * F) + j).long()  # [G, T, F]


def group_reshape(x: Tensor) -> Tensor:
    """x [B,T,G,F] -> [B,G,T,F] with the reference's element scramble."""
    B, T, G, F = x.shape
    idx = group_gather_index(T, G, F).reshape(-1)
    return x.reshape(B, T * G * F)[:, idx].reshape(B, G, T, F)


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
def conv1d_same_k5(x: Tensor, w: Tensor) -> Tensor:
    """Cross-correlation along time, zero 'same' padding, no bias.
    x [S,T,Cin], w [Cout,Cin,5] -> [S,T,Cout]   (models_new.py:192-198,228-230)."""
    S, T, Cin = x.shape
    xp = torch.nn.functional.pad(x, (0, 0, 2, 2))
    # unfold time: [S, T, 5, Cin]
    cols = torch.stack([xp[:, k:k + T, :] for k in range(5)], dim=2)
    return torch.einsum("stkc,ock->sto", cols, w)


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim, biased variance (models_new.py:206,214,298)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gru_direction(x: Tensor, lengths: Tensor, w_ih: Tensor, w_hh: Tensor,
                  b_ih: Tensor, b_hh: Tensor, reverse: bool) -> Tuple[Tensor, Tensor]:
    """One direction of torch.nn.GRU over packed (prefix-valid) sequences
    (models_new.py:243-249).

    x [S,T,I]; lengths [S] (valid prefix); returns (out [S,T,H] with zeros at padded
    steps, h_final [S,H]).  Gate order r,z,n.
    """
    S, T, _ = x.shape
    H = w_hh.shape[1]
    gi_all = x @ w_ih.t() + b_ih           # [S,T,3H]
    h = x.new_zeros(S, H)
    outs = [None] * T
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gi = gi_all[:, t]
        gh = h @ w_hh.t() + b_hh
        r = torch.sigmoid(gi[:, :H] + gh[:, :H])
        z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        h_new = (1.0 - z) * n + z * h
        valid = (t < lengths).to(x.dtype).unsqueeze(1)
        h = valid * h_new + (1.0 - valid) * h
        outs[t] = valid * h
    return torch.stack(outs, dim=1), h


# The reference runs its GRUs through torch.nn.GRU (ATen's fused CPU kernel).  For the timed
# CPU baseline (bench.py) the oracle can do the same: with USE_ATEN_GRU the dense case
# (all lengths == T, which is what pack_padded_sequence degenerates to) calls the very same
# ATen op; the explicit per-step restatement above stays the default and the checker.
USE_ATEN_GRU = False


def bigru_aten(x: Tensor, p: Dict[str, Tensor], prefix: str):
    names = ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")
    flat = [p[prefix + n] for n in names] + [p[prefix + n + "_reverse"] for n in names]
    H = flat[1].shape[1]
    h0 = x.new_zeros(2, x.shape[0], H)
    out, hn = torch._VF.gru(x, h0, flat, True, 1, 0.0, True, True, True)
    return out, hn.permute(1, 0, 2).reshape(x.shape[0], 2 * H)


def bigru(x: Tensor, lengths: Tensor, p: Dict[str, Tensor], prefix: str):
    """Bidirectional single-layer GRU; returns (out [S,T,2H], h_n [S,2H]=[fwd|bwd])."""
    if USE_ATEN_GRU and bool((lengths == x.shape[1]).all()):
        return bigru_aten(x.contiguous(), p, prefix)
    of, hf = gru_direction(x, lengths, p[prefix + "weight_ih_l0"], p[prefix + "weight_hh_l0"],
                           p[prefix + "bias_ih_l0"], p[prefix + "bias_hh_l0"], False)
    ob, hb = gru_direction(x, lengths, p[prefix + "weight_ih_l0_reverse"],
                           p[prefix + "weight_hh_l0_reverse"], p[prefix + "bias_ih_l0_reverse"],
                           p[prefix + "bias_hh_l0_reverse"], True)
    return torch.cat([of, ob], dim=-1), torch.cat([hf, hb], dim=-1)


def recurrent_block(seq: Tensor, p: Dict[str, Tensor], prefix: str, latent_dim: int) -> Tensor:
    """RecurrentBlockPT.forward (models_new.py:217-278) on sequences seq [S,T,F].
    Returns [S, 2*latent_dim]."""
    conv = torch.relu(conv1d_same_k5(seq, p[prefix + "conv1d.weight"]))       # :228-230
    mask = conv.abs().sum(dim=-1) > 0                                          # :233
    lengths = mask.sum(dim=1)                                                  # :234 (prefix semantics)
    g1, _ = bigru(conv, lengths, p, prefix + "gru1.")                          # :243-249
    n1 = layer_norm(g1, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], 1e-3)  # :252-253 (padded rows too)
    _, hn = bigru(n1, lengths, p, prefix + "gru2.")                            # :261-268
    out = layer_norm(hn, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], 1e-3)  # :271
    d_int = p[prefix + "gru2.weight_hh_l0"].shape[1]
    if d_int != latent_dim:                                                    # :274-275
        out = out @ p[prefix + "projection.weight"].t() + p[prefix + "projection.bias"]
    return out


def censnet(node: Tensor, edge: Tensor, lap: Tensor, elap: Tensor, inc: Tensor,
            p: Dict[str, Tensor], prefix: str) -> Tuple[Tensor, Tensor]:
    """CensNetConvPT.forward with activation='relu' (censNetConv_pt.py:92-158).
    node [B,N,Fn], edge [B,E,Fe] -> ([B,N,C], [B,E,C])."""
    we = (edge @ p[prefix + "edge_weights"]).squeeze(-1)                 # [B,E]
    Mv = torch.einsum("ie,be,je->bij", inc, we, inc) * lap              # :103-106
    on = torch.relu((Mv @ node) @ p[prefix + "node_kernel"] + p[prefix + "node_bias"])
    wn = (node @ p[prefix + "node_weights"]).squeeze(-1)                 # [B,N]
    Me = torch.einsum("ne,bn,nf->bef", inc, wn, inc) * elap             # :126-129
    oe = torch.relu((Me @ edge) @ p[prefix + "edge_kernel"] + p[prefix + "edge_bias"])
    return on, oe


def encoder_forward(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, latent_dim: int) -> Tensor:
    """RecurrentEncoderPT.forward, use_gnn=True (models_new.py:140-181)."""
    B, T, N, F = x.shape
    E = a.shape[2]
    lap, elap, inc = graph
    xs = group_reshape(x).reshape(B * N, T, F)
    es = group_reshape(a).reshape(B * E, T, a.shape[3])
    node = recurrent_block(xs, p, "encoder.node_recurrent_block.", latent_dim).reshape(B, N, -1)
    edge = recurrent_block(es, p, "encoder.edge_recurrent_block.", latent_dim).reshape(B, E, -1)
    on, oe = censnet(node, edge, lap.to(x.dtype), elap.to(x.dtype), inc.to(x.dtype), p,
                     "encoder.spatial_gnn_block.")
    # :165 compares a function with the string "relu" -> no second ReLU
    flat = torch.cat([on.reshape(B, -1), oe.reshape(B, -1)], dim=-1)
    return flat @ p["encoder.final_dense.weight"].t() + p["encoder.final_dense.bias"]


def latent_forward(enc: Tensor, p: Dict[str, Tensor], training: bool, eps: Optional[Tensor]):
    """GaussianMixtureLatentPT.forward (models_new.py:1761-1791) minus the kmeans term.
    Returns z, q, z_mean, z_log_var."""
    z_mean = enc @ p["latent_space.encoder_mean.weight"].t() + p["latent_space.encoder_mean.bias"]
    pre = enc @ p["latent_space.encoder_log_var.weight"].t() + p["latent_space.encoder_log_var.bias"]
    z_log_var = torch.nn.functional.softplus(pre)                         # :1766
    if training:
        if eps is None:
            eps = torch.randn_like(z_mean)
        z = z_mean + torch.exp(0.5 * z_log_var) * eps                     # :1739-1742
    else:
        z = z_mean                                                        # :1770
    q = gmm_posterior(z, p)
    return z, q, z_mean, z_log_var


def gmm_posterior(z: Tensor, p: Dict[str, Tensor]) -> Tensor:
    """_calculate_posterior (models_new.py:1745-1759)."""
    std = torch.exp(0.5 * p["latent_space.gmm_log_vars"]).clamp(min=1e-3)
    mu = p["latent_space.gmm_means"]
    d = (z.unsqueeze(1) - mu.unsqueeze(0)) / std.unsqueeze(0)
    logp = (-0.5 * d * d - torch.log(std).unsqueeze(0) - 0.5 * LOG_2PI).sum(dim=-1)
    logit = torch.log(p["latent_space.prior"].to(z.dtype) + 1e-9) + logp
    return torch.softmax(logit, dim=-1)


def kmeans_loss(z: Tensor, weight: float) -> Tensor:
    """compute_kmeans_loss_pt (losses.py:257-287): mean sqrt of the singular values of
    the Gram matrix z^T z / B, in fp64."""
    B = float(z.shape[0])
    gram = (z.t() @ z) / B
    sv = torch.linalg.svdvals(gram.to(torch.float64))
    pen = torch.sqrt(torch.clamp(sv, min=1e-9))
    return weight * pen.mean()


def decoder_forward(z: Tensor, x_flat: Tensor, p: Dict[str, Tensor]):
    """RecurrentDecoderPT.forward (models_new.py:326-373).  Returns (loc [B,T,3N],
    validity mask [B,T] as float)."""
    B, T, _ = x_flat.shape
    mask = ~torch.all(x_flat == 0.0, dim=2)                               # :330
    lengths = mask.sum(dim=1)                                             # :331
    gen = z.unsqueeze(1).expand(-1, T, -1)                                # :341
    g1, _ = bigru(gen, lengths, p, "decoder.gru1.")
    n1 = layer_norm(g1, p["decoder.norm1.weight"], p["decoder.norm1.bias"], 1e-3)
    g2, _ = bigru(n1, lengths, p, "decoder.gru2.")
    n2 = layer_norm(g2, p["decoder.norm2.weight"], p["decoder.norm2.bias"], 1e-3)
    conv = torch.relu(conv1d_same_k5(n2, p["decoder.conv1d.weight"]))     # :366-368
    n3 = layer_norm(conv, p["decoder.norm3.weight"], p["decoder.norm3.bias"], 1e-3)
    loc = n3 @ p["decoder.prob_decoder.loc_projection.weight"].t() \
        + p["decoder.prob_decoder.loc_projection.bias"]                   # :690-691
    loc = torch.nan_to_num(loc, nan=0.0, posinf=1e6, neginf=-1e6)         # :694
    return loc, mask.to(z.dtype)


def recon_log_prob(loc: Tensor, mask: Tensor, x_flat: Tensor) -> Tensor:
    """log_prob of AffineTransformedDistribution(Independent(Normal(loc,1),1),
    scale=mask[...,None]) at x  (models_new.py:696-708; torch TransformedDistribution)."""
    scale = mask.unsqueeze(-1)
    y = x_flat / scale
    base = (-0.5 * (y - loc) ** 2 - 0.5 * LOG_2PI).sum(dim=-1)
    ladj = torch.log(torch.abs(scale)).expand_as(x_flat).sum(dim=-1)
    return base - ladj                                                    # [B,T]


# ----------------------------------------------------------------------------
# VadeLoss (losses.py:567-797)
# ----------------------------------------------------------------------------
def _log_normal_diag(x, mean, log_var):
    return -0.5 * torch.sum(LOG_2PI + log_var + (x - mean) ** 2 * torch.exp(-log_var), dim=-1)


def monte_carlo_kl(z_mean, z_log_var, gmm_means, gmm_log_vars, prior, mc_eps, clamp):
    """_monte_carlo_kl + _log_mog (losses.py:506-545)."""
    z_log_var = torch.clamp(z_log_var, min=-4.0, max=4.0)
    scale_q = torch.exp(0.5 * z_log_var)
    zs = z_mean.unsqueeze(0) + mc_eps * scale_q.unsqueeze(0)              # [S,B,D]
    log_q = _log_normal_diag(zs, z_mean.unsqueeze(0), z_log_var.unsqueeze(0))
    glv = torch.clamp(gmm_log_vars, min=clamp[0], max=clamp[1])
    log_prior = torch.log(torch.clamp(prior, min=1e-8))
    C, D = gmm_means.shape
    lp = _log_normal_diag(zs.unsqueeze(2), gmm_means.view(1, 1, C, D), glv.view(1, 1, C, D))
    log_p = torch.logsumexp(log_prior.view(1, 1, C) + lp, dim=-1)
    return torch.clamp((log_q - log_p).mean(), min=0.0)


def vade_loss(loc, mask, z, q_in, kmeans, z_mean, z_log_var, p, x_flat, cfg: LossCfg,
              mc_eps: Optional[Tensor] = None, tau_batch: Optional[Tensor] = None,
              class_weight: Optional[Tensor] = None,
              teacher_marginal: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """VadeLoss.forward (losses.py:567-797).  ``tau_batch`` = tau_star[batch_indices]."""
    dt = z_mean.dtype
    zero = torch.zeros((), dtype=dt)
    recon = -(recon_log_prob(loc, mask, x_flat)).mean()                   # :586
    q = q_in.clamp_min(1e-8)
    q = q / q.sum(dim=-1, keepdim=True)                                   # :589-592
    activity = cfg.l1_activity_weight * torch.sum(torch.abs(z_log_var), dim=-1).mean()  # :595
    klw = float(cfg.kl_weight)
    lv = z_log_var.clamp(min=-4.0, max=2.0)                               # :604
    gmm_means = p["latent_space.gmm_means"]
    gmm_log_vars = p["latent_space.gmm_log_vars"]
    prior = p["latent_space.prior"].to(dt)
    if cfg.pretrain_mode:                                                 # :606-613
        kl_vec = 0.5 * (z_mean.pow(2) + lv.exp() - 1.0 - lv).sum(dim=-1) / lv.shape[-1]
        kl = klw * kl_vec.mean()
    else:                                                                 # :614-624
        if mc_eps is None:
            mc_eps = torch.randn(cfg.mc_samples, *z_mean.shape, dtype=dt)
        kl = klw * monte_carlo_kl(z_mean, lv, gmm_means, gmm_log_vars, prior, mc_eps,
                                  cfg.gmm_logvar_clamp)
    tf_cluster = prior_loss = cat_loss = scatter = repel = distill = temporal = nonempty = zero
    km = (cfg.kmeans_loss_weight * kmeans).to(dt)                         # :640-643
    if cfg.repel_weight > 0.0:                                            # :647-664
        qf = q.detach()
        pi_b = qf.sum(dim=0).clamp_min(1e-8)
        means = (qf.t() @ z) / pi_b.unsqueeze(1)
        C = means.size(0)
        diffs = means.unsqueeze(1) - means.unsqueeze(0)
        D2 = (diffs * diffs).sum(dim=-1)
        Kmat = torch.exp(-D2 / max(1e-9, 2.0 * (cfg.repel_length_scale ** 2)))
        Kmat = Kmat - torch.diag(torch.diag(Kmat))
        repel = cfg.repel_weight * (Kmat.sum() / float(max(1, C * C - C)))
    if cfg.nonempty_weight > 0.0:                                         # :668-684
        q_marg = q.mean(dim=0)
        if teacher_marginal is not None:
            floor_c = torch.max(cfg.nonempty_floor * torch.ones_like(teacher_marginal),
                                0.9 * teacher_marginal)
        else:
            floor_c = cfg.nonempty_floor * torch.ones_like(q_marg)
        nonempty = cfg.nonempty_weight * (floor_c - q_marg).clamp_min(0.0).pow(cfg.nonempty_p).sum()
    if not cfg.pretrain_mode:                                             # :689-726
        glv = torch.clamp(gmm_log_vars, min=cfg.gmm_logvar_clamp[0], max=cfg.gmm_logvar_clamp[1])
        scale = torch.exp(0.5 * glv).clamp(min=1e-3)
        d = (z.unsqueeze(1) - gmm_means.unsqueeze(0)) / scale.unsqueeze(0)
        logp = (-0.5 * d * d - torch.log(scale).unsqueeze(0) - 0.5 * LOG_2PI).sum(dim=-1)
        post_like = torch.softmax(logp, dim=-1)
        tf_cluster = -(q * post_like).sum(dim=-1).mean() * cfg.tf_cluster_weight
        C = gmm_means.shape[0]
        prior_loss = -(q * math.log(1.0 / max(1, C))).sum(dim=-1).mean()   # :700-702
        if cfg.reg_cat_clusters_weight > 0:                                # :705-706, :354-359
            mean_freq = q.mean(dim=0)
            uni = torch.ones(C, dtype=dt) / C
            kld = (uni * (torch.log(uni) - torch.log(mean_freq + 1e-9))).sum() / C  # batchmean over dim0 = C
            cat_loss = cfg.reg_cat_clusters_weight * kld
        if cfg.temporal_cohesion_weight > 0.0 and q.size(0) > 1:           # :709-712
            temporal = cfg.temporal_cohesion_weight * (q[1:] - q[:-1]).abs().sum(dim=-1).mean()
        if cfg.reg_scatter_weight > 0.0:                                   # :714-724
            pi_b = q.sum(dim=0).clamp_min(1e-8)
            mu = (q.t() @ z_mean) / pi_b.unsqueeze(1)
            diff = z_mean.unsqueeze(1) - mu.unsqueeze(0)
            scat_c = (q.unsqueeze(-1) * diff.pow(2)).sum(dim=0) / pi_b.unsqueeze(1)
            w = ((pi_b / pi_b.mean()).pow(-cfg.reg_scatter_beta)).unsqueeze(1)
            scatter = cfg.reg_scatter_weight * (w * scat_c).mean()
    if cfg.lambda_distill > 0.0 and tau_batch is not None:                 # :730-760
        tb = tau_batch
        if cfg.distill_sharpen_T is not None and cfg.distill_sharpen_T > 0.0:
            tb = torch.softmax(tb.clamp_min(1e-8).log() / float(cfg.distill_sharpen_T), dim=-1)
        ce = -(tb * q.clamp_min(1e-8).log()).sum(dim=-1)
        w_conf = None
        if cfg.distill_conf_weight:
            conf = tb.max(dim=1).values
            thr = float(cfg.distill_conf_thresh)
            w_conf = ((conf - thr) / max(1e-6, 1.0 - thr)).clamp(0.0, 1.0).detach()
        if class_weight is not None:
            w_class = tb @ class_weight.to(dt)
            w_class = (w_class / w_class.mean().clamp_min(1e-8)).detach()
            w_total = w_class if w_conf is None else w_class * w_conf
        else:
            w_total = w_conf
        distill = (w_total * ce).mean() if w_total is not None else ce.mean()
        distill = cfg.lambda_distill * distill
    total = (recon + kl + cat_loss + temporal + nonempty + tf_cluster + prior_loss + km
             + activity + scatter + repel + distill)                       # :766-779
    return {
        "total_loss": total, "reconstruct_loss": recon, "kl_div": kl,
        "kl_weight": torch.tensor(klw, dtype=dt), "tf_clust_loss": tf_cluster,
        "prior_loss": prior_loss, "kmeans_loss": km, "activity_l1": activity,
        "cat_clust_loss": cat_loss, "distill_loss": distill, "nonempty_loss": nonempty,
        "temporal_loss": temporal, "scatter_loss": scatter, "repel_loss": repel,
    }


LOG_KEYS = ("total_loss", "reconstruct_loss", "kl_div", "cat_clust_loss", "kmeans_loss",
            "activity_l1", "prior_loss", "distill_loss", "tf_clust_loss", "nonempty_loss",
            "temporal_loss", "scatter_loss", "repel_loss")  # step_vade, training.py:292-306


# ----------------------------------------------------------------------------
# full model forward / training step
# ----------------------------------------------------------------------------
def vade_forward(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, latent_dim: int,
                 training: bool, eps: Optional[Tensor] = None, model_kmeans_weight: float = 1.0):
    """VaDEPT.forward(return_gmm_params=True) (models_new.py:1841-1891).
    Returns dict(enc, z, q, z_mean, z_log_var, kmeans, loc, mask)."""
    B, T, N, F = x.shape
    enc = encoder_forward(x, a, p, graph, latent_dim)
    z, q, z_mean, z_log_var = latent_forward(enc, p, training, eps)
    if model_kmeans_weight > 0:                                           # :1787-1789
        km = kmeans_loss(z, model_kmeans_weight)
    else:
        km = torch.zeros((), dtype=x.dtype)
    loc, mask = decoder_forward(z, x.reshape(B, T, N * F), p)
    return dict(enc=enc, z=z, q=q, z_mean=z_mean, z_log_var=z_log_var, kmeans=km, loc=loc, mask=mask)


def embed(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, latent_dim: int):
    """Eval-mode judged outputs: (embedding = z_mean, q) — what embedding_per_video
    reads as model(x,a)[1], [2] (model_utils_new.py:610-617)."""
    enc = encoder_forward(x, a, p, graph, latent_dim)
    z, q, _, _ = latent_forward(enc, p, training=False, eps=None)
    return z, q


# parameters that never receive a gradient in the reference (grad is None, Adam skips
# them): latent_space.lens.* always; *.projection.* when internal_dim == latent_dim
def dead_parameter(name: str, p: Dict[str, Tensor], latent_dim: int) -> bool:
    if name.startswith("latent_space.lens."):
        return True
    if ".projection." in name:
        blk = name.split("projection.")[0]
        return p[blk + "gru2.weight_hh_l0"].shape[1] == latent_dim
    return False


BUFFER_NAMES = ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence",
                "latent_space.prior", "latent_space.pretrain")


def trainable_names(p: Dict[str, Tensor]):
    return [k for k in p.keys() if k not in BUFFER_NAMES]


def train_step(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, latent_dim: int, cfg: LossCfg,
               eps: Optional[Tensor] = None, mc_eps: Optional[Tensor] = None,
               tau_batch: Optional[Tensor] = None, class_weight: Optional[Tensor] = None,
               teacher_marginal: Optional[Tensor] = None):
    """One step_vade forward+backward (training.py:231-309,159-166).
    Returns (logs dict of python floats, grads dict name->Tensor|None, outputs)."""
    names = trainable_names(p)
    leaf = {}
    for k, v in p.items():
        if k in names and v.dtype.is_floating_point:
            leaf[k] = v.detach().clone().requires_grad_(True)
        else:
            leaf[k] = v
    B, T, N, F = x.shape
    out = vade_forward(x, a, leaf, graph, latent_dim, training=True, eps=eps,
                       model_kmeans_weight=cfg.model_kmeans_weight)
    losses = vade_loss(out["loc"], out["mask"], out["z"], out["q"], out["kmeans"], out["z_mean"],
                       out["z_log_var"], leaf, x.reshape(B, T, N * F), cfg, mc_eps=mc_eps,
                       tau_batch=tau_batch, class_weight=class_weight,
                       teacher_marginal=teacher_marginal)
    total = losses["total_loss"]
    plist = [leaf[k] for k in names]
    glist = torch.autograd.grad(total, plist, allow_unused=True)
    grads = {k: g for k, g in zip(names, glist)}
    logs = {k: float(losses[k].detach()) for k in LOG_KEYS}
    return logs, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


def adam_step(p: Dict[str, Tensor], grads: Dict[str, Optional[Tensor]], state: Dict[str, dict],
              lr_base: float, lr_gmm: float, clip: Optional[float] = 0.75,
              betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """clip_grad_value_(0.75) + torch.optim.Adam with the two VaDE param groups
    (training.py:164-166; losses.py:817-833).  In place on ``p`` / ``state``.
    Parameters whose grad is None are skipped, exactly like torch."""
    for k, g in grads.items():
        if g is None:
            continue
        g = g.to(p[k].dtype)
        if clip is not None:
            g = g.clamp(-clip, clip)
        st = state.setdefault(k, {"step": 0, "m": torch.zeros_like(p[k]), "v": torch.zeros_like(p[k])})
        st["step"] += 1
        lr = lr_gmm if k in ("latent_space.gmm_means", "latent_space.gmm_log_vars") else lr_base
        st["m"].mul_(betas[0]).add_(g, alpha=1 - betas[0])
        st["v"].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1 = 1 - betas[0] ** st["step"]
        bc2 = 1 - betas[1] ** st["step"]
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        p[k] = p[k] - (lr / bc1) * st["m"] / denom


def kl_weight_schedule(it: int, n_batches_per_epoch: int, mode: str, warmup_epochs: int,
                       max_weight: float, cooldown_epochs: int, end_weight: float,
                       at_max_epochs: int = 0) -> float:
    """Dynamic_weight_manager.get_weight at iteration ``it`` (losses.py:290-351)."""
    warm = max(1, warmup_epochs * n_batches_per_epoch)
    atmax = max(0, at_max_epochs * n_batches_per_epoch)
    cool = max(0, cooldown_epochs * n_batches_per_epoch)
    total = warm + atmax + cool

    def shape(pv):
        pv = float(max(0.0, min(1.0, pv)))
        if mode == "linear":
            return pv
        if mode == "sigmoid":
            return 1.0 / (1.0 + math.exp(-12.0 * (pv - 0.5)))
        if mode == "tf_sigmoid":
            denom = max(1e-2, pv - pv * pv)
            return 1.0 / (1.0 + math.exp(-((2.0 * pv - 1.0) / denom)))
        return pv

    if it >= total:
        return float(end_weight)
    if atmax > 0 and warm <= it < warm + atmax:
        return float(max_weight)
    if it <= warm:
        return float(max_weight) * shape(it / warm)
    if cool <= 0:
        return float(max_weight)
    pc = (it - (warm + atmax)) / cool
    return (1.0 - pc) * float(max_weight) + pc * float(end_weight)


# ----------------------------------------------------------------------------
# synthetic inputs: standardised xy/speed ~ N(0,1); edges recomputed
# from x as standardised log1p distances so the graph is self-consistent.
# ----------------------------------------------------------------------------
def default_adjacency(n_nodes: int) -> np.ndarray:
    """A connected mouse-like skeleton with E == N edges for N=14 (deepof_14-like):
    a spine chain plus one chord; deterministic."""
    A = np.zeros((n_nodes, n_nodes), dtype=np.float64)
    for i in range(n_nodes - 1):
        A[i, i + 1] = A[i + 1, i] = 1.0
    if n_nodes > 5:
        A[0, 5] = A[5, 0] = 1.0
    return A


def synthetic_windows(n_windows: int, T: int, adjacency: np.ndarray, seed: int,
                      dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    g = torch.Generator().manual_seed(seed)
    N = adjacency.shape[0]
    x = torch.randn(n_windows, T, N, 3, generator=g, dtype=torch.float32)
    rows, cols = np.nonzero(np.triu(adjacency))
    d = (x[:, :, rows, :2] - x[:, :, cols, :2]).norm(dim=-1)
    a = torch.log1p(d)
    a = (a - a.mean()) / a.std()
    return x.to(dtype), a.unsqueeze(-1).to(dtype)
