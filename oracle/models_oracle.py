"""CPU oracle for the VQ-VAE and contrastive paths of mlfpm/deepof (recurrent encoder / decoder).

TEST INFRASTRUCTURE — NOT THE PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` may import this module.

Plain torch CPU ops + autograd, restating (paths relative to /root/reference/deepof/clustering):

  VectorQuantizerPT                models_new.py:1330-1423
  VQVAEPT.forward                  models_new.py:1575-1635
  step_vqvae_distill               training.py:312-389      (teacher / distillation head off)
  ContrastivePT.forward            models_new.py:2063-2069
  step_contrastive_distill         training.py:482-589      (teacher / distillation head off)
  _make_augmented_view, _augment_* training.py:2128-2402, build_rotation_precomp :2064-2125
  recompute_edges, slice_time_per_sample   model_utils_new.py:332-364, 751-763
  nce_loss_pt + cosine similarity  losses.py:59-63, 130-141
  build_optimizer_generic          losses.py:805-814         (Adam, weight_decay 1e-4)

The encoder / decoder are those of ``oracle/vade_oracle.py``.  Pinned by ``tests/golden/vqvae_*.npz`` and
``tests/golden/contrastive_*.npz``, produced by the unmodified reference (``tests/golden/make_golden_models.py``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import vade_oracle as V

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# VQ-VAE
# ----------------------------------------------------------------------------
def vq_distances(z: Tensor, codebook: Tensor) -> Tensor:
    """get_code_indices (models_new.py:1406-1413): ||z||^2 + ||e_k||^2 - 2 z.e_k, codebook [D,K]."""
    sim = z @ codebook
    return (z ** 2).sum(dim=1, keepdim=True) + (codebook ** 2).sum(dim=0) - 2 * sim


def vq_forward(z: Tensor, codebook: Tensor, beta: float, kmeans_w: float = 0.0):
    """VectorQuantizerPT.forward.  Returns quantized, soft_counts, idx, dict(vq_loss, kmeans_loss) (floats:
    step_vqvae_distill takes them with float(), i.e. gradient-free, training.py:334-336)."""
    d = vq_distances(z, codebook)
    idx = torch.argmin(d, dim=1)                                          # :1422
    sim = (1 / d) ** 2                                                    # :1416-1418
    soft = sim / sim.sum(dim=1, keepdim=True)
    onehot = torch.nn.functional.one_hot(idx, codebook.shape[1]).to(z.dtype)
    quant = onehot @ codebook.t()                                         # :1382-1383 (differentiable w.r.t. the codebook)
    commit = beta * ((quant.detach() - z) ** 2).mean()                    # :1387-1391
    cb = ((quant - z.detach()) ** 2).mean()
    losses = {"vq_loss": float((commit + cb).detach()), "kmeans_loss": 0.0}
    if kmeans_w:
        losses["kmeans_loss"] = float(V.kmeans_loss(z, kmeans_w).detach())         # :1370-1373
    return quant, soft, idx, losses


VQ_LOG_KEYS = ("total_loss", "enc_rec_loss", "reconstruct_loss", "vq_loss", "kmeans_loss",
               "number_of_populated_clusters", "distill_loss")


def vqvae_forward(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, latent_dim: int, beta: float = 1.0,
                  kmeans_w: float = 0.0):
    B, T, N, F = x.shape
    enc = V.encoder_forward(x, a, p, graph, latent_dim)
    quant, soft, idx, vql = vq_forward(enc, p["vq_layer.codebook"], beta, kmeans_w)
    xf = x.reshape(B, T, N * F)
    loc_q, mask = V.decoder_forward(quant, xf, p)                         # decode from the quantized latents
    loc_e, _ = V.decoder_forward(enc, xf, p)                              # bypass path from the encoder output
    return dict(enc=enc, quant=quant, soft=soft, idx=idx, loc_q=loc_q, loc_e=loc_e, mask=mask, vq=vql)


def distill_loss(z: Tensor, head_w: Tensor, head_b: Tensor, tau_b: Tensor, lam: float, sharpen_T: float = 0.5,
                 conf_weight: bool = False, conf_thresh: float = 0.6) -> Tensor:
    """The teacher term shared by step_vqvae_distill (training.py:346-370) and step_contrastive_distill (:555-578):
    DiscriminativeHead logits (teacher_model.py:795-808), sharpened targets, _soft_ce_logits (training.py:392-400)."""
    logits = z @ head_w.t() + head_b
    if sharpen_T > 0.0:
        tau_b = torch.softmax(tau_b.clamp_min(1e-8).log() / sharpen_T, dim=-1)
    logp = torch.log_softmax(logits, dim=-1)
    per = -(tau_b.clamp(min=1e-8, max=1.0) * logp).sum(dim=-1)
    if conf_weight:
        conf = tau_b.max(dim=1).values
        w = ((conf - conf_thresh) / max(1e-6, 1.0 - conf_thresh)).clamp(0.0, 1.0).detach()
        return lam * (w * per).mean()
    return lam * per.mean()


def _distill_leaves(distill):
    """distill = dict(head_w, head_b, tau, lam, sharpen_T, conf_weight, conf_thresh) -> leaf copies of the head."""
    w = distill["head_w"].detach().clone().requires_grad_(True)
    b = distill["head_b"].detach().clone().requires_grad_(True)
    return w, b


def vqvae_train_step(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, latent_dim: int, beta: float = 1.0,
                     kmeans_w: float = 0.0, distill: Optional[dict] = None):
    """step_vqvae_distill forward + backward (teacher off unless `distill` is given: then the gradients of the head
    come back under the keys "head/fc.weight", "head/fc.bias").  Returns (logs, grads, outputs)."""
    names = [k for k in p if k not in V.BUFFER_NAMES]
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in names and v.dtype.is_floating_point else v) for k, v in p.items()}
    B, T, N, F = x.shape
    out = vqvae_forward(x, a, leaf, graph, latent_dim, beta, kmeans_w)
    xf = x.reshape(B, T, N * F)
    enc_rec = -(V.recon_log_prob(out["loc_q"], out["mask"], xf)).mean()   # training.py:331
    rec = -(V.recon_log_prob(out["loc_e"], out["mask"], xf)).mean()       # :332
    total = enc_rec + rec + (out["vq"]["vq_loss"] + out["vq"]["kmeans_loss"])
    dl, extra = 0.0, []
    if distill is not None and distill["lam"] > 0.0:
        hw, hb = _distill_leaves(distill)
        dlt = distill_loss(out["enc"], hw, hb, distill["tau"], distill["lam"], distill.get("sharpen_T", 0.5),
                           distill.get("conf_weight", False), distill.get("conf_thresh", 0.6))
        total = total + dlt
        dl, extra = float(dlt.detach()), [hw, hb]
    glist = torch.autograd.grad(total, [leaf[k] for k in names] + extra, allow_unused=True)
    logs = {"total_loss": float(total.detach()), "enc_rec_loss": float(enc_rec.detach()), "reconstruct_loss": float(rec.detach()),
            "vq_loss": out["vq"]["vq_loss"], "kmeans_loss": out["vq"]["kmeans_loss"],
            "number_of_populated_clusters": float(out["soft"].argmax(dim=-1).unique().numel()), "distill_loss": dl}
    grads = dict(zip(names, glist[:len(names)]))
    if extra:
        grads["head/fc.weight"], grads["head/fc.bias"] = glist[len(names)], glist[len(names) + 1]
    return logs, grads, {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


def adam_step_generic(p: Dict[str, Tensor], grads: Dict[str, Optional[Tensor]], state: Dict[str, dict], lr: float,
                      weight_decay: float = 1e-4, clip: Optional[float] = 0.75, betas=(0.9, 0.999), eps: float = 1e-8):
    """clip_grad_value_(0.75) (training.py:164) + torch.optim.Adam(lr, weight_decay) (losses.py:805-814).
    Parameters whose grad is None are skipped (no decay either), like torch."""
    for k, g in grads.items():
        if g is None:
            continue
        g = g.clamp(-clip, clip) if clip else g
        g = g + weight_decay * p[k]
        st = state.setdefault(k, dict(step=0, m=torch.zeros_like(p[k]), v=torch.zeros_like(p[k])))
        st["step"] += 1
        st["m"] = betas[0] * st["m"] + (1 - betas[0]) * g
        st["v"] = betas[1] * st["v"] + (1 - betas[1]) * g * g
        bc1, bc2 = 1 - betas[0] ** st["step"], 1 - betas[1] ** st["step"]
        p[k] = p[k] - lr / bc1 * st["m"] / ((st["v"] / bc2).sqrt() + eps)


# ----------------------------------------------------------------------------
# contrastive: views
# ----------------------------------------------------------------------------
def recompute_edges(x: Tensor, edge_index: Tensor) -> Tensor:
    """model_utils_new.py:332-364: Euclidean length of every edge from the (standardised) node coordinates."""
    c = x[..., 0:2]
    pi, pj = c.index_select(2, edge_index[:, 0].long()), c.index_select(2, edge_index[:, 1].long())
    return torch.sqrt(torch.clamp((pi - pj).pow(2).sum(dim=-1), min=1e-12)).unsqueeze(-1)


def slice_time(x: Tensor, start: Tensor, length: int) -> Tensor:
    t_idx = start.long()[:, None] + torch.arange(length)[None, :]
    return x[torch.arange(x.shape[0])[:, None], t_idx]


@dataclass
class RotationTable:
    """build_rotation_precomp (training.py:2064-2125): triplets (a, b, c) around every node b of degree >= 2 and
    the branch node sets reachable from a / c without crossing b."""
    triplets: List[Tuple[int, int, int]]
    branches_a: List[List[int]]
    branches_c: List[List[int]]


def rotation_table(edge_index: np.ndarray, n_nodes: int) -> RotationTable:
    adj = [[] for _ in range(n_nodes)]
    for u, v in np.asarray(edge_index).tolist():
        adj[u].append(v)
        adj[v].append(u)
    trip = []
    for b in range(n_nodes):
        nb = adj[b]
        for i in range(len(nb)):
            for j in range(i + 1, len(nb)):
                trip.append((nb[i], b, nb[j]))

    def branch(center, side):
        seen, stack = {side}, [side]
        while stack:
            u = stack.pop()
            for v in adj[u]:
                if v == center or v in seen:
                    continue
                seen.add(v)
                stack.append(v)
        return list(seen)

    return RotationTable(trip, [branch(b, a) for a, b, c in trip], [branch(b, c) for a, b, c in trip])


@dataclass
class AugCfg:
    """ContrastiveCfg.aug_* (model_utils_new.py:173-189)."""
    min_shift: int = 1
    max_shift: int = 6
    p_shift: float = 0.8
    max_rot: float = 30.0
    n_rot: int = 4
    p_rot: float = 0.0
    max_interp: int = 8
    min_interp: int = 3
    p_interp: float = 0.3
    noise_sigma: float = 0.03
    p_noise: float = 0.0


@dataclass
class AugParams:
    """The random decisions of one _make_augmented_view call, as plain arrays (what the CUDA kernel takes)."""
    start: Tensor                                  # [B] int   slice start of the augmented view
    rot_pivot: List[int] = field(default_factory=list)
    rot_nodes: List[List[int]] = field(default_factory=list)
    rot_theta: Optional[Tensor] = None             # [R,B] radians (0 where not applied)
    interp_t0: Optional[Tensor] = None             # [B] int, first replaced frame (in the half window)
    interp_len: Optional[Tensor] = None            # [B] int, 0 = not applied
    noise: Optional[Tensor] = None                 # [B,N,3] additive offsets (x, y, speed)


def draw_aug_params(B: int, T_full: int, N: int, cfg: AugCfg, rot: RotationTable) -> AugParams:
    """Replays, call for call, the draws _make_augmented_view makes from torch's GLOBAL generator
    (training.py:2128-2402) and turns them into AugParams.  Seed with torch.manual_seed first."""
    half = T_full // 2
    base = (T_full - half) // 2
    # _augment_time_shift :2128-2167
    apply = torch.rand(B) < cfg.p_shift
    mag = torch.randint(cfg.min_shift, cfg.max_shift + 1, (B,))
    sgn = torch.randint(0, 2, (B,)) * 2 - 1
    start = (base + mag * sgn * apply.long()).clamp(0, T_full - half)
    out = AugParams(start=start.int())
    # _augment_angle_rotations :2170-2256 (operates on the half window)
    M = len(rot.triplets)
    if cfg.n_rot > 0 and cfg.max_rot > 0.0 and cfg.p_rot > 0.0 and M > 0:
        app = (torch.rand(B) < cfg.p_rot).float()
        max_rad = float(cfg.max_rot) * math.pi / 180.0
        perm = torch.randperm(M)
        chosen, count = [], [0] * N
        for k in perm.tolist():
            b0 = rot.triplets[k][1]
            if count[b0] >= 2:
                continue
            count[b0] += 1
            chosen.append(k)
            if len(chosen) >= cfg.n_rot:
                break
        thetas = []
        for k in chosen:
            nodes = rot.branches_a[k] if torch.rand(()) < 0.5 else rot.branches_c[k]   # prefer_side is always 2
            if len(nodes) == 0:
                continue
            theta = (torch.rand(B) * 2.0 - 1.0) * max_rad * app
            out.rot_pivot.append(rot.triplets[k][1])
            out.rot_nodes.append(list(nodes))
            thetas.append(theta)
        if thetas:
            out.rot_theta = torch.stack(thetas)
    # _augment_linear_interpolate_segments :2299-2366
    if cfg.max_interp > 0 and cfg.p_interp > 0.0 and half >= 3:
        app = torch.rand(B) < cfg.p_interp
        L = torch.randint(cfg.min_interp, cfg.max_interp + 1, (B,))
        t0 = torch.randint(1, half - 1, (B,))
        t0 = torch.minimum(t0, (half - L - 1).clamp_min(1))
        out.interp_t0 = t0.int()
        out.interp_len = (L * app.long()).int()
    # _augment_noise_xys :2259-2296
    if cfg.noise_sigma > 0.0 and cfg.p_noise > 0.0:
        app = (torch.rand(B) < cfg.p_noise).float().view(B, 1).expand(B, N)
        axis = torch.randint(0, 2, (B, N))
        off = cfg.noise_sigma * torch.randn(B, N) * app
        ds = cfg.noise_sigma * torch.randn(B, N) * app
        out.noise = torch.stack([off * (axis == 0).float(), off * (axis == 1).float(), ds], dim=-1)
    return out


def augmented_view(x_full: Tensor, prm: AugParams) -> Tensor:
    """x_aug [B, T//2, N, 3] of _make_augmented_view given its random decisions."""
    B, T = x_full.shape[0], x_full.shape[1]
    half = T // 2
    x = slice_time(x_full, prm.start, half).clone()
    if prm.rot_theta is not None:
        coords = x[..., 0:2].clone()
        for r, (b0, nodes) in enumerate(zip(prm.rot_pivot, prm.rot_nodes)):
            th = prm.rot_theta[r]
            c, s = torch.cos(th).view(B, 1, 1), torch.sin(th).view(B, 1, 1)
            idx = torch.as_tensor(nodes, dtype=torch.long)
            pivot = coords[:, :, b0, :].unsqueeze(2)
            rel = coords.index_select(2, idx) - pivot
            rx = rel[..., 0] * c - rel[..., 1] * s
            ry = rel[..., 0] * s + rel[..., 1] * c
            coords[:, :, idx, :] = torch.stack([rx, ry], dim=-1) + pivot
        x[..., 0:2] = coords
    if prm.interp_len is not None:
        t0, L = prm.interp_t0.long(), prm.interp_len.long()
        Lr = torch.where(L > 0, L, torch.ones_like(L))
        bi = torch.arange(B)
        start, end = x[bi, t0 - 1], x[bi, (t0 + Lr).clamp(max=half - 1)]
        tt = torch.arange(half).view(1, half)
        mask = (tt >= t0.view(B, 1)) & (tt < (t0 + L).view(B, 1))
        alpha = ((tt.float() - (t0.view(B, 1).float() - 1.0)) / (Lr.view(B, 1).float() + 1.0)).clamp(0.0, 1.0)
        al = alpha.unsqueeze(-1).unsqueeze(-1)
        interp = (1.0 - al) * start.unsqueeze(1) + al * end.unsqueeze(1)
        x = torch.where(mask.unsqueeze(-1).unsqueeze(-1), interp, x)
    if prm.noise is not None:
        x = x + prm.noise.unsqueeze(1)
    return x


def contrastive_views(x_full: Tensor, edge_index: Tensor, prm: AugParams):
    """(x, a, x_aug, a_aug) of step_contrastive_distill (training.py:498-525)."""
    half = x_full.shape[1] // 2
    starts = (torch.ones(x_full.shape[0]) * half // 2).int()              # :519
    x = slice_time(x_full, starts, half)
    a = slice_time(recompute_edges(x_full, edge_index), starts, half)
    x_aug = augmented_view(x_full, prm)
    return x, a, x_aug, recompute_edges(x_aug, edge_index)


# ----------------------------------------------------------------------------
# contrastive: loss
# ----------------------------------------------------------------------------
def nce_loss(z: Tensor, z_aug: Tensor, temperature: float):
    """normalize (training.py:532-533) -> cosine similarity / temperature -> cross-entropy against the diagonal
    (losses.py:130-141).  Returns loss, mean positive similarity, mean negative similarity."""
    sim = _cos_sim(z, z_aug) / temperature
    n = sim.shape[0]
    loss = torch.nn.functional.cross_entropy(sim, torch.arange(n))
    pos = torch.diag(sim).mean() * temperature
    off = (sim * temperature)[~torch.eye(n, dtype=torch.bool)]
    neg = off.mean() if off.numel() else torch.zeros(())
    return loss, pos, neg


SIMILARITY = "cosine"     # module-level switch used by the tests: cosine | dot | euclidean | edit (losses.py:59-89)


def _cos_sim(z: Tensor, z_aug: Tensor) -> Tensor:
    zn = torch.nn.functional.normalize(z, dim=1)                          # training.py:532-533
    an = torch.nn.functional.normalize(z_aug, dim=1)
    if SIMILARITY == "dot":
        return zn @ an.t()                                                # losses.py:66-67
    if SIMILARITY in ("euclidean", "edit"):                               # losses.py:70-82
        d = torch.sqrt(torch.clamp(((zn.unsqueeze(1) - an.unsqueeze(0)) ** 2).sum(dim=2), min=0.0))
        return 1.0 / (1.0 + d)
    return torch.nn.functional.cosine_similarity(zn.unsqueeze(1), an.unsqueeze(0), dim=2)


def _off_diag(sim: Tensor) -> Tensor:
    n = sim.shape[0]
    return sim[~torch.eye(n, dtype=torch.bool)].reshape(n, n - 1)


def dcl_loss(z: Tensor, z_aug: Tensor, temperature: float, tau_plus: float):
    """dcl_loss_pt, debiased (losses.py:144-173)."""
    sim = _cos_sim(z, z_aug)
    n = sim.shape[0]
    pos = torch.exp(torch.diag(sim) / temperature)
    neg = _off_diag(sim)
    neg_sim = torch.exp(neg / temperature)
    n_eff = n - 1
    ng = (-tau_plus * n_eff * pos + neg_sim.sum(dim=-1)) / (1.0 - tau_plus)
    ng = torch.clamp(ng, min=n_eff * math.e ** (-1.0 / temperature), max=torch.finfo(z.dtype).max)
    loss = (-torch.log(pos / (pos + ng))).mean()
    return loss, torch.diag(sim).mean(), neg.mean()


def hard_loss(z: Tensor, z_aug: Tensor, temperature: float, tau_plus: float, beta: float):
    """hard_loss_pt, debiased (losses.py:213-249)."""
    sim = _cos_sim(z, z_aug)
    n = sim.shape[0]
    pos = torch.exp(torch.diag(sim) / temperature)
    neg = _off_diag(sim)
    neg_sim = torch.exp(neg / temperature)
    reweight = torch.ones_like(neg_sim) if beta == 0.0 else (beta * neg_sim) / neg_sim.mean(dim=1, keepdim=True)
    n_eff = n - 1
    ng = (-tau_plus * n_eff * pos + (reweight * neg_sim).sum(dim=-1)) / (1.0 - tau_plus)
    ng = torch.clamp(ng, min=math.e ** (-1.0 / temperature), max=torch.finfo(z.dtype).max)
    loss = (-torch.log(pos / (pos + ng))).mean()
    return loss, torch.diag(sim).mean(), neg.mean()


def fc_loss(z: Tensor, z_aug: Tensor, temperature: float, elimination_topk: float = 0.1):
    """fc_loss_pt (losses.py:176-210): the k = ceil(min(topk, 0.5) N) largest negatives of every row are dropped."""
    n = z.shape[0]
    k = int(math.ceil(min(elimination_topk, 0.5) * n)) or 1
    sim = _cos_sim(z, z_aug) / temperature
    pos = torch.exp(torch.diag(sim))
    mask = ~torch.eye(n, dtype=torch.bool)
    neg_raw = sim[mask].reshape(n, n - 1)                                  # _off_diagonal_rows, losses.py:91-101
    trimmed = torch.sort(neg_raw, dim=1).values[:, :max(n - 1 - k, 0)]
    neg = torch.exp(trimmed).sum(dim=1) if trimmed.numel() > 0 else torch.zeros(n)
    loss = (-torch.log(pos / (pos + neg))).mean()
    mean_neg = trimmed.mean() * temperature if trimmed.numel() > 0 else torch.tensor(0.0)
    return loss, torch.diag(sim).mean() * temperature, mean_neg


def contrastive_loss(z, z_aug, loss_fn: str, temperature: float, tau_plus: float = 0.1, beta: float = 0.1):
    """select_contrastive_loss_pt (losses.py:35-56) for the cosine similarity."""
    if loss_fn == "nce":
        return nce_loss(z, z_aug, temperature)
    if loss_fn == "dcl":
        return dcl_loss(z, z_aug, temperature, tau_plus)
    if loss_fn == "hard_dcl":
        return hard_loss(z, z_aug, temperature, tau_plus, beta)
    if loss_fn == "fc":
        return fc_loss(z, z_aug, temperature)
    raise ValueError(loss_fn)


CON_LOG_KEYS = ("total_loss", "pos_similarity", "neg_similarity", "distill_loss", "seperability")


def contrastive_train_step(x_full: Tensor, p: Dict[str, Tensor], graph, latent_dim: int, edge_index: Tensor,
                           prm: AugParams, temperature: float = 0.1, loss_fn: str = "nce", tau_plus: float = 0.1,
                           beta: float = 0.1, distill: Optional[dict] = None):
    """step_contrastive_distill forward + backward without labels; teacher off unless `distill` is given (the head sees
    the row-normalised embedding of the MAIN view, training.py:533, 556-557)."""
    names = [k for k in p if k not in V.BUFFER_NAMES]
    leaf = {k: (v.detach().clone().requires_grad_(True) if k in names and v.dtype.is_floating_point else v) for k, v in p.items()}
    x, a, xa, aa = contrastive_views(x_full, edge_index, prm)
    z = V.encoder_forward(x, a, leaf, graph, latent_dim)
    za = V.encoder_forward(xa, aa, leaf, graph, latent_dim)
    loss, pos, neg = contrastive_loss(z, za, loss_fn, temperature, tau_plus, beta)
    dl, extra = 0.0, []
    if distill is not None and distill["lam"] > 0.0:
        hw, hb = _distill_leaves(distill)
        # training.py:533, 556: z has been REASSIGNED to its row-normalised copy before `z_main = z`
        dlt = distill_loss(torch.nn.functional.normalize(z, dim=1), hw, hb, distill["tau"], distill["lam"],
                           distill.get("sharpen_T", 0.5), distill.get("conf_weight", False), distill.get("conf_thresh", 0.6))
        loss = loss + dlt
        dl, extra = float(dlt.detach()), [hw, hb]
    glist = torch.autograd.grad(loss, [leaf[k] for k in names] + extra, allow_unused=True)
    logs = {"total_loss": float(loss.detach()), "pos_similarity": float(pos.detach()), "neg_similarity": float(neg.detach()),
            "distill_loss": dl, "seperability": 0.0}
    grads = dict(zip(names, glist[:len(names)]))
    if extra:
        grads["head/fc.weight"], grads["head/fc.bias"] = glist[len(names)], glist[len(names) + 1]
    return logs, grads, dict(z=z.detach(), z_aug=za.detach(), x=x, a=a, x_aug=xa, a_aug=aa)
