"""CPU oracle for the TCN model family of mlfpm/deepof.

TEST INFRASTRUCTURE — NOT THE PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may
import this module, and only as the checker.  The product path (``deepof_b200``) never routes through it.

A restatement (own code, functional style, plain torch CPU ops) of deepof/clustering/models_new.py
  TemporalBlockPT :376-445, TCN1DPT :447-503, BatchNorm1dKerasFP32 :505-513, TCNEncoderPT :518-657, TCNDecoderPT :713-819
in eval AND train mode (batch statistics of every BatchNorm are returned so that the running buffers can be checked), and the
VaDE / VQ-VAE / contrastive training steps of the reference models built with ``encoder_type="TCN"`` (the latent space, the
vector quantiser and the losses are the ones of vade_oracle / models_oracle).  The TCN stacks have dropout_rate = 0 in every
caller of the reference (init_encoder_decoder :1467-1486, ContrastivePT :2041-2049), so a training step draws only the
reparameterisation / Monte-Carlo noise.

Parity pinning: ``tests/golden/tcn*.npz`` are outputs of the UNMODIFIED reference (``tests/golden/make_golden_tcn.py``);
``tests/test_oracle_tcn_golden.py`` checks this oracle against every one of them.  Parity is therefore pinned.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import vade_oracle as V

Tensor = torch.Tensor

ENC_DILATIONS = (1, 2, 4, 8, 1, 2, 4, 8)      # TCNEncoderPT: conv_stacks=2, conv_dilations=(1,2,4,8)  (:542-545)
DEC_DILATIONS = (8, 4, 2, 1)                  # TCNDecoderPT: conv_stacks=1, conv_dilations=(8,4,2,1)  (:736-737)
BLOCK_MOMENTUM = 0.1                          # nn.BatchNorm1d default inside TemporalBlockPT (:409, :413)
KERAS_MOMENTUM = 0.01                         # BatchNorm1dKerasFP32 (:506)


def batch_norm(v: Tensor, p: Dict[str, Tensor], pre: str, train: bool, stats: dict, dims) -> Tensor:
    """BatchNorm1d with eps 1e-3 over `dims` (all but the channel dimension 1).  train: batch statistics (biased variance),
    recorded as stats[pre] = (mean, biased var, count); eval: the running buffers."""
    shape = [1, -1] + [1] * (v.dim() - 2)
    if train:
        mu = v.mean(dim=dims)
        var = v.var(dim=dims, unbiased=False)
        n = v.numel() // v.shape[1]
        stats[pre] = (mu.detach(), var.detach(), n)
    else:
        mu, var = p[pre + "running_mean"].to(v.dtype), p[pre + "running_var"].to(v.dtype)
    return (v - mu.view(shape)) / torch.sqrt(var.view(shape) + 1e-3) * p[pre + "weight"].view(shape) + p[pre + "bias"].view(shape)


def causal_conv(x: Tensor, w: Tensor, b: Tensor, dilation: int) -> Tensor:
    """_causal_pad + Conv1d (:427-429, :433-437): x [S, C_in, T], left padding (k-1)*dilation."""
    pad = (w.shape[2] - 1) * dilation
    return torch.nn.functional.conv1d(torch.nn.functional.pad(x, (pad, 0)), w, b, dilation=dilation)


def temporal_block(x: Tensor, p: Dict[str, Tensor], pre: str, dilation: int, train: bool, stats: dict) -> Tuple[Tensor, Tensor]:
    """TemporalBlockPT.forward (:431-445), padding="causal", activation relu, batch norm on, dropout 0."""
    y = torch.relu(batch_norm(causal_conv(x, p[pre + "conv1.weight"], p[pre + "conv1.bias"], dilation), p, pre + "bn1.", train, stats, (0, 2)))
    y = torch.relu(batch_norm(causal_conv(y, p[pre + "conv2.weight"], p[pre + "conv2.bias"], dilation), p, pre + "bn2.", train, stats, (0, 2)))
    res = x
    if pre + "downsample.weight" in p:
        res = torch.nn.functional.conv1d(x, p[pre + "downsample.weight"], p[pre + "downsample.bias"])
    return torch.relu(y + res), y


def tcn(x: Tensor, p: Dict[str, Tensor], pre: str, dilations, return_sequences: bool, train: bool, stats: dict) -> Tensor:
    """TCN1DPT.forward with skip connections (:486-503): x [S, T, C_in] -> [S, T, C] or the last step [S, C]."""
    y = x.transpose(1, 2)
    skip_sum = None
    for i, d in enumerate(dilations):
        y, skip = temporal_block(y, p, pre + f"blocks.{i}.", d, train, stats)
        skip_sum = skip if skip_sum is None else skip_sum + skip
    out = torch.relu(skip_sum).transpose(1, 2)
    return out if return_sequences else out[:, -1, :]


def stabilize(v: Tensor) -> Tensor:
    """Per-sample RMS normalisation, clamp and nan_to_num (:640-650, :778-790)."""
    rms = v.pow(2).mean(dim=1, keepdim=True).sqrt()
    return torch.nan_to_num((v / rms.clamp(min=1.0)).clamp(min=-1e4, max=1e4), nan=0.0, posinf=1e4, neginf=-1e4)


def encoder_forward(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, train: bool, root: str = "encoder."):
    """TCNEncoderPT.forward, use_gnn=True (:609-657).  Returns (out [B, D], stats)."""
    B, T, N, F = x.shape
    E = a.shape[2]
    lap, elap, inc = graph
    stats = {}
    xn = V.group_reshape(x).reshape(B * N, T, F)
    xe = V.group_reshape(a).reshape(B * E, T, a.shape[3])
    nodes = tcn(xn, p, root + "node_tcn.", ENC_DILATIONS, False, train, stats).view(B, N, -1)
    edges = tcn(xe, p, root + "edge_tcn.", ENC_DILATIONS, False, train, stats).view(B, E, -1)
    gn, ge = V.censnet(nodes, edges, lap.to(x.dtype), elap.to(x.dtype), inc.to(x.dtype), p, root + "spatial_gnn_block.")
    h = stabilize(torch.cat([torch.relu(gn).reshape(B, -1), torch.relu(ge).reshape(B, -1)], dim=-1))
    h1 = batch_norm(torch.relu(h @ p[root + "head.0.weight"].t() + p[root + "head.0.bias"]), p, root + "head.2.", train, stats, (0,))
    h2 = batch_norm(torch.relu(h1 @ p[root + "head.3.weight"].t() + p[root + "head.3.bias"]), p, root + "head.5.", train, stats, (0,))
    return h2 @ p[root + "head.6.weight"].t() + p[root + "head.6.bias"], stats


def decoder_forward(z: Tensor, x_flat: Tensor, p: Dict[str, Tensor], train: bool, pre: str = "decoder."):
    """TCNDecoderPT.forward (:792-819).  Returns (loc [B, T, N*F], validity mask [B, T] as float, stats)."""
    B, T, _ = x_flat.shape
    mask = ~torch.all(x_flat == 0.0, dim=2)
    stats = {}
    g = stabilize(z)
    h = batch_norm(g @ p[pre + "fc0.weight"].t() + p[pre + "fc0.bias"], p, pre + "bn0.", train, stats, (0,))
    h = batch_norm(torch.relu(h @ p[pre + "fc1.weight"].t() + p[pre + "fc1.bias"]), p, pre + "bn1.", train, stats, (0,))
    h = batch_norm(torch.relu(h @ p[pre + "fc2.weight"].t() + p[pre + "fc2.bias"]), p, pre + "bn2.", train, stats, (0,))
    seq = tcn(h.unsqueeze(1).repeat(1, T, 1), p, pre + "tcn.", DEC_DILATIONS, True, train, stats)
    loc = seq @ p[pre + "prob_decoder.loc_projection.weight"].t() + p[pre + "prob_decoder.loc_projection.bias"]
    return torch.nan_to_num(loc, nan=0.0, posinf=1e6, neginf=-1e6), mask.to(z.dtype), stats


def momentum_of(name: str) -> float:
    """Running-statistics momentum of the BatchNorm whose parameter prefix is `name`."""
    return BLOCK_MOMENTUM if ".blocks." in name else KERAS_MOMENTUM


def running_after(p: Dict[str, Tensor], stats_list) -> Dict[str, Tensor]:
    """The running buffers after train-mode forward passes whose statistics are stats_list (in order)."""
    out = {}
    for stats in stats_list:
        for pre, (mu, var, n) in stats.items():
            m = momentum_of(pre)
            rm = out.get(pre + "running_mean", p[pre + "running_mean"])
            rv = out.get(pre + "running_var", p[pre + "running_var"])
            out[pre + "running_mean"] = (1 - m) * rm + m * mu
            out[pre + "running_var"] = (1 - m) * rv + m * var * (n / max(n - 1, 1))
            out[pre + "num_batches_tracked"] = out.get(pre + "num_batches_tracked", p[pre + "num_batches_tracked"]) + 1
    return out


def _leaves(p):
    names = [k for k in p if p[k].dtype.is_floating_point and k not in V.BUFFER_NAMES and "running_" not in k]
    return names, {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in p.items()}


def model_forward_eval(kind: str, x: Tensor, a: Tensor, p: Dict[str, Tensor], graph):
    """Eval-mode outputs of the TCN models: encoder output, latent heads / quantiser, decoder mean."""
    from . import models_oracle as MO
    B, T, N, F = x.shape
    enc, _ = encoder_forward(x, a, p, graph, False)
    out = dict(enc=enc)
    if kind == "contrastive":
        return out
    xf = x.reshape(B, T, N * F)
    if kind == "vade":
        z, q, z_mean, _ = V.latent_forward(enc, p, False, None)
        out.update(z=z_mean, q=q, loc=decoder_forward(z, xf, p, False)[0])
    else:
        quant, soft, idx, _ = MO.vq_forward(enc, p["vq_layer.codebook"], 1.0, 0.0)
        out.update(quant=quant, soft=soft, idx=idx, loc=decoder_forward(quant, xf, p, False)[0])
    return out


def vade_train_step(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, cfg, eps: Tensor, mc_eps=None, tau_batch=None,
                    class_weight=None, teacher_marginal=None):
    """step_vade forward + backward (training.py:231-309) for VaDEPT(encoder_type="TCN").  Returns (logs, grads, outputs)."""
    names, leaf = _leaves(p)
    B, T, N, F = x.shape
    enc, s_enc = encoder_forward(x, a, leaf, graph, True)
    z, q, z_mean, z_log_var = V.latent_forward(enc, leaf, True, eps)
    km = V.kmeans_loss(z, cfg.model_kmeans_weight) if cfg.model_kmeans_weight > 0 else torch.zeros((), dtype=x.dtype)
    xf = x.reshape(B, T, N * F)
    loc, mask, s_dec = decoder_forward(z, xf, leaf, True)
    losses = V.vade_loss(loc, mask, z, q, km, z_mean, z_log_var, leaf, xf, cfg, mc_eps=mc_eps, tau_batch=tau_batch,
                         class_weight=class_weight, teacher_marginal=teacher_marginal)
    glist = torch.autograd.grad(losses["total_loss"], [leaf[k] for k in names], allow_unused=True)
    logs = {k: float(losses[k].detach()) for k in V.LOG_KEYS}
    return logs, dict(zip(names, glist)), dict(enc=enc.detach(), z=z.detach(), q=q.detach(), loc=loc.detach(), bn=[s_enc, s_dec])


def vqvae_train_step(x: Tensor, a: Tensor, p: Dict[str, Tensor], graph, beta: float = 1.0, kmeans_w: float = 0.0):
    """step_vqvae_distill (training.py:312-389, teacher off) on VQVAEPT(encoder_type="TCN"): decoder(quantized) first, then
    decoder(encoder output) (models_new.py:1603-1612) — each pass with its own batch statistics."""
    from . import models_oracle as MO
    names, leaf = _leaves(p)
    B, T, N, F = x.shape
    enc, s_enc = encoder_forward(x, a, leaf, graph, True)
    quant, soft, idx, vql = MO.vq_forward(enc, leaf["vq_layer.codebook"], beta, kmeans_w)
    xf = x.reshape(B, T, N * F)
    loc_q, mask, s_q = decoder_forward(quant, xf, leaf, True)
    loc_e, _, s_e = decoder_forward(enc, xf, leaf, True)
    enc_rec = -(V.recon_log_prob(loc_q, mask, xf)).mean()
    rec = -(V.recon_log_prob(loc_e, mask, xf)).mean()
    total = enc_rec + rec + (vql["vq_loss"] + vql["kmeans_loss"])
    glist = torch.autograd.grad(total, [leaf[k] for k in names], allow_unused=True)
    logs = {"total_loss": float(total.detach()), "enc_rec_loss": float(enc_rec.detach()), "reconstruct_loss": float(rec.detach()),
            "vq_loss": vql["vq_loss"], "kmeans_loss": vql["kmeans_loss"],
            "number_of_populated_clusters": float(soft.argmax(dim=-1).unique().numel()), "distill_loss": 0.0}
    return logs, dict(zip(names, glist)), dict(enc=enc.detach(), quant=quant.detach(), soft=soft.detach(), idx=idx, bn=[s_enc, s_q, s_e])


def contrastive_views_step(x: Tensor, a: Tensor, xa: Tensor, aa: Tensor, p: Dict[str, Tensor], graph, temperature: float = 0.1,
                           loss_fn: str = "nce", tau_plus: float = 0.1, beta: float = 0.1):
    """The encoder half of step_contrastive_distill (training.py:527-545, teacher off) on ContrastivePT(encoder_type="TCN")
    given the two views: TWO encoder passes with separate batch statistics, NT-Xent on the embeddings."""
    from . import models_oracle as MO
    names, leaf = _leaves(p)
    z, s1 = encoder_forward(x, a, leaf, graph, True)
    za, s2 = encoder_forward(xa, aa, leaf, graph, True)
    loss, pos, neg = MO.contrastive_loss(z, za, loss_fn, temperature, tau_plus, beta)
    glist = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    logs = {"total_loss": float(loss.detach()), "pos_similarity": float(pos.detach()), "neg_similarity": float(neg.detach()),
            "distill_loss": 0.0, "seperability": 0.0}
    return logs, dict(zip(names, glist)), dict(z=z.detach(), z_aug=za.detach(), bn=[s1, s2])
