"""CPU: the transformer-encoder oracle (oracle/tfm_oracle.py, eval mode) vs golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden_tfm.py)."""
import numpy as np
import pytest
import torch

from oracle import tfm_oracle as TO
from oracle import vade_oracle as O
from helpers import golden_cases_of, load_golden_of, sub, rel_l2

TFM = golden_cases_of("tfm")


def test_goldens_present():
    assert len(TFM) >= 3


@pytest.mark.parametrize("case", TFM)
def test_tfm_encoder_eval(case):
    g = load_golden_of("tfm", case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    heads = int(g["meta"][6])
    with torch.no_grad():
        out = TO.encoder_forward_eval(x, a, p, graph, heads)
    B, N, E = x.shape[0], x.shape[2], a.shape[2]
    assert rel_l2(out["nodes"].reshape(B * N, -1), g["eval/nodes"]) < 5e-6
    assert rel_l2(out["edges"].reshape(B * E, -1), g["eval/edges"]) < 5e-6
    assert rel_l2(out["out"], g["eval/out"]) < 1e-5
    if case == "padded":      # the key-padding mask is really exercised
        xs = O.group_reshape(x).reshape(B * N, x.shape[1], 3)
        assert bool((xs == 0).all(-1).any())


@pytest.mark.parametrize("case", [c for c in golden_cases_of("tfmmodel") if c != "contrastive"])
def test_tfm_decoder_eval(case):
    """TFMDecoderPT in eval mode on the latent the reference fed it (VaDE: z_mean, VQ-VAE: encoder output / codes)."""
    g = load_golden_of("tfmmodel", case)
    p = sub(g, "p/")
    x = torch.from_numpy(g["x"])
    B, T, N, F = x.shape
    xf = x.reshape(B, T, N * F)
    with torch.no_grad():
        loc, mask = TO.decoder_forward_eval(torch.from_numpy(g["eval/emb"]), xf, p)
        assert rel_l2(loc, g["eval/loc"]) < 1e-5
        if "eval/loc_q" in g:
            loc_q, _ = TO.decoder_forward_eval(torch.from_numpy(g["eval/quant"]), xf, p)
            assert rel_l2(loc_q, g["eval/loc_q"]) < 1e-5
    assert bool(mask.all())


def _unpack_masks(g):
    masks = {}
    for k in g:
        if k.startswith("mask/"):
            shp = tuple(int(v) for v in g["mshape/" + k[5:]])
            n = int(np.prod(shp))
            masks[k[5:]] = torch.from_numpy(np.unpackbits(g[k])[:n].astype(np.float32)).reshape(shp)
    return masks


@pytest.mark.parametrize("case", golden_cases_of("tfmtrain"))
def test_tfm_encoder_train_mode_with_recorded_dropout(case):
    """TFMEncoderPT in train(): output, parameter gradients of sum(out * probe) and the BatchNorm running statistics
    after the step, with the reference's dropout masks as inputs (replayed from its generator when the golden was made)."""
    g = load_golden_of("tfmtrain", case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    names = [k[2:] for k in g if k.startswith("g/")]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    out, stats = TO.encoder_forward_train(x, a, leaf, graph, _unpack_masks(g), int(g["meta"][6]))
    assert rel_l2(out.detach(), g["train/out"]) < 1e-5
    grads = torch.autograd.grad((out * torch.from_numpy(g["probe"])).sum(), [leaf[k] for k in names])
    flat = torch.cat([v.flatten() for v in grads])
    ref = torch.cat([torch.from_numpy(g["g/" + k]).flatten() for k in names])
    assert rel_l2(flat, ref) < 2e-5
    B = x.shape[0]
    for i in (2, 5):
        mu, var = stats[f"head.{i}"]
        rm = 0.99 * p[f"head.{i}.running_mean"] + 0.01 * mu
        rv = 0.99 * p[f"head.{i}.running_var"] + 0.01 * var * B / (B - 1)
        assert rel_l2(rm, g[f"p1/head.{i}.running_mean"]) < 1e-5 and rel_l2(rv, g[f"p1/head.{i}.running_var"]) < 1e-5
    # about 10 % of every mask is dropped
    for k, m in _unpack_masks(g).items():
        assert 0.05 < 1.0 - float(m.mean()) < 0.16, k


@pytest.mark.parametrize("case", golden_cases_of("tfmvade"))
def test_vade_transformer_train_step(case):
    """step_vade on VaDEPT(encoder_type="transformer"): the 13 logged terms and every parameter gradient, with the
    reference's noise (eps, MC eps) and dropout masks (encoder + causal decoder) as inputs."""
    g = load_golden_of("tfmvade", case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    K = int(g["meta"][4])
    main = str(g["phase"]) == "main"
    cfg = (O.LossCfg.main_defaults if main else O.LossCfg.pretrain_defaults)(K, kl_weight=float(g["klw"]))
    logs, grads, _ = TO.vade_train_step(x, a, p, O.graph_operators(g["adjacency"]), cfg, _unpack_masks(g), torch.from_numpy(g["eps"]),
                                        mc_eps=torch.from_numpy(g["mc_eps"]) if main else None)
    for k in O.LOG_KEYS:
        ref = float(g["log/" + k])
        assert abs(logs[k] - ref) <= 5e-5 * max(1.0, abs(ref)), (k, logs[k], ref)
    names = [k[2:] for k in g if k.startswith("g/")]
    flat = torch.cat([grads[k].flatten() for k in names])
    ref = torch.cat([torch.from_numpy(g["g/" + k]).flatten() for k in names])
    assert rel_l2(flat, ref) < 5e-5
    for k, v in grads.items():
        assert (v is None) == (k not in names), k


@pytest.mark.parametrize("case", golden_cases_of("tfmstep"))
def test_vqvae_and_contrastive_transformer_steps(case):
    """step_vqvae_distill / step_contrastive_distill (teacher off) of the transformer family: logs and gradients of the
    reference with its dropout masks replayed (two decoder passes for VQ-VAE; two encoder passes with separate batch
    statistics for the contrastive model)."""
    g = load_golden_of("tfmstep", case)
    p = sub(g, "p/")
    graph = O.graph_operators(g["adjacency"])
    masks = _unpack_masks(g)
    if str(g["model"]) == "vqvae":
        logs, grads, _ = TO.vqvae_train_step(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]), p, graph, masks, float(g["beta"]), 0.0)
    else:
        m0 = {k: v for k, v in masks.items() if not k.startswith("aug.")}
        m1 = {k[4:]: v for k, v in masks.items() if k.startswith("aug.")}
        logs, grads, _ = TO.contrastive_views_step(*(torch.from_numpy(g[k]) for k in ("x", "a", "x_aug", "a_aug")), p, graph, m0, m1, 0.1)
    for k in logs:
        ref = float(g["log/" + k])
        assert abs(logs[k] - ref) <= 5e-5 * max(1.0, abs(ref)), (k, logs[k], ref)
    names = [k[2:] for k in g if k.startswith("g/")]
    flat = torch.cat([grads[k].flatten() for k in names])
    ref = torch.cat([torch.from_numpy(g["g/" + k]).flatten() for k in names])
    assert rel_l2(flat, ref) < 5e-5
