"""CPU: the VQ-VAE / contrastive oracle (oracle/models_oracle.py) vs golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden_models.py)."""
import numpy as np
import pytest
import torch

from oracle import models_oracle as MO
from oracle import vade_oracle as O
from helpers import golden_cases_of, load_golden_of, sub, rel_l2

VQ = golden_cases_of("vqvae")
CON = golden_cases_of("contrastive")


def test_goldens_present():
    assert len(VQ) >= 3 and len(CON) >= 3


def _check_grads(grads, gref, p, D):
    for k, gr in grads.items():
        if O.dead_parameter(k, p, D):
            assert gr is None and k not in gref, k
            continue
        assert k in gref, k
        err, scale = float((gr - gref[k]).abs().max()), float(gref[k].abs().max())
        assert err <= 3e-5 * max(scale, 1e-3) + 1e-7, (k, err, scale)
    flat = torch.cat([grads[k].flatten() for k in gref])
    flat_ref = torch.cat([gref[k].flatten() for k in gref])
    assert rel_l2(flat, flat_ref) < 1e-5


def _check_params(p, p2, lr=1e-3):
    """Parameters after two clip + Adam(weight_decay) steps.  Adam's first steps move every element by about
    lr * g / (|g| + 1e-8): where |g| is itself ~1e-8 (rounding noise of the gradient) the update is ill-conditioned,
    so single elements are allowed 10 % of one step while each tensor must agree to 5e-5 rel-L2."""
    for k in p2:
        if p2[k].dtype.is_floating_point and p2[k].numel() > 0:
            assert float((p[k] - p2[k]).abs().max()) <= 0.1 * lr, k
            assert rel_l2(p[k], p2[k]) < 5e-5, k


def _distill_of(g, hp):
    """Teacher-on goldens carry distill/*: the oracle's `distill` argument for the current head parameters."""
    if "distill/meta" not in g:
        return None
    Kt, lam, Tsh, cw, thr = (float(v) for v in g["distill/meta"])
    tau = torch.from_numpy(g["distill/tau_star"])[torch.from_numpy(g["idx"]).long()]
    return dict(head_w=hp["fc.weight"], head_b=hp["fc.bias"], tau=tau, lam=lam, sharpen_T=Tsh, conf_weight=bool(cw), conf_thresh=thr)


def _head_step(g, hp, grads, hstate, step):
    """gradient of the head after step 0, then its Adam step: same optimizer, but NOT clipped (training.py:165)."""
    hg = {"fc.weight": grads.pop("head/fc.weight"), "fc.bias": grads.pop("head/fc.bias")}
    if step == 0:
        for k, v in hg.items():
            assert rel_l2(v, g["distill/g/" + k]) < 1e-5, k
    MO.adam_step_generic(hp, hg, hstate, float(g["lr"]), clip=None)


@pytest.mark.parametrize("case", VQ)
def test_vqvae_eval_and_two_steps(case):
    g = load_golden_of("vqvae", case)
    T_, N, E, D, K, B = (int(v) for v in g["meta"])
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    beta, km = float(g["beta"]), float(g["kmeans"])
    with torch.no_grad():
        out = MO.vqvae_forward(x, a, p, graph, D, beta, km)
    assert rel_l2(out["enc"], g["eval/enc"]) < 2e-6
    assert torch.equal(out["idx"], torch.from_numpy(g["eval/idx"]))           # code indices bit-exact
    assert rel_l2(out["soft"], g["eval/soft"]) < 2e-5
    assert rel_l2(out["quant"], g["eval/quant"]) < 1e-6
    assert rel_l2(out["loc_q"], g["eval/loc_q"]) < 5e-6 and rel_l2(out["loc_e"], g["eval/loc_e"]) < 5e-6
    state, hstate, hp = {}, {}, sub(g, "distill/p/")
    for step in range(2):
        logs, grads, _ = MO.vqvae_train_step(x, a, p, graph, D, beta, km, distill=_distill_of(g, hp))
        for k in MO.VQ_LOG_KEYS:
            ref = float(g[f"s{step}/log/{k}"])
            assert abs(logs[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (step, k, logs[k], ref)
        if hp:
            _head_step(g, hp, grads, hstate, step)
        if step == 0:
            _check_grads(grads, sub(g, "g/"), p, D)
        MO.adam_step_generic(p, grads, state, float(g["lr"]))
    _check_params(p, sub(g, "p2/"))
    if hp:
        _check_params(hp, sub(g, "distill/p2/"))


def _aug_cfg(g):
    c = MO.AugCfg()
    for f in ("min_shift", "max_shift", "n_rot", "max_interp", "min_interp"):
        setattr(c, f, int(g["aug/" + f]))
    for f in ("p_shift", "max_rot", "p_rot", "p_interp", "noise_sigma", "p_noise"):
        setattr(c, f, float(g["aug/" + f]))
    return c


@pytest.mark.parametrize("case", CON)
def test_contrastive_views_and_two_steps(case, monkeypatch):
    g = load_golden_of("contrastive", case)
    Tf, N, E, D, B = (int(v) for v in g["meta"])
    p = sub(g, "p/")
    x_full = torch.from_numpy(g["x_full"])
    ei = torch.from_numpy(g["edge_index"])
    graph = O.graph_operators(g["adjacency"])
    rot = MO.rotation_table(g["edge_index_local"], N)
    cfg = _aug_cfg(g)
    state, hstate, hp = {}, {}, sub(g, "distill/p/")
    monkeypatch.setattr(MO, "SIMILARITY", str(g["similarity_function"]) if "similarity_function" in g else "cosine")
    for step in range(2):
        torch.manual_seed(int(g[f"s{step}/seed"]))
        prm = MO.draw_aug_params(B, Tf, N, cfg, rot)
        logs, grads, out = MO.contrastive_train_step(x_full, p, graph, D, ei, prm, float(g["temperature"]),
                                                     loss_fn=str(g["loss_function"]) if "loss_function" in g else "nce",
                                                     tau_plus=float(g["tau"]) if "tau" in g else 0.1,
                                                     beta=float(g["beta"]) if "beta" in g else 0.1,
                                                     distill=_distill_of(g, hp))
        # the four tensors the reference fed its encoder
        for k in ("x", "a", "x_aug", "a_aug"):
            np.testing.assert_allclose(out[k].numpy(), g[f"s{step}/{k}"], rtol=0, atol=2e-6, err_msg=k)
        assert rel_l2(out["z"], g[f"s{step}/z"]) < 3e-6 and rel_l2(out["z_aug"], g[f"s{step}/z_aug"]) < 3e-6
        for k in MO.CON_LOG_KEYS:
            ref = float(g[f"s{step}/log/{k}"])
            assert abs(logs[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (step, k, logs[k], ref)
        if hp:
            _head_step(g, hp, grads, hstate, step)
        if step == 0:
            _check_grads(grads, sub(g, "g/"), p, D)
        MO.adam_step_generic(p, grads, state, float(g["lr"]))
    _check_params(p, sub(g, "p2/"))
    if hp:
        _check_params(hp, sub(g, "distill/p2/"))
