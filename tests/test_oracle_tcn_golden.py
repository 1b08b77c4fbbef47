"""CPU: the TCN-family oracle (oracle/tcn_oracle.py, eval and train mode) vs golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden_tcn.py)."""
import pytest
import torch

from oracle import tcn_oracle as TC
from oracle import vade_oracle as O
from helpers import golden_cases_of, load_golden_of, sub, rel_l2


def test_goldens_present():
    assert len(golden_cases_of("tcnmodel")) >= 3 and len(golden_cases_of("tcnvade")) >= 3 and len(golden_cases_of("tcnstep")) >= 4


@pytest.mark.parametrize("case", golden_cases_of("tcnmodel"))
def test_tcn_models_eval(case):
    """Eval-mode (running statistics) encoder output, latent heads / quantiser and decoder mean of the three TCN models;
    window 1 has an all-zero second half (validity mask of the probabilistic decoder)."""
    g = load_golden_of("tcnmodel", case)
    p = sub(g, "p/")
    kind = str(g["model"])
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    with torch.no_grad():
        out = TC.model_forward_eval(kind, x, a, p, O.graph_operators(g["adjacency"]))
    assert rel_l2(out["enc"], g["eval/enc"]) < 1e-5
    if kind == "vade":
        assert rel_l2(out["z"], g["eval/emb"]) < 1e-5 and rel_l2(out["q"], g["eval/q"]) < 1e-5
        assert rel_l2(out["loc"], g["eval/loc"]) < 1e-5
    if kind == "vqvae":
        assert rel_l2(out["quant"], g["eval/quant"]) < 1e-5 and rel_l2(out["soft"], g["eval/q"]) < 1e-5
        assert rel_l2(out["loc"], g["eval/loc_q"]) < 1e-5


def _check_step(g, p, logs, grads, bn):
    for k in logs:
        ref = float(g["log/" + k])
        assert abs(logs[k] - ref) <= 5e-5 * max(1.0, abs(ref)), (k, logs[k], ref)
    names = [k[2:] for k in g if k.startswith("g/")]
    flat = torch.cat([grads[k].flatten() for k in names])
    ref = torch.cat([torch.from_numpy(g["g/" + k]).flatten() for k in names])
    # the TCN stacks (16 train-mode BatchNorms + ReLUs per branch) amplify fp32 rounding: the reference's own fp32 gradient
    # sits 1e-4 .. 5e-4 (relative, per tensor) away from an fp64 evaluation of the same step
    assert rel_l2(flat, ref) < 3e-4
    for k, v in grads.items():
        assert (v is None) == (k not in names), k
    run = TC.running_after(p, bn)
    assert len(run) == sum(1 for k in g if k.startswith("p1/"))
    for k, v in run.items():
        assert rel_l2(v, g["p1/" + k]) < 1e-5, k


@pytest.mark.parametrize("case", golden_cases_of("tcnvade"))
def test_vade_tcn_train_step(case):
    """step_vade on VaDEPT(encoder_type="TCN"): the 13 logged terms, every parameter gradient and the running statistics
    of all 46 BatchNorm layers after the step, with the reference's noise as input."""
    g = load_golden_of("tcnvade", case)
    p = sub(g, "p/")
    K = int(g["meta"][4])
    main = str(g["phase"]) == "main"
    cfg = (O.LossCfg.main_defaults if main else O.LossCfg.pretrain_defaults)(K, kl_weight=float(g["klw"]))
    logs, grads, out = TC.vade_train_step(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]), p, O.graph_operators(g["adjacency"]), cfg,
                                          torch.from_numpy(g["eps"]), mc_eps=torch.from_numpy(g["mc_eps"]) if main else None)
    _check_step(g, p, {k: logs[k] for k in O.LOG_KEYS}, grads, out["bn"])


@pytest.mark.parametrize("case", golden_cases_of("tcnstep"))
def test_vqvae_and_contrastive_tcn_steps(case):
    """step_vqvae_distill (two decoder passes, each updating the decoder's running statistics) and
    step_contrastive_distill (two encoder passes with separate batch statistics), teacher off."""
    g = load_golden_of("tcnstep", case)
    p = sub(g, "p/")
    graph = O.graph_operators(g["adjacency"])
    if str(g["model"]) == "vqvae":
        logs, grads, out = TC.vqvae_train_step(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]), p, graph, float(g["beta"]), 0.0)
    else:
        logs, grads, out = TC.contrastive_views_step(*(torch.from_numpy(g[k]) for k in ("x", "a", "x_aug", "a_aug")), p, graph, 0.1)
    _check_step(g, p, logs, grads, out["bn"])
