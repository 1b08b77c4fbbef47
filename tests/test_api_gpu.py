"""GPU: the reference-facing entry points (deepof_b200/api.py) — step functions with the reference signature checked
against the reference-generated goldens, train_deepof_model end to end for the three model kinds, and the checkpoint
bundle round trip (reference tests/test_build_models.py:245,677 style: types, key sets, save/load equality)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import vade_oracle as O
from helpers import golden_cases, load_golden, load_golden_of, sub

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["cfg2_main", "odd_pretrain"])
def test_step_vade_signature_vs_reference_golden(case):
    from deepof_b200 import VaDEB200, VadeLossCfg, step_vade, StepResult
    g = load_golden(case)
    d = g["dims"]
    m = VaDEB200((d["T"], d["N"], 3), (d["T"], d["E"], 1), g["adjacency"], d["D"], d["K"], max_batch=d["B"], seed=0)
    m.load_state_dict({k[2:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith("p/")})
    klw = float(g["s0/klw"])
    crit = VadeLossCfg.pretrain_defaults(d["K"], klw) if str(g["phase"]) == "pretrain" else VadeLossCfg.main_defaults(d["K"], klw)
    ctx = SimpleNamespace(criterion=crit, apply_distill=False, eps=torch.from_numpy(g["s0/eps"]),
                          mc_eps=torch.from_numpy(g["s0/mc_eps"]) if "s0/mc_eps" in g else None)
    batch = (torch.from_numpy(g["x"]), torch.from_numpy(g["a"]), torch.arange(d["B"]))
    res = step_vade(m, batch, ctx)
    assert isinstance(res, StepResult) and res.loss.backward() is None
    assert set(O.LOG_KEYS) <= set(res.logs)                                     # the reference's 13 log keys
    for k in O.LOG_KEYS:
        ref = float(g[f"s0/log/{k}"])
        assert abs(res.logs[k] - ref) <= 1e-4 * max(1.0, abs(ref)), (k, res.logs[k], ref)
    assert abs(res.loss.item() - float(g["s0/log/total_loss"])) <= 1e-4 * abs(float(g["s0/log/total_loss"]))
    total = sum(res.logs[k] for k in O.LOG_KEYS if k != "total_loss")         # loss-sum identity (reference tests: 1e-5)
    assert abs(total - res.logs["total_loss"]) <= 1e-4 * max(1.0, abs(total))


def _table_dicts(N, T, n_videos, nw, seed):
    adj = O.default_adjacency(N)
    td = {}
    for v in range(n_videos):
        x, a = O.synthetic_windows(nw, T, adj, seed=seed + v)
        nodes = torch.cat([x[..., 0], x[..., 1], x[..., 2]], dim=-1).numpy()      # [Nw,T,3N]: x.. | y.. | speed..
        td[f"video_{v}"] = (nodes, a[..., 0].numpy(), np.zeros((nw, T, 1), np.float32))
    return adj, td


@pytest.mark.parametrize("model_name,T", [("VaDE", 25), ("VQVAE", 25), ("Contrastive", 24)])
def test_train_deepof_model_end_to_end(model_name, T, tmp_path):
    from deepof_b200 import train_deepof_model, load_model_from_ckpt
    adj, train_td = _table_dicts(11, T, 2, 96, seed=1)
    _, val_td = _table_dicts(11, T, 1, 64, seed=9)
    if model_name != "VaDE":
        with pytest.raises(NotImplementedError):
            train_deepof_model((train_td, val_td), adj, None, encoder_type="recurrent", batch_size=64, latent_dim=6, epochs=1,
                               n_clusters=4, model_name=model_name)      # TURTLE teacher on (the reference default): VaDE only
    else:
        # teacher on: latents -> PCA views -> TURTLE -> tau* -> GMM init -> distillation in the main phase
        mv, ms, tinit, ls = train_deepof_model((train_td, val_td), adj, None, encoder_type="recurrent", batch_size=64,
                                               latent_dim=6, epochs=1, n_clusters=4, model_name=model_name, pretrain_epochs=1,
                                               random_seed=3, teacher_outer_steps=4, teacher_inner_steps=3,
                                               teacher_batch_size=64, save_weights=False)
        assert tinit is not None and tinit is not mv
        assert all(np.isfinite(l["total_loss"]) for l in ls["train_logs"]) and len(ls["train_logs"]) == 2
        assert ls["train_logs"][-1]["distill_loss"] > 0.0                 # the main phase saw tau*
        # the teacher-init snapshot holds the moment-matched mixture, the trained model moved on from it
        assert torch.isfinite(tinit.latent_space.gmm_log_vars).all()
        assert float(tinit.latent_space.gmm_log_vars.min()) >= float(np.log(0.01)) - 1e-5      # min_var = 0.01
    with pytest.raises(ValueError):
        train_deepof_model((train_td, val_td), adj, None, device="tpu", use_turtle_teacher=False)
    out = train_deepof_model((train_td, val_td), adj, None, encoder_type="recurrent", batch_size=64, latent_dim=6, epochs=2,
                             output_path=str(tmp_path), n_clusters=4, model_name=model_name, use_turtle_teacher=False,
                             pretrain_epochs=1, random_seed=3)
    model_val, model_score, teacher, log_summary = out                     # the reference's return tuple
    assert teacher is None and model_val is model_score
    logs = log_summary["train_logs"]
    assert len(logs) == (3 if model_name == "VaDE" else 2)
    assert all(np.isfinite(l["total_loss"]) for l in logs) and np.isfinite(log_summary["val_logs"][0]["total_loss"])
    if model_name != "Contrastive":
        assert logs[-1]["reconstruct_loss"] < logs[0]["reconstruct_loss"] + 1e-3     # it trains
    ckpt = os.path.join(str(tmp_path), f"{model_name.lower()}_final.pth")
    assert os.path.exists(ckpt) and os.path.exists(ckpt[:-4] + "_info.txt")
    bundle = torch.load(ckpt, map_location="cpu", weights_only=False)
    assert {"state_dict", "rebuild_spec", "log_summary"} <= set(bundle) and bundle["rebuild_spec"]["model_name"] == model_name.lower()
    m2, _ = load_model_from_ckpt(ckpt, max_batch=64)
    x, a = O.synthetic_windows(32, T if model_name != "Contrastive" else T // 2, adj, seed=4)
    if model_name == "Contrastive":
        assert torch.equal(model_val(x, a), m2(x, a))
    else:
        e1, q1 = model_val.embed(x, a)
        e2, q2 = m2.embed(x, a)
        assert torch.equal(e1, e2) and torch.equal(q1, q2)                 # save / load round trip
    m3, n, n2, ls = train_deepof_model(pretrained=ckpt, batch_size=64)
    assert n is None and n2 is None and ls["model_name"] == model_name.lower()


def test_pipelined_host_loop_equals_step_by_step():
    """trainer.train_steps (H2D of batch i+1 on a side stream while step i computes, non-blocking loss read-back) gives
    the losses and the parameters of the plain per-step host loop."""
    import numpy as np
    from deepof_b200.training import VQVAETrainer
    from oracle import vade_oracle as O
    T, N, D, K, B = 25, 14, 8, 6, 64
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = O.synthetic_windows(5 * B, T, adj, seed=3)
    xh, ah = x.pin_memory(), a.pin_memory()
    batches = [(xh[i * B:(i + 1) * B], ah[i * B:(i + 1) * B]) for i in range(5)]
    t1 = VQVAETrainer((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, seed=4)
    t2 = VQVAETrainer((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, seed=4)
    ref = [t1.train_step(xb, ab) for xb, ab in batches]
    got = t2.train_steps(batches)
    # the weight gradients are reduced with floating-point atomics, so two runs agree to rounding, not bitwise
    assert np.allclose(got, ref, rtol=1e-5, atol=0) and got[0] == ref[0]
    assert float((t1.model.state - t2.model.state).abs().max()) < 1e-5
    assert t2.train_steps([]) == []
