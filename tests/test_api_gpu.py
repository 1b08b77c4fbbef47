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


@pytest.mark.parametrize("model_name,T,enc", [("VaDE", 25, "recurrent"), ("VaDE", 12, "transformer"), ("VQVAE", 25, "recurrent"),
                                              ("VQVAE", 12, "transformer"), ("Contrastive", 24, "recurrent"),
                                              ("Contrastive", 24, "transformer"), ("VaDE", 12, "TCN"), ("VQVAE", 12, "TCN"),
                                              ("Contrastive", 24, "TCN")])
def test_train_deepof_model_end_to_end(model_name, T, enc, tmp_path):
    """The reference's run for the three model kinds and the three encoder families, TURTLE teacher ON (the reference default):
    per-epoch validation + diagnostics, log_summary in the reference's structure, best-val / best-score checkpoints under
    the reference's paths, the (model_val, model_score, teacher_init_model, log_summary) tuple, checkpoint round trip."""
    from deepof_b200 import train_deepof_model, load_model_from_ckpt
    name = model_name.lower()
    adj, train_td = _table_dicts(11, T, 2, 96, seed=1)
    _, val_td = _table_dicts(11, T, 1, 64, seed=9)
    with pytest.raises(ValueError):
        train_deepof_model((train_td, val_td), adj, None, device="tpu", batch_size=64, latent_dim=6, epochs=1)
    with pytest.raises(NotImplementedError):
        train_deepof_model((train_td, val_td), adj, None, encoder_type="LSTM", batch_size=64, latent_dim=6, epochs=1)
    with pytest.raises(NotImplementedError):
        train_deepof_model((train_td, val_td), adj, None, encoder_type=enc, batch_size=64, latent_dim=6, epochs=1, use_amp=True)
    epochs = 6
    out = train_deepof_model((train_td, val_td), adj, None, encoder_type=enc, batch_size=64, latent_dim=6, epochs=epochs,
                             output_path=str(tmp_path), n_clusters=4, model_name=model_name, pretrain_epochs=1, random_seed=3,
                             teacher_outer_steps=4, teacher_inner_steps=3, teacher_batch_size=64, run=2)
    model_val, model_score, tinit, ls = out
    assert model_val is not model_score
    assert (tinit is not None) == (name == "vade")
    # log summary: the reference's structure (logging.py:304-351), one entry per epoch
    # (the reference's _update_log_summary overwrites every top-level key that is not train / val with logs.get(key, nan):
    #  "model_type" is NaN after the first epoch there as well, logging.py:340-345)
    assert ls["model_type"] != ls["model_type"] and set(ls["train"]) == set(ls["val"])
    for split in ("train", "val"):
        assert len(ls[split]["total_loss"]) == epochs and all(np.isfinite(v) for v in ls[split]["total_loss"])
    assert all(np.isfinite(v) and 0.0 <= v <= 1.0 for v in ls["val"]["alignment_score"])       # teacher on -> diagnostics every epoch
    assert all(v > 0.0 for v in ls["train"]["distill_loss"][:1])                                # the step saw tau*
    if name == "contrastive":
        assert all(np.isfinite(v) for v in ls["train"]["pos_similarity"])
    # checkpoints under <output>/models/<model>/run_<run>/ (model_utils_new.py:368-374)
    d = os.path.join(str(tmp_path), "models", name, "run_2")
    best_val = os.path.join(d, "best_model_val.pth")
    assert os.path.exists(best_val) and os.path.exists(best_val[:-4] + "_info.txt")
    assert os.path.exists(os.path.join(d, "model_teacher_init.pth")) == (name == "vade")
    bundle = torch.load(best_val, map_location="cpu", weights_only=False)
    assert {"state_dict", "rebuild_spec", "log_summary"} <= set(bundle) and bundle["rebuild_spec"]["model_name"] == name
    assert bundle["rebuild_spec"]["encoder_type"] == enc
    m2, _ = load_model_from_ckpt(best_val, max_batch=64)
    x, a = O.synthetic_windows(32, T if name != "contrastive" else T // 2, adj, seed=4)
    if name == "contrastive":
        assert torch.equal(model_val(x, a), m2(x, a))
    else:
        e1, q1 = model_val.embed(x, a)
        e2, q2 = m2.embed(x, a)
        assert torch.equal(e1, e2) and torch.equal(q1, q2)                 # model_val IS the best-val checkpoint
    if name == "vade":
        assert torch.isfinite(tinit.latent_space.gmm_log_vars).all()
        assert float(tinit.latent_space.gmm_log_vars.min()) >= float(np.log(0.01)) - 1e-5      # min_var = 0.01
    m3, n, n2, ls3 = train_deepof_model(pretrained=best_val, batch_size=64)
    assert n is None and n2 is None and set(ls3) == set(ls)


def test_vade_without_teacher_initialises_the_gmm_from_data(tmp_path):
    """use_turtle_teacher=False: the mixture is fitted on the pretrained embeddings (initialize_gmm_from_data,
    models_new.py:1907-1947) before the main phase — epochs=0 returns exactly that state."""
    from deepof_b200 import train_deepof_model
    from deepof_b200.gmm_init import gmm_from_embeddings
    adj, train_td = _table_dicts(11, 25, 2, 96, seed=1)
    _, val_td = _table_dicts(11, 25, 1, 64, seed=9)
    mv, ms, tinit, ls = train_deepof_model((train_td, val_td), adj, None, encoder_type="recurrent", batch_size=64, latent_dim=6,
                                           epochs=0, output_path=str(tmp_path), n_clusters=4, model_name="VaDE", pretrain_epochs=1,
                                           random_seed=3, use_turtle_teacher=False, save_weights=False)
    assert tinit is None
    from deepof_b200.api import windows_from_table_dict
    x, a = windows_from_table_dict(train_td, "cuda")
    z = mv.embed(x, a)[0].cpu().numpy()
    np.random.seed(3)
    # numpy's global generator has been consumed identically up to the fit: batch shuffles use their own default_rng
    means, log_vars = gmm_from_embeddings(z, 4)
    got = mv.latent_space.gmm_means.cpu().numpy()
    assert np.isfinite(got).all() and np.allclose(np.sort(got, 0), np.sort(means.astype(np.float32), 0), atol=5e-2)


def test_train_from_frame_tables_equals_materialised_windows(tmp_path):
    """train_deepof_model over WindowLoaders (raw frames resident, windows built per batch by the loader kernel)
    gives the run it gives over the same windows materialised up front."""
    from deepof_b200 import WindowLoader, train_deepof_model
    from deepof_b200.api import WindowSource
    N, T = 11, 25
    adj = O.default_adjacency(N)
    rows, cols = np.nonzero(np.triu(adj))
    edges = np.stack([rows, cols], 1).astype(np.int32)
    g = torch.Generator().manual_seed(5)

    def video(n):
        centre = torch.cumsum(torch.randn(n, 1, 2, generator=g) * 1.5, 0) + 300.0
        body = torch.randn(1, N, 2, generator=g) * 18.0
        return (centre + body + torch.randn(n, N, 2, generator=g) * 0.4).float()

    tr = WindowLoader([video(150), video(120)], edges, T, 1, nose=0, tail_base=2, center_node=1, align_node=0, fps=25.0)
    va = WindowLoader([video(90)], edges, T, 1, nose=0, tail_base=2, center_node=1, align_node=0, fps=25.0,
                      global_scalers=tr.global_scalers)
    kw = dict(encoder_type="recurrent", batch_size=64, latent_dim=6, epochs=2, n_clusters=4, model_name="VQVAE", random_seed=3,
              use_turtle_teacher=False, save_weights=False)
    _, _, _, ls_f = train_deepof_model((tr, va), adj, None, output_path=str(tmp_path / "f"), **kw)
    xm, am = tr.load(0, len(tr))
    xv, av = va.load(0, len(va))
    src_t, src_v = WindowSource((xm.clone(), am.clone()), 64, 3), WindowSource((xv.clone(), av.clone()), 64, 3, shuffle=False)
    _, _, _, ls_m = train_deepof_model((src_t, src_v), adj, None, output_path=str(tmp_path / "m"), **kw)
    assert len(ls_f["train"]["total_loss"]) == 2
    for split in ("train", "val"):
        assert np.allclose(ls_f[split]["total_loss"], ls_m[split]["total_loss"], rtol=2e-5, atol=0), (ls_f[split], ls_m[split])


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
