"""Golden vectors for the TCN model family, produced by the UNMODIFIED reference on CPU.
Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_tcn.py

tcnmodel_<case>.npz : VaDEPT / VQVAEPT / ContrastivePT built with encoder_type="TCN" — state_dict after three train-mode
                      encoder passes (CensNet parameters exist, running statistics moved) and the EVAL-mode outputs.
tcnvade_<case>.npz  : one step_vade (main / pretrain) in train(): 13 logs, every gradient, the noise drawn, and the
                      BatchNorm running buffers after the step ("p1/").
tcnstep_<case>.npz  : one step_vqvae_distill / step_contrastive_distill (teacher off).
The TCN stacks carry no dropout (dropout_rate = 0 in every caller), so nothing but the noise has to be recorded.  Every
step case asserts that oracle/tcn_oracle.py reproduces the reference's logs, gradients and running buffers before the
file is written.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import refshim  # noqa: E402
from oracle import tcn_oracle as TC  # noqa: E402
from oracle import vade_oracle as O  # noqa: E402
from oracle.vade_oracle import default_adjacency, synthetic_windows  # noqa: E402

M, L, T, U = refshim.load()


def build(c, adj, E):
    xs, as_ = (c["T"], c["N"], 3), (c["T"], E, 1)
    if c["model"] == "vade":
        return M.VaDEPT(xs, as_, adj, c["D"], c["K"], encoder_type="TCN", use_gnn=True, kmeans_loss=1.0)
    if c["model"] == "vqvae":
        return M.VQVAEPT(xs, as_, adj, c["D"], c["K"], encoder_type="TCN", use_gnn=True, kmeans_loss=0.0, beta=c.get("beta", 1.0))
    return M.ContrastivePT(xs, as_, adj, c["D"], encoder_type="TCN", use_gnn=True, temperature=0.1)


def warm(model, c, adj, Tenc, seed0, n=3, B=24):
    """Train-mode passes: create the lazily built CensNet parameters, move every running statistic off 0 / 1; the decoder's
    BatchNorm layers get theirs from two decoder passes.  Then spread the BatchNorm scales / shifts so that they matter."""
    model.train()
    with torch.no_grad():
        for i in range(n):
            xi, ai = synthetic_windows(B, Tenc, adj, seed=seed0 + i)
            e = model.encoder(xi, ai)
            if hasattr(model, "decoder"):
                xf, _ = synthetic_windows(B, c["T"], adj, seed=seed0 + 50 + i)
                model.decoder(e, xf)
        for k, v in model.named_parameters():
            if (".bn" in k or "head.2." in k or "head.5." in k):
                v.add_(0.1 * torch.randn_like(v))


MODEL_CASES = {
    "vade": dict(model="vade", T=25, N=14, D=8, K=5, B=10, seed=161),
    "vqvae": dict(model="vqvae", T=24, N=11, D=6, K=7, B=9, seed=162),
    "contrastive": dict(model="contrastive", T=25, N=14, D=8, K=1, B=8, seed=163),
}


def run_model(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    model = build(c, adj, E)
    Tenc = c["T"] // 2 if c["model"] == "contrastive" else c["T"]
    warm(model, c, adj, Tenc, 17000 + 10 * c["seed"])
    with torch.no_grad():
        if c["model"] == "vade":
            model.latent_space.gmm_means.mul_(3.0)
        if c["model"] == "vqvae":
            model.vq_layer.codebook.copy_(0.5 * torch.randn(c["D"], c["K"]))
    model.eval()
    x, a = synthetic_windows(c["B"], Tenc, adj, seed=18000 + c["seed"])
    x[1, Tenc // 2:] = 0.0
    a[1, Tenc // 2:] = 0.0
    res = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64), "model": np.array(c["model"])}
    with torch.no_grad():
        res["eval/enc"] = model.encoder(x, a).numpy()
        if c["model"] == "vade":
            out = model(x, a)
            res["eval/emb"], res["eval/q"] = out[1].numpy(), out[2].numpy()
            res["eval/loc"] = out[0].base_dist.base_dist.loc.numpy()
        elif c["model"] == "vqvae":
            out = model(x, a, return_all_outputs=True)
            res["eval/emb"], res["eval/q"] = out[4].numpy(), out[3].numpy()
            res["eval/quant"] = out[2].numpy()
            res["eval/loc_q"] = out[0].base_dist.base_dist.loc.numpy()
            res["eval/loc"] = out[1].base_dist.base_dist.loc.numpy()
        else:
            res["eval/emb"] = model(x, a).numpy()
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    o = TC.model_forward_eval(c["model"], x, a, p0, O.graph_operators(adj))
    assert float((o["enc"] - torch.from_numpy(res["eval/enc"])).abs().max()) < 2e-5
    if c["model"] == "vade":
        assert float((o["loc"] - torch.from_numpy(res["eval/loc"])).abs().max()) < 5e-5
        assert float((o["q"] - torch.from_numpy(res["eval/q"])).abs().max()) < 2e-5
    if c["model"] == "vqvae":
        assert float((o["loc"] - torch.from_numpy(res["eval/loc_q"])).abs().max()) < 5e-5
    for k, v in p0.items():
        res["p/" + k] = v.numpy().copy()
    path = os.path.join(HERE, f"tcnmodel_{name}.npz")
    np.savez_compressed(path, **res)
    print("tcnmodel", name, "%.1f KB" % (os.path.getsize(path) / 1024), "enc", float(np.abs(res["eval/enc"]).mean()))


def check_and_pack(out, model, p0, res_logs, logs, grads, bn_stats, path, tag):
    for k, v in res_logs.items():
        assert abs(logs[k] - v) <= 5e-5 * max(1.0, abs(v)), (k, logs[k], v)
    for k, prm in model.named_parameters():
        assert (prm.grad is None) == (grads.get(k) is None), k
        if prm.grad is not None:
            assert float((grads[k] - prm.grad).abs().max()) <= 5e-4 * max(1.0, float(prm.grad.abs().max())), (k, float((grads[k] - prm.grad).abs().max()))
    run = TC.running_after(p0, bn_stats)
    sd = model.state_dict()
    for k, v in run.items():
        assert float((v.double() - sd[k].double()).abs().max()) <= 2e-5 * max(1.0, float(sd[k].double().abs().max())), k
    for k, v in p0.items():
        out["p/" + k] = v.numpy().copy()
    for k, v in sd.items():
        if "running" in k or "num_batches" in k:
            out["p1/" + k] = v.detach().numpy().copy()
    for k, v in res_logs.items():
        out["log/" + k] = np.array(v, dtype=np.float64)
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            out["g/" + k] = prm.grad.detach().numpy().copy()
    np.savez_compressed(path, **out)
    print(tag, "%.1f KB" % (os.path.getsize(path) / 1024), {k: round(v, 4) for k, v in list(res_logs.items())[:4]})


VADE_CASES = {
    "main": dict(model="vade", T=12, N=11, D=6, K=5, B=8, seed=181, phase="main"),
    "pretrain": dict(model="vade", T=25, N=14, D=8, K=4, B=6, seed=182, phase="pretrain"),
    "cfg2": dict(model="vade", T=25, N=14, D=16, K=8, B=16, seed=183, phase="main"),      # 4 D = 64 = conv_filters: no downsample
}


def run_vade(name, c):
    from make_golden import NoiseTape
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    model = build(c, adj, E)
    warm(model, c, adj, c["T"], 19100 + 10 * c["seed"], n=2, B=16)
    with torch.no_grad():
        model.latent_space.gmm_means.mul_(3.0)
    model.train()
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=19600 + c["seed"])
    common = U.CommonFitCfg(latent_dim=c["D"], n_components=c["K"])
    vcfg, tcfg = U.VaDECfg(), U.TurtleTeacherCfg()
    crit = L.VadeLoss(common, vcfg, tcfg)
    nb = 10
    if c["phase"] == "pretrain":
        model.set_pretrain_mode(True)
        crit.set_mode("pretrain")
        sched = L.Dynamic_weight_manager(nb, mode=vcfg.kl_annealing_mode_pretrain, warmup_epochs=vcfg.kl_warmup_pretrain,
                                         max_weight=vcfg.kl_max_weight_pretrain, cooldown_epochs=vcfg.kl_cooldown_pretrain,
                                         end_weight=vcfg.kl_end_weight_pretrain)
        crit.set_kl_scheduler(sched)
        sched.current_iteration = 90
    else:
        model.set_pretrain_mode(False)
        crit.set_mode("main")
        sched = L.Dynamic_weight_manager(nb, mode=vcfg.kl_annealing_mode, warmup_epochs=vcfg.kl_warmup, max_weight=vcfg.kl_max_weight,
                                         cooldown_epochs=vcfg.kl_cooldown, end_weight=vcfg.kl_end_weight)
        crit.set_kl_scheduler(sched)
        sched.current_iteration = 30
    crit.train()
    ctx = types.SimpleNamespace(criterion=crit, apply_distill=False, train=True)
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    torch.manual_seed(19950 + c["seed"])
    with NoiseTape() as tape:
        res = T.step_vade(model, (x, a, torch.arange(c["B"])), ctx)
    res.loss.backward()
    draws = tape.draws
    eps = draws[0]
    mc = draws[1] if c["phase"] == "main" else None
    klw = float(sched.get_weight())
    ocfg = (O.LossCfg.main_defaults(c["K"], kl_weight=klw) if c["phase"] == "main" else O.LossCfg.pretrain_defaults(c["K"], kl_weight=klw))
    logs, grads, oo = TC.vade_train_step(x, a, p0, O.graph_operators(adj), ocfg, eps, mc_eps=mc)
    out = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(), "phase": np.array(c["phase"]), "klw": np.array(klw),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64), "eps": eps.numpy()}
    if mc is not None:
        out["mc_eps"] = mc.numpy()
    check_and_pack(out, model, p0, res.logs, logs, grads, oo["bn"], os.path.join(HERE, f"tcnvade_{name}.npz"), "tcnvade " + name)


STEP_CASES = {
    "vq_small": dict(model="vqvae", T=12, N=11, D=6, K=7, B=8, seed=191, beta=1.0),
    "vq_cfg3": dict(model="vqvae", T=25, N=14, D=16, K=64, B=16, seed=196, beta=0.25),
    "con_small": dict(model="contrastive", T=24, N=11, D=6, K=1, B=8, seed=193),
    "con_cfg": dict(model="contrastive", T=50, N=14, D=8, K=1, B=6, seed=194),
}


def run_step(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    rows, cols = np.nonzero(np.triu(adj))
    E = len(rows)
    con = c["model"] == "contrastive"
    model = build(c, adj, E)
    Tenc = c["T"] // 2 if con else c["T"]
    warm(model, c, adj, Tenc, 19200 + 10 * c["seed"], n=2, B=16)
    if not con:
        with torch.no_grad():
            xi, ai = synthetic_windows(c["B"], Tenc, adj, seed=19700 + c["seed"])
            model.eval()
            e0 = model.encoder(xi, ai)
            model.vq_layer.codebook.copy_(e0.mean(0, keepdim=True).t() + e0.std() * 1.5 * torch.randn(c["D"], c["K"]))
    model.train()
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=19700 + c["seed"])
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    seen = []
    hook = model.encoder.register_forward_pre_hook(lambda m, i: seen.append([t.detach().clone() for t in i]))
    torch.manual_seed(19960 + c["seed"])
    if con:
        names = [f"B_n{i}" for i in range(c["N"])]
        meta = {"node_columns": [(n, "x") for n in names] + [(n, "y") for n in names] + names,
                "edge_columns": [(names[i], names[j]) for i, j in zip(rows, cols)]}
        eg, el, _ = T._build_edge_from_metainfo(meta, torch.device("cpu"), c["N"])
        rot = T.build_rotation_precomp(edge_index=el, n_nodes=c["N"], device=torch.device("cpu"))
        ccfg = U.ContrastiveCfg()
        ccfg.aug_p_interp = 0.6
        ctx = types.SimpleNamespace(apply_distill=False, edge_index=eg, edge_index_local=el, contrastive_cfg=ccfg, rot_precomp=rot)
        res = T.step_contrastive_distill(model, (x, a, torch.arange(c["B"])), ctx)
    else:
        ctx = types.SimpleNamespace(apply_distill=False)
        res = T.step_vqvae_distill(model, (x, a, torch.arange(c["B"])), ctx)
    res.loss.backward()
    hook.remove()
    graph = O.graph_operators(adj)
    out = {"adjacency": adj, "model": np.array(c["model"]), "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64)}
    if con:
        assert len(seen) == 2
        (xv, av), (xav, aav) = seen
        logs, grads, oo = TC.contrastive_views_step(xv, av, xav, aav, p0, graph, 0.1)
        out.update({"x": xv.numpy(), "a": av.numpy(), "x_aug": xav.numpy(), "a_aug": aav.numpy()})
    else:
        logs, grads, oo = TC.vqvae_train_step(x, a, p0, graph, c["beta"], 0.0)
        out.update({"x": x.numpy(), "a": a.numpy(), "beta": np.array(c["beta"])})
    check_and_pack(out, model, p0, res.logs, logs, grads, oo["bn"], os.path.join(HERE, f"tcnstep_{name}.npz"), "tcnstep " + name)


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in MODEL_CASES.items():
        if not only or ("model_" + name) in only:
            run_model(name, c)
    for name, c in VADE_CASES.items():
        if not only or ("vade_" + name) in only:
            run_vade(name, c)
    for name, c in STEP_CASES.items():
        if not only or name in only:
            run_step(name, c)
