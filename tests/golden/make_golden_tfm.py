"""Golden vectors for the transformer encoder, produced by the UNMODIFIED reference on CPU.
Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_tfm.py

tfm_<case>.npz: models_new.TFMEncoderPT — state_dict AFTER three train-mode forward passes (they create the lazily built
CensNet parameters and move the BatchNorm running statistics off their initial 0 / 1), inputs x, a, and the EVAL-mode
outputs: the last-step outputs of the node / edge transformer cores (recorded by forward hooks), and the encoder output.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from oracle.vade_oracle import default_adjacency, synthetic_windows  # noqa: E402

M, L, T, U = refshim.load()

CASES = {
    "cfg5": dict(T=25, N=14, D=16, B=12, seed=51, zero_from=None),          # key_dim 40, head_dim 10
    "odd": dict(T=24, N=11, D=6, B=9, seed=52, zero_from=None),             # key_dim 32, head_dim 8
    "padded": dict(T=25, N=14, D=8, B=10, seed=53, zero_from=9),            # all-zero frames -> key-padding mask
}


def run(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=5000 + c["seed"])
    if c["zero_from"] is not None:
        x[::2, c["zero_from"]:] = 0.0
        a[::2, c["zero_from"]:] = 0.0
    enc = M.TFMEncoderPT((c["T"], c["N"], 3), (c["T"], E, 1), adj, c["D"])
    enc.train()
    with torch.no_grad():
        for i in range(3):
            xi, ai = synthetic_windows(32, c["T"], adj, seed=6000 + 10 * c["seed"] + i)
            enc(xi, ai)
        # spread the head so that the BatchNorm statistics matter
        for k, v in enc.state_dict().items():
            if k.endswith("running_mean"):
                v.add_(0.05 * torch.randn_like(v))
    enc.eval()
    seen = {}
    hooks = [enc.node_tf.register_forward_hook(lambda m, i, o: seen.__setitem__("nodes", o.detach().clone())),
             enc.edge_tf.register_forward_hook(lambda m, i, o: seen.__setitem__("edges", o.detach().clone()))]
    with torch.no_grad():
        out = enc(x, a)
    for h in hooks:
        h.remove()
    assert torch.isfinite(out).all() and torch.isfinite(seen["nodes"]).all()
    res = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["B"], enc.key_dim, 4, 128, 2], dtype=np.int64),
           "eval/nodes": seen["nodes"].numpy(), "eval/edges": seen["edges"].numpy(), "eval/out": out.numpy()}
    for k, v in enc.state_dict().items():
        res["p/" + k] = v.detach().numpy().copy()
    path = os.path.join(HERE, f"tfm_{name}.npz")
    np.savez_compressed(path, **res)
    print("tfm", name, "%.1f KB" % (os.path.getsize(path) / 1024), "out", float(out.abs().mean()))


MODEL_CASES = {
    # full reference models with encoder_type="transformer" in eval mode: what embedding_per_video reads from them
    "vade": dict(model="vade", T=25, N=14, D=8, K=5, B=10, seed=61),
    "vqvae": dict(model="vqvae", T=24, N=11, D=6, K=7, B=9, seed=62),
    "contrastive": dict(model="contrastive", T=25, N=14, D=8, K=1, B=8, seed=63),
}


def run_model(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    xs, as_ = (c["T"], c["N"], 3), (c["T"], E, 1)
    if c["model"] == "vade":
        model = M.VaDEPT(xs, as_, adj, c["D"], c["K"], encoder_type="transformer")
    elif c["model"] == "vqvae":
        model = M.VQVAEPT(xs, as_, adj, c["D"], c["K"], encoder_type="transformer", use_gnn=True)
    else:
        model = M.ContrastivePT(xs, as_, adj, c["D"], encoder_type="transformer", use_gnn=True)
    model.train()
    Tenc = c["T"] // 2 if c["model"] == "contrastive" else c["T"]     # ContrastivePT encodes half windows (:2013)
    with torch.no_grad():
        for i in range(3):                                   # builds the CensNet parameters, moves the BN statistics
            xi, ai = synthetic_windows(32, Tenc, adj, seed=7000 + 10 * c["seed"] + i)
            model.encoder(xi, ai)
        if c["model"] == "vade":
            model.latent_space.gmm_means.mul_(3.0)
        if c["model"] == "vqvae":
            model.vq_layer.codebook.copy_(0.5 * torch.randn(c["D"], c["K"]))
    model.eval()
    x, a = synthetic_windows(c["B"], Tenc, adj, seed=8000 + c["seed"])
    res = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64), "model": np.array(c["model"])}
    with torch.no_grad():
        if c["model"] == "vade":
            out = model(x, a)                                 # (dist, emb, q, kmeans): model_utils_new.py:585-596
            res["eval/emb"], res["eval/q"] = out[1].numpy(), out[2].numpy()
            res["eval/loc"] = out[0].base_dist.base_dist.loc.numpy()          # TFMDecoderPT(z = z_mean in eval)
        elif c["model"] == "vqvae":
            out = model(x, a, return_all_outputs=True)        # soft counts [3], encoder output [4]
            res["eval/emb"], res["eval/q"] = out[4].numpy(), out[3].numpy()
            res["eval/quant"] = out[2].numpy()
            res["eval/loc_q"] = out[0].base_dist.base_dist.loc.numpy()        # decoder(quantized)
            res["eval/loc"] = out[1].base_dist.base_dist.loc.numpy()          # decoder(encoder output)
        else:
            res["eval/emb"] = model(x, a).numpy()
    for k, v in model.state_dict().items():
        res["p/" + k] = v.detach().numpy().copy()
    path = os.path.join(HERE, f"tfmmodel_{name}.npz")
    np.savez_compressed(path, **res)
    print("tfmmodel", name, "%.1f KB" % (os.path.getsize(path) / 1024))


TRAIN_CASES = {
    # TFMEncoderPT in train(): one forward + backward of sum(out * probe) with every dropout mask recorded
    "train_small": dict(T=12, N=11, D=6, B=7, seed=71),
    "train_cfg5": dict(T=25, N=14, D=16, B=5, seed=72),
}


def run_train(name, c):
    """The reference draws its dropout masks from the global generator inside C++ (scaled_dot_product_attention) and
    nn.Dropout; they are REPLAYED here: re-seed, then draw bernoulli tensors of the same shapes in the same order with
    the same kernel (empty_like().bernoulli_(1 - p), what at::dropout does).  The oracle fed with the replayed masks
    must reproduce the reference's output and gradients — that check runs right here, so a wrong replay cannot be
    committed."""
    from oracle import tfm_oracle as TO
    from oracle import vade_oracle as O
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    enc = M.TFMEncoderPT((c["T"], c["N"], 3), (c["T"], E, 1), adj, c["D"])
    enc.train()
    with torch.no_grad():
        for i in range(2):
            xi, ai = synthetic_windows(16, c["T"], adj, seed=9000 + 10 * c["seed"] + i)
            enc(xi, ai)
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=9500 + c["seed"])
    x[1, c["T"] // 2:] = 0.0
    a[1, c["T"] // 2:] = 0.0
    probe = torch.randn(c["B"], c["D"])
    p0 = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    seed = 9900 + c["seed"]
    torch.manual_seed(seed)
    out = enc(x, a)
    (out * probe).sum().backward()
    # replay
    torch.manual_seed(seed)
    masks = {}
    dk, heads, layers, rate = enc.key_dim, 4, 2, 0.1
    for core, S in (("node", c["B"] * c["N"]), ("edge", c["B"] * E)):
        for nm, shp in TO.dropout_mask_shapes(S, c["T"], dk, heads, layers):
            masks[f"{core}.{nm}"] = torch.empty(shp).bernoulli_(1.0 - rate)
    graph = O.graph_operators(adj)
    leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and k in dict(enc.named_parameters()) else v) for k, v in p0.items()}
    o2, stats = TO.encoder_forward_train(x, a, leaf, graph, masks)
    assert float((o2 - out).abs().max()) < 2e-5, float((o2 - out).abs().max())
    names = [k for k, _ in enc.named_parameters()]
    gl = torch.autograd.grad((o2 * probe).sum(), [leaf[k] for k in names], allow_unused=True)
    for k, g2 in zip(names, gl):
        gr = dict(enc.named_parameters())[k].grad
        assert (gr is None) == (g2 is None), k
        if gr is not None:
            assert float((g2 - gr).abs().max()) <= 2e-4 * max(1.0, float(gr.abs().max())), (k, float((g2 - gr).abs().max()))
    res = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(), "probe": probe.numpy(),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["B"], dk, heads, 128, layers], dtype=np.int64),
           "train/out": out.detach().numpy()}
    for k, v in p0.items():
        res["p/" + k] = v.numpy().copy()
    for k, v in enc.state_dict().items():
        if "running" in k:
            res["p1/" + k] = v.detach().numpy().copy()       # running statistics after the step
    for k, prm in enc.named_parameters():
        if prm.grad is not None:
            res["g/" + k] = prm.grad.detach().numpy().copy()
    for k, m in masks.items():
        res["mask/" + k] = np.packbits(m.numpy().astype(np.uint8).reshape(-1))
        res["mshape/" + k] = np.array(m.shape, dtype=np.int64)
    path = os.path.join(HERE, f"tfmtrain_{name}.npz")
    np.savez_compressed(path, **res)
    print("tfmtrain", name, "%.1f KB" % (os.path.getsize(path) / 1024), "out", float(out.abs().mean()))


VADE_TRAIN_CASES = {
    "vade_main": dict(T=12, N=11, D=6, K=5, B=8, seed=81, phase="main"),
    "vade_pretrain": dict(T=25, N=14, D=8, K=4, B=6, seed=82, phase="pretrain"),
}


def run_vade_train(name, c):
    """One reference step_vade on VaDEPT(encoder_type="transformer"): logs and gradients, with the noise tensors recorded
    (torch.randn / randn_like are wrapped; the wrappers call the originals) and the dropout masks replayed from the
    generator as in run_train.  The oracle's step must reproduce logs and gradients right here."""
    import types
    from make_golden import NoiseTape
    from oracle import tfm_oracle as TO
    from oracle import vade_oracle as O
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    model = M.VaDEPT((c["T"], c["N"], 3), (c["T"], E, 1), adj, c["D"], c["K"], encoder_type="transformer", use_gnn=True, kmeans_loss=1.0)
    model.train()
    with torch.no_grad():
        for i in range(2):
            xi, ai = synthetic_windows(16, c["T"], adj, seed=9100 + 10 * c["seed"] + i)
            model.encoder(xi, ai)
        model.latent_space.gmm_means.mul_(3.0)
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=9600 + c["seed"])
    common = U.CommonFitCfg(latent_dim=c["D"], n_components=c["K"])
    vcfg, tcfg = U.VaDECfg(), U.TurtleTeacherCfg()
    crit = L.VadeLoss(common, vcfg, tcfg)
    nb = 10
    if c["phase"] == "pretrain":
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
    seed = 9950 + c["seed"]
    torch.manual_seed(seed)
    with NoiseTape() as tape:
        res = T.step_vade(model, (x, a, torch.arange(c["B"])), ctx)
    res.loss.backward()
    draws = tape.draws
    # ---- replay the generator: encoder masks, reparameterisation noise, decoder masks, (MC noise)
    torch.manual_seed(seed)
    masks = {}
    dk = model.encoder.key_dim
    for core, S in (("node", c["B"] * c["N"]), ("edge", c["B"] * E)):
        for nm, shp in TO.dropout_mask_shapes(S, c["T"], dk, 4, 2):
            masks[f"{core}.{nm}"] = torch.empty(shp).bernoulli_(0.9)
    eps = torch.randn_like(draws[0])
    assert torch.equal(eps, draws[0]), "the reparameterisation noise is not where the replay expects it"
    for nm, shp in TO.decoder_mask_shapes(c["B"], c["T"], 4 * c["D"], 8, 128, 2):
        masks[nm] = torch.empty(shp).bernoulli_(0.8)
    mc = draws[1] if c["phase"] == "main" else None
    klw = float(sched.get_weight())
    ocfg = (O.LossCfg.main_defaults(c["K"], kl_weight=klw) if c["phase"] == "main" else O.LossCfg.pretrain_defaults(c["K"], kl_weight=klw))
    logs, grads, _ = TO.vade_train_step(x, a, p0, O.graph_operators(adj), ocfg, masks, eps, mc_eps=mc)
    for k, v in res.logs.items():
        assert abs(logs[k] - v) <= 5e-5 * max(1.0, abs(v)), (k, logs[k], v)
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            assert float((grads[k] - prm.grad).abs().max()) <= 5e-4 * max(1.0, float(prm.grad.abs().max())), k
    out = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(), "phase": np.array(c["phase"]), "klw": np.array(klw),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64), "eps": eps.numpy()}
    if mc is not None:
        out["mc_eps"] = mc.numpy()
    for k, v in p0.items():
        out["p/" + k] = v.numpy().copy()
    for k, v in res.logs.items():
        out["log/" + k] = np.array(v, dtype=np.float64)
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            out["g/" + k] = prm.grad.detach().numpy().copy()
    for k, m in masks.items():
        out["mask/" + k] = np.packbits(m.numpy().astype(np.uint8).reshape(-1))
        out["mshape/" + k] = np.array(m.shape, dtype=np.int64)
    path = os.path.join(HERE, f"tfmvade_{name}.npz")
    np.savez_compressed(path, **out)
    print("tfmvade", name, "%.1f KB" % (os.path.getsize(path) / 1024), {k: round(v, 4) for k, v in list(res.logs.items())[:4]})


STEP_CASES = {
    # step_vqvae_distill / step_contrastive_distill (teacher off) on the transformer model family
    "vq_small": dict(model="vqvae", T=12, N=11, D=6, K=7, B=8, seed=91, beta=1.0),
    "vq_cfg3": dict(model="vqvae", T=25, N=14, D=16, K=64, B=6, seed=92, beta=0.25),
    "con_small": dict(model="contrastive", T=24, N=11, D=6, K=1, B=8, seed=93),
    "con_cfg": dict(model="contrastive", T=50, N=14, D=8, K=1, B=6, seed=94),
}


def run_step(name, c):
    """One reference training step of VQVAEPT / ContrastivePT with encoder_type="transformer".  Dropout masks are replayed
    from the generator state captured right before every encoder / decoder call (hooks record torch.get_rng_state());
    the oracle fed with the replayed masks must reproduce logs and gradients before anything is written."""
    import types
    from oracle import tfm_oracle as TO
    from oracle import vade_oracle as O
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    rows, cols = np.nonzero(np.triu(adj))
    E = len(rows)
    xs, as_ = (c["T"], c["N"], 3), (c["T"], E, 1)
    con = c["model"] == "contrastive"
    if con:
        model = M.ContrastivePT(xs, as_, adj, c["D"], encoder_type="transformer", use_gnn=True, temperature=0.1)
    else:
        model = M.VQVAEPT(xs, as_, adj, c["D"], c["K"], encoder_type="transformer", use_gnn=True, kmeans_loss=0.0, beta=c["beta"])
    Tenc = c["T"] // 2 if con else c["T"]
    model.train()
    with torch.no_grad():
        for i in range(2):
            xi, ai = synthetic_windows(16, Tenc, adj, seed=9200 + 10 * c["seed"] + i)
            model.encoder(xi, ai)
        if not con:
            xi, ai = synthetic_windows(c["B"], Tenc, adj, seed=9700 + c["seed"])
            model.eval()
            e0 = model.encoder(xi, ai)
            model.train()
            model.vq_layer.codebook.copy_(e0.mean(0, keepdim=True).t() + e0.std() * 1.5 * torch.randn(c["D"], c["K"]))
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=9700 + c["seed"])
    p0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    states = []
    hooks = [model.encoder.register_forward_pre_hook(lambda m, i: states.append(("enc", torch.get_rng_state(), [t.detach().clone() for t in i])))]
    if not con:
        hooks.append(model.decoder.register_forward_pre_hook(lambda m, i: states.append(("dec", torch.get_rng_state(), None))))
    seed = 9960 + c["seed"]
    torch.manual_seed(seed)
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
    for h in hooks:
        h.remove()
    dk = model.encoder.key_dim
    graph = O.graph_operators(adj)

    def enc_masks(state, B):
        torch.set_rng_state(state)
        mk = {}
        for core, S in (("node", B * c["N"]), ("edge", B * E)):
            for nm, shp in TO.dropout_mask_shapes(S, Tenc, dk, 4, 2):
                mk[f"{core}.{nm}"] = torch.empty(shp).bernoulli_(0.9)
        return mk

    def dec_masks(state, prefix):
        torch.set_rng_state(state)
        return {nm.replace("dec.", prefix): torch.empty(shp).bernoulli_(0.8)
                for nm, shp in TO.decoder_mask_shapes(c["B"], c["T"], 4 * c["D"], 8, 128, 2)}

    out = {"adjacency": adj, "model": np.array(c["model"]), "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64)}
    if con:
        assert [s[0] for s in states] == ["enc", "enc"], [s[0] for s in states]
        (xv, av), (xav, aav) = states[0][2], states[1][2]
        m0, m1 = enc_masks(states[0][1], c["B"]), enc_masks(states[1][1], c["B"])
        logs, grads, _ = TO.contrastive_views_step(xv, av, xav, aav, p0, graph, m0, m1, 0.1)
        out.update({"x": xv.numpy(), "a": av.numpy(), "x_aug": xav.numpy(), "a_aug": aav.numpy()})
        masks = {**m0, **{"aug." + k: v for k, v in m1.items()}}
    else:
        assert [s[0] for s in states] == ["enc", "dec", "dec"], [s[0] for s in states]
        masks = {**enc_masks(states[0][1], c["B"]), **dec_masks(states[1][1], "dec."), **dec_masks(states[2][1], "dec1.")}
        logs, grads, _ = TO.vqvae_train_step(x, a, p0, graph, masks, c["beta"], 0.0)
        out.update({"x": x.numpy(), "a": a.numpy(), "beta": np.array(c["beta"])})
    for k, v in res.logs.items():
        assert abs(logs[k] - v) <= 5e-5 * max(1.0, abs(v)), (k, logs[k], v)
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            assert float((grads[k] - prm.grad).abs().max()) <= 5e-4 * max(1.0, float(prm.grad.abs().max())), k
    for k, v in p0.items():
        out["p/" + k] = v.numpy().copy()
    for k, v in model.state_dict().items():
        if "running" in k:
            out["p1/" + k] = v.detach().numpy().copy()
    for k, v in res.logs.items():
        out["log/" + k] = np.array(v, dtype=np.float64)
    for k, prm in model.named_parameters():
        if prm.grad is not None:
            out["g/" + k] = prm.grad.detach().numpy().copy()
    for k, m in masks.items():
        out["mask/" + k] = np.packbits(m.numpy().astype(np.uint8).reshape(-1))
        out["mshape/" + k] = np.array(m.shape, dtype=np.int64)
    path = os.path.join(HERE, f"tfmstep_{name}.npz")
    np.savez_compressed(path, **out)
    print("tfmstep", name, "%.1f KB" % (os.path.getsize(path) / 1024), {k: round(v, 4) for k, v in list(res.logs.items())[:4]})


if __name__ == "__main__":
    only = sys.argv[1:]
    sys.path.insert(0, HERE)
    for name, c in STEP_CASES.items():
        if not only or name in only:
            run_step(name, c)
    for name, c in VADE_TRAIN_CASES.items():
        if not only or name in only:
            run_vade_train(name, c)
    for name, c in TRAIN_CASES.items():
        if not only or name in only:
            run_train(name, c)
    for name, c in CASES.items():
        if not only or name in only:
            run(name, c)
    for name, c in MODEL_CASES.items():
        if not only or ("model_" + name) in only:
            run_model(name, c)
