"""Golden vectors for the VQ-VAE and contrastive paths (recurrent encoder), produced by the UNMODIFIED
reference on CPU.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_models.py

vqvae_<case>.npz       models_new.VQVAEPT + training.step_vqvae_distill (teacher off): state_dict, x, a, eval
                       outputs (encoder output, soft counts, code indices, quantized latents, both decoder means),
                       two training steps (logs of each, raw gradient of the first, parameters after the second;
                       clip_grad_value_(0.75) + losses.build_optimizer_generic = Adam lr, weight_decay 1e-4).
contrastive_<case>.npz models_new.ContrastivePT + training.step_contrastive_distill (teacher off): state_dict,
                       x_full, edge_index, the augmentation config and the seed of torch's global generator set
                       right before each step; the four tensors the reference fed its encoder (x, a, x_aug, a_aug,
                       recorded by wrapping model.forward — the wrapper calls the original and keeps copies), z,
                       z_aug, logs, raw gradient of step 1, parameters after step 2.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from oracle.vade_oracle import default_adjacency, synthetic_windows  # noqa: E402

M, L, T, U = refshim.load()

VQ_CASES = {
    # distill: teacher on — ctx.distill_head (DiscriminativeHead), tau_star, lambda_scheduler (training.py:341-370)
    "distill": dict(T=25, N=14, D=8, K=6, B=16, seed=24, beta=1.0, kmeans=0.0,
                    distill=dict(Kt=5, lam=0.7, T=0.5, conf_weight=False, thr=0.6)),
    "distill_conf": dict(T=24, N=11, D=6, K=5, B=9, seed=25, beta=1.0, kmeans=0.0,
                         distill=dict(Kt=4, lam=1.3, T=0.0, conf_weight=True, thr=0.3)),
    "cfg3r": dict(T=25, N=14, D=16, K=64, B=12, seed=21, beta=1.0, kmeans=0.0),
    "small_kmeans": dict(T=25, N=14, D=8, K=6, B=16, seed=22, beta=0.25, kmeans=1.0),
    "odd": dict(T=24, N=11, D=6, K=5, B=9, seed=23, beta=1.0, kmeans=0.0),
}
CON_CASES = {
    # T is the FULL window; the encoder sees T // 2
    "cfg4": dict(T=50, N=22, D=16, B=12, seed=31, aug=dict(p_rot=0.7, p_noise=1.0, p_interp=0.6, n_rot=3)),
    "defaults": dict(T=50, N=14, D=8, B=16, seed=32, aug=dict()),
    "odd": dict(T=24, N=11, D=6, B=9, seed=33, aug=dict(p_rot=1.0, p_noise=0.5, p_interp=1.0, max_shift=3)),
    "dcl": dict(T=50, N=14, D=8, B=16, seed=34, aug=dict(p_interp=0.6), loss="dcl"),
    "hard": dict(T=50, N=14, D=8, B=16, seed=35, aug=dict(p_interp=0.6), loss="hard_dcl"),
    # p_noise=1: with the euclidean similarity a window that draws no augmentation has distance 0 to its own view and
    # the reference's sqrt backward turns every gradient into NaN (seen with p_noise=0.5, seed 36, step 2)
    "distill": dict(T=50, N=14, D=8, B=16, seed=38, aug=dict(p_interp=0.6),
                    distill=dict(Kt=5, lam=0.9, T=0.5, conf_weight=True, thr=0.25)),
    "fc": dict(T=50, N=14, D=8, B=16, seed=39, aug=dict(p_interp=0.6), loss="fc"),
    "fc_euclid": dict(T=24, N=11, D=6, B=9, seed=40, aug=dict(p_interp=0.6, p_noise=1.0, max_shift=3), loss="fc", sim="euclidean"),
    "euclid": dict(T=50, N=14, D=8, B=16, seed=36, aug=dict(p_interp=0.6, p_noise=1.0), loss="nce", sim="euclidean"),
    "dot_dcl": dict(T=24, N=11, D=6, B=9, seed=37, aug=dict(p_interp=0.6), loss="dcl", sim="dot"),
}


class _Lambda:
    """stands in for the reference's lambda scheduler: only get_weight() is read by the step functions"""

    def __init__(self, w):
        self.w = float(w)

    def get_weight(self):
        return self.w


def distill_ctx(c, out, n_windows, seed):
    """ctx fields of the teacher-on step + the head module; everything the tests need goes into `out`."""
    d = c.get("distill")
    if d is None:
        return {}, None
    import deepof.clustering.teacher_model as TM
    g = torch.Generator().manual_seed(9000 + seed)
    tau = torch.softmax(2.0 * torch.randn(n_windows, d["Kt"], generator=g), dim=-1)
    head = TM.DiscriminativeHead(c["D"], d["Kt"])
    out["distill/tau_star"] = tau.numpy()
    out["distill/meta"] = np.array([d["Kt"], d["lam"], d["T"], float(d["conf_weight"]), d["thr"]], dtype=np.float64)
    for k, v in head.state_dict().items():
        out["distill/p/" + k] = v.detach().numpy().copy()
    ctx = dict(apply_distill=True, distill_head=head, tau_star=tau, lambda_scheduler=_Lambda(d["lam"]), distill_sharpen_T=d["T"],
               distill_conf_weight=d["conf_weight"], distill_conf_thresh=d["thr"])
    return ctx, head


def two_animal_adjacency(n_half):
    """Two copies of the chain graph + one cross-animal edge (Nose-Nose), like a 2-animal project."""
    a = default_adjacency(n_half)
    A = np.zeros((2 * n_half, 2 * n_half))
    A[:n_half, :n_half] = a
    A[n_half:, n_half:] = a
    A[0, n_half] = A[n_half, 0] = 1.0
    return A


def run_vq(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=2000 + c["seed"])
    model = M.VQVAEPT((c["T"], c["N"], 3), (c["T"], E, 1), adj, c["D"], c["K"], encoder_type="recurrent",
                      use_gnn=True, kmeans_loss=c["kmeans"], beta=c["beta"])
    with torch.no_grad():     # spread the codebook over the range of the encoder output so several codes are used
        enc0 = model.encoder(x, a)
        model.vq_layer.codebook.copy_(enc0.mean(0, keepdim=True).t() + enc0.std() * 1.5 * torch.randn(c["D"], c["K"]))
    out = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64),
           "beta": np.array(c["beta"]), "kmeans": np.array(c["kmeans"])}
    for k, v in model.state_dict().items():
        out["p/" + k] = v.detach().numpy().copy()
    model.eval()
    with torch.no_grad():
        enc_rec, rec, quant, soft, enc, _ = model(x, a, return_losses=True, return_all_outputs=True)
        out["eval/enc"] = enc.numpy()
        out["eval/soft"] = soft.numpy()
        out["eval/idx"] = model.vq_layer.get_code_indices(enc).numpy()
        out["eval/quant"] = quant.numpy()
        out["eval/loc_q"] = enc_rec.base_dist.base_dist.loc.numpy()
        out["eval/loc_e"] = rec.base_dist.base_dist.loc.numpy()
    lr = 1e-3
    dctx, head = distill_ctx(c, out, c["B"] + 5, c["seed"])
    opt = L.build_optimizer_generic(model, head, base_lr=lr, weight_decay=1e-4)
    out["lr"] = np.array(lr)
    model.train()
    ctx = types.SimpleNamespace(**(dctx or dict(apply_distill=False)))
    idx = torch.arange(c["B"]) if head is None else torch.randperm(c["B"] + 5)[:c["B"]]
    out["idx"] = idx.numpy()
    for step in range(2):
        res = T.step_vqvae_distill(model, (x, a, idx), ctx)
        opt.zero_grad(set_to_none=True)
        res.loss.backward()
        for k, v in res.logs.items():
            out[f"s{step}/log/{k}"] = np.array(v, dtype=np.float64)
        if step == 0:
            for k, prm in model.named_parameters():
                if prm.grad is not None:
                    out["g/" + k] = prm.grad.detach().numpy().copy()
            if head is not None:
                for k, prm in head.named_parameters():
                    out["distill/g/" + k] = prm.grad.detach().numpy().copy()
        torch.nn.utils.clip_grad_value_(model.parameters(), 0.75)
        opt.step()
    for k, v in model.state_dict().items():
        out["p2/" + k] = v.detach().numpy().copy()
    if head is not None:
        for k, v in head.state_dict().items():
            out["distill/p2/" + k] = v.detach().numpy().copy()
    path = os.path.join(HERE, f"vqvae_{name}.npz")
    np.savez_compressed(path, **out)
    print("vqvae", name, "%.1f KB" % (os.path.getsize(path) / 1024), {k: round(v, 5) for k, v in res.logs.items()})


def run_con(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    N = c["N"]
    adj = two_animal_adjacency(N // 2) if name == "cfg4" else default_adjacency(N)
    rows, cols = np.nonzero(np.triu(adj))
    E = len(rows)
    x_full, a_full = synthetic_windows(c["B"], c["T"], adj, seed=3000 + c["seed"])
    model = M.ContrastivePT((c["T"], N, 3), (c["T"], E, 1), adj, c["D"], encoder_type="recurrent", use_gnn=True,
                            temperature=0.1, similarity_function=c.get("sim", "cosine"), loss_function=c.get("loss", "nce"))
    names = ([f"B_n{i}" for i in range(N // 2)] + [f"W_n{i}" for i in range(N - N // 2)]) if name == "cfg4" \
        else [f"B_n{i}" for i in range(N)]
    meta = {"node_columns": [(n, "x") for n in names] + [(n, "y") for n in names] + names,
            "edge_columns": [(names[i], names[j]) for i, j in zip(rows, cols)]}
    eg, el, _ = T._build_edge_from_metainfo(meta, torch.device("cpu"), N)
    rot = T.build_rotation_precomp(edge_index=el, n_nodes=N, device=torch.device("cpu"))
    ccfg = U.ContrastiveCfg()
    for k, v in c["aug"].items():
        setattr(ccfg, "aug_" + k, v)
    out = {"adjacency": adj, "x_full": x_full.numpy(), "a_full": a_full.numpy(), "edge_index": eg.numpy(),
           "edge_index_local": el.numpy(), "meta": np.array([c["T"], N, E, c["D"], c["B"]], dtype=np.int64),
           "temperature": np.array(0.1), "loss_function": np.array(c.get("loss", "nce")),
           "similarity_function": np.array(c.get("sim", "cosine")), "tau": np.array(model.tau),
           "beta": np.array(model.beta)}
    for f in ("min_shift", "max_shift", "p_shift", "max_rot", "n_rot", "p_rot", "max_interp", "min_interp", "p_interp",
              "noise_sigma", "p_noise"):
        out["aug/" + f] = np.array(getattr(ccfg, "aug_" + f), dtype=np.float64)
    for k, v in model.state_dict().items():
        out["p/" + k] = v.detach().numpy().copy()
    lr = 1e-3
    dctx, head = distill_ctx(c, out, c["B"] + 5, c["seed"])
    opt = L.build_optimizer_generic(model, head, base_lr=lr, weight_decay=1e-4)
    out["lr"] = np.array(lr)
    model.train()
    ctx = types.SimpleNamespace(**{**(dctx or dict(apply_distill=False)), "edge_index": eg, "edge_index_local": el,
                                   "contrastive_cfg": ccfg, "rot_precomp": rot})
    idx = torch.arange(c["B"]) if head is None else torch.randperm(c["B"] + 5)[:c["B"]]
    out["idx"] = idx.numpy()
    seen = []
    orig_forward = model.forward

    def recording_forward(xx, aa):
        z = orig_forward(xx, aa)
        seen.append((xx.detach().clone(), aa.detach().clone(), z.detach().clone()))
        return z

    model.forward = recording_forward
    for step in range(2):
        seed = 4000 + 10 * c["seed"] + step
        torch.manual_seed(seed)
        out[f"s{step}/seed"] = np.array(seed, dtype=np.int64)
        seen.clear()
        res = T.step_contrastive_distill(model, (x_full, a_full, idx), ctx)
        opt.zero_grad(set_to_none=True)
        res.loss.backward()
        (x, a, z), (xa, aa, za) = seen
        for k, v in (("x", x), ("a", a), ("z", z), ("x_aug", xa), ("a_aug", aa), ("z_aug", za)):
            out[f"s{step}/{k}"] = v.numpy()
        for k, v in res.logs.items():
            out[f"s{step}/log/{k}"] = np.array(v, dtype=np.float64)
        if step == 0:
            for k, prm in model.named_parameters():
                if prm.grad is not None:
                    out["g/" + k] = prm.grad.detach().numpy().copy()
            if head is not None:
                for k, prm in head.named_parameters():
                    out["distill/g/" + k] = prm.grad.detach().numpy().copy()
        torch.nn.utils.clip_grad_value_(model.parameters(), 0.75)
        opt.step()
    for k, v in model.state_dict().items():
        out["p2/" + k] = v.detach().numpy().copy()
    if head is not None:
        for k, v in head.state_dict().items():
            out["distill/p2/" + k] = v.detach().numpy().copy()
    path = os.path.join(HERE, f"contrastive_{name}.npz")
    np.savez_compressed(path, **out)
    print("contrastive", name, "%.1f KB" % (os.path.getsize(path) / 1024), {k: round(v, 5) for k, v in res.logs.items()})


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in VQ_CASES.items():
        if not only or ("vqvae_" + name) in only:
            run_vq(name, c)
    for name, c in CON_CASES.items():
        if not only or ("contrastive_" + name) in only:
            run_con(name, c)
