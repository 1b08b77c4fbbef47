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


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in CASES.items():
        if not only or name in only:
            run(name, c)
    for name, c in MODEL_CASES.items():
        if not only or ("model_" + name) in only:
            run_model(name, c)
