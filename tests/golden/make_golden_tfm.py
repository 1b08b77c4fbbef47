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


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in CASES.items():
        if not only or name in only:
            run(name, c)
