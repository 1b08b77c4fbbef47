"""Golden vectors for the host-side pieces of the fit loop, produced by the UNMODIFIED reference on CPU (build container
only):  python tests/golden/make_golden_fit.py   ->  fit_helpers.npz

  gmm/*    VaDEPT.initialize_gmm_from_data (models_new.py:1907-1947) on a stub model whose encoder returns given embeddings
  diag*/   logging.compute_diagnostics (logging.py:149-301) on given soft assignments, with and without a teacher
  summary  init_log_summary / _update_log_summary (logging.py:304-351): key lists and what lands where
  edges/*  training._build_edge_from_metainfo (training.py:1936-2004) on a two-animal meta_info
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

M, L, T, U = refshim.load()
LG = sys.modules["deepof.clustering.logging"]

out = {}

# ---- initialize_gmm_from_data
g = torch.Generator().manual_seed(3)
K, D, n = 4, 6, 900
centres = torch.randn(K, D, generator=g) * 3.0
emb = (centres[torch.randint(0, K, (n,), generator=g)] + torch.randn(n, D, generator=g) * torch.rand(D, generator=g)).float()


class _Latent:
    n_components = K
    gmm_means = types.SimpleNamespace(data=None)
    gmm_log_vars = types.SimpleNamespace(data=None)

    def _encode(self, enc):
        return enc, None


class _Stub:
    latent_space = _Latent()

    def eval(self):
        return self

    def parameters(self):
        return iter([torch.zeros(1)])

    def encoder(self, x, a):
        return x


loader = [(emb[i:i + 256], emb[i:i + 256]) for i in range(0, n, 256)]
np.random.seed(11)
M.VaDEPT.initialize_gmm_from_data(_Stub(), loader, n_samples=700)
out["gmm/emb"] = emb.numpy()
out["gmm/means"] = _Stub.latent_space.gmm_means.data.numpy()
out["gmm/log_vars"] = _Stub.latent_space.gmm_log_vars.data.numpy()
out["gmm/meta"] = np.array([K, 700, 11], dtype=np.int64)          # n_components, n_samples, numpy seed


# ---- compute_diagnostics
class _M(torch.nn.Module):
    pass


Kq = 5
qb = [torch.softmax(torch.randn(40, Kq, generator=g) * 2.0, dim=-1) for _ in range(6)]
tau = torch.softmax(torch.randn(300, Kq, generator=g) * 1.5, dim=-1)
dl = [(q, q) for q in qb]
for tag, kw in (("diag_teacher", dict(tau_star=tau, distill_sharpen_T=0.5, distill_conf_weight=True, distill_conf_thresh=0.3)),
                ("diag_noteacher", dict(tau_star=None)),
                ("diag_T0", dict(tau_star=tau, distill_sharpen_T=0.0, distill_conf_weight=False))):
    d = LG.compute_diagnostics(model=_M(), dataloader=dl, q_fn=lambda m, x, a: x, device=torch.device("cpu"), n_components=Kq,
                               max_batches=4, **kw)
    out[tag + "/json"] = np.array(json.dumps(d))
out["diag/q"] = torch.stack(qb).numpy()
out["diag/tau"] = tau.numpy()

# ---- log summary
ls = LG.init_log_summary("vade")
out["summary/top_keys"] = np.array(json.dumps(list(ls.keys())))
out["summary/train_keys"] = np.array(json.dumps(list(ls["train"].keys())))
tl = {"total_loss": 1.5, "reconstruct_loss": 1.0, "kl_div": 0.2, "kmeans_loss": 0.1, "distill_loss": 0.05, "model_type": "x"}
vl = {"total_loss": 2.5, "alignment_score": 0.3, "conf_norm": 0.5, "bal_norm": 0.6}
ls = LG._update_log_summary(ls, tl, vl)
out["summary/after"] = np.array(json.dumps(ls))

# ---- edges from meta_info (two animals, one cross-animal edge)
names = ["B_Nose", "B_Spine", "B_Tail", "W_Nose", "W_Spine", "W_Tail"]
meta = {"node_columns": [(n_, "x") for n_ in names] + [(n_, "y") for n_ in names] + names,
        "edge_columns": [("B_Nose", "B_Spine"), ("B_Spine", "B_Tail"), ("B_Nose", "W_Nose"), ("W_Nose", "W_Spine"), ("W_Spine", "W_Tail")]}
eg, el, _ = T._build_edge_from_metainfo(meta, torch.device("cpu"), len(names))
out["edges/meta"] = np.array(json.dumps({"node_columns": [list(c) if isinstance(c, tuple) else c for c in meta["node_columns"]],
                                         "edge_columns": [list(c) for c in meta["edge_columns"]]}))
out["edges/global"] = eg.numpy()
out["edges/local"] = el.numpy()

np.savez_compressed(os.path.join(HERE, "fit_helpers.npz"), **out)
print("fit_helpers.npz", {k: getattr(v, "shape", None) for k, v in out.items()})
