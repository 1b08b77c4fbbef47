"""CPU: host-side pieces of the fit loop (deepof_b200/api.py, gmm_init.py) against goldens produced by the UNMODIFIED
reference (tests/golden/make_golden_fit.py): GMM initialisation from embeddings, compute_diagnostics / alignment score,
the log-summary structure, edge lists from meta_info; plus the oracle-side parameter tables against the library layout."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR

G = np.load(os.path.join(GOLDEN_DIR, "fit_helpers.npz"), allow_pickle=False)


def test_gmm_init_from_embeddings_matches_reference():
    from deepof_b200.gmm_init import gmm_from_embeddings
    K, n, seed = (int(v) for v in G["gmm/meta"])
    np.random.seed(seed)
    means, log_vars = gmm_from_embeddings(G["gmm/emb"][:n], K)
    assert np.allclose(means.astype(np.float32), G["gmm/means"], rtol=0, atol=1e-6)
    assert np.allclose(log_vars.astype(np.float32), G["gmm/log_vars"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag,kw", [("diag_teacher", dict(tau=True, distill_sharpen_T=0.5, distill_conf_weight=True, distill_conf_thresh=0.3)),
                                    ("diag_noteacher", dict(tau=False)),
                                    ("diag_T0", dict(tau=True, distill_sharpen_T=0.0, distill_conf_weight=False))])
def test_compute_diagnostics_matches_reference(tag, kw):
    from deepof_b200.api import compute_diagnostics
    ref = json.loads(str(G[tag + "/json"]))
    q = torch.from_numpy(G["diag/q"])
    kw = dict(kw)
    tau = torch.from_numpy(G["diag/tau"]) if kw.pop("tau") else None
    got = compute_diagnostics(None, [(qb, qb) for qb in q], lambda m, x, a: x, q.shape[2], tau_star=tau, max_batches=4, **kw)
    assert set(got) == set(ref)
    for k, v in ref.items():
        if v != v:
            assert got[k] != got[k], k
        else:
            assert abs(got[k] - v) <= 1e-6 * max(1.0, abs(v)), (k, got[k], v)


def test_log_summary_structure_matches_reference():
    from deepof_b200.api import init_log_summary, update_log_summary
    ls = init_log_summary("vade")
    assert list(ls.keys()) == json.loads(str(G["summary/top_keys"]))
    assert list(ls["train"].keys()) == json.loads(str(G["summary/train_keys"])) == list(ls["val"].keys())
    tl = {"total_loss": 1.5, "reconstruct_loss": 1.0, "kl_div": 0.2, "kmeans_loss": 0.1, "distill_loss": 0.05, "model_type": "x"}
    vl = {"total_loss": 2.5, "alignment_score": 0.3, "conf_norm": 0.5, "bal_norm": 0.6}
    ls = update_log_summary(ls, tl, vl)
    ref = json.loads(str(G["summary/after"]))

    def same(a, b):
        if isinstance(a, dict):
            return set(a) == set(b) and all(same(a[k], b[k]) for k in a)
        if isinstance(a, list):
            return len(a) == len(b) and all(same(u, v) for u, v in zip(a, b))
        if isinstance(a, float) and a != a:
            return b != b
        return a == b
    assert same(ls, ref), (ls, ref)


def test_edges_from_meta_info_match_reference():
    from deepof_b200.api import build_edge_from_metainfo
    meta = json.loads(str(G["edges/meta"]))
    meta["node_columns"] = [tuple(c) if isinstance(c, list) else c for c in meta["node_columns"]]
    meta["edge_columns"] = [tuple(c) for c in meta["edge_columns"]]
    eg, el = build_edge_from_metainfo(meta, 6)
    assert np.array_equal(eg, G["edges/global"]) and np.array_equal(el, G["edges/local"])


def test_train_deepof_model_signature_is_the_references():
    """Keyword names and defaults of train_deepof_model = the reference's (training.py:592-719), so a misspelled keyword
    fails exactly as there and every default is the reference's default."""
    import inspect
    from deepof_b200 import train_deepof_model
    sig = inspect.signature(train_deepof_model)
    ref = {"preprocessed_object": None, "encoder_type": None, "n_clusters": 10, "learning_rate": 1e-3, "run": 0, "freeze_gmm_epochs": 0,
           "freeze_decoder_epochs": 0, "interaction_regularization": 0.0003, "use_turtle_teacher": True, "teacher_gamma": 8.0,
           "teacher_outer_steps": 500, "teacher_inner_steps": 100, "lambda_distill": 4.0, "lambda_decay_start": 10, "lambda_end_weight": 0.2,
           "lambda_cooldown": 10, "teacher_refresh_every": False, "teacher_freeze_at": 10, "teacher_batch_size": 2048, "pretrain_epochs": 10,
           "kl_warmup_pretrain": 15, "kl_max_weight_pretrain": 0.2, "nonempty_weight": 2e-2, "distill_conf_thresh": 0.3,
           "pca_nodes_dim": 32, "diag_max_batches": 4, "model_name": "VaDE", "generic_distill_conf_thresh": 0.6, "temperature": 0.1,
           "aug_max_shift": 3, "aug_p_rot": 0.8, "h5_dataset_folder": None, "bootstrap_block_len": 250, "random_seed": 0}
    for k, v in ref.items():
        assert k in sig.parameters and sig.parameters[k].default == v, k
    assert len(sig.parameters) == 108 and not any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())
    with pytest.raises(TypeError):
        train_deepof_model(latent_dimm=3)


def test_oracle_parameter_tables_match_the_library_layout():
    """oracle/params.py (what the CPU reference arm of bench.py builds its parameters from) lists exactly the entries of
    the library's state layout = the reference's state_dict, for every model kind and encoder family."""
    import ctypes as C
    from deepof_b200 import _lib
    from deepof_b200.vade import state_layout
    from oracle import params as P
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    for kind, mi in (("vade", 0), ("vqvae", 1), ("contrastive", 2)):
        for enc, ei in (("recurrent", 0), ("transformer", 1)):
            lay = state_layout(_lib.DofConfig(25, 14, 14, 3, 1, 16, 8, mi, ei))
            sh = (P.transformer_shapes if enc == "transformer" else P.recurrent_shapes)(kind, 14, 14, 3, 1, 16, 8)
            assert [(n, tuple(s)) for n, _, _, s, _ in lay] == [(n, tuple(s)) for n, s in sh.items()], (kind, enc)
