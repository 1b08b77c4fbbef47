"""CPU: the oracle restatement vs golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import numpy as np
import pytest
import torch

from oracle import vade_oracle as O
from helpers import golden_cases, load_golden, sub, rel_l2

CASES = golden_cases()


def _cfg(g):
    d = g["dims"]
    klw = float(g["s0/klw"])
    if str(g["phase"]) == "pretrain":
        cfg = O.LossCfg.pretrain_defaults(d["K"], kl_weight=klw)
    else:
        cfg = O.LossCfg.main_defaults(d["K"], kl_weight=klw)
    if "tau_star" in g:
        cfg.lambda_distill = float(g["lambda_distill"])
    return cfg


def test_golden_present():
    assert len(CASES) >= 4


@pytest.mark.parametrize("case", CASES)
def test_graph_operators_match_reference_buffers(case):
    g = load_golden(case)
    lap, elap, inc = O.graph_operators(g["adjacency"])
    assert torch.equal(inc, torch.from_numpy(g["p/encoder.incidence"]))
    np.testing.assert_allclose(lap.numpy(), g["p/encoder.laplacian"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(elap.numpy(), g["p/encoder.edge_laplacian"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("case", CASES)
def test_eval_outputs(case):
    """Judged outputs: embedding + q within 1e-4 rel-L2 (here: ~1e-6), argmax exact."""
    g = load_golden(case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    D = g["dims"]["D"]
    with torch.no_grad():
        out = O.vade_forward(x, a, p, graph, D, training=False)
    assert rel_l2(out["enc"], g["eval/enc"]) < 2e-6
    assert rel_l2(out["z"], g["eval/emb"]) < 2e-6
    assert rel_l2(out["q"], g["eval/q"]) < 2e-5
    assert rel_l2(out["loc"], g["eval/loc"]) < 5e-6
    assert torch.equal(out["q"].argmax(1), torch.from_numpy(g["eval/q"]).argmax(1))


@pytest.mark.parametrize("case", CASES)
def test_two_training_steps(case):
    """13 logged loss terms per step, raw gradient of step 1, parameters after 2 Adam steps."""
    g = load_golden(case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    d = g["dims"]
    cfg = _cfg(g)
    lr_base, lr_gmm = (float(v) for v in g["lr"])
    state = {}
    tau = torch.from_numpy(g["tau_star"]) if "tau_star" in g else None
    cw = torch.from_numpy(g["class_weight"]) if "class_weight" in g else None
    tm = torch.from_numpy(g["teacher_marginal"]) if "teacher_marginal" in g else None
    for step in range(2):
        cfg.kl_weight = float(g[f"s{step}/klw"])
        eps = torch.from_numpy(g[f"s{step}/eps"])
        mc = torch.from_numpy(g[f"s{step}/mc_eps"]) if f"s{step}/mc_eps" in g else None
        logs, grads, _ = O.train_step(x, a, p, graph, d["D"], cfg, eps=eps, mc_eps=mc,
                                      tau_batch=tau, class_weight=cw, teacher_marginal=tm)
        for k in O.LOG_KEYS:
            ref = float(g[f"s{step}/log/{k}"])
            assert abs(logs[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (step, k, logs[k], ref)
        if step == 0:
            gref = sub(g, "g/")
            for k, gr in grads.items():
                if O.dead_parameter(k, p, d["D"]):
                    assert gr is None and k not in gref, k
                    continue
                assert k in gref, k
                err = float((gr - gref[k]).abs().max())
                scale = float(gref[k].abs().max())
                assert err <= 2e-5 * max(scale, 1e-3) + 1e-7, (k, err, scale)
            flat = torch.cat([grads[k].flatten() for k in gref])
            flat_ref = torch.cat([gref[k].flatten() for k in gref])
            assert rel_l2(flat, flat_ref) < 1e-5
        O.adam_step(p, grads, state, lr_base, lr_gmm)
    p2 = sub(g, "p2/")
    for k in O.trainable_names(p):
        err = float((p[k] - p2[k]).abs().max())
        assert err <= 5e-6, (k, err)


def test_kl_schedule_matches_reference_values():
    # tf_sigmoid(p) = sigmoid((2p-1)/max(0.01, p-p^2))  (losses.py:317-321)
    g = load_golden("cfg1_main")
    nb, it0 = (int(v) for v in g["sched"])
    for step in range(2):
        w = O.kl_weight_schedule(it0 + step, nb, "tf_sigmoid", 5, 1.0, 5, 0.2)
        assert abs(w - float(g[f"s{step}/klw"])) < 1e-12
    g = load_golden("cfg1_pretrain")
    nb, it0 = (int(v) for v in g["sched"])
    for step in range(2):
        w = O.kl_weight_schedule(it0 + step, nb, "tf_sigmoid", 15, 0.2, 10, 0.2)
        assert abs(w - float(g[f"s{step}/klw"])) < 1e-12


def test_group_reshape_law_examples():
    # SURVEY A.1 worked example: T=25,G=14,F=3 -> out[b,0,0..5,0] reads (t,j) =
    # (0,0),(14,0),(3,1),(17,1),(6,2),(20,2)
    idx = O.group_gather_index(25, 14, 3)
    got = [(int(i) // 42, int(i) % 42) for i in idx[0, :6, 0]]
    assert got == [(0, 0), (14, 0), (3, 1), (17, 1), (6, 2), (20, 2)]
    # it is a permutation of the window
    assert sorted(idx.flatten().tolist()) == list(range(25 * 14 * 3))


@pytest.mark.parametrize("case", ["cfg2_main", "odd_pretrain"])
def test_aten_gru_variant_matches_golden(case):
    """The ATen-GRU variant of the oracle (used only as the timed CPU baseline in bench.py)
    reproduces the reference's logged loss terms and raw gradient too."""
    g = load_golden(case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    cfg = _cfg(g)
    cfg.kl_weight = float(g["s0/klw"])
    mc = torch.from_numpy(g["s0/mc_eps"]) if "s0/mc_eps" in g else None
    O.USE_ATEN_GRU = True
    try:
        logs, grads, _ = O.train_step(x, a, p, graph, g["dims"]["D"], cfg, eps=torch.from_numpy(g["s0/eps"]), mc_eps=mc)
    finally:
        O.USE_ATEN_GRU = False
    for k in O.LOG_KEYS:
        ref = float(g[f"s0/log/{k}"])
        assert abs(logs[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (k, logs[k], ref)
    gref = sub(g, "g/")
    flat = torch.cat([grads[k].flatten() for k in gref])
    flat_ref = torch.cat([gref[k].flatten() for k in gref])
    assert rel_l2(flat, flat_ref) < 1e-5
