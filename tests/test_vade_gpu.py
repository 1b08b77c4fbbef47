"""GPU parity of the CUDA VaDE/recurrent path (through the C-ABI) against
 (1) the golden vectors produced by the UNMODIFIED reference (tests/golden), and
 (2) the CPU oracle on larger seeded batches, incl. zero-padded windows.
Tolerances: embeddings / q within 1e-4 rel-L2 and argmax(q) exact (BASELINE.json
north_star); loss terms 1e-4; gradients 1e-3 rel-L2 per tensor (fp32 reassociation)."""
import numpy as np
import pytest
import torch

from oracle import vade_oracle as O
from helpers import golden_cases, load_golden, sub, rel_l2

pytestmark = pytest.mark.gpu

CASES = golden_cases()


def _model(g, max_batch=None, training=True):
    from deepof_b200 import VaDEB200
    d = g["dims"]
    m = VaDEB200((d["T"], d["N"], 3), (d["T"], d["E"], 1), g["adjacency"], d["D"], d["K"],
                 max_batch=max_batch or d["B"], training=training, seed=0)
    m.load_state_dict({k[2:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith("p/")})
    return m


def _cfg(g, step):
    from deepof_b200 import VadeLossCfg
    d = g["dims"]
    klw = float(g[f"s{step}/klw"])
    cfg = VadeLossCfg.pretrain_defaults(d["K"], klw) if str(g["phase"]) == "pretrain" else \
        VadeLossCfg.main_defaults(d["K"], klw)
    if "tau_star" in g:
        cfg.lambda_distill = float(g["lambda_distill"])
    return cfg


@pytest.mark.parametrize("case", CASES)
def test_state_buffers_match_reference(case):
    g = load_golden(case)
    from deepof_b200 import VaDEB200
    d = g["dims"]
    m = VaDEB200((d["T"], d["N"], 3), (d["T"], d["E"], 1), g["adjacency"], d["D"], d["K"], max_batch=4,
                 training=False, seed=0)
    sd = m.state_dict()
    assert list(sd.keys()) == [k[2:] for k in g if k.startswith("p/")]
    for k in ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence", "latent_space.prior"):
        np.testing.assert_allclose(sd[k].cpu().numpy(), g["p/" + k], rtol=0, atol=1e-7)


@pytest.mark.parametrize("case", CASES)
def test_eval_outputs_vs_reference_golden(case):
    g = load_golden(case)
    m = _model(g, training=False)
    enc, emb, q, loc = m.forward_eval(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
    errs = dict(enc=rel_l2(enc.cpu(), g["eval/enc"]), emb=rel_l2(emb.cpu(), g["eval/emb"]),
                q=rel_l2(q.cpu(), g["eval/q"]), loc=rel_l2(loc.cpu(), g["eval/loc"]))
    print(case, errs)
    assert errs["emb"] < 1e-4 and errs["q"] < 1e-4 and errs["enc"] < 1e-4 and errs["loc"] < 1e-4, errs
    assert torch.equal(q.cpu().argmax(1), torch.from_numpy(g["eval/q"]).argmax(1))
    emb2, q2 = m.embed(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
    assert torch.equal(emb2, emb) and torch.equal(q2, q)


@pytest.mark.parametrize("case", CASES)
def test_two_training_steps_vs_reference_golden(case):
    g = load_golden(case)
    m = _model(g)
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    lr_base, lr_gmm = (float(v) for v in g["lr"])
    tau = torch.from_numpy(g["tau_star"]) if "tau_star" in g else None
    cw = torch.from_numpy(g["class_weight"]) if "class_weight" in g else None
    tm = torch.from_numpy(g["teacher_marginal"]) if "teacher_marginal" in g else None
    bad = []
    for step in range(2):
        cfg = _cfg(g, step)
        mc = torch.from_numpy(g[f"s{step}/mc_eps"]) if f"s{step}/mc_eps" in g else None
        m.loss_grad(x, a, cfg, eps=torch.from_numpy(g[f"s{step}/eps"]), mc_eps=mc, tau_batch=tau, class_weight=cw,
                    teacher_marginal=tm)
        logs = m.logs_dict()
        for k in O.LOG_KEYS:
            ref = float(g[f"s{step}/log/{k}"])
            if abs(logs[k] - ref) > 1e-4 * max(1.0, abs(ref)):
                bad.append(("log", step, k, logs[k], ref))
        if step == 0:
            gref = sub(g, "g/")
            gd = m.grad_dict()
            for k, gr in gref.items():
                err = float((gd[k].cpu() - gr).norm())
                scale = float(gr.norm())
                if err > 1e-3 * scale + 1e-7:
                    bad.append(("grad", k, err, scale))
            for k, v in gd.items():   # dead parameters / buffers keep a zero gradient
                if k not in gref:
                    assert float(v.abs().max()) == 0.0, k
            flat = torch.cat([gd[k].cpu().flatten() for k in gref])
            flat_ref = torch.cat([gref[k].flatten() for k in gref])
            print(case, "flat grad rel-L2", rel_l2(flat, flat_ref))
            if rel_l2(flat, flat_ref) > 2e-4:
                bad.append(("flatgrad", rel_l2(flat, flat_ref)))
        m.adam_step(lr_base, lr_gmm)
    assert not bad, bad
    p2 = sub(g, "p2/")
    sd = m.state_dict()
    worst = max((float((sd[k].cpu() - p2[k]).abs().max()), k) for k in p2)
    print(case, "post-Adam worst abs diff", worst)
    # Adam normalises the step: a parameter whose gradient is ~0 can flip sign of m/sqrt(v);
    # bound by 2 steps * lr plus tolerance on everything else
    for k in p2:
        err = float((sd[k].cpu() - p2[k]).abs().max())
        assert err <= 2e-4, (k, err)


def _oracle_case(T, N, D, K, B, seed, pad=False):
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = O.synthetic_windows(B, T, adj, seed=seed)
    if pad:   # zero tails: exercises the packed-sequence (valid-length) semantics
        g = torch.Generator().manual_seed(seed)
        cut = torch.randint(T // 2, T + 1, (B,), generator=g)
        for b in range(B):
            x[b, cut[b]:] = 0.0
            a[b, cut[b]:] = 0.0
    return adj, E, x, a


@pytest.mark.parametrize("T,N,D,K,B,pad", [(25, 14, 16, 8, 96, False), (25, 14, 8, 4, 64, False),
                                           (24, 11, 6, 5, 33, False), (25, 14, 16, 8, 40, True),
                                           (12, 5, 4, 3, 7, False), (25, 14, 32, 16, 16, False),
                                           # B*T >= 4096 rows: the decoder's tensor-core weight-gradient GEMMs, incl. the
                                           # five time-shifted per-tap GEMMs of the Conv1d(k5) weight gradient
                                           (25, 14, 16, 8, 192, False)])
def test_eval_and_step_vs_oracle(T, N, D, K, B, pad):
    from deepof_b200 import VaDEB200, VadeLossCfg
    adj, E, x, a = _oracle_case(T, N, D, K, B, seed=100 + D + B, pad=pad)
    m = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, training=True, seed=D * 7 + K)
    with torch.no_grad():
        m.latent_space.gmm_means.mul_(3.0)
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    graph = O.graph_operators(adj)
    enc, emb, q, loc = m.forward_eval(x, a)
    with torch.no_grad():
        ref = O.vade_forward(x, a, p, graph, D, training=False)
    errs = dict(enc=rel_l2(enc.cpu(), ref["enc"]), emb=rel_l2(emb.cpu(), ref["z"]), q=rel_l2(q.cpu(), ref["q"]))
    if not pad:
        errs["loc"] = rel_l2(loc.cpu(), ref["loc"])
    print("eval", (T, N, D, K, B, pad), errs)
    assert max(errs.values()) < 1e-4, errs
    # argmax(q) exact wherever the top-2 margin is not a numerical tie
    top2 = ref["q"].topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(q.cpu().argmax(1)[clear], ref["q"].argmax(1)[clear])
    if pad:
        return   # the reference loss is NaN on all-zero rows (log|scale|=-inf); forward-only check
    gen = torch.Generator().manual_seed(5)
    eps = torch.randn(B, D, generator=gen)
    mc = torch.randn(32, B, D, generator=gen)
    for phase in ("pretrain", "main"):
        if phase == "pretrain":
            cfg, ocfg = VadeLossCfg.pretrain_defaults(K, 0.15), O.LossCfg.pretrain_defaults(K, 0.15)
        else:
            cfg, ocfg = VadeLossCfg.main_defaults(K, 0.8), O.LossCfg.main_defaults(K, 0.8)
            cfg.reg_cat_clusters_weight = ocfg.reg_cat_clusters_weight = 0.3
            cfg.temporal_cohesion_weight = ocfg.temporal_cohesion_weight = 0.2
        m.loss_grad(x, a, cfg, eps=eps, mc_eps=mc)
        logs = m.logs_dict()
        ologs, ograds, _ = O.train_step(x, a, p, graph, D, ocfg, eps=eps, mc_eps=mc)
        bad = []
        for k, v in ologs.items():
            if abs(logs[k] - v) > 1e-4 * max(1.0, abs(v)):
                bad.append(("log", k, logs[k], v))
        gd = m.grad_dict()
        for k, gv in ograds.items():
            if gv is None:
                continue
            err, scale = float((gd[k].cpu() - gv).norm()), float(gv.norm())
            if err > 1e-3 * scale + 1e-7:
                bad.append(("grad", k, err, scale))
        names = [k for k, gv in ograds.items() if gv is not None]
        fr = rel_l2(torch.cat([gd[k].cpu().flatten() for k in names]), torch.cat([ograds[k].flatten() for k in names]))
        print(phase, (T, N, D, K, B), "flat grad rel-L2", fr)
        assert not bad and fr < 2e-4, (phase, fr, bad)


@pytest.mark.parametrize("phase", ["main", "pretrain"])
def test_full_size_step_vs_oracle(phase):
    """BASELINE cfg2 at the benchmarked size — B = 4096 windows in ONE step, every default batch-level term of the phase
    on (main: MC-KL S = 32, non-empty floor; pretrain: Gram-SVD k-means loss, repel between soft centroids, non-empty
    floor) — directly against the CPU oracle: 13 logs (1e-4), flat gradient (2e-4), eval embeddings / q (1e-4) and the
    argmax flip count over all 4096 windows (flips only inside numerical ties)."""
    from deepof_b200 import VaDEB200, VadeLossCfg
    O.USE_ATEN_GRU = True                      # the reference's nn.GRU kernels: the python-loop GRU of the oracle is 20x slower
    try:
        T, N, D, K, B = 25, 14, 16, 8, 4096
        adj, E, x, a = _oracle_case(T, N, D, K, B, seed=4096)
        m = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, training=True, seed=21)
        with torch.no_grad():
            m.latent_space.gmm_means.mul_(3.0)
        p = {k: v.cpu() for k, v in m.state_dict().items()}
        graph = O.graph_operators(adj)
        gen = torch.Generator().manual_seed(6)
        eps, mc = torch.randn(B, D, generator=gen), torch.randn(32, B, D, generator=gen)
        if phase == "main":
            cfg, ocfg = VadeLossCfg.main_defaults(K, 0.8), O.LossCfg.main_defaults(K, 0.8)
        else:
            cfg, ocfg = VadeLossCfg.pretrain_defaults(K, 0.15), O.LossCfg.pretrain_defaults(K, 0.15)
            m.set_pretrain_mode(True)
            p = {k: v.cpu() for k, v in m.state_dict().items()}
        m.loss_grad(x, a, cfg, eps=eps, mc_eps=mc)
        logs = m.logs_dict()
        ologs, ograds, _ = O.train_step(x, a, p, graph, D, ocfg, eps=eps, mc_eps=mc)
        for k, v in ologs.items():
            assert abs(logs[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, logs[k], v)
        gd = m.grad_dict()
        names = [k for k, gv in ograds.items() if gv is not None]
        fr = rel_l2(torch.cat([gd[k].cpu().flatten() for k in names]), torch.cat([ograds[k].flatten() for k in names]))
        print(phase, "B=4096 flat grad rel-L2", fr)
        assert fr < 2e-4, fr
        if phase == "main":
            enc, emb, q, _ = m.forward_eval(x, a, want_loc=False)
            with torch.no_grad():
                ref = O.vade_forward(x, a, p, graph, D, training=False)
            assert rel_l2(emb.cpu(), ref["z"]) < 1e-4 and rel_l2(q.cpu(), ref["q"]) < 1e-4
            flips = q.cpu().argmax(1) != ref["q"].argmax(1)
            top2 = ref["q"].topk(2, dim=1).values
            margin = top2[:, 0] - top2[:, 1]
            print("argmax(q) flips at B=4096 (unconstrained data):", int(flips.sum()), "of", B,
                  "; smallest reference top-2 margin:", float(margin.min()))
            assert bool((margin[flips] < 1e-4).all())
    finally:
        O.USE_ATEN_GRU = False


def test_batch_chunking_and_determinism():
    from deepof_b200 import VaDEB200
    T, N, D, K, B = 25, 14, 16, 8, 70
    adj, E, x, a = _oracle_case(T, N, D, K, B, seed=9)
    m = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=32, training=False, seed=1)
    emb1, q1 = m.embed(x, a)           # 3 chunks
    emb2, q2 = m.embed(x, a)
    assert torch.equal(emb1, emb2) and torch.equal(q1, q2)
    m2 = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=128, training=False, seed=1)
    m2.load_state_dict(m.state_dict())
    emb3, q3 = m2.embed(x, a)
    assert torch.equal(emb1, emb3) and torch.equal(q1, q3)


def test_full_size_batch_properties():
    """BASELINE cfg2 batch (4096 windows/GPU), size-independent properties that tie the full-size launch geometry
    (896-CTA GRU grids, tensor-core GEMM tiles, 1.4 M-row LayerNorm / conv launches) to the sizes pinned to the oracle:
    (1) embeddings / soft assignments of the full batch equal those of the same windows pushed through in chunks of 96;
    (2) with the batch-level loss terms switched off the loss is a mean over windows, so the gradient of the full batch
        equals the mean of the gradients of its 32 chunks of 128 windows (linearity)."""
    from deepof_b200 import VaDEB200, VadeLossCfg
    T, N, D, K, B = 25, 14, 16, 8, 4096
    adj, E, x, a = _oracle_case(T, N, D, K, B, seed=77)
    m = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, training=True, seed=3)
    emb, q = m.embed(x, a)
    small = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=96, training=False, seed=3)
    small.load_state_dict(m.state_dict())
    emb_c, q_c = small.embed(x, a)          # 43 chunks
    assert rel_l2(emb.cpu(), emb_c.cpu()) < 1e-5 and rel_l2(q.cpu(), q_c.cpu()) < 1e-5
    top2 = q_c.cpu().topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(q.cpu().argmax(1)[clear], q_c.cpu().argmax(1)[clear])

    cfg = VadeLossCfg(pretrain_mode=True, kl_weight=0.15, kmeans_loss_weight=0.0, model_kmeans_weight=0.0,
                      repel_weight=0.0, nonempty_weight=0.0)
    eps = torch.randn(B, D, generator=torch.Generator().manual_seed(5))
    m.loss_grad(x, a, cfg, eps=eps)
    g_full = m.grad.clone()
    loss_full = m.logs_dict()["total_loss"] if "total_loss" in m.logs_dict() else None
    acc = torch.zeros_like(g_full, dtype=torch.float64)
    CH = 128
    for s0 in range(0, B, CH):
        m.loss_grad(x[s0:s0 + CH], a[s0:s0 + CH], cfg, eps=eps[s0:s0 + CH])
        acc += m.grad.double()
    acc /= B // CH
    err = float((g_full.double() - acc).norm() / acc.norm())
    print("full-size gradient linearity rel-L2", err, "loss", loss_full)
    assert err < 1e-4, err


def test_errors_are_loud():
    from deepof_b200 import VaDEB200, DofError, VadeLossCfg
    adj = O.default_adjacency(5)
    with pytest.raises(NotImplementedError):
        VaDEB200((12, 5, 3), (12, 4, 1), adj, 4, 3, encoder_type="LSTM")
    m = VaDEB200((12, 5, 3), (12, 4, 1), adj, 4, 3, max_batch=4, training=False)
    x, a = O.synthetic_windows(8, 12, adj, seed=1)
    with pytest.raises(DofError):
        m.loss_grad(x[:4], a[:4], VadeLossCfg.main_defaults(3))
    m3 = VaDEB200((12, 5, 3), (12, 4, 1), adj, 4, 3, max_batch=4, training=True)
    cfg = VadeLossCfg.main_defaults(3)
    cfg.tf_cluster_weight = 1.0
    with pytest.raises(DofError):
        m3.loss_grad(x[:4], a[:4], cfg)


def test_in_kernel_philox_noise():
    """eps / mc_eps = None: the kernels draw the reparameterisation noise and the 32 Monte-Carlo samples from the
    counter-based Philox stream (nothing materialised).  The recovered eps = (z - z_mean) / exp(lv / 2) is standard
    normal, changes with the step, and forward / backward use the SAME noise: the step equals the step with that eps fed
    back explicitly (MC noise pinned by feeding explicit samples to both)."""
    from deepof_b200 import VaDEB200, VadeLossCfg
    T, N, D, K, B = 25, 14, 16, 8, 2048
    adj, E, x, a = _oracle_case(T, N, D, K, B, seed=77)
    m = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, training=True, seed=5)
    cfg = VadeLossCfg.main_defaults(K, 0.8)
    mc = torch.randn(32, B, D, generator=torch.Generator().manual_seed(1))

    def eps_of_step():
        z, zm, lv = m.debug("z").view(B, D), m.debug("z_mean").view(B, D), m.debug("z_log_var").view(B, D)
        return ((z - zm) / torch.exp(0.5 * lv)).cpu()

    m.loss_grad(x, a, cfg, mc_eps=mc)                      # eps from Philox
    e1, g1, l1 = eps_of_step(), m.grad.clone(), m.logs_dict()
    m.loss_grad(x, a, cfg, mc_eps=mc)
    e2 = eps_of_step()
    assert abs(float(e1.mean())) < 0.02 and abs(float(e1.std()) - 1.0) < 0.02
    assert abs(float((e1 ** 4).mean()) - 3.0) < 0.15        # kurtosis of a normal
    assert float((e1 - e2).abs().mean()) > 0.5              # a new stream every step
    m.loss_grad(x, a, cfg, eps=e1, mc_eps=mc)               # same noise, explicit
    l3 = m.logs_dict()
    for k in l1:
        assert abs(l1[k] - l3[k]) <= 2e-5 * max(1.0, abs(l1[k])), (k, l1[k], l3[k])
    assert rel_l2(m.grad.cpu(), g1.cpu()) < 2e-4            # eps recovered through exp / division: ~1e-6 relative
    m.loss_grad(x, a, cfg)                                  # MC noise from Philox too
    lm = m.logs_dict()
    assert np.isfinite(lm["total_loss"]) and abs(lm["kl_div"] - l1["kl_div"]) < 0.05 * max(1.0, abs(l1["kl_div"]))
