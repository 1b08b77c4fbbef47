"""GPU parity of the CUDA VQ-VAE and contrastive paths (recurrent encoder, through the C-ABI) against
 (1) golden vectors produced by the UNMODIFIED reference (tests/golden/vqvae_*.npz, contrastive_*.npz) and
 (2) the CPU oracle (oracle/models_oracle.py) on larger seeded batches.
Tolerances: encoder outputs / soft counts / embeddings 1e-4 rel-L2, VQ code indices bit-exact, logged loss terms
1e-4, gradients 1e-3 rel-L2 per tensor and 2e-4 flat (fp32 reassociation + 3xTF32 GEMMs)."""
import numpy as np
import pytest
import torch

from oracle import models_oracle as MO
from oracle import vade_oracle as O
from helpers import golden_cases_of, load_golden_of, sub, rel_l2

pytestmark = pytest.mark.gpu

VQ = golden_cases_of("vqvae")
CON = golden_cases_of("contrastive")


def _load(m, g):
    m.load_state_dict({k[2:]: torch.from_numpy(np.array(v)) for k, v in g.items() if k.startswith("p/")})
    return m


def _check_grads(m, gref, bad, tag):
    gd = m.grad_dict()
    for k, gr in gref.items():
        err, scale = float((gd[k].cpu() - gr).norm()), float(gr.norm())
        if err > 1e-3 * scale + 1e-7:
            bad.append((tag, "grad", k, err, scale))
    for k, v in gd.items():            # dead parameters / buffers keep a zero gradient
        if k not in gref:
            assert float(v.abs().max()) == 0.0, k
    flat = torch.cat([gd[k].cpu().flatten() for k in gref])
    flat_ref = torch.cat([gref[k].flatten() for k in gref])
    r = rel_l2(flat, flat_ref)
    print(tag, "flat grad rel-L2", r)
    if r > 2e-4:
        bad.append((tag, "flatgrad", r))


def _check_params(m, p2, lr):
    """Two clip + Adam steps: single elements whose gradient is rounding noise may move differently (Adam normalises
    the step), so elements get 20 % of one step and every tensor 1e-4 rel-L2."""
    sd = m.state_dict()
    for k in p2:
        if p2[k].dtype.is_floating_point and p2[k].numel():
            assert float((sd[k].cpu() - p2[k]).abs().max()) <= 0.2 * lr, k
            assert rel_l2(sd[k].cpu(), p2[k]) < 1e-4, k


# ---------------------------------------------------------------------------------------------- VQ-VAE
def _vq_model(g, max_batch=None, training=True):
    from deepof_b200 import VQVAEB200
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    m = VQVAEB200((T, N, 3), (T, E, 1), g["adjacency"], D, K, kmeans_loss=float(g["kmeans"]), beta=float(g["beta"]),
                  max_batch=max_batch or B, training=training, seed=0)
    assert list(m.state_dict().keys()) == [k[2:] for k in g if k.startswith("p/")]
    return _load(m, g)


@pytest.mark.parametrize("case", VQ)
def test_vqvae_eval_vs_reference_golden(case):
    g = load_golden_of("vqvae", case)
    m = _vq_model(g, training=False)
    enc, quant, soft, idx, lq, le = m.forward_eval(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
    errs = dict(enc=rel_l2(enc.cpu(), g["eval/enc"]), soft=rel_l2(soft.cpu(), g["eval/soft"]),
                quant=rel_l2(quant.cpu(), g["eval/quant"]), lq=rel_l2(lq.cpu(), g["eval/loc_q"]), le=rel_l2(le.cpu(), g["eval/loc_e"]))
    print(case, errs)
    assert all(v < 1e-4 for v in errs.values()), errs
    assert torch.equal(idx.cpu().long(), torch.from_numpy(g["eval/idx"]))            # bit-exact code indices
    assert torch.equal(soft.argmax(1).cpu(), torch.from_numpy(g["eval/soft"]).argmax(1))
    e2, s2 = m.embed(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
    assert torch.equal(e2, enc) and torch.equal(s2, soft)
    out = m(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
    assert torch.equal(out[3], soft) and torch.equal(out[4], enc)                    # reference tuple positions [3], [4]


class _Lambda:
    def __init__(self, w):
        self.w = w

    def get_weight(self):
        return self.w

    def step(self):
        pass


def _distill_ctx(g):
    """Teacher-on goldens (distill/*): the ctx fields step_*_distill reads, with a DistillHeadB200 holding the
    reference head's parameters.  Teacher-off goldens give apply_distill=False."""
    from types import SimpleNamespace
    from deepof_b200 import DistillHeadB200
    if "distill/meta" not in g:
        return SimpleNamespace(apply_distill=False), None
    Kt, lam, Tsh, cw, thr = (float(v) for v in g["distill/meta"])
    D = g["distill/p/fc.weight"].shape[1]
    head = DistillHeadB200(D, int(Kt))
    head.load_state_dict(sub(g, "distill/p/"))
    ctx = SimpleNamespace(apply_distill=True, distill_head=head, tau_star=torch.from_numpy(g["distill/tau_star"]),
                          lambda_scheduler=_Lambda(lam), distill_sharpen_T=Tsh, distill_conf_weight=bool(cw),
                          distill_conf_thresh=thr)
    return ctx, head


def _check_head(head, g, lr, step, bad):
    if step == 0:
        for k, v in head.grad_dict().items():
            if rel_l2(v.cpu(), g["distill/g/" + k]) > 1e-4:
                bad.append(("head grad", k, rel_l2(v.cpu(), g["distill/g/" + k])))
    head.adam_step(lr)                                                      # same Adam, weight_decay 1e-4, NOT clipped
    if step == 1:
        for k, v in head.state_dict().items():
            ref = torch.from_numpy(g["distill/p2/" + k])
            assert float((v.cpu() - ref).abs().max()) <= 0.1 * lr and rel_l2(v.cpu(), ref) < 1e-4, k


@pytest.mark.parametrize("case", VQ)
def test_vqvae_two_training_steps_vs_reference_golden(case):
    from deepof_b200 import step_vqvae_distill
    from deepof_b200.models import VQ_LOG_KEYS
    g = load_golden_of("vqvae", case)
    m = _vq_model(g)
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    idx = torch.from_numpy(g["idx"]) if "idx" in g else torch.arange(x.shape[0])
    ctx, head = _distill_ctx(g)
    lr = float(g["lr"])
    bad = []
    for step in range(2):
        step_vqvae_distill(m, (x, a, idx), ctx)
        logs = m.logs_dict()
        for k in VQ_LOG_KEYS:
            ref = float(g[f"s{step}/log/{k}"])
            if abs(logs[k] - ref) > 1e-4 * max(1.0, abs(ref)):
                bad.append(("log", step, k, logs[k], ref))
        if step == 0:
            _check_grads(m, sub(g, "g/"), bad, case)
        m.adam_step(lr)
        if head is not None:
            _check_head(head, g, lr, step, bad)
    assert not bad, bad
    _check_params(m, sub(g, "p2/"), lr)


def test_vqvae_vs_oracle_large_batch_with_kmeans():
    from deepof_b200 import VQVAEB200
    T, N, D, K, B = 25, 14, 16, 64, 300
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = O.synthetic_windows(B, T, adj, seed=91)
    m = VQVAEB200((T, N, 3), (T, E, 1), adj, D, K, kmeans_loss=0.7, beta=0.25, max_batch=128, seed=4)   # eval is chunked
    with torch.no_grad():
        enc0 = m.encode(x, a)
        g = torch.Generator().manual_seed(3)
        m._views["vq_layer.codebook"].copy_((enc0.mean(0, keepdim=True).t().cpu() + enc0.std().cpu() * 1.5 * torch.randn(D, K, generator=g)).to(m.device))
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    graph = O.graph_operators(adj)
    with torch.no_grad():
        ref = MO.vqvae_forward(x, a, p, graph, D, 0.25, 0.7)
    enc, quant, soft, idx, lq, le = m.forward_eval(x, a)
    assert rel_l2(enc.cpu(), ref["enc"]) < 1e-4 and rel_l2(soft.cpu(), ref["soft"]) < 1e-4
    # code indices: exact wherever the two nearest codes are not within fp32 rounding of each other
    d = MO.vq_distances(ref["enc"], p["vq_layer.codebook"])
    top2 = d.topk(2, dim=1, largest=False).values
    safe = (top2[:, 1] - top2[:, 0]) > 1e-4 * top2[:, 1].abs()
    assert safe.float().mean() > 0.95
    assert torch.equal(idx.cpu().long()[safe], ref["idx"][safe])
    m2 = VQVAEB200((T, N, 3), (T, E, 1), adj, D, K, kmeans_loss=0.7, beta=0.25, max_batch=B, seed=4)
    m2.load_state_dict(p)
    m2.loss_grad(x, a)
    logs, grads, _ = MO.vqvae_train_step(x, a, p, graph, D, 0.25, 0.7)
    got = m2.logs_dict()
    for k, v in logs.items():
        assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, got[k], v)
    bad = []
    _check_grads(m2, {k: v for k, v in grads.items() if v is not None}, bad, "oracle-B300")
    assert not bad, bad
    assert float(m2.grad_dict()["vq_layer.codebook"].abs().max()) > 0.0


# ---------------------------------------------------------------------------------------------- contrastive
def _aug_cfg(g, cls):
    c = cls()
    for f in ("min_shift", "max_shift", "n_rot", "max_interp", "min_interp"):
        setattr(c, f, int(g["aug/" + f]))
    for f in ("p_shift", "max_rot", "p_rot", "p_interp", "noise_sigma", "p_noise"):
        setattr(c, f, float(g["aug/" + f]))
    return c


def _con_model(g, max_batch=None):
    from deepof_b200 import ContrastiveB200
    Tf, N, E, D, B = (int(v) for v in g["meta"])
    m = ContrastiveB200((Tf, N, 3), (Tf, E, 1), g["adjacency"], D, temperature=float(g["temperature"]),
                        loss_function=str(g["loss_function"]) if "loss_function" in g else "nce",
                        similarity_function=str(g["similarity_function"]) if "similarity_function" in g else "cosine",
                        tau=float(g["tau"]) if "tau" in g else 0.1, beta=float(g["beta"]) if "beta" in g else 0.1,
                        edge_index=g["edge_index"], edge_index_local=g["edge_index_local"], max_batch=max_batch or B, seed=0)
    assert list(m.state_dict().keys()) == [k[2:] for k in g if k.startswith("p/")]
    return _load(m, g)


def _to_product_params(prm):
    from deepof_b200 import AugParams
    return AugParams(start=prm.start, rot_pivot=prm.rot_pivot, rot_nodes=prm.rot_nodes, rot_theta=prm.rot_theta,
                     interp_t0=prm.interp_t0, interp_len=prm.interp_len, noise=prm.noise)


@pytest.mark.parametrize("case", CON)
def test_contrastive_two_training_steps_vs_reference_golden(case):
    from deepof_b200.models import CON_LOG_KEYS
    g = load_golden_of("contrastive", case)
    Tf, N, E, D, B = (int(v) for v in g["meta"])
    m = _con_model(g)
    # the rotation table the product builds == the reference's (same triplets / branches)
    rot = MO.rotation_table(g["edge_index_local"], N)
    assert m.rotations.triplets == rot.triplets and m.rotations.branches_a == rot.branches_a and m.rotations.branches_c == rot.branches_c
    from deepof_b200 import step_contrastive_distill
    x_full = torch.from_numpy(g["x_full"])
    idx = torch.from_numpy(g["idx"]) if "idx" in g else torch.arange(B)
    ctx, head = _distill_ctx(g)
    cfg = _aug_cfg(g, MO.AugCfg)
    lr = float(g["lr"])
    bad = []
    for step in range(2):
        torch.manual_seed(int(g[f"s{step}/seed"]))                      # replay the reference's draws (CPU generator)
        ctx.aug_params = _to_product_params(MO.draw_aug_params(B, Tf, N, cfg, rot))
        step_contrastive_distill(m, (x_full, None, idx), ctx)
        x2, a2 = m._x2[:2 * B].cpu(), m._a2[:2 * B].cpu()
        for name, got in (("x", x2[:B]), ("a", a2[:B]), ("x_aug", x2[B:]), ("a_aug", a2[B:])):
            np.testing.assert_allclose(got.numpy(), g[f"s{step}/{name}"], rtol=0, atol=1e-5, err_msg=f"{case} {name}")
        z = m.z_all[:2 * B].cpu()
        assert rel_l2(z[:B], g[f"s{step}/z"]) < 1e-4 and rel_l2(z[B:], g[f"s{step}/z_aug"]) < 1e-4
        logs = m.logs_dict()
        for k in CON_LOG_KEYS:
            ref = float(g[f"s{step}/log/{k}"])
            if abs(logs[k] - ref) > 1e-4 * max(1.0, abs(ref)):
                bad.append(("log", step, k, logs[k], ref))
        if step == 0:
            _check_grads(m, sub(g, "g/"), bad, case)
        m.adam_step(lr)
        if head is not None:
            _check_head(head, g, lr, step, bad)
    assert not bad, bad
    _check_params(m, sub(g, "p2/"), lr)
    # ContrastivePT.forward on half windows == the z of the main view
    z_main = m(torch.from_numpy(g["s1/x"]), torch.from_numpy(g["s1/a"]))
    assert z_main.shape == (B, D)


def test_contrastive_vs_oracle_multi_tile_batch():
    """B = 300 > one 128-row NT-Xent tile, not a multiple of the 8-row CTA; all four augmentations on."""
    from deepof_b200 import ContrastiveB200
    Tf, N, D, B = 50, 14, 16, 300
    adj = O.default_adjacency(N)
    r, c = np.nonzero(np.triu(adj))
    ei = np.stack([r, c], 1)
    x_full, _ = O.synthetic_windows(B, Tf, adj, seed=77)
    m = ContrastiveB200((Tf, N, 3), (Tf, len(r), 1), adj, D, temperature=0.1, max_batch=B, seed=6)
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    graph = O.graph_operators(adj)
    rot = MO.rotation_table(ei, N)
    cfg = MO.AugCfg(p_rot=0.7, p_noise=1.0, p_interp=0.6, n_rot=3)
    torch.manual_seed(5)
    prm = MO.draw_aug_params(B, Tf, N, cfg, rot)
    logs, grads, out = MO.contrastive_train_step(x_full, p, graph, D, torch.from_numpy(ei), prm, 0.1)
    m.loss_grad(x_full, _to_product_params(prm))
    z = m.z_all[:2 * B].cpu()
    assert rel_l2(z[:B], out["z"]) < 1e-4 and rel_l2(z[B:], out["z_aug"]) < 1e-4
    got = m.logs_dict()
    for k, v in logs.items():
        assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, got[k], v)
    bad = []
    _check_grads(m, {k: v for k, v in grads.items() if v is not None}, bad, "oracle-B300")
    assert not bad, bad


@pytest.mark.parametrize("loss_fn,beta", [("dcl", 0.1), ("hard_dcl", 0.1), ("hard_dcl", 0.0), ("hard_dcl", 1.0), ("fc", 0.1)])
def test_contrastive_debiased_losses_vs_oracle(loss_fn, beta):
    """dcl / hard_dcl / fc (losses.py:144-249) on a multi-tile batch, incl. beta = 0 (no re-weighting)."""
    from deepof_b200 import ContrastiveB200
    Tf, N, D, B = 50, 14, 16, 200
    adj = O.default_adjacency(N)
    r, c = np.nonzero(np.triu(adj))
    ei = np.stack([r, c], 1)
    x_full, _ = O.synthetic_windows(B, Tf, adj, seed=78)
    m = ContrastiveB200((Tf, N, 3), (Tf, len(r), 1), adj, D, temperature=0.1, loss_function=loss_fn, tau=0.1, beta=beta,
                        max_batch=B, seed=7)
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    graph = O.graph_operators(adj)
    rot = MO.rotation_table(ei, N)
    torch.manual_seed(6)
    prm = MO.draw_aug_params(B, Tf, N, MO.AugCfg(), rot)
    logs, grads, out = MO.contrastive_train_step(x_full, p, graph, D, torch.from_numpy(ei), prm, 0.1, loss_fn=loss_fn,
                                                 tau_plus=0.1, beta=beta)
    m.loss_grad(x_full, _to_product_params(prm))
    got = m.logs_dict()
    for k, v in logs.items():
        assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, got[k], v)
    bad = []
    _check_grads(m, {k: v for k, v in grads.items() if v is not None}, bad, f"{loss_fn}-beta{beta}")
    assert not bad, bad


@pytest.mark.parametrize("sim,loss_fn", [("dot", "nce"), ("euclidean", "nce"), ("edit", "dcl"), ("euclidean", "hard_dcl"),
                                         ("euclidean", "fc")])
def test_contrastive_similarities_vs_oracle(sim, loss_fn):
    """dot / euclidean / edit similarities (losses.py:66-89) under the three losses."""
    from deepof_b200 import ContrastiveB200
    Tf, N, D, B = 24, 11, 8, 150
    adj = O.default_adjacency(N)
    r, c = np.nonzero(np.triu(adj))
    ei = np.stack([r, c], 1)
    x_full, _ = O.synthetic_windows(B, Tf, adj, seed=79)
    m = ContrastiveB200((Tf, N, 3), (Tf, len(r), 1), adj, D, temperature=0.2, similarity_function=sim, loss_function=loss_fn,
                        tau=0.1, beta=0.5, max_batch=B, seed=8)
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    graph = O.graph_operators(adj)
    rot = MO.rotation_table(ei, N)
    torch.manual_seed(9)
    prm = MO.draw_aug_params(B, Tf, N, MO.AugCfg(max_shift=3), rot)
    MO.SIMILARITY = sim
    try:
        logs, grads, out = MO.contrastive_train_step(x_full, p, graph, D, torch.from_numpy(ei), prm, 0.2, loss_fn=loss_fn,
                                                     tau_plus=0.1, beta=0.5)
    finally:
        MO.SIMILARITY = "cosine"
    m.loss_grad(x_full, _to_product_params(prm))
    got = m.logs_dict()
    for k, v in logs.items():
        assert abs(got[k] - v) <= 1e-4 * max(1.0, abs(v)), (k, got[k], v)
    bad = []
    _check_grads(m, {k: v for k, v in grads.items() if v is not None}, bad, f"{sim}-{loss_fn}")
    assert not bad, bad


def test_contrastive_euclidean_zero_distance_is_finite():
    """A window that draws no augmentation has distance 0 to its own view.  The reference's sqrt backward makes every
    gradient NaN there (losses.py:70-82; seen while generating contrastive_euclid.npz); the kernel takes the zero
    sub-gradient for that pair instead, so the step stays finite.  Documented deviation."""
    from deepof_b200 import AugParams, ContrastiveB200
    Tf, N, D, B = 24, 11, 6, 20
    adj = O.default_adjacency(N)
    x_full, _ = O.synthetic_windows(B, Tf, adj, seed=80)
    m = ContrastiveB200((Tf, N, 3), (Tf, int(np.count_nonzero(np.triu(adj))), 1), adj, D, similarity_function="euclidean",
                        max_batch=B, seed=3)
    base = (Tf - Tf // 2) // 2
    prm = AugParams(start=torch.full((B,), base, dtype=torch.int32, device="cuda"))      # the view IS the centre crop
    m.loss_grad(x_full, prm)
    logs = m.logs_dict()
    assert abs(logs["pos_similarity"] - 1.0) < 1e-6 and np.isfinite(logs["total_loss"])
    assert all(bool(torch.isfinite(v).all()) for v in m.grad_dict().values())


def test_contrastive_draw_augmentation_ranges():
    """The product's own draws (device RNG) respect the reference's ranges (training.py:2128-2366)."""
    from deepof_b200 import ContrastiveB200, ContrastiveAugCfg
    Tf, N, D, B = 50, 14, 8, 512
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    m = ContrastiveB200((Tf, N, 3), (Tf, E, 1), adj, D, max_batch=B, seed=1)
    cfg = ContrastiveAugCfg(p_rot=0.7, p_noise=1.0, p_interp=0.6, n_rot=3)
    g = torch.Generator(device="cuda").manual_seed(3)
    prm = m.draw_augmentation(B, cfg, generator=g, host_generator=torch.Generator().manual_seed(4))
    half, base = Tf // 2, (Tf - Tf // 2) // 2
    st = prm.start.cpu()
    assert int(st.min()) >= base - cfg.max_shift and int(st.max()) <= base + cfg.max_shift
    assert 0.6 < float((st != base).float().mean()) < 0.95                        # p_shift = 0.8
    L, t0 = prm.interp_len.cpu(), prm.interp_t0.cpu()
    on = L > 0
    assert 0.45 < float(on.float().mean()) < 0.75 and int(L[on].min()) >= cfg.min_interp and int(L.max()) <= cfg.max_interp
    assert int(t0.min()) >= 1 and int((t0 + L).max()) <= half - 1
    assert prm.rot_theta is not None and prm.rot_theta.shape[0] <= cfg.n_rot
    assert float(prm.rot_theta.abs().max()) <= np.pi / 6 + 1e-6
    assert prm.noise.shape == (B, N, 3) and float(prm.noise[..., :2].abs().min(dim=-1).values.max()) == 0.0   # one axis per node
    x_full, _ = O.synthetic_windows(B, Tf, adj, seed=8)
    m.loss_grad(x_full, prm)
    m.adam_step(1e-3)
    logs = m.logs_dict()
    assert np.isfinite(logs["total_loss"]) and -1.0 <= logs["neg_similarity"] <= logs["pos_similarity"] <= 1.0 + 1e-6
