"""GPU: the TCN model family through the C-ABI vs
(1) the goldens produced by the UNMODIFIED reference — ``tcnmodel_*.npz`` (eval-mode outputs of VaDEPT / VQVAEPT / ContrastivePT
    built with encoder_type="TCN"), ``tcnvade_*.npz`` (step_vade: 13 logs, every parameter gradient, running statistics of all
    BatchNorm layers after the step) and ``tcnstep_*.npz`` (step_vqvae_distill / step_contrastive_distill);
(2) the CPU oracle on fresh seeds at sizes that take the tcgen05 GEMM paths (>= 2048 rows per convolution, >= 4096 per weight
    gradient).
Tolerances: embeddings / reconstructions rel-L2 <= 1e-4, logs |d| <= 1e-4 max(1, |v|), running statistics 1e-5.

Gradients.  The TCN step is CHAOTIC at fp32 resolution: 16 ReLUs behind train-mode BatchNorms per branch, and only the last step of
a sequence feeds the encoder output, so a single pre-activation within ~1e-6 of zero at such a position flips under ANY change of
the summation order and moves the gradient of everything upstream of it by 1e-3 .. 1e-2.  Measured: the reference's own fp32
gradient is 2e-4 .. 1.1e-3 away from an fp64 evaluation of the same step at B = 16 .. 192; one flipped unit of
encoder.edge_tcn.blocks.0.bn2 put 4 % into that channel's weight gradient while every tensor not upstream of it agreed to 1e-4;
over five seeds of the B = 192 case the CUDA path sat 1e-4 .. 4e-3 from fp64.  The golden cases therefore pass when the flat
gradient agrees to 5e-4 OR the disagreement is explained by isolated flips (median per-tensor error below 3e-4, flat below 2e-2);
against the oracle (B = 176 / 192, tcgen05 paths) the yardstick is the fp64 evaluation: the CUDA path must be within
max(1e-3, 5 x the fp32 CPU oracle's own distance from fp64).  The tcgen05 GEMMs are 3xTF32 with the correction terms accumulated
first and the remainder rounded to nearest (3.9e-7 relative error of a K = 128 product vs 1.3e-6 before)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from helpers import golden_cases_of, load_golden_of, sub, rel_l2

pytestmark = pytest.mark.gpu
SEED_OFF = int(os.environ.get("DOF_TEST_SEED_OFFSET", "0"))      # fresh-seed tests: shift every seed (sensitivity studies)

FLAT_TOL, P50_TOL, FLIP_FLAT_TOL = 5e-4, 3e-4, 2e-2


def _grad_cmp(gd, ref, names):
    num = sum(float((gd[k].cpu().double() - ref[k].double()).pow(2).sum()) for k in names)
    den = sum(float(ref[k].double().pow(2).sum()) for k in names)
    return (num / max(den, 1e-300)) ** 0.5


def _grad_check(gd, ref, names, tag, strict=True):
    """Flat error and the distribution of per-tensor relative errors (tensors whose reference norm is not negligible)."""
    flat = _grad_cmp(gd, ref, names)
    scale = max(float(ref[k].double().norm()) for k in names)
    per = sorted((rel_l2(gd[k].cpu(), ref[k]), k) for k in names if float(ref[k].double().norm()) > 1e-4 * scale)
    vals = np.array([v for v, _ in per])
    p50, p75 = float(np.percentile(vals, 50)), float(np.percentile(vals, 75))
    print(tag, "grad flat %.2e  per tensor p50 %.2e p75 %.2e max %.2e (%s)" % (flat, p50, p75, per[-1][0], per[-1][1]))
    if strict:
        assert flat < FLAT_TOL or (p50 < P50_TOL and flat < FLIP_FLAT_TOL), (tag, flat, p50, p75, per[-3:])
    return flat


def _check_running(m, g):
    """dof_clip_adam with lr 0 moves the running buffers of every BatchNorm (and nothing else)."""
    before = m.state.clone()
    m.adam_step(0.0, 0.0) if hasattr(m, "latent_space") else m.adam_step(0.0)
    sd = m.state_dict()
    n = 0
    for k in g:
        if k.startswith("p1/"):
            n += 1
            if k.endswith("num_batches_tracked"):
                assert int(sd[k[3:]]) == int(g[k]), k
            else:
                assert rel_l2(sd[k[3:]].cpu(), g[k]) < 1e-5, (k, rel_l2(sd[k[3:]].cpu(), g[k]))
    assert n > 0
    changed = (m.state != before).nonzero().flatten().cpu().numpy()
    ok = np.zeros(m.state.numel(), bool)
    for k, off, num, *_ in m.layout:
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            ok[off:off + num] = True
    assert ok[changed].all()


@pytest.mark.parametrize("case", golden_cases_of("tcnmodel"))
def test_tcn_models_eval_vs_reference_golden(case):
    from deepof_b200 import ContrastiveB200, VaDEB200, VQVAEB200
    g = load_golden_of("tcnmodel", case)
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    kind = str(g["model"])
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    if kind == "vade":
        m = VaDEB200((T, N, 3), (T, E, 1), g["adjacency"], D, K, encoder_type="TCN", max_batch=16, training=False)
        assert [k for k, *_ in m.layout] == [k[2:] for k in g if k.startswith("p/")]       # the reference's state_dict order
        m.load_state_dict(sub(g, "p/"))
        enc, emb, q, loc = m.forward_eval(x, a)
        assert rel_l2(enc.cpu(), g["eval/enc"]) < 1e-4
        assert rel_l2(emb.cpu(), g["eval/emb"]) < 1e-4 and rel_l2(q.cpu(), g["eval/q"]) < 1e-4
        assert torch.equal(q.cpu().argmax(1), torch.from_numpy(g["eval/q"]).argmax(1))
        assert rel_l2(loc.cpu(), g["eval/loc"]) < 1e-4
    elif kind == "vqvae":
        m = VQVAEB200((T, N, 3), (T, E, 1), g["adjacency"], D, K, encoder_type="TCN", max_batch=16, training=False)
        assert [k for k, *_ in m.layout] == [k[2:] for k in g if k.startswith("p/")]
        m.load_state_dict(sub(g, "p/"))
        enc, quant, soft, idx, lq, le = m.forward_eval(x, a)
        assert rel_l2(enc.cpu(), g["eval/emb"]) < 1e-4 and rel_l2(soft.cpu(), g["eval/q"]) < 1e-4
        assert rel_l2(quant.cpu(), g["eval/quant"]) < 1e-5
        assert rel_l2(lq.cpu(), g["eval/loc_q"]) < 1e-4 and rel_l2(le.cpu(), g["eval/loc"]) < 1e-4
    else:
        m = ContrastiveB200((T, N, 3), (T, E, 1), g["adjacency"], D, encoder_type="TCN", max_batch=16, training=False)
        assert [k for k, *_ in m.layout] == [k[2:] for k in g if k.startswith("p/")]
        m.load_state_dict(sub(g, "p/"))
        assert rel_l2(m(x, a).cpu(), g["eval/emb"]) < 1e-4


@pytest.mark.parametrize("case", golden_cases_of("tcnvade"))
def test_vade_tcn_step_vs_reference_golden(case):
    from deepof_b200 import VaDEB200, VadeLossCfg
    from deepof_b200._lib import LOG_KEYS
    g = load_golden_of("tcnvade", case)
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    m = VaDEB200((T, N, 3), (T, E, 1), g["adjacency"], D, K, encoder_type="TCN", max_batch=B, training=True, seed=1)
    assert [k for k, *_ in m.layout] == [k[2:] for k in g if k.startswith("p/")]
    m.load_state_dict(sub(g, "p/"))
    main = str(g["phase"]) == "main"
    cfg = (VadeLossCfg.main_defaults if main else VadeLossCfg.pretrain_defaults)(K, kl_weight=float(g["klw"]))
    m.set_pretrain_mode(not main)
    m.loss_grad(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]), cfg, eps=torch.from_numpy(g["eps"]),
                mc_eps=torch.from_numpy(g["mc_eps"]) if main else None)
    logs = m.logs_dict()
    for k in LOG_KEYS:
        ref = float(g["log/" + k])
        assert abs(logs[k] - ref) <= 1e-4 * max(1.0, abs(ref)), (k, logs[k], ref)
    gd = m.grad_dict()
    gnames = [k[2:] for k in g if k.startswith("g/")]
    ref = {k: torch.from_numpy(g["g/" + k]) for k in gnames}
    _grad_check(gd, ref, gnames, case)
    for k, off, n, shape, grp in m.layout:        # parameters the reference leaves without a gradient stay at zero
        if k not in gnames:
            assert float(m.grad[off:off + n].abs().max()) == 0.0, k
    _check_running(m, g)


@pytest.mark.parametrize("case", golden_cases_of("tcnstep"))
def test_vqvae_and_contrastive_tcn_steps_vs_reference_golden(case):
    """step_vqvae_distill (two decoder passes, each with its own batch statistics) and step_contrastive_distill (two encoder passes =
    two statistics groups over 2B windows) of the TCN models vs the reference."""
    from deepof_b200 import ContrastiveB200, VQVAEB200, _lib
    from deepof_b200.models import CON_LOG_KEYS, VQ_LOG_KEYS
    g = load_golden_of("tcnstep", case)
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    gnames = [k[2:] for k in g if k.startswith("g/")]
    ref = {k: torch.from_numpy(g["g/" + k]) for k in gnames}
    if str(g["model"]) == "vqvae":
        m = VQVAEB200((T, N, 3), (T, E, 1), g["adjacency"], D, K, encoder_type="TCN", beta=float(g["beta"]), max_batch=B, training=True, seed=1)
        assert [k for k, *_ in m.layout] == [k[2:] for k in g if k.startswith("p/")]
        m.load_state_dict(sub(g, "p/"))
        m.loss_grad(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
        keys = VQ_LOG_KEYS
    else:
        m = ContrastiveB200((T, N, 3), (T, E, 1), g["adjacency"], D, encoder_type="TCN", max_batch=B, training=True, seed=1)
        assert [k for k, *_ in m.layout] == [k[2:] for k in g if k.startswith("p/")]
        m.load_state_dict(sub(g, "p/"))
        x2 = torch.cat([torch.from_numpy(g["x"]), torch.from_numpy(g["x_aug"])]).cuda().contiguous()
        a2 = torch.cat([torch.from_numpy(g["a"]), torch.from_numpy(g["a_aug"])]).cuda().contiguous()
        _lib.check(m.L.dof_contrastive_loss_grad(m.handle, _lib.ptr(m.state), _lib.ptr(m.grad), _lib.ptr(x2), _lib.ptr(a2), B, 0, 0, 0.1, 0.1, 0.1,
                                                 _lib.ptr(m.logs), None, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        keys = CON_LOG_KEYS
    logs = m.logs_dict()
    for k in keys:
        r = float(g["log/" + k])
        assert abs(logs[k] - r) <= 1e-4 * max(1.0, abs(r)), (k, logs[k], r)
    _grad_check(m.grad_dict(), ref, gnames, case)
    _check_running(m, g)


@pytest.mark.parametrize("geom", ["cfg2", "small_latent"])
def test_vade_tcn_step_vs_oracle_tensor_core_sizes(geom):
    """Fresh seeds, B large enough that the dilated convolutions, their input gradients (>= 2048 rows) and weight gradients
    (>= 4096 rows) take the tcgen05 kernels: logs, embeddings and the flat gradient against the oracle.  "small_latent" has
    4 D = 24 != 64 decoder input channels: the 1x1 residual projection of the decoder's first block."""
    from deepof_b200 import VaDEB200, VadeLossCfg
    from deepof_b200._lib import LOG_KEYS
    from oracle import tcn_oracle as TC
    from oracle import vade_oracle as O
    T, N = 25, 14
    D, K, B = (16, 8, 192) if geom == "cfg2" else (6, 5, 176)
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = O.synthetic_windows(B, T, adj, seed=177 + SEED_OFF)
    m = VaDEB200((T, N, 3), (T, E, 1), adj, D, K, encoder_type="TCN", max_batch=B, training=True, seed=21 + SEED_OFF)
    pg = torch.Generator().manual_seed(27)
    with torch.no_grad():
        m.latent_space.gmm_means.mul_(3.0)
        for k, v in m._views.items():                       # move the affine parameters off their init so that they matter
            if k.endswith("bias") and v.dim() == 1:
                v.add_(0.05 * torch.randn(v.shape, generator=pg).to(v.device))
            if ".bn" in k and k.endswith("weight"):
                v.add_(0.1 * torch.randn(v.shape, generator=pg).to(v.device))
    m.set_pretrain_mode(False)
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(13)
    eps, mc = torch.randn(B, D, generator=gen), torch.randn(32, B, D, generator=gen)
    cfg = VadeLossCfg.main_defaults(K, kl_weight=0.6)
    m.loss_grad(x, a, cfg, eps=eps, mc_eps=mc)
    logs = m.logs_dict()
    ocfg = O.LossCfg.main_defaults(K, kl_weight=0.6)
    ologs, ograds, oo = TC.vade_train_step(x, a, p, O.graph_operators(adj), ocfg, eps, mc_eps=mc)
    assert rel_l2(m.debug("enc")[:B * D].cpu(), oo["enc"]) < 1e-4
    for k in LOG_KEYS:
        assert abs(logs[k] - ologs[k]) <= 1e-4 * max(1.0, abs(ologs[k])), (k, logs[k], ologs[k])
    # the truth is the fp64 evaluation of the same step: at this size two fp32 evaluations of the TCN gradient (ReLU masks behind
    # 16 train-mode BatchNorms per branch) differ by ~1e-3, so the CUDA path must be as close to fp64 as the fp32 CPU oracle is
    p64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in p.items()}
    _, g64, _ = TC.vade_train_step(x.double(), a.double(), p64, tuple(t.double() for t in O.graph_operators(adj)), ocfg, eps.double(),
                                   mc_eps=mc.double())
    names = [k for k, v in g64.items() if v is not None]
    err = _grad_check(m.grad_dict(), g64, names, geom + " vs fp64", strict=False)
    err32 = _grad_cmp({k: v for k, v in ograds.items() if v is not None}, g64, names)
    print(geom, "fp32 oracle vs fp64", err32)
    assert err < max(1e-3, 5.0 * err32) and err < FLIP_FLAT_TOL, (err, err32)
    run = TC.running_after(p, oo["bn"])
    m.adam_step(0.0, 0.0)
    sd = m.state_dict()
    for k, v in run.items():
        if not k.endswith("num_batches_tracked"):
            assert rel_l2(sd[k].cpu(), v) < 1e-5, k


def test_tcn_trainers_run_and_learn():
    """The three trainers accept encoder_type="TCN": a few steps each, finite logs, the VaDE pretraining loss goes down."""
    from deepof_b200.training import ContrastiveTrainer, VaDETrainer, VQVAETrainer
    from oracle import vade_oracle as O
    T, N, D, K, B = 24, 11, 8, 4, 32
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = O.synthetic_windows(B, T, adj, seed=9)
    tr = VaDETrainer((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, seed=3, encoder_type="TCN")
    tr.set_phase("pretrain", kl_weight=0.0, lr_base=1e-3)
    losses = [float(tr.train_step_device(x.cuda(), a.cuda())[0]) for _ in range(15)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
    for cls in (VQVAETrainer, ContrastiveTrainer):
        t2 = cls((T, N, 3), (T, E, 1), adj, D, *( (K,) if cls is VQVAETrainer else () ), max_batch=B, seed=3, encoder_type="TCN")
        vals = [float(t2.train_step_device(x.cuda(), a.cuda())[0]) for _ in range(3)]
        assert np.isfinite(vals).all(), (cls.__name__, vals)
