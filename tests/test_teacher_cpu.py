"""deepof_b200.teacher.initialize_gmm_from_teacher against vectors produced by the UNMODIFIED reference function
(teacher_model.py:394-460; generator tests/golden/make_golden_teacher.py).  Host-side torch code: runs on CPU here and
on device tensors (the views of VaDEB200's flat state) on the GPU box."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from deepof_b200.teacher import gmm_moments_from_teacher, initialize_gmm_from_teacher

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "teacher_gmm_init.npz"))


def _stub_model(C, D):
    ls = SimpleNamespace(gmm_means=torch.zeros(C, D), gmm_log_vars=torch.zeros(C, D), prior=torch.full((C,), 1.0 / C))
    return SimpleNamespace(latent_space=ls)


@pytest.mark.parametrize("case", ["dense", "empty", "sharp"])
def test_gmm_init_matches_reference(case):
    z, tau = torch.from_numpy(G[f"{case}/z"]), torch.from_numpy(G[f"{case}/tau"])
    C, D = tau.shape[1], z.shape[1]
    m = _stub_model(C, D)
    initialize_gmm_from_teacher(m, z, tau, verbose=False)
    means, lv, prior = (torch.from_numpy(G[f"{case}/{k}"]) for k in ("means", "log_vars", "prior"))
    # the reference accumulates in fp32 (two-pass variance); here fp64 moments rounded once
    assert float((m.latent_space.gmm_means - means).abs().max()) <= 2e-5 * float(means.abs().max())
    assert float((m.latent_space.gmm_log_vars - lv).abs().max()) <= 1e-4
    assert float((m.latent_space.prior - prior).abs().max()) <= 1e-6
    if case == "empty":        # cluster 2 has no teacher mass: global moments of z (teacher_model.py:439-443)
        assert torch.allclose(m.latent_space.gmm_means[2], z.mean(0), atol=1e-5)
    if case == "sharp":        # cluster 1's variance is below min_var: clamped
        assert torch.allclose(m.latent_space.gmm_log_vars[1], torch.full((D,), float(np.log(1e-4))), atol=1e-4)


def test_gmm_init_chunking_is_exact_enough_and_parameters_are_written_in_place():
    z, tau = torch.from_numpy(G["dense/z"]), torch.from_numpy(G["dense/tau"])
    a = gmm_moments_from_teacher(z, tau)
    b = gmm_moments_from_teacher(z, tau, chunk=257)          # ragged last chunk
    for u, v in zip(a, b):
        assert float((u - v).abs().max()) < 1e-12
    ls = torch.nn.Module()
    ls.gmm_means = torch.nn.Parameter(torch.zeros(8, 16))
    ls.gmm_log_vars = torch.nn.Parameter(torch.zeros(8, 16))
    ls.register_buffer("prior", torch.zeros(8))
    ptr = ls.gmm_means.data_ptr()
    initialize_gmm_from_teacher(SimpleNamespace(latent_space=ls), z, tau, verbose=False)
    assert ls.gmm_means.data_ptr() == ptr and float(ls.gmm_means.abs().sum()) > 0 and abs(float(ls.prior.sum()) - 1.0) < 1e-6


def test_gmm_init_rejects_mismatched_inputs():
    with pytest.raises(ValueError):
        gmm_moments_from_teacher(torch.zeros(10, 4), torch.zeros(9, 3))


# ---- TURTLE teacher: batched closed-form heads vs the reference's per-view autograd loops -------------------------
from deepof_b200.teacher import TurtleTeacherB200, run_turtle_teacher_on_views  # noqa: E402

T = np.load(os.path.join(os.path.dirname(__file__), "golden", "teacher_turtle.npz"))


def _turtle_inputs(name):
    meta = [int(v) for v in T[f"{name}/meta"]]
    N, K, seed, outer, inner, batch = meta[:6]
    dims = meta[6:]
    views = {f"view{i}": torch.from_numpy(T[f"{name}/view{i}"]) for i in range(len(dims))}
    return views, dims, K, seed, outer, inner, batch


@pytest.mark.parametrize("name", ["three_views", "one_view"])
def test_turtle_teacher_matches_reference(name):
    """Same seed -> same initial heads / task encoder and the same shuffled batches as `run_turtle_teacher_on_views`
    (teacher_model.py:710-792); tau* and every fitted parameter agree to fp32 reassociation (measured 2e-7)."""
    views, dims, K, seed, outer, inner, batch = _turtle_inputs(name)
    torch.manual_seed(seed + 100)
    teacher, tau = run_turtle_teacher_on_views(views, K, outer_steps=outer, inner_steps=inner, batch_size=batch, verbose=False)
    ref = torch.from_numpy(T[f"{name}/tau_star"])
    assert tau.shape == ref.shape
    assert float((tau - ref).abs().max()) < 1e-5
    assert torch.equal(tau.argmax(1), ref.argmax(1))
    assert float((tau.sum(1) - 1).abs().max()) < 1e-5
    for i, d in enumerate(dims):
        assert float((teacher.Wh[i, :, :d] - torch.from_numpy(T[f"{name}/head{i}_w"])).abs().max()) < 1e-5
        assert float((teacher.bh[i] - torch.from_numpy(T[f"{name}/head{i}_b"])).abs().max()) < 1e-5
        assert float((teacher.Wp[i, :, :d].detach() - torch.from_numpy(T[f"{name}/proj{i}_w"])).abs().max()) < 1e-5
        assert float((teacher.bp[i].detach() - torch.from_numpy(T[f"{name}/proj{i}_b"])).abs().max()) < 1e-5
        # zero padding of narrower views never leaks into the parameters
        assert float(teacher.Wh[i, :, d:].abs().sum()) == 0.0 and float(teacher.Wp[i, :, d:].detach().abs().sum()) == 0.0


def test_turtle_none_views_are_skipped_and_empty_is_an_error():
    views, dims, K, seed, outer, inner, batch = _turtle_inputs("one_view")
    torch.manual_seed(1)
    _, tau_a = run_turtle_teacher_on_views({"nodes": views["view0"], "angles": None}, K, outer_steps=3, inner_steps=2,
                                           batch_size=batch, verbose=False)
    torch.manual_seed(1)
    _, tau_b = run_turtle_teacher_on_views(views, K, outer_steps=3, inner_steps=2, batch_size=batch, verbose=False)
    assert torch.equal(tau_a, tau_b)
    with pytest.raises(AssertionError):
        run_turtle_teacher_on_views({"nodes": None}, K)


def test_turtle_teacher_feeds_gmm_init():
    """tau* -> initialize_gmm_from_teacher: the plumbing of the VaDE main phase (training.py teacher branch)."""
    views, dims, K, seed, outer, inner, batch = _turtle_inputs("one_view")
    ref = torch.from_numpy(T["one_view/tau_star"])
    z = views["view0"][:, :8].contiguous()
    m = _stub_model(K, 8)
    initialize_gmm_from_teacher(m, z, ref, verbose=False)
    assert torch.isfinite(m.latent_space.gmm_means).all() and torch.isfinite(m.latent_space.gmm_log_vars).all()
    assert abs(float(m.latent_space.prior.sum()) - 1.0) < 1e-6


# ---- the teacher's PCA views ---------------------------------------------------------------------------------------
from deepof_b200.teacher import IncrementalPCAB200, pca_view, teacher_views_from_windows  # noqa: E402

VW = np.load(os.path.join(os.path.dirname(__file__), "golden", "teacher_views.npz"))


def test_teacher_views_match_reference_functions():
    """fit_nodes_pca / extract_pca_edges_view / fit_angles_pca (teacher_model.py:464-708, run unmodified on a stub
    dataset by the golden script) vs the device restatement; the reference runs LAPACK in float32, so 2e-4 of scale."""
    x, a, ang = (torch.from_numpy(VW[k]) for k in ("x", "a", "ang"))
    v = teacher_views_from_windows(x, a, ang, pca_nodes_dim=12, pca_edges_dim=8, pca_angles_dim=6, batch_size_nodes=400,
                                   batch_size_edges=512, batch_size_angles=300)
    assert v["z"] is None
    for k in ("pca_pos", "pca_spd", "pca_edges", "pca_angles"):
        ref = torch.from_numpy(VW[k])
        assert v[k].shape == ref.shape and v[k].dtype == torch.float32
        assert float((v[k] - ref).abs().max()) <= 2e-4 * float(ref.abs().max()), k
    off = teacher_views_from_windows(x, a, None, include_nodes_view=False, pca_edges_dim=8, batch_size_edges=512)
    assert off["pca_pos"] is None and off["pca_spd"] is None and off["pca_angles"] is None and off["pca_edges"] is not None


def test_incremental_pca_is_sklearns_algorithm_in_float64():
    """Pins the restated update to the third-party implementation the reference calls (scikit-learn IncrementalPCA):
    float64 in, same batch partition -> components, singular values, mean and scores agree to 1e-9."""
    sk = pytest.importorskip("sklearn.decomposition")
    g = np.random.default_rng(5)
    X = g.normal(size=(900, 40)) @ g.normal(size=(40, 40)) + g.normal(size=(1, 40)) * 3.0
    ref = sk.IncrementalPCA(n_components=9)
    mine = IncrementalPCAB200(9)
    for i in range(0, 900, 250):                       # ragged last batch (150 rows)
        ref.partial_fit(X[i:i + 250])
        mine.partial_fit(torch.from_numpy(X[i:i + 250]))
    assert np.abs(mine.components_.numpy() - ref.components_).max() < 1e-9
    assert np.abs(mine.singular_values_.numpy() - ref.singular_values_).max() < 1e-8
    assert np.abs(mine.mean_.numpy() - ref.mean_).max() < 1e-12 and mine.n_samples_seen_ == ref.n_samples_seen_
    Z = ((torch.from_numpy(X) - mine.mean_) @ mine.components_.t()).numpy()
    assert np.abs(Z - ref.transform(X)).max() < 1e-8


def test_pca_view_max_samples_and_errors():
    X = torch.from_numpy(VW["a"]).reshape(1100, -1)
    p_all, _ = pca_view(X, 8, 512)
    p_cut, f_cut = pca_view(X, 8, 512, max_samples=700)          # second batch truncated to 188 rows, third skipped
    assert p_all.n_samples_seen_ == 1100 and p_cut.n_samples_seen_ == 700 and f_cut.shape == (1100, 8)
    with pytest.raises(ValueError):
        IncrementalPCAB200(50).partial_fit(torch.zeros(20, 100))  # fewer rows than components (sklearn raises too)
    with pytest.raises(ValueError):
        teacher_views_from_windows(torch.zeros(4, 5, 3, 2), torch.zeros(4, 5, 2, 1))


def test_build_turtle_teacher_end_to_end():
    """windows -> PCA views -> teacher -> tau* -> GMM init, all on one device; view order and defaults of the reference."""
    from deepof_b200.teacher import build_turtle_teacher
    x, a = torch.from_numpy(VW["x"]), torch.from_numpy(VW["a"])
    z = torch.from_numpy(VW["pca_pos"])[:, :8].contiguous()           # stands in for the model's latent view
    with pytest.raises(ValueError):
        build_turtle_teacher(x, a, 5)                                 # latent view is on by default, like the reference
    torch.manual_seed(3)
    teacher, tau, views = build_turtle_teacher(x, a, 5, latent_view=z, pca_nodes_dim=12, batch_size_nodes=400,
                                               teacher_outer_steps=6, teacher_inner_steps=5, teacher_batch_size=256,
                                               verbose=False)
    assert list(views) == ["z", "pca_pos", "pca_spd", "pca_edges", "pca_angles"]
    assert views["pca_edges"] is None and views["pca_angles"] is None and teacher.dims == [8, 12, 12]
    assert tau.shape == (1100, 5) and float((tau.sum(1) - 1).abs().max()) < 1e-5 and not tau.requires_grad
    m = _stub_model(5, 8)
    initialize_gmm_from_teacher(m, z, tau, verbose=False)
    assert torch.isfinite(m.latent_space.gmm_log_vars).all()


def test_pca_view_folds_a_short_tail_into_the_previous_batch():
    X = torch.from_numpy(VW["a"]).reshape(1100, -1)[:1029]           # batches of 512: 512 + 512 + 5 rows, 8 components
    p, f = pca_view(X, 8, 512)
    assert p.n_samples_seen_ == 1029 and f.shape == (1029, 8)
    ref = IncrementalPCAB200(8).partial_fit(X[:512]).partial_fit(X[512:])
    assert float((p.components_ - ref.components_).abs().max()) == 0.0


def test_teacher_context_matches_reference_set_teacher():
    """VadeLoss.set_teacher (losses.py:460-491) called unbound on a bare namespace in the build container; elsewhere the
    closed form is checked."""
    from deepof_b200.teacher import teacher_context
    tau = torch.from_numpy(T["three_views/tau_star"])
    tc = teacher_context(tau, True, 1.0, 3.0)
    pi = tau.mean(0).clamp_min(1e-8)
    w = pi.pow(-1.0)
    w = (w / w.mean()).clamp_max(3.0)
    assert torch.allclose(tc["class_weight"], w) and torch.allclose(tc["teacher_marginal"], pi)
    assert teacher_context(tau, False)["class_weight"] is None
    from oracle import refshim
    if refshim.available():
        _, L, _, _ = refshim.load()
        ns = SimpleNamespace(distill_use_class_reweight=True, distill_class_reweight_beta=1.0, distill_class_reweight_cap=3.0)
        L.VadeLoss.set_teacher(ns, tau, 4.0, None)
        assert torch.equal(ns.class_weight, tc["class_weight"]) and torch.equal(ns.teacher_marginal, tc["teacher_marginal"])


def test_lambda_schedule_is_the_reference_weight_manager():
    """The distillation weight of the VaDE main phase (training.py:1672-1676): KLSchedule with warm-up 0, flat for
    lambda_decay_start epochs, linear cool-down — against Dynamic_weight_manager in the build container."""
    from deepof_b200.training import KLSchedule
    nb = 7
    mine = KLSchedule(nb, "tf_sigmoid", 0, 4.0, 10, 0.2, at_max_epochs=10)
    vals = []
    for _ in range(nb * 22):
        vals.append(mine.get_weight())
        mine.step()
    assert vals[0] < 1e-30 and vals[1] == 4.0 and vals[nb * 10] == 4.0 and abs(vals[-1] - 0.2) < 1e-12
    assert all(b <= a + 1e-12 for a, b in zip(vals[1:], vals[2:]))            # flat, then monotone decay
    from oracle import refshim
    if refshim.available():
        _, L, _, _ = refshim.load()
        ref = L.Dynamic_weight_manager(nb, mode="tf_sigmoid", warmup_epochs=0, at_max_epochs=10, max_weight=4.0,
                                       cooldown_epochs=10, end_weight=0.2)
        for v in vals:
            assert abs(ref.get_weight() - v) < 1e-12
            ref.step()
