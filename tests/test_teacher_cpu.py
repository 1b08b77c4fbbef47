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
