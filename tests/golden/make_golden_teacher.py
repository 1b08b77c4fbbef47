"""Golden vectors for deepof_b200.teacher, produced by the UNMODIFIED reference function
`initialize_gmm_from_teacher` (deepof/clustering/teacher_model.py:394-460).  Run in the build container:

    python tests/golden/make_golden_teacher.py

Cases: (dense) every cluster has mass; (empty) one cluster gets ~zero teacher mass -> the global-moment guard;
(sharp) near one-hot assignments with a tiny-variance cluster -> the min_var clamp."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.load()
import deepof.clustering.teacher_model as TM  # noqa: E402


class _Latent(torch.nn.Module):
    def __init__(self, C, D):
        super().__init__()
        self.gmm_means = torch.nn.Parameter(torch.zeros(C, D))
        self.gmm_log_vars = torch.nn.Parameter(torch.zeros(C, D))
        self.register_buffer("prior", torch.full((C,), 1.0 / C))


class _Model(torch.nn.Module):
    def __init__(self, C, D):
        super().__init__()
        self.latent_space = _Latent(C, D)


def case(name, N, D, C, seed):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(C, D, generator=g) * 2.0
    lab = torch.randint(0, C, (N,), generator=g)
    z = centres[lab] + torch.randn(N, D, generator=g) * 0.5
    logits = torch.randn(N, C, generator=g)
    logits[torch.arange(N), lab] += 3.0
    if name == "empty":
        logits[:, 2] = -60.0                     # cluster 2: teacher mass ~ 1e-26 per window
    if name == "sharp":
        logits = logits * 8.0
        logits[lab != 1, 1] = -80.0              # no teacher mass leaks into cluster 1 from the other windows
        z[lab == 1] = centres[1] + torch.randn(int((lab == 1).sum()), D, generator=g) * 1e-3   # variance below min_var
    tau = torch.softmax(logits, 1)
    m = _Model(C, D)
    TM.initialize_gmm_from_teacher(m, z, tau)
    return {f"{name}/z": z.numpy(), f"{name}/tau": tau.numpy(),
            f"{name}/means": m.latent_space.gmm_means.detach().numpy(),
            f"{name}/log_vars": m.latent_space.gmm_log_vars.detach().numpy(),
            f"{name}/prior": m.latent_space.prior.numpy()}


if __name__ == "__main__":
    out = {}
    out.update(case("dense", 3000, 16, 8, 11))
    out.update(case("empty", 2000, 8, 4, 12))
    out.update(case("sharp", 2500, 16, 6, 13))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "teacher_gmm_init.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
