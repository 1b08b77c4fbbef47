"""Golden vectors for deepof_b200.teacher, produced by the UNMODIFIED reference functions
`initialize_gmm_from_teacher` (deepof/clustering/teacher_model.py:394-460) -> teacher_gmm_init.npz and
`run_turtle_teacher_on_views` (:710-792, TurtleTeacher :152-351) -> teacher_turtle.npz.  Run in the build container:

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


def main_gmm_init():
    out = {}
    out.update(case("dense", 3000, 16, 8, 11))
    out.update(case("empty", 2000, 8, 4, 12))
    out.update(case("sharp", 2500, 16, 6, 13))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "teacher_gmm_init.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def turtle_case(name, N, dims, K, seed, outer, inner, batch):
    """Reference `run_turtle_teacher_on_views` (teacher_model.py:710-792) on synthetic clustered views; the global
    generator is seeded right before the call, the test re-seeds it the same way."""
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, K, (N,), generator=g)
    views = {}
    for i, d in enumerate(dims):
        centres = torch.randn(K, d, generator=g) * 1.5
        views[f"view{i}"] = centres[lab] + torch.randn(N, d, generator=g) * 0.7
    torch.manual_seed(seed + 100)
    teacher, tau = TM.run_turtle_teacher_on_views(views, K, outer_steps=outer, inner_steps=inner, batch_size=batch, verbose=False)
    out = {f"{name}/tau_star": tau.numpy(), f"{name}/meta": np.array([N, K, seed, outer, inner, batch] + list(dims))}
    for i, d in enumerate(dims):
        out[f"{name}/view{i}"] = views[f"view{i}"].numpy()
        out[f"{name}/head{i}_w"] = teacher.heads.heads[i].weight.detach().numpy()
        out[f"{name}/head{i}_b"] = teacher.heads.heads[i].bias.detach().numpy()
        out[f"{name}/proj{i}_w"] = teacher.task_encoder.projs[i].weight.detach().numpy()
        out[f"{name}/proj{i}_b"] = teacher.task_encoder.projs[i].bias.detach().numpy()
    return out


if __name__ == "__main__":
    main_gmm_init()
    out = {}
    out.update(turtle_case("three_views", 2000, [16, 12, 8], 6, 21, outer=40, inner=30, batch=512))
    out.update(turtle_case("one_view", 1500, [10], 4, 22, outer=25, inner=20, batch=256))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "teacher_turtle.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if "tau" in k})


class _StubDataset:
    """What fit_nodes_pca / extract_pca_edges_view / fit_angles_pca need from BatchDictDataset: make_loader(batch_size=...)
    yielding (x, a[, angles]) batches in order (h5py is not installed here, the functions themselves run unmodified)."""
    return_angles = True

    def __init__(self, x, a, ang):
        self.x, self.a, self.ang = x, a, ang

    def make_loader(self, batch_size, **kw):
        return [(self.x[i:i + batch_size], self.a[i:i + batch_size], self.ang[i:i + batch_size])
                for i in range(0, self.x.shape[0], batch_size)]


def views_case():
    g = torch.Generator().manual_seed(31)
    N, T, Nn, E, A = 1100, 10, 6, 7, 4
    base = torch.randn(N, 1, Nn, 3, generator=g)
    x = base + 0.3 * torch.cumsum(torch.randn(N, T, Nn, 3, generator=g), 1)
    a = torch.randn(N, T, E, 1, generator=g) * torch.linspace(0.5, 2.0, E).view(1, 1, E, 1)
    ang = torch.randn(N, T, A, generator=g)
    ds = _StubDataset(x, a, ang)
    _, pos, _, spd = TM.fit_nodes_pca(ds, n_components_pos=12, n_components_spd=12, batch_size=400)
    edges = TM.extract_pca_edges_view(ds, n_components=8, batch_size=512)
    _, angles = TM.fit_angles_pca(ds, n_components=6, batch_size=300)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "teacher_views.npz")
    np.savez_compressed(path, x=x.numpy(), a=a.numpy(), ang=ang.numpy(), pca_pos=pos.numpy(), pca_spd=spd.numpy(),
                        pca_edges=edges.numpy(), pca_angles=angles.numpy())
    print("wrote", path, pos.shape, spd.shape, edges.shape, angles.shape)


if __name__ == "__main__":
    views_case()
