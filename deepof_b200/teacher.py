"""Teacher-side plumbing of the VaDE main phase that sits next to the step path.

`initialize_gmm_from_teacher` mirrors the reference function of the same name
(deepof/clustering/teacher_model.py:394-460): moment-match the mixture to the teacher's soft
assignments tau* over the training-set embeddings and write the result into
``model.latent_space.{gmm_means, gmm_log_vars, prior}`` (for a `VaDEB200` these are views of
the flat state buffer the kernels read, so no copy-back step exists).

The reference materialises the [N, C, D] tensor of deviations; here the three weighted moments
(sum tau, sum tau z, sum tau z^2) are accumulated in float64 over chunks of windows, so N = 10 M
embeddings (SURVEY cfg5) need no 5 GB intermediate.  The TURTLE teacher itself (tau*) is an input.
"""
from typing import Optional

import torch


@torch.no_grad()
def gmm_moments_from_teacher(z_all: torch.Tensor, tau_star: torch.Tensor, min_var: float = 1e-4,
                             min_mass: float = 1e-6, chunk: int = 1 << 20):
    """(means [C,D], log_vars [C,D], prior [C]) as teacher_model.py:425-448 computes them (float64 internally)."""
    if z_all.dim() != 2 or tau_star.dim() != 2 or z_all.shape[0] != tau_star.shape[0]:
        raise ValueError(f"z_all [N,D] and tau_star [N,C] expected, got {tuple(z_all.shape)} and {tuple(tau_star.shape)}")
    dev = z_all.device
    N, D = z_all.shape
    C = tau_star.shape[1]
    s0 = torch.zeros(C, dtype=torch.float64, device=dev)
    s1 = torch.zeros(C, D, dtype=torch.float64, device=dev)
    zsum = torch.zeros(D, dtype=torch.float64, device=dev)
    for i in range(0, N, chunk):
        z = z_all[i:i + chunk].to(torch.float64)
        t = tau_star[i:i + chunk].to(device=dev, dtype=torch.float64)
        s0 += t.sum(0)
        s1 += t.t() @ z
        zsum += z.sum(0)
    mass = s0 + min_mass                                              # :426
    prior = (mass / mass.sum()).clamp(min=1e-8, max=1.0)              # :427
    means = s1 / mass.unsqueeze(1)                                    # :430
    # second pass for the centred second moment (the reference's two-pass variance, :433-434)
    s2 = torch.zeros(C, D, dtype=torch.float64, device=dev)
    gvar = torch.zeros(D, dtype=torch.float64, device=dev)
    gmean = zsum / max(N, 1)
    for i in range(0, N, chunk):
        z = z_all[i:i + chunk].to(torch.float64)
        t = tau_star[i:i + chunk].to(device=dev, dtype=torch.float64)
        # sum_i tau_ic (z_id - mu_cd)^2 = sum tau z^2 - 2 mu sum tau z + mu^2 sum tau, per chunk with the FINAL means
        s2 += t.t() @ (z * z) - 2.0 * means * (t.t() @ z) + means * means * t.sum(0).unsqueeze(1)
        gvar += ((z - gmean) ** 2).sum(0)
    vars_ = (s2 / mass.unsqueeze(1)).clamp(min=min_var)               # :434-435
    log_vars = vars_.log()
    tiny = mass <= 1e-4                                               # :439-443: empty clusters fall back to the global moments
    if bool(tiny.any()):
        means[tiny] = gmean
        log_vars[tiny] = (gvar / max(N, 1)).clamp(min=min_var).log()
    return means, log_vars, prior


@torch.no_grad()
def initialize_gmm_from_teacher(model, z_all: torch.Tensor, tau_star: torch.Tensor, min_var: float = 1e-4,
                                min_mass: float = 1e-6, verbose: bool = True) -> None:
    """Same signature and effect as teacher_model.py:394 (writes gmm_means, gmm_log_vars and prior in place)."""
    ls = model.latent_space
    tgt = ls.gmm_means
    z = z_all.to(device=tgt.device)
    means, log_vars, prior = gmm_moments_from_teacher(z, tau_star.to(tgt.device), min_var, min_mass)
    data = lambda p: p.data if isinstance(p, torch.nn.Parameter) else p
    data(ls.gmm_means).copy_(means.to(tgt.dtype))
    data(ls.gmm_log_vars).copy_(log_vars.to(tgt.dtype))
    pr: Optional[torch.Tensor] = getattr(ls, "prior", None)
    if pr is not None:
        data(pr).copy_(prior.to(pr.dtype))
    if verbose:
        ent = float(-(prior * prior.clamp_min(1e-9).log()).sum())
        print("Initialized GMM from teacher τ*: "
              f"mean |μ|={float(means.norm(dim=1).mean()):.3f}, mean σ²={float(log_vars.exp().mean()):.5f}, entropy(π)={ent:.3f}")


# ---------------------------------------------------------------------------------------------------------------------
# TURTLE teacher (teacher_model.py:43-351, 710-792; Gadetsky et al., arXiv 2406.07236) on device-resident views.
#
# The reference runs V x M tiny autograd graphs per outer step (V views, M = 100-200 inner SGD steps): ~25 kernel
# launches per head and step.  Here the V per-view heads are ONE batched problem (views zero-padded to a common width,
# [V, B, Dmax] x [V, Dmax, K] bmm) and the inner SGD uses the closed-form gradient of the soft cross-entropy of a linear
# head, so an inner step is 6 launches for all views; only the single task-encoder update per outer step goes through
# autograd.  Batch order and initial weights consume torch's global generator exactly like the reference (same
# nn.Linear construction order, a DataLoader over window INDICES with the same batch size / shuffle / drop_last), so
# with the same seed tau* agrees with the reference up to fp32 reassociation.
# ---------------------------------------------------------------------------------------------------------------------
class TurtleTeacherB200:
    """Mirror of ``TurtleTeacher`` (teacher_model.py:152-351): ``fit(batches)``, ``predict(batches)``."""

    def __init__(self, feature_dims, n_components: int, gamma: float = 10.0, alpha_sample_entropy: float = 0.1,
                 inner_lr: float = 0.1, inner_steps: int = 100, head_wd: float = 1e-4, head_temp: float = 0.5,
                 task_temp: float = 0.5, normalize_feats: bool = True, lr_theta: float = 5e-3,
                 delta_death_barrier: float = 40.0, device="cpu"):
        self.dims = [int(d) for d in feature_dims]
        self.V, self.K, self.Dmax = len(self.dims), int(n_components), max(self.dims)
        self.gamma, self.alpha, self.delta = float(gamma), float(alpha_sample_entropy), float(delta_death_barrier)
        self.inner_lr, self.M, self.head_wd = float(inner_lr), int(inner_steps), float(head_wd)
        self.head_temp, self.task_temp, self.normalize_feats = float(head_temp), float(task_temp), bool(normalize_feats)
        self.device = torch.device(device)
        # same construction order as the reference (heads first, then the task encoder) -> same draws from the global RNG
        heads = [torch.nn.Linear(d, self.K) for d in self.dims]
        projs = [torch.nn.Linear(d, self.K) for d in self.dims]
        self.Wh, self.bh = self._stack(heads)                   # [V, K, Dmax], [V, K]   per-view heads (inner SGD)
        Wp, bp = self._stack(projs)
        self.Wp = Wp.requires_grad_(True)                       # task encoder (outer Adam)
        self.bp = bp.requires_grad_(True)
        self.opt_theta = torch.optim.Adam([self.Wp, self.bp], lr=lr_theta)

    def _stack(self, linears):
        W = torch.zeros(self.V, self.K, self.Dmax, device=self.device)
        b = torch.zeros(self.V, self.K, device=self.device)
        with torch.no_grad():
            for v, lin in enumerate(linears):
                W[v, :, :self.dims[v]] = lin.weight.to(self.device)
                b[v] = lin.bias.to(self.device)
        return W, b

    def _pack(self, feats_list):
        """[V, B, Dmax] zero-padded float32 stack of one batch of views."""
        B = feats_list[0].shape[0]
        X = torch.zeros(self.V, B, self.Dmax, device=self.device)
        for v, f in enumerate(feats_list):
            X[v, :, :self.dims[v]] = f.to(self.device, non_blocking=True).float()
        return X

    def _tau(self, X):
        logits = (torch.baddbmm(self.bp.unsqueeze(1), X, self.Wp.transpose(1, 2)) / self.task_temp).sum(0)
        return torch.softmax(logits / max(self.V, 1), dim=-1)               # teacher_model.py:132-150

    @torch.no_grad()
    def _inner_fit(self, Xn, tau):
        """M SGD steps (lr, weight decay on weight AND bias like torch.optim.SGD over head.parameters()) of every head
        on  mean_b -sum_k clamp(tau,1e-8,1)_bk log softmax(head(x)/T)_bk  (teacher_model.py:84-105, :32-40)."""
        t = tau.detach().clamp(min=1e-8, max=1.0)
        tsum = t.sum(-1, keepdim=True)
        B = Xn.shape[1]
        scale = 1.0 / (B * self.head_temp)
        XT = Xn.transpose(1, 2).contiguous()
        for _ in range(self.M):
            p = torch.softmax(torch.baddbmm(self.bh.unsqueeze(1), Xn, self.Wh.transpose(1, 2)) / self.head_temp, dim=-1)
            g = (p * tsum - t) * scale                                       # d loss / d (x W^T + b)   [V, B, K]
            dW = torch.bmm(XT, g).transpose(1, 2)                            # [V, K, Dmax]
            self.Wh.sub_(self.inner_lr * (dW + self.head_wd * self.Wh))
            self.bh.sub_(self.inner_lr * (g.sum(1) + self.head_wd * self.bh))

    def fit(self, loader, outer_steps: int = 200, rho: float = 0.04, verbose: bool = True):
        """teacher_model.py:240-351.  ``loader`` yields lists of per-view feature batches [B, D_v] (any device)."""
        import math
        ent = lambda p: -(p.clamp_min(1e-9) * p.clamp_min(1e-9).log()).sum(-1)
        it = iter(loader)
        K = self.K
        for step in range(outer_steps):
            try:
                feats = next(it)
            except StopIteration:
                it = iter(loader)
                feats = next(it)
            X = self._pack(feats)
            Xn = torch.nn.functional.normalize(X, dim=-1) if self.normalize_feats else X
            tau = self._tau(X)
            self._inner_fit(Xn, tau)
            with torch.no_grad():
                logp = torch.log_softmax(torch.baddbmm(self.bh.unsqueeze(1), Xn, self.Wh.transpose(1, 2)) / self.head_temp, -1)
            ce = -(tau.clamp(min=1e-8, max=1.0).unsqueeze(0) * logp).sum(-1).mean(1).sum() / max(self.V, 1)
            sample_entropy = ent(tau).mean()
            marginal = tau.mean(0)
            marg_gap = torch.relu(math.log(K) - ent(marginal.unsqueeze(0)).mean())
            frac = 1.0 - float(step) / float(max(1, outer_steps))
            gamma_t = self.gamma * frac
            dead_floor = max(1e-4, 0.1 / K)
            usage = (tau.clamp_min(1e-8) ** 2.0).mean(0)
            dead_pen = torch.relu(dead_floor - usage).sum() / (dead_floor * K)
            delta_t = self.delta * max(0.5, 0.6 + 0.4 * frac)
            loss = ce + self.alpha * sample_entropy + gamma_t * marg_gap + delta_t * dead_pen
            if (step % 2) != 0 and rho > 0.0:                                # batch-local smoothness on odd steps
                loss = loss + rho * (tau[1:] - tau[:-1]).abs().sum(-1).mean()
            self.opt_theta.zero_grad(set_to_none=True)
            loss.backward()
            self.opt_theta.step()
            if verbose and (step % 20 == 0 or step == outer_steps - 1):
                print(f"[Teacher] step {step:03d} | loss {float(loss):.4f} | CE {float(ce):.4f} | E[H(τ)] {float(sample_entropy):.4f} | "
                      f"mean max_p {float(tau.max(1).values.mean()):.3f} | dead_pen {float(dead_pen):.3f}")

    @torch.no_grad()
    def predict(self, loader) -> torch.Tensor:
        """Soft assignments [N, K] in loader order (teacher_model.py:219-238); stays on the teacher's device."""
        return torch.cat([self._tau(self._pack(feats)) for feats in loader], dim=0)


def run_turtle_teacher_on_views(views_dict: dict, n_components: int, gamma: float = 6.0, alpha_sample_entropy: float = 1.0,
                                outer_steps: int = 200, inner_steps: int = 200, normalize_feats: bool = True,
                                verbose: bool = True, device=None, head_temp: float = 0.3, task_temp: float = 0.3,
                                batch_size: int = 2048):
    """Same signature and result as teacher_model.py:710-792: ``(teacher, tau_star [N, K])``.  The views stay resident
    on ``device``; only window indices go through the (shuffling, drop_last) DataLoader, which keeps the consumption of
    torch's global generator — hence the batch order — identical to the reference's loader over the view tensors."""
    from torch.utils.data import DataLoader, TensorDataset
    device = torch.device(device) if device is not None else torch.device("cpu")
    views = [v.to(device).float() for v in views_dict.values() if v is not None]
    assert len(views) > 0, "No active views found."
    N = views[0].shape[0]
    index_set = TensorDataset(torch.arange(N))
    loader = DataLoader(index_set, batch_size=batch_size, shuffle=True, num_workers=0, drop_last=True)
    teacher = TurtleTeacherB200([v.shape[1] for v in views], n_components, gamma=gamma,
                                alpha_sample_entropy=alpha_sample_entropy, inner_lr=0.1, inner_steps=inner_steps,
                                head_wd=1e-4, head_temp=head_temp, task_temp=task_temp, normalize_feats=normalize_feats,
                                lr_theta=1e-3, device=device)

    class _Gather:
        def __init__(self, ld):
            self.ld = ld

        def __iter__(self):
            for (idx,) in self.ld:
                idx = idx.to(device)
                yield [v.index_select(0, idx) for v in views]

    teacher.fit(_Gather(loader), outer_steps=outer_steps, rho=0.04, verbose=verbose)
    seq = DataLoader(index_set, batch_size=batch_size * 2, shuffle=False, num_workers=0)
    return teacher, teacher.predict(_Gather(seq))


# ---------------------------------------------------------------------------------------------------------------------
# The teacher's PCA views (teacher_model.py:464-708): fit_nodes_pca / fit_angles_pca / extract_pca_edges_view run
# scikit-learn's IncrementalPCA (pinned in this image: scikit-learn 1.9) over the flattened windows in batches and then
# transform every window.  Restated here on device tensors: the published incremental-SVD update (Ross et al. 2008, as
# sklearn.decomposition.IncrementalPCA.partial_fit implements it: running column mean, SVD of
# [diag(S) Vt ; X - batch mean ; sqrt(n_seen n_batch / n_total) (mean - batch mean)], deterministic sign, truncation to
# n_components) with the same batch partition, in float64 (the reference runs LAPACK in float32; the views agree to
# ~1e-4 of their scale).  The windows stay where they are (HBM); nothing is materialised on the host.
# ---------------------------------------------------------------------------------------------------------------------
class IncrementalPCAB200:
    """partial_fit / transform of sklearn.decomposition.IncrementalPCA (whiten=False) on torch tensors."""

    def __init__(self, n_components: int):
        self.n_components = int(n_components)
        self.components_ = None          # [k, D] float64
        self.singular_values_ = None
        self.mean_ = None
        self.n_samples_seen_ = 0

    @torch.no_grad()
    def partial_fit(self, X: torch.Tensor):
        X = X.to(torch.float64)
        n, D = X.shape
        k = self.n_components
        if self.components_ is None and not (1 <= k <= D):
            raise ValueError(f"n_components={k} invalid for n_features={D}")
        if k > n:
            raise ValueError(f"n_components={k} must be less or equal to the batch number of samples {n}")
        batch_mean = X.mean(0)
        if self.n_samples_seen_ == 0:
            col_mean, total = batch_mean, n
            M = X - batch_mean
        else:
            total = self.n_samples_seen_ + n
            col_mean = (self.mean_ * self.n_samples_seen_ + batch_mean * n) / total
            corr = (self.n_samples_seen_ / total * n) ** 0.5 * (self.mean_ - batch_mean)
            M = torch.cat([self.singular_values_.unsqueeze(1) * self.components_, X - batch_mean, corr.unsqueeze(0)], 0)
        U, S, Vt = torch.linalg.svd(M, full_matrices=False)
        # sklearn.utils.extmath.svd_flip(u_based_decision=False): the largest-|.| entry of every row of Vt is positive
        idx = Vt.abs().argmax(dim=1)
        sign = torch.sign(Vt[torch.arange(Vt.shape[0], device=Vt.device), idx])
        sign[sign == 0] = 1.0
        Vt = Vt * sign.unsqueeze(1)
        self.components_, self.singular_values_ = Vt[:k].contiguous(), S[:k].contiguous()
        self.mean_, self.n_samples_seen_ = col_mean, total
        return self

    @torch.no_grad()
    def transform(self, X: torch.Tensor) -> torch.Tensor:
        return ((X.to(torch.float64) - self.mean_) @ self.components_.t()).float()


@torch.no_grad()
def pca_view(flat: torch.Tensor, n_components: int, batch_size: int, max_samples: Optional[int] = None):
    """One teacher view: IncrementalPCA fitted over consecutive batches of ``flat`` [N, D] (truncated to ``max_samples``
    windows like teacher_model.py:509-516), then every window transformed.  Returns (pca, feats [N, n_components])."""
    pca = IncrementalPCAB200(n_components)
    N = flat.shape[0]
    bounds = list(range(0, N, batch_size)) + [N]
    if len(bounds) > 2 and bounds[-1] - bounds[-2] < n_components and max_samples is None:
        # a ragged tail with fewer rows than components makes scikit-learn (and the reference with it) raise; fold it
        # into the previous batch instead — the only place this deviates from the reference's batch partition
        del bounds[-2]
    seen = 0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        Xb = flat[lo:hi]
        if max_samples is not None and seen >= max_samples:
            break
        if max_samples is not None and seen + Xb.shape[0] > max_samples:
            Xb = Xb[:max(1, max_samples - seen)]
        pca.partial_fit(Xb)
        seen += Xb.shape[0]
    feats = torch.cat([pca.transform(flat[i:i + batch_size]) for i in range(0, N, batch_size)], 0)
    return pca, feats


@torch.no_grad()
def teacher_views_from_windows(x: torch.Tensor, a: torch.Tensor, angles: Optional[torch.Tensor] = None, *,
                               pca_nodes_dim: int = 32, pca_edges_dim: int = 16, pca_angles_dim: int = 32,
                               batch_size_nodes: int = 4096, batch_size_edges: int = 8192, batch_size_angles: int = 8192,
                               include_nodes_view: bool = True, include_edges_view: bool = True,
                               latent_view: Optional[torch.Tensor] = None) -> dict:
    """The ``views`` dict of ``maybe_build_turtle_teacher`` (teacher_model.py:811-905) from resident windows
    x [N,T,Nn,F>=3], a [N,T,E,Fe] (and angles [N,T,A]): keys z / pca_pos / pca_spd / pca_edges / pca_angles, None = off."""
    if x.shape[-1] < 3:
        raise ValueError(f"Expected at least 3 channels (x,y,speed); got F={x.shape[-1]}")
    N = x.shape[0]
    views = {"z": latent_view, "pca_pos": None, "pca_spd": None, "pca_edges": None, "pca_angles": None}
    if include_nodes_view:                                                   # fit_nodes_pca :464-573
        views["pca_pos"] = pca_view(x[..., :2].reshape(N, -1), pca_nodes_dim, batch_size_nodes)[1]
        views["pca_spd"] = pca_view(x[..., 2:3].reshape(N, -1), pca_nodes_dim, batch_size_nodes)[1]
    if include_edges_view:                                                   # extract_pca_edges_view :638-708
        views["pca_edges"] = pca_view(a.reshape(N, -1), pca_edges_dim, batch_size_edges)[1]
    if angles is not None:                                                   # fit_angles_pca :576-635
        views["pca_angles"] = pca_view(angles.reshape(N, -1), pca_angles_dim, batch_size_angles)[1]
    return views


def build_turtle_teacher(x: torch.Tensor, a: torch.Tensor, n_components: int, *, angles: Optional[torch.Tensor] = None,
                         latent_view: Optional[torch.Tensor] = None, device=None, include_latent_view: bool = True,
                         include_nodes_view: bool = True, include_edges_view: bool = False,
                         include_angles_view: bool = False, pca_nodes_dim: int = 32, pca_edges_dim: int = 32,
                         pca_angles_dim: int = 32, batch_size_nodes: int = 4096, batch_size_edges: int = 8192,
                         batch_size_angles: int = 8192, teacher_gamma: float = 8.0, teacher_alpha_sample_entropy: float = 2.0,
                         teacher_outer_steps: int = 500, teacher_inner_steps: int = 100,
                         teacher_normalize_feats: bool = True, teacher_head_temp: float = 0.35,
                         teacher_task_temp: float = 0.35, teacher_batch_size: int = 2048, verbose: bool = True):
    """``maybe_build_turtle_teacher`` (teacher_model.py:811-905) on resident windows, keyword defaults = the reference's
    ``TurtleTeacherCfg`` (model_utils_new.py:78-125).  Returns ``(teacher, tau_star [N, K], views)``; the view order
    (z, pca_pos, pca_spd, pca_edges, pca_angles) is the reference's, it fixes which head draws which initial weights."""
    if include_latent_view and latent_view is None:
        raise ValueError("include_latent_view=True but latent_view=None")
    if include_angles_view and angles is None:
        raise ValueError("include_angles_view=True but angles=None")
    device = torch.device(device) if device is not None else x.device
    views = teacher_views_from_windows(
        x.to(device), a.to(device), angles.to(device) if include_angles_view else None, pca_nodes_dim=pca_nodes_dim,
        pca_edges_dim=pca_edges_dim, pca_angles_dim=pca_angles_dim, batch_size_nodes=batch_size_nodes,
        batch_size_edges=batch_size_edges, batch_size_angles=batch_size_angles, include_nodes_view=include_nodes_view,
        include_edges_view=include_edges_view, latent_view=latent_view.to(device) if include_latent_view else None)
    teacher, tau_star = run_turtle_teacher_on_views(
        views, n_components, gamma=teacher_gamma, alpha_sample_entropy=teacher_alpha_sample_entropy,
        outer_steps=teacher_outer_steps, inner_steps=teacher_inner_steps, normalize_feats=teacher_normalize_feats,
        verbose=verbose, device=device, head_temp=teacher_head_temp, task_temp=teacher_task_temp,
        batch_size=teacher_batch_size)
    return teacher, tau_star.detach(), views


@torch.no_grad()
def teacher_context(tau_star: torch.Tensor, class_reweight: bool = True, class_reweight_beta: float = 1.0,
                    class_reweight_cap: Optional[float] = 3.0) -> dict:
    """What ``VadeLoss.set_teacher`` derives from tau* (losses.py:460-491): the inverse-marginal class weights
    (``pi^-beta`` normalised to mean 1, capped) and the clamped teacher marginal that lifts the non-empty floor."""
    pi = tau_star.mean(dim=0).clamp_min(1e-8)
    out = {"tau_star": tau_star, "teacher_marginal": pi, "class_weight": None}
    if class_reweight:
        w = pi.pow(-class_reweight_beta)
        w = w / w.mean()
        if class_reweight_cap is not None:
            w = w.clamp_max(class_reweight_cap)
        out["class_weight"] = w
    return out
