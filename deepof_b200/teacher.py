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
