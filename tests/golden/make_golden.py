"""Generate golden vectors by running the UNMODIFIED reference (mlfpm/deepof) on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/vade_<case>.npz.  Each file holds, for one seeded VaDE/recurrent
model of the reference (models_new.VaDEPT, use_gnn=True):
  * the full state_dict (``p/<key>``), adjacency, synthetic windows x, a
  * eval-mode outputs: encoder output, embedding (z_mean), q, decoder loc
  * two consecutive reference training steps (training.step_vade + backward +
    clip_grad_value_(0.75) + Adam from losses.build_optimizer_vade) with the noise
    tensors the reference drew (reparam eps, MC-KL eps) recorded by patching
    torch.randn / torch.randn_like, the 13 logged loss terms per step, the clipped
    ... no: the RAW gradient of step 1 (``g/<key>``), and parameters after step 2
    (``p2/<key>``).
Noise recording does not alter the reference's arithmetic: the patched functions call
the originals and keep a copy of what they returned.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from oracle.vade_oracle import default_adjacency, synthetic_windows  # noqa: E402

M, L, T, U = refshim.load()

CASES = {
    # name: (T, N, D, K, B, phase, seed, distill)
    "cfg1_main": dict(T=25, N=14, D=8, K=4, B=16, phase="main", seed=11, distill=False),
    "cfg1_pretrain": dict(T=25, N=14, D=8, K=4, B=16, phase="pretrain", seed=12, distill=False),
    "cfg2_main": dict(T=25, N=14, D=16, K=8, B=12, phase="main", seed=13, distill=False),
    "cfg2_distill": dict(T=25, N=14, D=16, K=8, B=12, phase="main", seed=14, distill=True),
    "odd_main": dict(T=24, N=11, D=6, K=5, B=9, phase="main", seed=15, distill=False),
    "odd_pretrain": dict(T=24, N=11, D=6, K=5, B=9, phase="pretrain", seed=16, distill=False),
}


class NoiseTape:
    """Records every tensor torch.randn / torch.randn_like hands out."""

    def __init__(self):
        self.draws = []
        self._randn, self._randn_like = torch.randn, torch.randn_like

    def __enter__(self):
        def randn(*a, **k):
            t = self._randn(*a, **k)
            self.draws.append(t.detach().clone())
            return t

        def randn_like(*a, **k):
            t = self._randn_like(*a, **k)
            self.draws.append(t.detach().clone())
            return t

        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def run_case(name, c):
    torch.manual_seed(c["seed"])
    torch.set_num_threads(1)
    adj = default_adjacency(c["N"])
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = synthetic_windows(c["B"], c["T"], adj, seed=1000 + c["seed"])
    model = M.VaDEPT((c["T"], c["N"], 3), (c["T"], E, 1), adj, c["D"], c["K"],
                     encoder_type="recurrent", use_gnn=True, kmeans_loss=1.0)
    # move the GMM off its tiny xavier init so q is not uniform
    with torch.no_grad():
        model.latent_space.gmm_means.mul_(3.0)
    out = {"adjacency": adj, "x": x.numpy(), "a": a.numpy(),
           "meta": np.array([c["T"], c["N"], E, c["D"], c["K"], c["B"]], dtype=np.int64),
           "phase": np.array(c["phase"])}
    for k, v in model.state_dict().items():
        out["p/" + k] = v.detach().numpy().copy()

    # ---- eval outputs (the judged ones)
    model.eval()
    with torch.no_grad():
        enc = model.encoder(x, a)
        dist, emb, q, _km = model(x, a)
        out["eval/enc"] = enc.numpy()
        out["eval/emb"] = emb.numpy()
        out["eval/q"] = q.numpy()
        out["eval/loc"] = dist.base_dist.base_dist.loc.numpy()

    # ---- two training steps
    common = U.CommonFitCfg(latent_dim=c["D"], n_components=c["K"])
    vcfg = U.VaDECfg()
    tcfg = U.TurtleTeacherCfg()
    crit = L.VadeLoss(common, vcfg, tcfg)
    nb = 10
    if c["phase"] == "pretrain":
        sched = L.Dynamic_weight_manager(nb, mode=vcfg.kl_annealing_mode_pretrain,
                                         warmup_epochs=vcfg.kl_warmup_pretrain,
                                         max_weight=vcfg.kl_max_weight_pretrain,
                                         cooldown_epochs=vcfg.kl_cooldown_pretrain,
                                         end_weight=vcfg.kl_end_weight_pretrain)
        crit.set_kl_scheduler(sched)
        sched.current_iteration = 90
        lr_base, lr_gmm = vcfg.learning_rate_pretrain, 0.0
    else:
        model.set_pretrain_mode(False)
        crit.set_mode("main")
        sched = L.Dynamic_weight_manager(nb, mode=vcfg.kl_annealing_mode, warmup_epochs=vcfg.kl_warmup,
                                         max_weight=vcfg.kl_max_weight, cooldown_epochs=vcfg.kl_cooldown,
                                         end_weight=vcfg.kl_end_weight)
        crit.set_kl_scheduler(sched)
        sched.current_iteration = 30
        lr_base, lr_gmm = 5e-4, 2e-4   # training.py:1750-1755
    sched_args = dict(n_batches_per_epoch=nb, it0=sched.current_iteration)
    out["sched"] = np.array([sched_args["n_batches_per_epoch"], sched_args["it0"]], dtype=np.int64)
    out["lr"] = np.array([lr_base, lr_gmm], dtype=np.float64)
    idx = torch.arange(c["B"], dtype=torch.long)
    if c["distill"]:
        g = torch.Generator().manual_seed(77)
        tau = torch.softmax(2.0 * torch.randn(c["B"], c["K"], generator=g), dim=-1)
        crit.set_teacher(tau_star=tau, lambda_distill=tcfg.lambda_distill, lambda_scheduler=None)
        out["tau_star"] = tau.numpy()
        out["class_weight"] = crit.class_weight.numpy()
        out["teacher_marginal"] = crit.teacher_marginal.numpy()
        out["lambda_distill"] = np.array(tcfg.lambda_distill)
    opt = L.build_optimizer_vade(model, base_lr=lr_base, gmm_lr=lr_gmm)
    model.train()
    crit.train()
    ctx = types.SimpleNamespace(criterion=crit, apply_distill=c["distill"], train=True)
    for step in range(2):
        with NoiseTape() as tape:
            res = T.step_vade(model, (x, a, idx), ctx)
        opt.zero_grad(set_to_none=True)
        res.loss.backward()
        draws = tape.draws
        out[f"s{step}/eps"] = draws[0].numpy()               # reparam eps [B,D]
        if c["phase"] == "main":
            assert len(draws) == 2 and draws[1].shape[0] == 32
            out[f"s{step}/mc_eps"] = draws[1].numpy()        # [32,B,D]
        else:
            assert len(draws) == 1
        out[f"s{step}/klw"] = np.array(float(sched.get_weight()))
        for k, v in res.logs.items():
            out[f"s{step}/log/{k}"] = np.array(v, dtype=np.float64)
        if step == 0:
            for k, prm in model.named_parameters():
                if prm.grad is not None:
                    out["g/" + k] = prm.grad.detach().numpy().copy()
        torch.nn.utils.clip_grad_value_(model.parameters(), 0.75)
        opt.step()
        sched.step()
    for k, v in model.state_dict().items():
        out["p2/" + k] = v.detach().numpy().copy()
    path = os.path.join(HERE, f"vade_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024),
          {k: round(v, 5) for k, v in res.logs.items() if v != 0.0})


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        run_case(name, c)
