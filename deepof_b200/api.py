"""Reference-facing entry points of the B200 path: the step functions, the epoch loop, ``train_deepof_model`` and the
checkpoint bundle, with the names, argument meaning and return shapes of ``deepof/clustering/training.py`` and
``model_utils_new.py``, for the configurations this library implements
(``encoder_type="recurrent" | "transformer" | "TCN"``, GNN path).

* ``step_vade / step_vqvae_distill / step_contrastive_distill(model, (x, a, idx), ctx) -> StepResult(loss, logs)``
  (reference ``training.py:231-309, 312-389, 482-589``).  The CUDA library fuses forward, loss and backward, so the
  gradient is already in ``model.grad`` when the step function returns; ``StepResult.loss`` is a device scalar whose
  ``backward()`` is a no-op kept for loop compatibility, ``logs`` holds the reference's keys.
* ``train_one_epoch_indexed`` (``training.py:104-187``): step, [all-reduce], clip_grad_value_(0.75) + Adam, schedulers.
* ``train_deepof_model`` (``training.py:592-905``) -> ``(model_val, model_score, teacher_init_model, log_summary)``.
* ``save_model_info / load_model_from_ckpt`` (``model_utils_new.py:263-329, 367-417``): ``torch.save({"state_dict",
  "rebuild_spec", "log_summary"})`` with the reference's keys, so checkpoints move between the two implementations.

Options the library does not implement raise (AMP, bootstrap training, tf-cluster / scatter / prior loss weights, the
TCN encoder); the keyword list is the reference's, so a misspelled keyword is a TypeError like there.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np
import torch

from .loader import batch_starts
from .models import (AugParams, CON_LOG_KEYS, ContrastiveAugCfg, ContrastiveB200, DistillHeadB200, Distillation, VQ_LOG_KEYS,
                     VQVAEB200)
from .tfm import TFMModelB200
from .training import KLSchedule
from .vade import VaDEB200, VadeLossCfg


class _Loss:
    """Device scalar of a finished step.  The backward pass already ran inside the CUDA library."""

    def __init__(self, t: torch.Tensor):
        self._t = t

    def backward(self, *a, **k):
        return None

    def detach(self):
        return self._t.detach()

    def item(self) -> float:
        return float(self._t.item())

    def __float__(self) -> float:
        return self.item()


class StepResult(SimpleNamespace):
    """``StepResult(loss, logs)`` (reference ``training.py:72-75``)."""


def _logs(model) -> Dict[str, float]:
    return model.logs_dict()


def step_vade(model: VaDEB200, batch, ctx: SimpleNamespace) -> StepResult:
    """``ctx.criterion`` is a :class:`VadeLossCfg` (plays the role of the reference ``VadeLoss`` module); optional
    ``ctx.kl_scheduler`` (:class:`KLSchedule`), ``ctx.tau_star`` [Nw,K] with ``ctx.apply_distill``, ``ctx.eps`` /
    ``ctx.mc_eps`` to inject the noise (tests)."""
    x, a, idx = batch
    cfg: VadeLossCfg = ctx.criterion
    sched = getattr(ctx, "kl_scheduler", None)
    if sched is not None:
        cfg.kl_weight = sched.get_weight()
    lsched = getattr(ctx, "lambda_scheduler", None)
    if lsched is not None:
        # training.py:264-265 — only when the CALLER puts a lambda_scheduler into ctx; fit_VADE does not (its ctx is
        # (criterion, scheduler_per_batch), :1805), so the distillation weight of a VaDE fit stays at lambda_distill
        cfg.lambda_distill = float(lsched.get_weight())
    tau = None
    tau_star = getattr(ctx, "tau_star", None)
    if tau_star is not None and getattr(ctx, "apply_distill", True) and cfg.lambda_distill > 0.0:
        tau = tau_star[idx.to(model.device).long()]
    if not getattr(ctx, "train", True):          # validate_one_epoch_indexed: model.eval(), criterion.eval() (training.py:190-229)
        logs = model.loss_eval(x, a, cfg, mc_eps=getattr(ctx, "mc_eps", None), tau_batch=tau,
                               class_weight=getattr(ctx, "class_weight", None), teacher_marginal=getattr(ctx, "teacher_marginal", None))
    else:
        logs = model.loss_grad(x, a, cfg, eps=getattr(ctx, "eps", None), mc_eps=getattr(ctx, "mc_eps", None), tau_batch=tau,
                               class_weight=getattr(ctx, "class_weight", None), teacher_marginal=getattr(ctx, "teacher_marginal", None),
                               dropout_masks=getattr(ctx, "dropout_masks", None))
    return StepResult(loss=_Loss(logs[0]), logs=_logs(model) if getattr(ctx, "read_logs", True) else {})


def _distillation(model, idx, ctx: SimpleNamespace) -> Optional[Distillation]:
    """The teacher part of ``ctx`` exactly as ``step_vqvae_distill`` / ``step_contrastive_distill`` read it
    (training.py:341-370, 550-578): ``distill_head`` (:class:`DistillHeadB200`), ``tau_star`` [Nw,K],
    ``lambda_scheduler.get_weight()``, ``distill_sharpen_T`` (0.5), ``distill_conf_weight`` / ``distill_conf_thresh``."""
    sched = getattr(ctx, "lambda_scheduler", None)
    lam = float(sched.get_weight()) if sched is not None else 0.0
    head = getattr(ctx, "distill_head", None)
    if not getattr(ctx, "apply_distill", True) or head is None or lam <= 0.0:
        return None
    if not isinstance(head, DistillHeadB200):
        raise TypeError("ctx.distill_head must be a deepof_b200.DistillHeadB200")
    tau = ctx.tau_star.to(model.device)[idx.to(model.device).long()]
    return Distillation(head, tau, lam, float(getattr(ctx, "distill_sharpen_T", 0.5)),
                        bool(getattr(ctx, "distill_conf_weight", False)), float(getattr(ctx, "distill_conf_thresh", 0.6)))


def step_vqvae_distill(model: VQVAEB200, batch, ctx: SimpleNamespace) -> StepResult:
    x, a, idx = batch
    if not getattr(ctx, "train", True):
        logs = model.loss_eval(x, a)
    else:
        logs = model.loss_grad(x, a, distill=_distillation(model, idx, ctx), dropout_masks=getattr(ctx, "dropout_masks", None))
    return StepResult(loss=_Loss(logs[0]), logs=_logs(model) if getattr(ctx, "read_logs", True) else {})


def step_contrastive_distill(model: ContrastiveB200, batch, ctx: SimpleNamespace) -> StepResult:
    """``batch[0]`` is the FULL window ``x_full`` [B,T,N,3]; the edge tensor is recomputed from it like the reference
    (``training.py:497``).  ``ctx.contrastive_cfg`` is a :class:`ContrastiveAugCfg`; ``ctx.aug_params`` injects the
    augmentation decisions (tests), otherwise they are drawn from ``ctx.generator`` / ``ctx.host_generator``."""
    x_full = batch[0]
    prm: Optional[AugParams] = getattr(ctx, "aug_params", None)
    if prm is None:
        cfg = getattr(ctx, "contrastive_cfg", None) or ContrastiveAugCfg()
        prm = model.draw_augmentation(x_full.shape[0], cfg, getattr(ctx, "generator", None), getattr(ctx, "host_generator", None))
    if not getattr(ctx, "train", True):
        logs = model.loss_eval(x_full, prm)
    else:
        logs = model.loss_grad(x_full, prm, distill=_distillation(model, batch[2] if len(batch) > 2 else None, ctx))
    return StepResult(loss=_Loss(logs[0]), logs=_logs(model) if getattr(ctx, "read_logs", True) else {})


def average_logs(logs_list: Iterable[Dict[str, float]]) -> Dict[str, float]:
    logs_list = list(logs_list)
    if not logs_list:
        return {}
    return {k: float(np.mean([l[k] for l in logs_list if k in l])) for k in logs_list[0]}


_LOG_KEYS_OF = {"vade": None, "vqvae": VQ_LOG_KEYS, "contrastive": CON_LOG_KEYS}


def _log_keys(model_name: str):
    from ._lib import LOG_KEYS
    return _LOG_KEYS_OF.get(model_name) or LOG_KEYS


def train_one_epoch_indexed(model, model_name: str, dataloader, optimizer, step_fn, epoch: int = 0, num_epochs: int = 1,
                            grad_clip_value: Optional[float] = 0.75, ctx: Optional[SimpleNamespace] = None,
                            world_size: int = 1, log_every: int = 0):
    """One epoch (reference ``training.py:104-187``).  ``dataloader`` yields ``(x, a, idx)``; ``optimizer`` is a dict of
    the Adam hyper-parameters: ``{"lr": .., "gmm_lr": .., "weight_decay": .., "active": (enc, dec, gmm)}``.  Returns
    ``(averaged logs, kl weight at mid epoch, distillation weight at mid epoch)`` like the reference.  The reference
    averages the logs of EVERY step (one ``.item()`` sync per term and step); here the log vector of every step is added
    into a device accumulator and read back once per epoch — the same mean without per-step synchronisation."""
    ctx = ctx or SimpleNamespace()
    mid_kl, mid_lambda = 0.0, 0.0
    batches = list(dataloader) if not hasattr(dataloader, "__len__") else dataloader
    n = len(batches)
    acc = torch.zeros_like(model.logs, dtype=torch.float64)
    steps = 0
    for step, batch in enumerate(batches):
        sub = SimpleNamespace(**{**ctx.__dict__, "train": True, "epoch": epoch, "num_epochs": num_epochs, "read_logs": False})
        res = step_fn(model, batch, sub)
        res.loss.backward()
        acc.add_(model.logs)
        steps += 1
        scale = 1.0
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(model.grad, op=dist.ReduceOp.SUM)
            scale = 1.0 / world_size
        if model_name == "vade":
            model.adam_step(optimizer["lr"], optimizer.get("gmm_lr", 0.0), clip=grad_clip_value or 0.0, grad_scale=scale,
                            active=optimizer.get("active", (True, True, True)))
            sched = getattr(ctx, "kl_scheduler", None)
            if sched is not None:
                sched.step()
                if step == n // 2:
                    mid_kl = sched.get_weight()
            # criterion.lambda_scheduler of the reference: stepped and reported, never applied (training.py:173-176)
            lsched = getattr(ctx, "criterion_lambda_scheduler", None) or getattr(ctx, "lambda_scheduler", None)
            if lsched is not None:
                lsched.step()
                if step == n // 2:
                    mid_lambda = lsched.get_weight()
        else:
            model.adam_step(optimizer["lr"], clip=grad_clip_value or 0.0, grad_scale=scale,
                            weight_decay=optimizer.get("weight_decay", 1e-4))
            head = getattr(ctx, "distill_head", None)
            sched = getattr(ctx, "lambda_scheduler", None)
            if head is not None and getattr(ctx, "apply_distill", True) and sched is not None and sched.get_weight() > 0.0:
                # build_optimizer_generic puts the head into the same Adam (losses.py:811-814); DDP averages its gradient
                if world_size > 1:
                    dist.all_reduce(head.grad, op=dist.ReduceOp.SUM)
                head.adam_step(optimizer["lr"], weight_decay=optimizer.get("weight_decay", 1e-4), grad_scale=scale)
            if sched is not None:
                sched.step()                                              # training.py:177-181
                if step == n // 2:
                    mid_lambda = sched.get_weight()
    if steps == 0:
        return {}, mid_kl, mid_lambda
    v = (acc / steps).cpu().tolist()
    return {k: v[i] for i, k in enumerate(_log_keys(model_name))}, mid_kl, mid_lambda


def validate_one_epoch_indexed(model, model_name: str, dataloader, step_fn, epoch: int = 0, num_epochs: int = 1,
                               ctx: Optional[SimpleNamespace] = None, world_size: int = 1, rank: int = 0) -> Dict[str, float]:
    """``validate_one_epoch_indexed`` (reference ``training.py:190-229``): every batch through ``step_fn`` with
    ``ctx.train = False`` (the model in eval mode, no gradient), the logs averaged over batches.  The reference runs the
    whole validation set on every rank; here the batches are dealt round-robin to the ranks and the log sums are
    all-reduced — the result is the same mean."""
    ctx = ctx or SimpleNamespace()
    acc = torch.zeros(model.logs.numel() + 1, dtype=torch.float64, device=model.logs.device)
    for i, batch in enumerate(dataloader):
        if world_size > 1 and i % world_size != rank:
            continue
        sub = SimpleNamespace(**{**ctx.__dict__, "train": False, "epoch": epoch, "num_epochs": num_epochs, "read_logs": False})
        step_fn(model, batch, sub)
        acc[:-1].add_(model.logs)
        acc[-1] += 1.0
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    v = acc.cpu().tolist()
    if v[-1] <= 0:
        return {}
    return {k: v[i] / v[-1] for i, k in enumerate(_log_keys(model_name))}


# ---- diagnostics, log summary (deepof/clustering/logging.py) -----------------------------------------------------------
def _clip01(x: float) -> float:
    return max(0.0, min(1.0, x))


@torch.no_grad()
def compute_diagnostics(model, dataloader, q_fn, n_components: int, tau_star: Optional[torch.Tensor] = None,
                        distill_sharpen_T: float = 0.5, distill_conf_weight: bool = False, distill_conf_thresh: float = 0.55,
                        max_batches: int = 4, extra_stats_fn=None) -> Dict[str, float]:
    """``compute_diagnostics`` (reference ``logging.py:149-301``): confidence / balance / alignment score from the soft
    assignments ``q_fn(model, x, a) -> [B, K]`` of the first ``max_batches`` batches (device tensors, one read-back)."""
    import math
    if n_components < 2:
        raise ValueError(f"n_components must be >= 2, got {n_components}")
    total, sum_ent, sum_max, sum_q = 0, None, None, None
    for bi, batch in enumerate(dataloader):
        if bi >= max_batches:
            break
        q = q_fn(model, batch[0], batch[1]).float().clamp_min(1e-8)
        q = q / q.sum(dim=-1, keepdim=True).clamp_min(1e-8)
        ent = -(q * q.log()).sum(dim=-1)
        e, m, qs = ent.sum().double(), q.max(dim=-1).values.sum().double(), q.sum(dim=0).double()
        sum_ent, sum_max, sum_q = (e, m, qs) if sum_q is None else (sum_ent + e, sum_max + m, sum_q + qs)
        total += q.shape[0]
    nan = float("nan")
    out = {"diag/q_mean_entropy": nan, "diag/q_marginal_entropy": nan, "diag/q_mean_max_prob": nan,
           "diag/teacher_marginal_entropy": nan, "diag/teacher_conf_mean": nan, "diag/teacher_weight_mean": nan,
           "diag/kl_marg_q_to_tau": nan, "conf_norm": nan, "bal_norm": nan, "alignment_score": nan}
    if total > 0:
        q_marg = (sum_q / total).float().clamp_min(1e-9)
        mean_ent, mean_max = float(sum_ent) / total, float(sum_max) / total
        q_marg_ent = float(-(q_marg * q_marg.log()).sum())
        logK = math.log(float(n_components))
        conf_norm = _clip01(1.0 - mean_ent / max(1e-9, logK))
        out.update({"diag/q_mean_entropy": mean_ent, "diag/q_marginal_entropy": q_marg_ent, "diag/q_mean_max_prob": mean_max})
        if tau_star is not None:
            tau = tau_star.detach().to(q_marg.device, q_marg.dtype)
            tau_marg = tau.mean(dim=0).clamp_min(1e-9)
            kl = max(0.0, float((q_marg * (q_marg.log() - tau_marg.log())).sum()))
            bal_norm = _clip01(1.0 - kl / max(1e-9, logK))
            T = float(distill_sharpen_T)
            tau_sharp = torch.softmax(tau.clamp_min(1e-8).log() / T, dim=-1) if T > 0.0 else tau
            conf = tau_sharp.max(dim=1).values
            if distill_conf_weight:
                thr = float(distill_conf_thresh)
                wmean = float(((conf - thr) / max(1e-6, 1.0 - thr)).clamp(0.0, 1.0).mean())
            else:
                wmean = 1.0
            out.update({"diag/teacher_marginal_entropy": float(-(tau_marg * tau_marg.log()).sum()),
                        "diag/teacher_conf_mean": float(conf.mean()), "diag/teacher_weight_mean": wmean, "diag/kl_marg_q_to_tau": kl})
        else:
            bal_norm = _clip01(q_marg_ent / max(1e-9, logK))
        out.update({"conf_norm": conf_norm, "bal_norm": bal_norm, "alignment_score": conf_norm * bal_norm})
    if extra_stats_fn is not None:
        out.update(extra_stats_fn(model))
    return out


def compute_vade_specific_diagnostics(model) -> Dict[str, float]:
    """``compute_vade_specific_diagnostics`` (``logging.py:118-146``)."""
    lv = model.latent_space.gmm_log_vars.detach()
    prior = model.latent_space.prior.detach().clamp_min(1e-9)
    return {"diag/gmm_logvar_min": float(lv.min()), "diag/gmm_logvar_max": float(lv.max()),
            "diag/prior_entropy": float(-(prior * prior.log()).sum())}


def get_q_vade(model, x, a):
    return model.embed(x, a)[1]                                                # logging.py:36-57 (renormalised by the caller)


def get_q_vqvae(model, x, a, *, distill_head):
    return torch.softmax(distill_head(model.encode(x, a)), dim=-1)             # logging.py:60-80


def get_q_contrastive(model, x_full, a_full, *, distill_head):
    x, a = model.main_view(x_full)                                             # logging.py:83-115
    return torch.softmax(distill_head(torch.nn.functional.normalize(model(x, a), dim=1)), dim=-1)


_SUMMARY_KEYS = ("total_loss", "reconstruction_loss", "kl_divergence", "cat_cluster_loss", "kmeans_loss", "distill_loss",
                 "temporal_loss", "scatter_loss", "nonempty_loss", "repel_loss", "tf_cluster_loss", "prior_loss", "activity_l1",
                 "pos_similarity", "neg_similarity", "conf_norm", "bal_norm", "alignment_score")


def init_log_summary(model_name: str) -> Dict[str, Any]:
    """``init_log_summary`` (``logging.py:304-332``), key for key — including the keys that do not match any logged name
    (``reconstruction_loss``, ``kl_divergence``, ``cat_cluster_loss``, ``tf_cluster_loss``), which therefore collect NaN in
    the reference as well."""
    return {"model_type": model_name, "train": {k: [] for k in _SUMMARY_KEYS}, "val": {k: [] for k in _SUMMARY_KEYS}}


def update_log_summary(log_summary: Dict[str, Any], train_logs: Dict[str, float], val_logs: Dict[str, float]) -> Dict[str, Any]:
    """``_update_log_summary`` (``logging.py:335-351``)."""
    for data_type, logs in (("train", train_logs), ("val", val_logs)):
        if data_type == "train":
            for key in list(log_summary.keys()):
                if key in ("train", "val", "test"):
                    continue
                log_summary[key] = logs.get(key, np.nan)
        for key in log_summary[data_type]:
            log_summary[data_type][key].append(logs.get(key, np.nan))
    return log_summary


# ---- data -----------------------------------------------------------------------------------------------------
def windows_from_table_dict(td: Dict[str, Any], device) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(nodes [Nw,T,3N], edges [Nw,T,E], ...)`` tuples of a preprocessed TableDict (reference ``data.py:2877-2888``)
    -> ``x [Nw,T,N,3]`` (``reorder_and_reshape``, ``dataset.py:16-26``), ``a [Nw,T,E,1]``, videos concatenated in key
    order like ``BatchDictDataset._build_hdf5`` (``dataset.py:183-290``)."""
    xs, as_ = [], []
    for key in td:
        v = td[key]
        nodes = torch.as_tensor(np.asarray(v[0]), dtype=torch.float32)
        edges = torch.as_tensor(np.asarray(v[1]), dtype=torch.float32)
        assert nodes.shape[2] % 3 == 0, "Error! Number of columns is not a multiple of 3 (x, y, speed)!"
        n = nodes.shape[2] // 3
        xs.append(torch.stack([nodes[:, :, :n], nodes[:, :, n:2 * n], nodes[:, :, 2 * n:]], dim=-1))
        as_.append(edges.unsqueeze(-1))
    return torch.cat(xs).to(device).contiguous(), torch.cat(as_).to(device).contiguous()


class WindowSource:
    """The window store ``train_deepof_model`` iterates — ``BatchDictDataset`` + ``_H5BatchIterableDataset`` of the
    reference (``dataset.py:183-290, 561-671``) — over one of two backings:

    * a :class:`~deepof_b200.loader.WindowLoader`: only the raw pose frames are resident, every batch is produced by the
      loader kernel (2.2 GB instead of 56 GB at 10 M windows), or
    * materialised windows ``x [Nw,T,N,3]``, ``a [Nw,T,E,1]`` (a preprocessed TableDict).

    Batches are contiguous runs of windows; batch starts are shuffled with ``default_rng(seed + epoch)`` and dealt
    ``starts[rank::world]`` exactly as in the reference."""

    def __init__(self, backing, batch_size: int, seed: int = 0, rank: int = 0, world: int = 1, shuffle: bool = True, device=None):
        from .loader import WindowLoader
        self.loader = backing if isinstance(backing, WindowLoader) else None
        if self.loader is None:
            self.x, self.a = backing
            self.n = int(self.x.shape[0])
            self.x_shape, self.a_shape = tuple(self.x.shape[1:]), tuple(self.a.shape[1:])
            self.device = self.x.device
        else:
            L = self.loader
            self.n = len(L)
            self.x_shape, self.a_shape = (L.T, L.N, 3), (L.T, L.E, 1)
            self.device = L.device
        self.bs, self.seed, self.rank, self.world, self.shuffle = int(batch_size), seed, rank, world, shuffle
        self.epoch = 0

    def get(self, s: int, n: int):
        if self.loader is not None:
            return self.loader.load(int(s), int(n))
        return self.x[s:s + n], self.a[s:s + n]

    def __iter__(self):
        """One training epoch for this rank."""
        self.epoch += 1
        for s in batch_starts(self.n, self.bs, self.epoch, self.seed, self.shuffle, self.rank, self.world):
            n = int(min(self.bs, self.n - int(s)))
            x, a = self.get(int(s), n)
            yield x, a, torch.arange(int(s), int(s) + n, device=self.device)

    def __len__(self):
        return len(batch_starts(self.n, self.bs, 1, self.seed, False, self.rank, self.world))

    def sequential(self, batch_size: Optional[int] = None, min_rows: int = 1):
        """All windows in store order (validation, embeddings): ``(x, a, idx)``."""
        bs = int(batch_size or self.bs)
        for s in range(0, self.n, bs):
            n = min(bs, self.n - s)
            if n < min_rows:
                continue
            x, a = self.get(s, n)
            yield x, a, torch.arange(s, s + n, device=self.device)

    def all_windows(self):
        """Every window materialised on the device (the TURTLE teacher's PCA views are fitted on window tensors)."""
        if self.loader is None:
            return self.x, self.a
        return self.loader.load(0, self.n)


def _as_source(obj, device, batch_size, seed, rank, world, shuffle=True) -> "WindowSource":
    from .loader import WindowLoader
    if isinstance(obj, WindowSource):
        return obj
    if isinstance(obj, WindowLoader):
        return WindowSource(obj, batch_size, seed, rank, world, shuffle)
    return WindowSource(windows_from_table_dict(obj, device), batch_size, seed, rank, world, shuffle)


# ---- checkpoints ----------------------------------------------------------------------------------------------
def ckpt_paths(model_name: str, output_path: str, run: int = 0):
    """``ckpt_paths`` (``model_utils_new.py:368-374``)."""
    ckpt_dir = os.path.join(output_path, "models", model_name.lower(), f"run_{run}")
    os.makedirs(ckpt_dir, exist_ok=True)
    return (ckpt_dir, os.path.join(ckpt_dir, "best_model_val.pth"), os.path.join(ckpt_dir, "best_model_score.pth"),
            os.path.join(ckpt_dir, "model_teacher_init.pth"))


def save_model_info(ckpt_path: str, *, model, rebuild_spec: Dict[str, Any], log_summary: Optional[Dict[str, Any]] = None,
                    stage: str = "final", save_weights: bool = True, **info) -> None:
    """The reference's checkpoint bundle (``model_utils_new.py:263-329``)."""
    os.makedirs(os.path.dirname(os.path.abspath(ckpt_path)), exist_ok=True)
    if save_weights:
        payload = {"state_dict": {k: v.cpu() for k, v in model.state_dict().items()}, "rebuild_spec": rebuild_spec}
        if log_summary is not None:
            payload["log_summary"] = log_summary
        torch.save(payload, ckpt_path)
    with open(os.path.splitext(ckpt_path)[0] + "_info.txt", "w", encoding="utf-8") as f:
        f.write(f"stage: {stage}\n" + "".join(f"{k}: {v}\n" for k, v in info.items() if v is not None)
                + "\n[checkpoint_format]\nckpt_contains: bundle\nbundle_keys: state_dict, rebuild_spec"
                + (", log_summary" if log_summary is not None else "") + "\n")


def build_model(rebuild_spec: Dict[str, Any], max_batch: int = 4096, training: bool = False, device: Optional[int] = None):
    """Model object from a reference ``rebuild_spec`` (``model_utils_new.py:367-417``)."""
    name = str(rebuild_spec["model_name"]).lower()
    etype = rebuild_spec.get("encoder_type", "recurrent")
    if etype not in ("recurrent", "transformer", "TCN") or not rebuild_spec.get("use_gnn", True):
        raise NotImplementedError("deepof_b200 implements encoder_type='recurrent', 'transformer' and 'TCN', use_gnn=True")
    xs, as_ = tuple(rebuild_spec["x_shape"]), tuple(rebuild_spec["a_shape"])
    adj, D, K = np.asarray(rebuild_spec["adjacency_matrix"]), int(rebuild_spec["latent_dim"]), int(rebuild_spec.get("n_components", 1))
    kw = dict(max_batch=max_batch, training=training, device=device, encoder_type=etype)
    if name == "vade":
        return VaDEB200(xs, as_, adj, D, K, kmeans_loss=float(rebuild_spec.get("kmeans_loss", 0.0)), **kw)
    if name == "vqvae":
        return VQVAEB200(xs, as_, adj, D, K, kmeans_loss=float(rebuild_spec.get("kmeans_loss", 0.0)), **kw)
    if name == "contrastive":
        return ContrastiveB200(xs, as_, adj, D, **kw)
    raise ValueError(f"unknown model_name {rebuild_spec['model_name']!r}")


def load_model_from_ckpt(ckpt_path: str, max_batch: int = 4096, training: bool = False, device: Optional[int] = None):
    bundle = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    model = build_model(bundle["rebuild_spec"], max_batch=max_batch, training=training, device=device)
    model.load_state_dict(bundle["state_dict"], strict=False)
    return model, bundle.get("log_summary")


def load_best_checkpoints(model, rebuild_spec, best_path_val: str, best_path_score: str, save_weights: bool, max_batch: int):
    """``load_best_checkpoints`` (``model_utils_new.py:907-940``): ``model_score`` starts as a copy of the last-epoch
    model, the best-val / best-score checkpoints are loaded into ``model`` / ``model_score`` when they exist."""
    model_score = build_model(rebuild_spec, max_batch=max_batch, training=False)
    model_score.load_state_dict(model.state_dict())
    if save_weights and os.path.exists(best_path_val):
        model.load_state_dict(torch.load(best_path_val, map_location="cpu", weights_only=False)["state_dict"], strict=False)
    if save_weights and os.path.exists(best_path_score):
        model_score.load_state_dict(torch.load(best_path_score, map_location="cpu", weights_only=False)["state_dict"], strict=False)
    return model, model_score


def build_edge_from_metainfo(meta_info: Dict[str, Any], n_nodes: int):
    """``_build_edge_from_metainfo`` (``training.py:1936-2004``): (edge_index_global [E,2], edge_index_local [E',2]) from
    ``meta_info["node_columns"] / ["edge_columns"]``; local = edges whose two nodes share the animal prefix."""
    names = []
    for c in meta_info["node_columns"]:
        if isinstance(c, tuple) and len(c) == 2 and c[1] == "x":
            names.append(c[0])
            if len(names) == n_nodes:
                break
    if len(names) != n_nodes:
        raise RuntimeError(f"Failed to infer {n_nodes} node names from meta_info['node_columns']. Got {len(names)}.")
    idx = {nm: i for i, nm in enumerate(names)}
    pairs = []
    for u, v in meta_info["edge_columns"]:
        if u not in idx or v not in idx:
            raise RuntimeError(f"Edge ({u},{v}) contains node(s) not found in inferred node list.")
        pairs.append((idx[u], idx[v]))
    eg = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    key = [nm.split("_", 1)[0] if "_" in nm else "" for nm in names]
    same = np.array([key[u] == key[v] for u, v in pairs], dtype=bool) if pairs else np.zeros(0, bool)
    return eg, eg[same]


# ---- train entry ------------------------------------------------------------------------------------------------
def _teacher_for(source: "WindowSource", n_clusters: int, t: Dict[str, Any], dev, world: int, latent_view=None):
    """``maybe_build_turtle_teacher`` (``teacher_model.py:811-905``) on the training windows -> (tau_star [Nw,K], views)."""
    from .teacher import build_turtle_teacher
    x, a = source.all_windows()
    Nw, T, N = x.shape[0], x.shape[1], x.shape[2]
    _, tau_star, views = build_turtle_teacher(
        x, a, n_clusters, latent_view=latent_view, device=dev, include_latent_view=latent_view is not None,
        include_nodes_view=t["include_nodes_view"], include_edges_view=t["include_edges_view"],
        pca_nodes_dim=min(int(t["pca_nodes_dim"]), Nw, T * N), pca_edges_dim=min(int(t["pca_edges_dim"]), Nw, T * a.shape[2]),
        teacher_gamma=t["teacher_gamma"], teacher_alpha_sample_entropy=t["teacher_alpha_sample_entropy"],
        teacher_outer_steps=t["teacher_outer_steps"], teacher_inner_steps=t["teacher_inner_steps"],
        teacher_normalize_feats=t["teacher_normalize_feats"], teacher_head_temp=t["teacher_head_temp"],
        teacher_task_temp=t["teacher_task_temp"], teacher_batch_size=min(int(t["teacher_batch_size"]), Nw),
        batch_size_nodes=min(4096, Nw), verbose=False)
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(tau_star, src=0)                                       # one teacher for all ranks
    return tau_star, views


def train_deepof_model(preprocessed_object=None, adjacency_matrix=None, meta_info=None, encoder_type: str = None,
                       batch_size: int = None, latent_dim: int = None, epochs: int = None, output_path: str = None,
                       n_clusters: int = 10, learning_rate: float = 1e-3, log_history: bool = True, data_path: str = ".",
                       pretrained: Optional[str] = None, save_weights: bool = True, run: int = 0, reg_cat_clusters: float = 0.0,
                       recluster: bool = False, freeze_gmm_epochs: int = 0, freeze_decoder_epochs: int = 0,
                       prior_loss_weight: float = 0.0, gmm_learning_rate: float = 1e-3, learning_rate_pretrain: float = 1e-3,
                       interaction_regularization: float = 0.0003, kmeans_loss: float = 0.0, num_workers: int = 0,
                       prefetch_factor: int = 0, use_amp: bool = False, use_turtle_teacher: bool = True, teacher_gamma: float = 8.0,
                       teacher_outer_steps: int = 500, teacher_inner_steps: int = 100, teacher_normalize_feats: bool = True,
                       lambda_distill: float = 4.0, lambda_decay_start: int = 10, lambda_end_weight: float = 0.2,
                       lambda_cooldown: int = 10, teacher_refresh_every: Optional[int] = False, teacher_freeze_at: Optional[int] = 10,
                       teacher_head_temp: float = 0.5, teacher_task_temp: float = 0.5, teacher_alpha_sample_entropy: float = 2.0,
                       teacher_batch_size: int = 2048, pretrain_epochs: int = 10, kmeans_loss_pretrain: float = 1.0,
                       repel_weight_pretrain: float = 0.5, repel_length_scale_pretrain: float = 0.5,
                       nonempty_weight_pretrain: float = 2e-2, nonempty_p_pretrain: float = 2.0,
                       nonempty_floor_percent_pretrain: float = 0.05, kl_annealing_mode: str = "tf_sigmoid", kl_max_weight: float = 1,
                       kl_warmup: int = 5, kl_end_weight: float = 0.2, kl_cooldown: int = 5,
                       kl_annealing_mode_pretrain: str = "tf_sigmoid", kl_max_weight_pretrain: float = 0.2,
                       kl_warmup_pretrain: int = 15, kl_end_weight_pretrain: float = 0.2, kl_cooldown_pretrain: int = 10,
                       reg_scatter_weight: float = 0, temporal_cohesion_weight: float = 0, reg_scatter_beta: float = 1.0,
                       repel_weight: float = 0, repel_length_scale: float = 1.0, main_clustering_loss: float = 0.0,
                       nonempty_weight: float = 2e-2, nonempty_floor_percent: float = 0.05, nonempty_p: float = 2.0,
                       distill_conf_weight: bool = False, distill_conf_thresh: float = 0.3, distill_sharpen_T: float = 0.5,
                       include_edges_view: bool = False, include_nodes_view: bool = True, pca_nodes_dim: int = 32,
                       pca_edges_dim: int = 32, include_angles_view: bool = False, pca_angles_dim: int = 32,
                       reinit_gmm_on_refresh: bool = False, diag_max_batches: int = 4, model_name: str = "VaDE",
                       generic_lambda_distill: float = 2.0, generic_distill_sharpen_T: float = 0.5,
                       generic_distill_conf_weight: bool = True, generic_distill_conf_thresh: float = 0.6,
                       generic_distill_warmup_epochs: int = 1, distill_class_reweight_beta: float = 1,
                       distill_class_reweight_cap: float = 3, temperature: float = 0.1,
                       contrastive_similarity_function: str = "cosine", contrastive_loss_function: str = "nce", beta: float = 0.1,
                       tau: float = 0.1, aug_min_shift: int = 1, aug_max_shift: int = 3, aug_p_shift: float = 0.4,
                       aug_max_rot: float = 30, aug_n_rot: int = 3, aug_p_rot: float = 0.8, aug_max_interp: int = 8,
                       aug_min_interp: int = 3, aug_p_interp: float = 0.4, aug_noise_sigma: float = 0.03, aug_p_noise: float = 0.4,
                       device: Optional[str] = None, h5_dataset_folder: Optional[str] = None, bootstrap_training: Optional[bool] = False,
                       bootstrap_block_len: int = 250, random_seed: int = 0):
    """Drop-in for ``deepof.clustering.training.train_deepof_model`` (``training.py:592-905``) with the reference's
    keyword list and defaults, on one rank or one rank of a torchrun job (RANK / WORLD_SIZE are honoured).

    ``preprocessed_object = (train, val)``: each either a preprocessed TableDict ``td[key] = (nodes [Nw,T,3N], edges
    [Nw,T,E], ...)`` (materialised windows, as in the reference) or a :class:`~deepof_b200.loader.WindowLoader` over raw
    pose frames (windows are produced per batch by the loader kernel).

    Returns ``(model_val, model_score, teacher_init_model, log_summary)`` like ``fit_VADE / fit_VQVAE / fit_contrastive``
    (``training.py:1087-1918``): per-epoch validation, ``compute_diagnostics`` / alignment score, best-val and best-score
    checkpoints under ``<output_path>/models/<model>/run_<run>/``; with ``pretrained=path`` ``(model, None, None,
    log_summary)``.

    No effect here (data plumbing of the reference's CPU loader, TensorBoard): ``log_history, data_path, num_workers,
    prefetch_factor, h5_dataset_folder, interaction_regularization, recluster, generic_lambda_distill,
    generic_distill_warmup_epochs, bootstrap_block_len`` (the last three are unused by the reference's fit functions too).
    """
    import math
    if device == "cpu":
        raise ValueError("deepof_b200 has no CPU path (device must be None or 'gpu')")
    if device not in (None, "gpu"):
        raise ValueError("If a device is given, it needs to be either cpu or gpu!")      # training.py:935
    if pretrained is not None:
        model, log_summary = load_model_from_ckpt(pretrained, max_batch=int(batch_size or 4096))
        return model, None, None, log_summary
    name = str(model_name).lower()
    encoder_type = {"recurrent": "recurrent", "transformer": "transformer", "tcn": "TCN"}.get(str(encoder_type or "recurrent").lower())
    if encoder_type is None:
        raise NotImplementedError('invalid encoder type, try "recurrent", "TCN" or "transformer"')         # models_new.py:1499-1502
    unsupported = {"use_amp": use_amp, "bootstrap_training": bootstrap_training, "reg_scatter_weight": reg_scatter_weight,
                   "main_clustering_loss": main_clustering_loss, "prior_loss_weight": prior_loss_weight,
                   "include_angles_view": include_angles_view}
    bad = [k for k, v in unsupported.items() if v]
    if bad:
        raise NotImplementedError(f"options not implemented by deepof_b200 (reference defaults are off / 0): {bad}")
    assert batch_size and latent_dim and epochs is not None and adjacency_matrix is not None and preprocessed_object is not None
    if output_path is None:
        output_path = "."
    adj = np.asarray(adjacency_matrix, dtype=np.float64)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    is_main = rank == 0
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(random_seed)                                             # training.py:947-948
    torch.cuda.manual_seed_all(random_seed)
    np.random.seed(random_seed)
    if world > 1:
        import torch.distributed as dist
    bs, K, D = int(batch_size), int(n_clusters), int(latent_dim)
    train_obj, val_obj = preprocessed_object
    train = _as_source(train_obj, dev, bs, random_seed, rank, world, shuffle=True)
    val = _as_source(val_obj if val_obj else train_obj, dev, bs, random_seed, 0, 1, shuffle=False)
    (T, N, F), (_, E, _) = train.x_shape, train.a_shape
    nb = max(1, len(train))
    tc = dict(teacher_gamma=teacher_gamma, teacher_outer_steps=teacher_outer_steps, teacher_inner_steps=teacher_inner_steps,
              teacher_normalize_feats=teacher_normalize_feats, teacher_head_temp=teacher_head_temp, teacher_task_temp=teacher_task_temp,
              teacher_alpha_sample_entropy=teacher_alpha_sample_entropy, teacher_batch_size=teacher_batch_size,
              include_nodes_view=include_nodes_view, include_edges_view=include_edges_view, pca_nodes_dim=pca_nodes_dim,
              pca_edges_dim=pca_edges_dim)
    refresh = None if teacher_refresh_every is False else teacher_refresh_every
    rebuild_spec = {"model_name": name, "x_shape": (T, N, F), "a_shape": (T, E, 1), "adjacency_matrix": adj.astype("float32"),
                    "latent_dim": D, "n_components": K, "encoder_type": encoder_type, "use_gnn": True,
                    "interaction_regularization": float(interaction_regularization)}
    if name == "vade":
        rebuild_spec.update(kmeans_loss=float(kmeans_loss), lens_enabled=False)
    _, best_path_val, best_path_score, teacher_init_path = ckpt_paths(name, output_path, run)
    log_summary = init_log_summary(name)
    score_start_epoch = max(3, math.ceil(0.1 * int(epochs)))
    score_tol = 0.01
    mk = dict(encoder_type=encoder_type, max_batch=bs, training=True, seed=random_seed)

    def lambda_schedule(mode):
        return KLSchedule(nb, mode, 0, lambda_distill, lambda_cooldown, lambda_end_weight, at_max_epochs=lambda_decay_start)

    def save(path, stage, model, epoch, val_total, score_value=None):
        if save_weights and is_main:
            save_model_info(path, model=model, rebuild_spec=rebuild_spec, log_summary=log_summary, stage=stage, epoch=epoch,
                            train_steps=(epoch + 1) * nb, val_total=val_total, score_value=score_value, save_weights=True)

    val_batches = lambda: val.sequential(bs, min_rows=2)

    # ------------------------------------------------------------------------------------------------ VaDE (fit_VADE)
    if name == "vade":
        model = VaDEB200((T, N, F), (T, E, 1), adj, D, K, kmeans_loss=kmeans_loss_pretrain, **mk)
        if world > 1:
            dist.broadcast(model.state, src=0)
        crit = VadeLossCfg.pretrain_defaults(K)
        crit.kmeans_loss_weight, crit.model_kmeans_weight = float(kmeans_loss_pretrain), float(kmeans_loss_pretrain)
        crit.repel_weight, crit.repel_length_scale = float(repel_weight_pretrain), float(repel_length_scale_pretrain)
        crit.nonempty_weight, crit.nonempty_p = float(nonempty_weight_pretrain), int(nonempty_p_pretrain)
        crit.nonempty_floor = max(1e-4, float(nonempty_floor_percent_pretrain) / K)
        model.set_pretrain_mode(True)
        ctx = SimpleNamespace(criterion=crit, apply_distill=False,
                              kl_scheduler=KLSchedule(nb, str(kl_annealing_mode_pretrain).lower(), kl_warmup_pretrain,
                                                      kl_max_weight_pretrain, kl_cooldown_pretrain, kl_end_weight_pretrain))
        opt = {"lr": learning_rate_pretrain, "gmm_lr": 0.0}
        for ep in range(int(pretrain_epochs)):                                # training.py:1617-1636 (pretrain logs go to TensorBoard only)
            train_one_epoch_indexed(model, "vade", train, opt, step_vade, ep, pretrain_epochs, 0.75, ctx, world)
        model.set_pretrain_mode(False)                                        # training.py:1643-1653
        crit = VadeLossCfg.main_defaults(K)
        crit.kmeans_loss_weight, crit.model_kmeans_weight = float(kmeans_loss), float(kmeans_loss_pretrain)   # training.py:1556
        crit.repel_weight, crit.repel_length_scale = float(repel_weight), float(repel_length_scale)
        crit.nonempty_weight, crit.nonempty_p = float(nonempty_weight), int(nonempty_p)
        crit.nonempty_floor = max(1e-4, float(nonempty_floor_percent) / K)
        crit.temporal_cohesion_weight, crit.reg_cat_clusters_weight = float(temporal_cohesion_weight), float(reg_cat_clusters)
        crit.distill_sharpen_T, crit.distill_conf_weight, crit.distill_conf_thresh = distill_sharpen_T, bool(distill_conf_weight), float(distill_conf_thresh)
        kl_sched = KLSchedule(nb, str(kl_annealing_mode).lower(), kl_warmup, kl_max_weight, kl_cooldown, kl_end_weight)
        model.adam_m.zero_(); model.adam_v.zero_(); model.adam_steps = [0, 0, 0, 0]        # the optimizer is rebuilt
        opt = {"lr": learning_rate, "gmm_lr": gmm_learning_rate, "active": (True, True, True)}
        teacher_ctx: Dict[str, Any] = {}
        teacher_init_model, tau_star, views, lam_sched = None, None, None, None
        if use_turtle_teacher:                                                # training.py:1664-1712
            from .teacher import initialize_gmm_from_teacher, run_turtle_teacher_on_views, teacher_context
            z_all = torch.cat([model.embed(x, a)[0] for x, a, _ in train.sequential(2048)])     # extract_latents
            lam_sched = lambda_schedule(str(kl_annealing_mode).lower())
            tau_star, views = _teacher_for(train, K, tc, dev, world, latent_view=z_all)
            initialize_gmm_from_teacher(model, z_all, tau_star, min_var=0.01, verbose=False)

            def set_teacher(ts):
                t = teacher_context(ts, True, distill_class_reweight_beta, distill_class_reweight_cap)
                crit.lambda_distill = float(lambda_distill)                   # constant: fit_VADE's ctx has no lambda_scheduler
                teacher_ctx.update(tau_star=t["tau_star"], class_weight=t["class_weight"], teacher_marginal=t["teacher_marginal"],
                                   criterion_lambda_scheduler=lam_sched, apply_distill=True)
            set_teacher(tau_star)
            teacher_init_model = build_model(rebuild_spec, max_batch=bs, training=False)
            teacher_init_model.load_state_dict(model.state_dict())
            if save_weights and is_main:
                save_model_info(teacher_init_path, model=model, rebuild_spec=rebuild_spec, log_summary=log_summary, stage="teacher_init",
                                epoch=int(pretrain_epochs) - 1, train_steps=int(pretrain_epochs) * nb,
                                note="after pretrain + teacher + GMM init, before main training")
        else:                                                                 # training.py:1719-1722
            model.initialize_gmm_from_data(train_obj_iter(train))
        best_val, best_score, best_score_val = -float("inf"), -float("inf"), float("inf")
        val_tol, val_top_reached = 0.01, False
        for epoch in range(int(epochs)):
            if epoch == 0 and freeze_gmm_epochs > 0:                          # training.py:1746-1767
                opt["active"] = (opt["active"][0], opt["active"][1], False)
            if epoch == freeze_gmm_epochs:
                opt.update(lr=5e-4, gmm_lr=2e-4, active=(opt["active"][0], opt["active"][1], True))
            if epoch == 0 and freeze_decoder_epochs > 0:
                opt["active"] = (opt["active"][0], False, opt["active"][2])
            if epoch == freeze_decoder_epochs:
                opt["active"] = (opt["active"][0], True, opt["active"][2])
            if (epoch > 0 and use_turtle_teacher and refresh is not None and refresh > 0 and epoch % refresh == 0
                    and (teacher_freeze_at is None or epoch <= teacher_freeze_at)):          # training.py:1770-1802
                z_curr = torch.cat([model.embed(x, a)[0] for x, a, _ in train.sequential(2048)])
                vd = {"z": z_curr}
                vd.update({k: v for k, v in views.items() if k != "z"})
                _, tau_star = run_turtle_teacher_on_views(
                    vd, K, gamma=teacher_gamma, alpha_sample_entropy=teacher_alpha_sample_entropy,
                    outer_steps=max(200, int(teacher_outer_steps)), inner_steps=teacher_inner_steps,
                    normalize_feats=teacher_normalize_feats, verbose=False, device=dev, head_temp=teacher_head_temp,
                    task_temp=teacher_task_temp, batch_size=min(int(teacher_batch_size), z_curr.shape[0]))
                tau_star = tau_star.detach()
                if world > 1:
                    dist.broadcast(tau_star, src=0)
                set_teacher(tau_star)
                if reinit_gmm_on_refresh:
                    initialize_gmm_from_teacher(model, z_curr, tau_star, min_var=1e-4, verbose=False)
            ctx = SimpleNamespace(criterion=crit, apply_distill=False, kl_scheduler=kl_sched)
            ctx.__dict__.update(teacher_ctx)
            train_logs, klw, lambda_d = train_one_epoch_indexed(model, "vade", train, opt, step_vade, epoch, epochs, 0.75, ctx, world)
            val_logs = validate_one_epoch_indexed(model, "vade", val_batches(), step_vade, epoch, epochs,
                                                  SimpleNamespace(criterion=crit, apply_distill=False, kl_scheduler=kl_sched), world, rank)
            val_logs.update(compute_diagnostics(model, val_batches(), get_q_vade, K, tau_star=teacher_ctx.get("tau_star"),
                                                distill_sharpen_T=float(distill_sharpen_T), distill_conf_weight=bool(distill_conf_weight),
                                                distill_conf_thresh=float(distill_conf_thresh), max_batches=int(diag_max_batches),
                                                extra_stats_fn=compute_vade_specific_diagnostics))
            val_total = float(val_logs.get("total_loss", float("inf")))
            score_value = float(val_logs["alignment_score"])
            train_logs = dict(train_logs, kl_weight_mid=klw, lambda_distill_mid=lambda_d)
            log_summary = update_log_summary(log_summary, train_logs, val_logs)
            improved_val = (val_total + val_tol) < best_val                   # training.py:1836-1847 (VaDE's "top, then improve" rule)
            if not improved_val and not val_top_reached:
                best_val = val_total
            improved_score = math.isfinite(score_value) and ((score_value > best_score) or
                                                             (abs(score_value - best_score) <= score_tol and val_total < best_score_val))
            if improved_val:
                val_top_reached, best_val, val_tol = True, val_total, 0.0
                save(best_path_val, "best_val", model, epoch, val_total)
            if improved_score and epoch > score_start_epoch:
                best_score, best_score_val = score_value, val_total
                save(best_path_score, "best_score", model, epoch, val_total, score_value)
        if world > 1:
            dist.barrier()
        model_val, model_score = load_best_checkpoints(model, rebuild_spec, best_path_val, best_path_score, save_weights, bs)
        return model_val, model_score, teacher_init_model, log_summary

    # ------------------------------------------------------------------- VQ-VAE / contrastive (fit_VQVAE / fit_contrastive)
    if name == "vqvae":
        model = VQVAEB200((T, N, F), (T, E, 1), adj, D, K, kmeans_loss=kmeans_loss, **mk)
        step_fn, base_ctx = step_vqvae_distill, {}
        q_fn = lambda head: (lambda m, x, a: get_q_vqvae(m, x, a, distill_head=head))
    elif name == "contrastive":
        eg = el = None
        if meta_info is not None and "node_columns" in meta_info and "edge_columns" in meta_info:
            eg, el = build_edge_from_metainfo(meta_info, N)                   # training.py:1322-1328
        model = ContrastiveB200((T, N, F), (T, E, 1), adj, D, temperature=temperature,
                                similarity_function=str(contrastive_similarity_function).lower(),
                                loss_function=str(contrastive_loss_function).lower(), beta=beta, tau=tau, edge_index=eg,
                                edge_index_local=el, **mk)
        aug = ContrastiveAugCfg(aug_min_shift, aug_max_shift, aug_p_shift, aug_max_rot, aug_n_rot, aug_p_rot, aug_max_interp,
                                aug_min_interp, aug_p_interp, aug_noise_sigma, aug_p_noise)
        gen = torch.Generator(device=dev).manual_seed(random_seed + 7919 * rank)
        hgen = torch.Generator().manual_seed(random_seed + 7919 * rank)
        step_fn, base_ctx = step_contrastive_distill, dict(contrastive_cfg=aug, generator=gen, host_generator=hgen)
        q_fn = lambda head: (lambda m, x, a: get_q_contrastive(m, x, a, distill_head=head))
    else:
        raise ValueError(f"Unsupported model: {model_name}")
    if world > 1:
        dist.broadcast(model.state, src=0)
    tau_star = None
    if use_turtle_teacher:                                                    # training.py:1099-1124, 1342-1368: no latent view
        tau_star, _ = _teacher_for(train, K, tc, dev, world, latent_view=None)
    apply_distill = tau_star is not None
    lam_sched = lambda_schedule("tf_sigmoid") if apply_distill else None
    head = DistillHeadB200(D, K, device=dev, seed=random_seed)
    if world > 1:
        dist.broadcast(head.state, src=0)
    opt = {"lr": learning_rate, "weight_decay": 1e-4}
    best_val, best_score, best_score_val = float("inf"), -float("inf"), float("inf")
    for epoch in range(int(epochs)):
        ctx = SimpleNamespace(tau_star=tau_star, distill_head=head, lambda_scheduler=lam_sched,
                              distill_sharpen_T=generic_distill_sharpen_T, distill_conf_weight=generic_distill_conf_weight,
                              distill_conf_thresh=generic_distill_conf_thresh, apply_distill=apply_distill, **base_ctx)
        train_logs, _, lam = train_one_epoch_indexed(model, name, train, opt, step_fn, epoch, epochs, 0.75, ctx, world)
        val_logs = validate_one_epoch_indexed(model, name, val_batches(), step_fn, epoch, epochs,
                                              SimpleNamespace(apply_distill=False, **base_ctx), world, rank)
        v_total = float(val_logs.get("total_loss", float("inf")))
        score_value = float("nan")
        if apply_distill:
            val_logs.update(compute_diagnostics(model, val_batches(), q_fn(head), K, tau_star=tau_star,
                                                distill_sharpen_T=generic_distill_sharpen_T, distill_conf_weight=generic_distill_conf_weight,
                                                distill_conf_thresh=generic_distill_conf_thresh, max_batches=int(diag_max_batches)))
            score_value = float(val_logs["alignment_score"])
        else:
            val_logs.update(alignment_score=float("nan"), conf_norm=float("nan"), bal_norm=float("nan"))
        log_summary = update_log_summary(log_summary, dict(train_logs, lambda_distill_mid=lam), val_logs)
        if v_total < best_val:
            best_val = v_total
            save(best_path_val, "best_val", model, epoch, v_total)
        improved_score = apply_distill and math.isfinite(score_value) and (
            (score_value > best_score) or (abs(score_value - best_score) <= score_tol and v_total < best_score_val))
        if improved_score and epoch > score_start_epoch:
            best_score, best_score_val = score_value, v_total
            save(best_path_score, "best_score", model, epoch, v_total, score_value)
    if world > 1:
        dist.barrier()
    model_val, model_score = load_best_checkpoints(model, rebuild_spec, best_path_val, best_path_score, save_weights, bs)
    model_val.distill_head = head
    return model_val, model_score, None, log_summary


def train_obj_iter(source: "WindowSource"):
    """The un-shuffled pass ``initialize_gmm_from_data`` makes over the train loader's first batches.  (The reference
    iterates its shuffled loader; which 10 000 windows seed the mixture is not part of the parity contract, the
    scikit-learn call and what is stored are.)"""
    return source.sequential(source.bs)
