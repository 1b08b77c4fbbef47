"""Reference-facing entry points of the B200 path: the step functions, the epoch loop, ``train_deepof_model`` and the
checkpoint bundle, with the names, argument meaning and return shapes of ``deepof/clustering/training.py`` and
``model_utils_new.py``, for the configurations this library implements
(``encoder_type="recurrent"``, GNN path, TURTLE teacher off).

* ``step_vade / step_vqvae_distill / step_contrastive_distill(model, (x, a, idx), ctx) -> StepResult(loss, logs)``
  (reference ``training.py:231-309, 312-389, 482-589``).  The CUDA library fuses forward, loss and backward, so the
  gradient is already in ``model.grad`` when the step function returns; ``StepResult.loss`` is a device scalar whose
  ``backward()`` is a no-op kept for loop compatibility, ``logs`` holds the reference's keys.
* ``train_one_epoch_indexed`` (``training.py:104-187``): step, [all-reduce], clip_grad_value_(0.75) + Adam, schedulers.
* ``train_deepof_model`` (``training.py:592-905``) -> ``(model_val, model_score, teacher_init_model, log_summary)``.
* ``save_model_info / load_model_from_ckpt`` (``model_utils_new.py:263-329, 367-417``): ``torch.save({"state_dict",
  "rebuild_spec", "log_summary"})`` with the reference's keys, so checkpoints move between the two implementations.

Unsupported options raise instead of being ignored (``use_turtle_teacher=True`` outside the VaDE run, teacher refresh,
other encoders, AMP).
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
        cfg.lambda_distill = float(lsched.get_weight())                       # training.py:264-265
    tau = None
    tau_star = getattr(ctx, "tau_star", None)
    if tau_star is not None and getattr(ctx, "apply_distill", True) and cfg.lambda_distill > 0.0:
        tau = tau_star[idx.to(model.device).long()]
    logs = model.loss_grad(x, a, cfg, eps=getattr(ctx, "eps", None), mc_eps=getattr(ctx, "mc_eps", None), tau_batch=tau,
                           class_weight=getattr(ctx, "class_weight", None), teacher_marginal=getattr(ctx, "teacher_marginal", None))
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
    logs = model.loss_grad(x, a, distill=_distillation(model, idx, ctx))
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
    logs = model.loss_grad(x_full, prm, distill=_distillation(model, batch[2] if len(batch) > 2 else None, ctx))
    return StepResult(loss=_Loss(logs[0]), logs=_logs(model) if getattr(ctx, "read_logs", True) else {})


def average_logs(logs_list: Iterable[Dict[str, float]]) -> Dict[str, float]:
    logs_list = list(logs_list)
    if not logs_list:
        return {}
    return {k: float(np.mean([l[k] for l in logs_list if k in l])) for k in logs_list[0]}


def train_one_epoch_indexed(model, model_name: str, dataloader, optimizer, step_fn, epoch: int = 0, num_epochs: int = 1,
                            grad_clip_value: Optional[float] = 0.75, ctx: Optional[SimpleNamespace] = None,
                            world_size: int = 1, log_every: int = 0):
    """One epoch (reference ``training.py:104-187``).  ``dataloader`` yields ``(x, a, idx)``; ``optimizer`` is a dict of
    the Adam hyper-parameters: ``{"lr": .., "gmm_lr": .., "weight_decay": ..}``.  Returns
    ``(averaged logs, kl weight at mid epoch, distillation weight at mid epoch)`` like the reference.  Logs are read back every ``log_every``
    steps (0: only the last step of the epoch) so the loop does not synchronise per step."""
    ctx = ctx or SimpleNamespace()
    logs_accum, mid_kl, mid_lambda = [], 0.0, 0.0
    batches = list(dataloader) if not hasattr(dataloader, "__len__") else dataloader
    n = len(batches)
    for step, batch in enumerate(batches):
        read = (log_every > 0 and step % log_every == 0) or step == n - 1
        sub = SimpleNamespace(**{**ctx.__dict__, "train": True, "epoch": epoch, "num_epochs": num_epochs, "read_logs": read})
        res = step_fn(model, batch, sub)
        res.loss.backward()
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
            lsched = getattr(ctx, "lambda_scheduler", None)
            if lsched is not None:                                            # training.py:173-176
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
        if res.logs:
            logs_accum.append(res.logs)
    return average_logs(logs_accum), mid_kl, mid_lambda


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


class _Batches:
    """Contiguous batches with the epoch-seeded shuffle of batch starts and ``starts[rank::world]``
    (``dataset.py:561-671``)."""

    def __init__(self, x, a, batch_size, seed, rank=0, world=1, shuffle=True):
        self.x, self.a, self.bs, self.seed, self.rank, self.world, self.shuffle = x, a, int(batch_size), seed, rank, world, shuffle
        self.epoch = 0

    def __iter__(self):
        self.epoch += 1
        n = self.x.shape[0]
        for s in batch_starts(n, self.bs, self.epoch, self.seed, self.shuffle, self.rank, self.world):
            e = min(n, int(s) + self.bs)
            yield self.x[s:e], self.a[s:e], torch.arange(int(s), e, device=self.x.device)

    def __len__(self):
        return len(batch_starts(self.x.shape[0], self.bs, 1, self.seed, False, self.rank, self.world))


# ---- checkpoints ----------------------------------------------------------------------------------------------
def save_model_info(ckpt_path: str, *, model, rebuild_spec: Dict[str, Any], log_summary: Optional[Dict[str, Any]] = None,
                    stage: str = "final", save_weights: bool = True) -> None:
    """The reference's checkpoint bundle (``model_utils_new.py:263-329``)."""
    os.makedirs(os.path.dirname(os.path.abspath(ckpt_path)), exist_ok=True)
    if save_weights:
        payload = {"state_dict": {k: v.cpu() for k, v in model.state_dict().items()}, "rebuild_spec": rebuild_spec}
        if log_summary is not None:
            payload["log_summary"] = log_summary
        torch.save(payload, ckpt_path)
    with open(os.path.splitext(ckpt_path)[0] + "_info.txt", "w", encoding="utf-8") as f:
        f.write(f"stage: {stage}\n\n[checkpoint_format]\nckpt_contains: bundle\nbundle_keys: state_dict, rebuild_spec"
                + (", log_summary" if log_summary is not None else "") + "\n")


def build_model(rebuild_spec: Dict[str, Any], max_batch: int = 4096, training: bool = False, device: Optional[int] = None):
    """Model object from a reference ``rebuild_spec`` (``model_utils_new.py:367-417``)."""
    name = str(rebuild_spec["model_name"]).lower()
    etype = rebuild_spec.get("encoder_type", "recurrent")
    if etype not in ("recurrent", "transformer") or not rebuild_spec.get("use_gnn", True):
        raise NotImplementedError("deepof_b200 implements encoder_type='recurrent' (training + inference) and 'transformer' "
                                  "(inference), use_gnn=True")
    xs, as_ = tuple(rebuild_spec["x_shape"]), tuple(rebuild_spec["a_shape"])
    adj, D, K = np.asarray(rebuild_spec["adjacency_matrix"]), int(rebuild_spec["latent_dim"]), int(rebuild_spec.get("n_components", 1))
    if etype == "transformer":
        if training:
            raise NotImplementedError("the training step of the transformer model family is not built (inference only)")
        return TFMModelB200(name, xs, as_, adj, D, K, max_batch=max_batch, device=device)
    kw = dict(max_batch=max_batch, training=training, device=device)
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
    model.load_state_dict(bundle["state_dict"])
    return model, bundle.get("log_summary")


# ---- train entry ------------------------------------------------------------------------------------------------
# teacher / distillation keywords of the reference's train_deepof_model with its defaults (training.py:624-640, 691-692)
_TEACHER_KWARGS = dict(teacher_gamma=8.0, teacher_outer_steps=500, teacher_inner_steps=100, teacher_normalize_feats=True,
                       lambda_distill=4.0, lambda_decay_start=10, lambda_end_weight=0.2, lambda_cooldown=10,
                       teacher_refresh_every=False, teacher_freeze_at=10, teacher_head_temp=0.5, teacher_task_temp=0.5,
                       teacher_alpha_sample_entropy=2.0, teacher_batch_size=2048, distill_class_reweight_beta=1.0,
                       distill_class_reweight_cap=3.0)
def train_deepof_model(preprocessed_object=None, adjacency_matrix=None, meta_info=None, encoder_type: str = "recurrent",
                       batch_size: int = 1024, latent_dim: int = 8, epochs: int = 10, output_path: Optional[str] = None,
                       n_clusters: int = 10, learning_rate: float = 1e-3, pretrained: Optional[str] = None,
                       save_weights: bool = True, gmm_learning_rate: float = 1e-3, learning_rate_pretrain: float = 1e-3,
                       kmeans_loss: float = 0.0, use_amp: bool = False, use_turtle_teacher: bool = True,
                       pretrain_epochs: int = 10, kmeans_loss_pretrain: float = 1.0, repel_weight_pretrain: float = 0.5,
                       repel_length_scale_pretrain: float = 0.5, nonempty_weight_pretrain: float = 2e-2,
                       nonempty_p_pretrain: float = 2.0, nonempty_floor_percent_pretrain: float = 0.05,
                       kl_annealing_mode: str = "tf_sigmoid", kl_max_weight: float = 1, kl_warmup: int = 5,
                       kl_end_weight: float = 0.2, kl_cooldown: int = 5, kl_annealing_mode_pretrain: str = "tf_sigmoid",
                       kl_max_weight_pretrain: float = 0.2, kl_warmup_pretrain: int = 15, kl_end_weight_pretrain: float = 0.2,
                       kl_cooldown_pretrain: int = 10, temporal_cohesion_weight: float = 0, reg_cat_clusters: float = 0.0,
                       repel_weight: float = 0, repel_length_scale: float = 1.0, nonempty_weight: float = 2e-2,
                       nonempty_floor_percent: float = 0.05, nonempty_p: float = 2.0, model_name: str = "VaDE",
                       temperature: float = 0.1, contrastive_similarity_function: str = "cosine",
                       contrastive_loss_function: str = "nce", beta: float = 0.1, tau: float = 0.1, aug_min_shift: int = 1,
                       aug_max_shift: int = 3, aug_p_shift: float = 0.4, aug_max_rot: float = 30, aug_n_rot: int = 3,
                       aug_p_rot: float = 0.8, aug_max_interp: int = 8, aug_min_interp: int = 3, aug_p_interp: float = 0.4,
                       aug_noise_sigma: float = 0.03, aug_p_noise: float = 0.4, device: Optional[str] = None,
                       random_seed: int = 0, freeze_gmm_epochs: int = 0, **unsupported):
    """Drop-in for ``deepof.clustering.training.train_deepof_model`` (``training.py:592-905``) on one rank (or one rank
    of a torchrun job: RANK / WORLD_SIZE are honoured).  ``preprocessed_object = (train_td, val_td)`` with
    ``td[key] = (nodes [Nw,T,3N], edges [Nw,T,E], ...)``.  Returns ``(model_val, model_score, teacher_init_model,
    log_summary)``; with ``pretrained=path`` ``(model, None, None, log_summary)``."""
    if device == "cpu":
        raise ValueError("deepof_b200 has no CPU path (device must be None or 'gpu')")
    if device not in (None, "gpu"):
        raise ValueError(f"Invalid device '{device}'")                      # training.py:935
    if pretrained is not None:
        model, log_summary = load_model_from_ckpt(pretrained, max_batch=int(batch_size or 4096))
        return model, None, None, log_summary
    if encoder_type != "recurrent":
        raise NotImplementedError("deepof_b200 implements encoder_type='recurrent' only (transformer / TCN: next round)")
    if use_turtle_teacher and model_name.lower() != "vade":
        raise NotImplementedError("the TURTLE teacher is wired into the VaDE run only: call with use_turtle_teacher=False "
                                  "(a precomputed tau_star goes through ctx of the step functions)")
    tk = {k: unsupported.pop(k) for k in list(unsupported) if k in _TEACHER_KWARGS}
    tcfg = {**_TEACHER_KWARGS, **tk}
    if tcfg["teacher_refresh_every"]:
        raise NotImplementedError("teacher_refresh_every is not mirrored (the reference default is no refresh)")
    if use_amp:
        raise NotImplementedError("AMP is not used: the B200 path computes in fp32-class precision (3xTF32)")
    bad = [k for k, v in unsupported.items() if k in ("main_clustering_loss", "reg_scatter_weight") and v]
    if bad:
        raise NotImplementedError(f"unsupported non-zero options: {bad}")
    train_td, val_td = preprocessed_object
    adj = np.asarray(adjacency_matrix, dtype=np.float64)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", torch.cuda.current_device())
    x, a = windows_from_table_dict(train_td, dev)
    xv, av = windows_from_table_dict(val_td, dev) if val_td else (x, a)
    Nw, T, N, F = x.shape
    E = a.shape[2]
    name = model_name.lower()
    loader = _Batches(x, a, batch_size, random_seed, rank, world)
    nb = max(1, len(loader))
    log_summary: Dict[str, Any] = {"model_name": name, "epochs": int(epochs), "train_logs": [], "val_logs": []}
    rebuild_spec = {"model_name": name, "x_shape": (T, N, F), "a_shape": (T, E, 1), "adjacency_matrix": adj.astype("float32"),
                    "latent_dim": int(latent_dim), "n_components": int(n_clusters), "encoder_type": encoder_type, "use_gnn": True,
                    "kmeans_loss": float(kmeans_loss), "interaction_regularization": 0.0, "lens_enabled": False}
    if world > 1:
        import torch.distributed as dist

    def validate(model, step_fn, ctx):
        vl = []
        bs = int(batch_size)
        for s in range(0, xv.shape[0], bs):
            xb, ab = xv[s:s + bs], av[s:s + bs]
            if xb.shape[0] < 2:
                continue
            res = step_fn(model, (xb, ab, torch.arange(s, s + xb.shape[0], device=dev)), ctx)
            vl.append(res.logs)
        return average_logs(vl)

    if name == "vade":
        model = VaDEB200((T, N, F), (T, E, 1), adj, latent_dim, n_clusters, kmeans_loss=kmeans_loss_pretrain, max_batch=int(batch_size),
                         training=True, seed=random_seed)
        if world > 1:
            dist.broadcast(model.state, src=0)
        crit = VadeLossCfg.pretrain_defaults(n_clusters)
        crit.kmeans_loss_weight, crit.model_kmeans_weight = float(kmeans_loss_pretrain), float(kmeans_loss_pretrain)
        crit.repel_weight, crit.repel_length_scale = float(repel_weight_pretrain), float(repel_length_scale_pretrain)
        crit.nonempty_weight, crit.nonempty_p = float(nonempty_weight_pretrain), int(nonempty_p_pretrain)
        crit.nonempty_floor = max(1e-4, float(nonempty_floor_percent_pretrain) / n_clusters)
        model.set_pretrain_mode(True)
        ctx = SimpleNamespace(criterion=crit, apply_distill=False,
                              kl_scheduler=KLSchedule(nb, kl_annealing_mode_pretrain, kl_warmup_pretrain, kl_max_weight_pretrain,
                                                      kl_cooldown_pretrain, kl_end_weight_pretrain))
        opt = {"lr": learning_rate_pretrain, "gmm_lr": 0.0}
        for ep in range(int(pretrain_epochs)):                                # training.py:1617-1636
            logs, _, _ = train_one_epoch_indexed(model, "vade", loader, opt, step_vade, ep, pretrain_epochs, 0.75, ctx, world)
            log_summary["train_logs"].append({"phase": "pretrain", "epoch": ep, **logs})
        model.set_pretrain_mode(False)                                        # training.py:1643-1653
        teacher_ctx, teacher_init_model = {}, None
        if use_turtle_teacher:                                                # training.py:1664-1712
            from .teacher import build_turtle_teacher, initialize_gmm_from_teacher, teacher_context
            z_all = model.embed(x, a)[0]                                      # extract_latents: z_mean in eval mode
            lam = KLSchedule(nb, kl_annealing_mode, 0, tcfg["lambda_distill"], tcfg["lambda_cooldown"],
                             tcfg["lambda_end_weight"], at_max_epochs=tcfg["lambda_decay_start"])
            _, tau_star, _ = build_turtle_teacher(
                x, a, n_clusters, latent_view=z_all, device=dev, include_latent_view=True,
                teacher_gamma=tcfg["teacher_gamma"], teacher_alpha_sample_entropy=tcfg["teacher_alpha_sample_entropy"],
                teacher_outer_steps=tcfg["teacher_outer_steps"], teacher_inner_steps=tcfg["teacher_inner_steps"],
                teacher_normalize_feats=tcfg["teacher_normalize_feats"], teacher_head_temp=tcfg["teacher_head_temp"],
                teacher_task_temp=tcfg["teacher_task_temp"], teacher_batch_size=min(tcfg["teacher_batch_size"], Nw),
                batch_size_nodes=min(4096, Nw), pca_nodes_dim=min(32, Nw, T * N), verbose=False)
            if world > 1:
                dist.broadcast(tau_star, src=0)                               # one teacher for all ranks
            initialize_gmm_from_teacher(model, z_all, tau_star, min_var=0.01, verbose=False)
            tc = teacher_context(tau_star, True, tcfg["distill_class_reweight_beta"], tcfg["distill_class_reweight_cap"])
            teacher_ctx = dict(tau_star=tc["tau_star"], class_weight=tc["class_weight"], teacher_marginal=tc["teacher_marginal"],
                               lambda_scheduler=lam, apply_distill=True)
            teacher_init_model = build_model(rebuild_spec, max_batch=int(batch_size), training=False)
            teacher_init_model.load_state_dict(model.state_dict())
        crit = VadeLossCfg.main_defaults(n_clusters)
        crit.kmeans_loss_weight, crit.model_kmeans_weight = float(kmeans_loss), float(kmeans_loss_pretrain)   # training.py:1556
        crit.repel_weight, crit.repel_length_scale = float(repel_weight), float(repel_length_scale)
        crit.nonempty_weight, crit.nonempty_p = float(nonempty_weight), int(nonempty_p)
        crit.nonempty_floor = max(1e-4, float(nonempty_floor_percent) / n_clusters)
        crit.temporal_cohesion_weight, crit.reg_cat_clusters_weight = float(temporal_cohesion_weight), float(reg_cat_clusters)
        ctx = SimpleNamespace(criterion=crit, apply_distill=False,
                              kl_scheduler=KLSchedule(nb, kl_annealing_mode, kl_warmup, kl_max_weight, kl_cooldown, kl_end_weight))
        ctx.__dict__.update(teacher_ctx)
        model.adam_m.zero_(); model.adam_v.zero_(); model.adam_steps = [0, 0, 0, 0]        # the optimizer is rebuilt
        opt = {"lr": learning_rate, "gmm_lr": gmm_learning_rate}
        for ep in range(int(epochs)):
            if ep == 0 and freeze_gmm_epochs > 0:
                opt["active"] = (True, True, False)
            if ep == freeze_gmm_epochs:                                       # training.py:1750-1755
                opt.update(lr=5e-4, gmm_lr=2e-4, active=(True, True, True))
            logs, klw, _ = train_one_epoch_indexed(model, "vade", loader, opt, step_vade, ep, epochs, 0.75, ctx, world)
            log_summary["train_logs"].append({"phase": "main", "epoch": ep, "kl_weight": klw, **logs})
        step_fn, vctx = step_vade, SimpleNamespace(criterion=crit, apply_distill=False)
    elif name == "vqvae":
        model = VQVAEB200((T, N, F), (T, E, 1), adj, latent_dim, n_clusters, kmeans_loss=kmeans_loss, max_batch=int(batch_size),
                          training=True, seed=random_seed)
        if world > 1:
            dist.broadcast(model.state, src=0)
        ctx = SimpleNamespace(apply_distill=False)
        opt = {"lr": learning_rate, "weight_decay": 1e-4}
        for ep in range(int(epochs)):
            logs, _, _ = train_one_epoch_indexed(model, "vqvae", loader, opt, step_vqvae_distill, ep, epochs, 0.75, ctx, world)
            log_summary["train_logs"].append({"epoch": ep, **logs})
        step_fn, vctx = step_vqvae_distill, ctx
    elif name == "contrastive":
        model = ContrastiveB200((T, N, F), (T, E, 1), adj, latent_dim, temperature=temperature,
                                similarity_function=contrastive_similarity_function, loss_function=contrastive_loss_function,
                                beta=beta, tau=tau, max_batch=int(batch_size), training=True, seed=random_seed)
        if world > 1:
            dist.broadcast(model.state, src=0)
        aug = ContrastiveAugCfg(aug_min_shift, aug_max_shift, aug_p_shift, aug_max_rot, aug_n_rot, aug_p_rot, aug_max_interp,
                                aug_min_interp, aug_p_interp, aug_noise_sigma, aug_p_noise)
        gen = torch.Generator(device=dev).manual_seed(random_seed + 7919 * rank)
        hgen = torch.Generator().manual_seed(random_seed + 7919 * rank)
        ctx = SimpleNamespace(apply_distill=False, contrastive_cfg=aug, generator=gen, host_generator=hgen)
        opt = {"lr": learning_rate, "weight_decay": 1e-4}
        for ep in range(int(epochs)):
            logs, _, _ = train_one_epoch_indexed(model, "contrastive", loader, opt, step_contrastive_distill, ep, epochs, 0.75, ctx, world)
            log_summary["train_logs"].append({"epoch": ep, **logs})
        step_fn, vctx = step_contrastive_distill, ctx
    else:
        raise ValueError(f"unknown model_name {model_name!r}")
    # validation logs of the final model (the reference validates every epoch on every rank, training.py:190-229);
    # the step functions fill model.grad as a side effect, parameters are untouched
    log_summary["val_logs"].append(validate(model, step_fn, vctx))
    if output_path and save_weights and rank == 0:
        save_model_info(os.path.join(output_path, f"{name}_final.pth"), model=model, rebuild_spec=rebuild_spec, log_summary=log_summary)
    return model, model, (teacher_init_model if name == "vade" else None), log_summary
