"""Host-side training loop pieces for the B200-native VaDE path.

``VaDETrainer`` plays the role of the reference's ``step_vade`` + ``train_one_epoch_indexed``
body (reference ``deepof/clustering/training.py:104-187, 231-309``): per batch
forward -> VadeLoss -> backward -> [gradient all-reduce] -> clip_grad_value_(0.75) -> Adam,
with the KL weight schedule of ``Dynamic_weight_manager`` (``losses.py:290-351``).

Data parallel contract = the reference's DDP (``model_utils_new.py:196-226``):
one process per GPU, replicas hold identical parameters, every batch statistic is
rank-local, ONE all-reduce(sum) of the flat fp32 gradient buffer per step; the 1/world
scaling is folded into the fused clip+Adam kernel.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional

import torch

from . import _lib
from .vade import VaDEB200, VadeLossCfg


class KLSchedule:
    """``Dynamic_weight_manager`` (reference losses.py:290-351), stepped once per batch."""

    def __init__(self, n_batches_per_epoch: int, mode: str = "tf_sigmoid", warmup_epochs: int = 5,
                 max_weight: float = 1.0, cooldown_epochs: int = 5, end_weight: float = 0.2,
                 at_max_epochs: int = 0):
        self.mode = mode
        self.warm = max(1, warmup_epochs * n_batches_per_epoch)
        self.atmax = max(0, at_max_epochs * n_batches_per_epoch)
        self.cool = max(0, cooldown_epochs * n_batches_per_epoch)
        self.max_weight, self.end_weight = float(max_weight), float(end_weight)
        self.current_iteration = 0

    def _shape(self, p: float) -> float:
        p = float(max(0.0, min(1.0, p)))
        if self.mode == "sigmoid":
            return 1.0 / (1.0 + math.exp(-12.0 * (p - 0.5)))
        if self.mode == "tf_sigmoid":
            return 1.0 / (1.0 + math.exp(-((2.0 * p - 1.0) / max(1e-2, p - p * p))))
        return p

    def get_weight(self) -> float:
        it = self.current_iteration
        total = self.warm + self.atmax + self.cool
        if it >= total:
            return self.end_weight
        if self.atmax > 0 and self.warm <= it < self.warm + self.atmax:
            return self.max_weight
        if it <= self.warm:
            return self.max_weight * self._shape(it / self.warm)
        if self.cool <= 0:
            return self.max_weight
        pc = (it - (self.warm + self.atmax)) / self.cool
        return (1.0 - pc) * self.max_weight + pc * self.end_weight

    def step(self):
        self.current_iteration += 1


class _HostPipeline:
    """``train_steps``: the epoch loop a host caller runs (``train_one_epoch_indexed`` fed by a pinned-memory DataLoader,
    reference training.py:104-187) with the copies off the critical path: the host->device copy of batch i+1 runs on a
    side stream while step i computes (two staging buffers, events in both directions), the loss of every step is read
    back with a non-blocking device->host copy and the host synchronises once at the end.  Needs ``_xs``, ``_as`` (device
    staging) and ``train_step_device``."""

    def _pipeline_init(self):
        if getattr(self, "_copy_stream", None) is None:
            dev = self._xs.device
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [(self._xs, self._as), (torch.empty_like(self._xs), torch.empty_like(self._as) if self._as is not None else None)]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]

    def train_steps(self, batches) -> list:
        """``batches``: sequence of ``(x_host, a_host)`` (pinned) or ``(x_host, a_host, idx)``.  Returns the total_loss of
        every step as python floats."""
        self._pipeline_init()
        batches = list(batches)
        n = len(batches)
        if n == 0:
            return []
        losses = torch.zeros(n, pin_memory=True)
        main, side = torch.cuda.current_stream(), self._copy_stream
        side.wait_stream(main)

        def upload(i):
            k = i & 1
            xh, ah = batches[i][0], batches[i][1]
            B = xh.shape[0]
            with torch.cuda.stream(side):
                if i >= 2:
                    side.wait_event(self._free[k])               # step i-2 has consumed this staging buffer
                xs, as_ = self._stage[k]
                xs[:B].copy_(xh, non_blocking=True)
                if ah is not None and as_ is not None:
                    as_[:B].copy_(ah, non_blocking=True)
                self._ready[k].record(side)

        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            k = i & 1
            B = batches[i][0].shape[0]
            main.wait_event(self._ready[k])
            xs, as_ = self._stage[k]
            idx = batches[i][2] if len(batches[i]) > 2 else None
            logs = self.train_step_device(xs[:B], as_[:B] if as_ is not None else None, idx)
            self._free[k].record(main)
            losses[i:i + 1].copy_(logs[:1], non_blocking=True)
        main.synchronize()
        return losses.tolist()


class PeerGradExchange:
    """The per-step gradient all-reduce of the data-parallel run over NVLink PEER MEMORY instead of an NCCL
    call: the model's flat gradient lives in a symmetric buffer mapped into every rank (torch symmetric memory = plumbing),
    and ``all_reduce_`` enqueues  barrier -> one-shot reduce (every rank sums all peers' gradients in rank order, so the
    result is bit-identical everywhere) -> barrier -> copy back  on the current stream (csrc/peer.cuh).  For the 86 KB - 4 MB
    payloads of this path the NCCL all-reduce costs 0.26 - 0.32 ms of pure latency per step; this is four tiny kernels.
    ``DOF_ALLREDUCE=nccl`` (or a failed symmetric-memory rendezvous, e.g. the gloo CPU tests) selects
    ``torch.distributed.all_reduce`` instead; ``self.kind`` says which one runs."""

    def __init__(self, model, world_size: int, rank: int):
        import torch.distributed as dist
        self.model, self.world, self.rank, self.kind, self.epoch = model, int(world_size), int(rank), "nccl", 0
        self.why = ""
        if self.world <= 1 or os.environ.get("DOF_ALLREDUCE", "peer") == "nccl" or not model.state.is_cuda:
            self.why = "disabled"
            return
        ok = 0
        try:
            import torch.distributed._symmetric_memory as symm_mem
            n = model.state.numel()
            self.n, self.n4 = n, (n + 3) // 4 * 4
            pad_floats = 64                                   # >= world int32 signal slots, keeps 16-byte alignment
            buf = symm_mem.empty(self.n4 + pad_floats, dtype=torch.float32, device=model.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            assert len(ptrs) == self.world and self.world <= 16
            self._buf, self._hdl = buf, hdl
            self._peers = (C.c_void_p * self.world)(*ptrs)
            self._pads = (C.c_void_p * self.world)(*[p + self.n4 * 4 for p in ptrs])
            self._sum = torch.empty(self.n4, device=model.device)
            ok = 1
        except Exception as ex:                               # no symmetric memory on this box / backend
            self.why = f"{type(ex).__name__}: {ex}"[:200]
        # every rank must take the same path
        flag = torch.tensor([ok], device=model.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            model.grad = self._buf[:self.n]                   # the backward kernels write into the exchange buffer
            torch.cuda.synchronize()
            dist.barrier()
            self.kind = "peer_memory"

    def all_reduce_(self):
        """model.grad <- sum over ranks (in place, on the current stream)."""
        m = self.model
        if self.kind != "peer_memory":
            import torch.distributed as dist
            dist.all_reduce(m.grad, op=dist.ReduceOp.SUM)
            return
        L = _lib.lib()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.epoch += 1
        _lib.check(L.dof_peer_barrier(self._pads, self.world, self.rank, self.epoch, st))
        _lib.check(L.dof_peer_reduce(self._peers, self.world, C.c_void_p(self._sum.data_ptr()), self.n4, st))
        self.epoch += 1
        _lib.check(L.dof_peer_barrier(self._pads, self.world, self.rank, self.epoch, st))
        m.grad.copy_(self._sum[:self.n])


class VaDETrainer(_HostPipeline):
    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int, n_components: int,
                 max_batch: int = 4096, seed: Optional[int] = None, world_size: int = 1, rank: int = 0,
                 kmeans_loss: float = 1.0, device: Optional[int] = None, encoder_type: str = "recurrent"):
        self.model = VaDEB200(input_shape, edge_feature_shape, adjacency_matrix, latent_dim, n_components,
                              encoder_type=encoder_type, kmeans_loss=kmeans_loss, device=device, max_batch=max_batch,
                              training=True, seed=seed)
        self.world_size, self.rank = int(world_size), int(rank)
        self.loss_cfg = VadeLossCfg.pretrain_defaults(n_components)
        self.loss_cfg.model_kmeans_weight = float(kmeans_loss)
        self.lr_base, self.lr_gmm = 1e-3, 0.0
        self.kl_schedule: Optional[KLSchedule] = None
        self.active = (True, True, True)   # encoder+latent / decoder / GMM groups
        self.tau_star = None
        self.class_weight = None
        self.teacher_marginal = None
        dev = self.model.device
        T, N, F = self.model.input_shape
        _, E, Fe = self.model.edge_feature_shape
        self._xs = torch.empty(max_batch, T, N, F, device=dev)
        self._as = torch.empty(max_batch, T, E, Fe, device=dev)
        self._loss_host = torch.zeros(16, pin_memory=True)
        self.exchange = None
        if self.world_size > 1:
            import torch.distributed as dist
            dist.broadcast(self.model.state, src=0)   # DDP constructor broadcast (reference training.py:1567)
            self.exchange = PeerGradExchange(self.model, self.world_size, self.rank)

    # ---- phase control (reference training.py:1579-1653, 1746-1767)
    def set_phase(self, phase: str, kl_weight: Optional[float] = None, lr_base: Optional[float] = None,
                  lr_gmm: Optional[float] = None, kl_schedule: Optional[KLSchedule] = None):
        K = self.model.n_components
        mk = self.loss_cfg.model_kmeans_weight
        self.loss_cfg = VadeLossCfg.pretrain_defaults(K) if phase == "pretrain" else VadeLossCfg.main_defaults(K)
        self.loss_cfg.model_kmeans_weight = mk
        self.model.set_pretrain_mode(phase == "pretrain")
        if kl_weight is not None:
            self.loss_cfg.kl_weight = float(kl_weight)
        self.kl_schedule = kl_schedule
        if lr_base is not None:
            self.lr_base = float(lr_base)
        if lr_gmm is not None:
            self.lr_gmm = float(lr_gmm)
        # rebuilding the optimizer resets Adam state in the reference (training.py:1648-1653)
        self.model.adam_m.zero_()
        self.model.adam_v.zero_()
        self.model.adam_steps = [0, 0, 0, 0]

    def set_teacher(self, tau_star: torch.Tensor, lambda_distill: float, class_weight=None, teacher_marginal=None):
        self.tau_star = torch.as_tensor(tau_star, dtype=torch.float32).to(self.model.device)
        self.loss_cfg.lambda_distill = float(lambda_distill)
        self.class_weight, self.teacher_marginal = class_weight, teacher_marginal

    # ---- one step
    def train_step_device(self, x: torch.Tensor, a: torch.Tensor, idx: Optional[torch.Tensor] = None, eps=None,
                          mc_eps=None) -> torch.Tensor:
        """x, a already on the device.  Returns the device log vector (no host sync)."""
        m = self.model
        if self.kl_schedule is not None:
            self.loss_cfg.kl_weight = self.kl_schedule.get_weight()
        tau = None
        if self.tau_star is not None and self.loss_cfg.lambda_distill > 0.0 and idx is not None:
            tau = self.tau_star[idx.to(m.device)]
        logs = m.loss_grad(x, a, self.loss_cfg, eps=eps, mc_eps=mc_eps, tau_batch=tau, class_weight=self.class_weight,
                           teacher_marginal=self.teacher_marginal)
        scale = 1.0
        if self.world_size > 1:
            self.exchange.all_reduce_()
            scale = 1.0 / self.world_size
        m.adam_step(self.lr_base, self.lr_gmm, grad_scale=scale, active=self.active)
        if self.kl_schedule is not None:
            self.kl_schedule.step()
        return logs

    def train_step(self, x_host: torch.Tensor, a_host: torch.Tensor, idx: Optional[torch.Tensor] = None) -> float:
        """Public end-to-end step from HOST tensors: H2D copy, step, D2H read of total_loss."""
        B = x_host.shape[0]
        xs, as_ = self._xs[:B], self._as[:B]
        xs.copy_(x_host, non_blocking=True)
        as_.copy_(a_host, non_blocking=True)
        logs = self.train_step_device(xs, as_, idx)
        self._loss_host.copy_(logs, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._loss_host[0])

    def logs(self) -> Dict[str, float]:
        return self.model.logs_dict()


class _GenericTrainer:
    """Shared step plumbing of ``fit_VQVAE`` / ``fit_contrastive`` (reference training.py:1087-1200, 1321-1460):
    loss.backward -> [all-reduce(sum) of the flat gradient] -> clip_grad_value_(0.75) -> Adam(lr, weight_decay=1e-4)
    (``build_optimizer_generic``, losses.py:805-814)."""

    def __init__(self, model, world_size: int = 1, rank: int = 0, lr: float = 1e-3, weight_decay: float = 1e-4):
        self.model, self.world_size, self.rank = model, int(world_size), int(rank)
        self.lr, self.weight_decay = float(lr), float(weight_decay)
        self._loss_host = torch.zeros(16, pin_memory=True)
        self.exchange = None
        if self.world_size > 1:
            import torch.distributed as dist
            dist.broadcast(self.model.state, src=0)
            self.exchange = PeerGradExchange(self.model, self.world_size, self.rank)

    def _finish(self, logs):
        m = self.model
        scale = 1.0
        if self.world_size > 1:
            self.exchange.all_reduce_()
            scale = 1.0 / self.world_size
        m.adam_step(self.lr, grad_scale=scale, weight_decay=self.weight_decay)
        return logs

    def _read_loss(self, logs) -> float:
        self._loss_host.copy_(logs, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._loss_host[0])

    def logs(self) -> Dict[str, float]:
        return self.model.logs_dict()


class VQVAETrainer(_GenericTrainer, _HostPipeline):
    """``step_vqvae_distill`` + the optimizer step (teacher off)."""

    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int, n_components: int,
                 max_batch: int = 4096, seed: Optional[int] = None, world_size: int = 1, rank: int = 0,
                 kmeans_loss: float = 0.0, beta: float = 1.0, lr: float = 1e-3, device: Optional[int] = None,
                 encoder_type: str = "recurrent"):
        from .models import VQVAEB200
        m = VQVAEB200(input_shape, edge_feature_shape, adjacency_matrix, latent_dim, n_components, encoder_type=encoder_type,
                      kmeans_loss=kmeans_loss, beta=beta, device=device, max_batch=max_batch, training=True, seed=seed)
        super().__init__(m, world_size, rank, lr)
        T, N, F = m.input_shape
        _, E, Fe = m.edge_feature_shape
        self._xs = torch.empty(max_batch, T, N, F, device=m.device)
        self._as = torch.empty(max_batch, T, E, Fe, device=m.device)

    def train_step_device(self, x: torch.Tensor, a: torch.Tensor, idx=None) -> torch.Tensor:
        return self._finish(self.model.loss_grad(x, a))

    def train_step(self, x_host: torch.Tensor, a_host: torch.Tensor, idx=None) -> float:
        B = x_host.shape[0]
        xs, as_ = self._xs[:B], self._as[:B]
        xs.copy_(x_host, non_blocking=True)
        as_.copy_(a_host, non_blocking=True)
        return self._read_loss(self.train_step_device(xs, as_))


class ContrastiveTrainer(_GenericTrainer, _HostPipeline):
    """``step_contrastive_distill`` + the optimizer step (teacher off): per batch draw the augmentation decisions,
    build both views on the device, encode them in one pass, NT-Xent, backward, clip + Adam."""

    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int, max_batch: int = 4096,
                 seed: Optional[int] = None, world_size: int = 1, rank: int = 0, temperature: float = 0.1,
                 aug=None, lr: float = 1e-3, edge_index=None, edge_index_local=None, device: Optional[int] = None,
                 encoder_type: str = "recurrent"):
        from .models import ContrastiveAugCfg, ContrastiveB200
        m = ContrastiveB200(input_shape, edge_feature_shape, adjacency_matrix, latent_dim, encoder_type=encoder_type,
                            temperature=temperature,
                            edge_index=edge_index, edge_index_local=edge_index_local, device=device, max_batch=max_batch,
                            training=True, seed=seed)
        super().__init__(m, world_size, rank, lr)
        self.aug = aug if aug is not None else ContrastiveAugCfg()
        self.gen = torch.Generator(device=m.device)
        self.host_gen = torch.Generator()
        s = (seed if seed is not None else 0) + 7919 * rank
        self.gen.manual_seed(s)
        self.host_gen.manual_seed(s)
        Tf, N = m.full_time_steps, m.input_shape[1]
        self._xf = torch.empty(max_batch, Tf, N, 3, device=m.device)
        self._xs, self._as = self._xf, None          # staging of the pipelined host loop (_HostPipeline): full windows only

    def train_step_device(self, x_full: torch.Tensor, a_full=None, idx=None, aug_params=None) -> torch.Tensor:
        """a_full is accepted for signature parity and ignored: the reference recomputes the edge lengths from
        the coordinates (training.py:497)."""
        m = self.model
        prm = aug_params if aug_params is not None else m.draw_augmentation(x_full.shape[0], self.aug, self.gen, self.host_gen)
        return self._finish(m.loss_grad(x_full, prm))

    def train_step(self, x_host: torch.Tensor, a_host=None, idx=None) -> float:
        B = x_host.shape[0]
        xf = self._xf[:B]
        xf.copy_(x_host, non_blocking=True)
        return self._read_loss(self.train_step_device(xf))
