"""Host-side mirrors of the reference ``VQVAEPT`` and ``ContrastivePT`` (recurrent encoder, GNN path) on top of
the deepof_b200 C-ABI.  They share the state / optimizer plumbing of ``VaDEB200``.

Reference (``deepof/clustering``): ``models_new.py:1510-1640`` (VQVAEPT), ``:1978-2069`` (ContrastivePT),
``training.py:312-389`` (step_vqvae_distill), ``:482-589`` (step_contrastive_distill), ``:2064-2402`` (rotation
table + augmentations), ``model_utils_new.py:173-189`` (ContrastiveCfg).  torch is device memory + RNG only.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DofViewsCfg, check, ptr
from .vade import TFM_BUFFERS, VaDEB200, _stream

VQ_LOG_KEYS = ("total_loss", "enc_rec_loss", "reconstruct_loss", "vq_loss", "kmeans_loss",
               "number_of_populated_clusters", "distill_loss")       # step_vqvae_distill, training.py:380-388
CON_LOG_KEYS = ("total_loss", "pos_similarity", "neg_similarity", "distill_loss", "seperability")   # :582-588


class DistillHeadB200:
    """``DiscriminativeHead`` (teacher_model.py:795-808): one ``Linear(latent_dim, n_components)`` on the encoder output,
    trained together with the VQ-VAE / contrastive model by ``build_optimizer_generic`` (losses.py:805-814) and NOT
    clipped (training.py:165 clips ``model.parameters()`` only).  Its forward + soft cross-entropy + backward run inside
    ``dof_*_loss_grad_distill``; this object owns the parameters, their gradient and the Adam moments."""

    def __init__(self, latent_dim: int, n_components: int, device="cuda", seed: Optional[int] = None):
        self.latent_dim, self.n_components = int(latent_dim), int(n_components)
        self.device = torch.device(device)
        n = self.n_components * self.latent_dim + self.n_components
        g = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        bound = 1.0 / math.sqrt(self.latent_dim)                       # nn.Linear's default init
        self.state = ((torch.rand(n, generator=g) * 2.0 - 1.0) * bound).to(self.device)
        self.grad = torch.zeros(n, device=self.device)
        self.adam_m, self.adam_v = torch.zeros_like(self.state), torch.zeros_like(self.state)
        self.step = 0
        self.L = _lib.lib()

    @property
    def weight(self) -> torch.Tensor:
        return self.state[: self.n_components * self.latent_dim].view(self.n_components, self.latent_dim)

    @property
    def bias(self) -> torch.Tensor:
        return self.state[self.n_components * self.latent_dim:]

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {"fc.weight": self.weight.clone(), "fc.bias": self.bias.clone()}

    def load_state_dict(self, sd) -> None:
        self.weight.copy_(torch.as_tensor(sd["fc.weight"]).to(self.device, torch.float32))
        self.bias.copy_(torch.as_tensor(sd["fc.bias"]).to(self.device, torch.float32))

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        kd = self.n_components * self.latent_dim
        return {"fc.weight": self.grad[:kd].view(self.n_components, self.latent_dim), "fc.bias": self.grad[kd:]}

    def __call__(self, z: torch.Tensor) -> torch.Tensor:
        """Logits for inference-time cluster read-outs (``get_q_vqvae``): plain torch on the device, off the step path."""
        return z.to(self.device, torch.float32) @ self.weight.t() + self.bias

    def adam_step(self, lr: float, weight_decay: float = 1e-4, grad_scale: float = 1.0, beta1: float = 0.9,
                  beta2: float = 0.999, eps: float = 1e-8) -> None:
        self.step += 1
        check(self.L.dof_adam_flat(ptr(self.state), ptr(self.grad), ptr(self.adam_m), ptr(self.adam_v), self.state.numel(),
                                   float(lr), beta1, beta2, eps, float(weight_decay), self.step, float(grad_scale), _stream()))


@dataclass
class Distillation:
    """What ``step_*_distill`` reads from ``ctx`` when the teacher is on (training.py:341-370)."""
    head: DistillHeadB200
    tau_batch: torch.Tensor              # [B, K] = ctx.tau_star[idx]
    lambda_distill: float
    sharpen_T: float = 0.5
    conf_weight: bool = False
    conf_thresh: float = 0.6

    def cfg(self, B: int) -> "_lib.DofDistillCfg":
        tau = self.tau_batch.to(self.head.device, torch.float32).contiguous()
        if tuple(tau.shape) != (B, self.head.n_components):
            raise ValueError(f"tau_batch must be [{B}, {self.head.n_components}], got {tuple(tau.shape)}")
        self._keep = tau
        return _lib.DofDistillCfg(self.head.state.data_ptr(), self.head.grad.data_ptr(), tau.data_ptr(), self.head.n_components,
                                  float(self.lambda_distill), float(self.sharpen_T or 0.0), int(bool(self.conf_weight)),
                                  float(self.conf_thresh))


class VQVAEB200(VaDEB200):
    """Stand-in for ``VQVAEPT(encoder_type="recurrent", use_gnn=True)``."""
    _MODEL = _lib.MODEL_VQVAE
    _BUFFERS = ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence") + TFM_BUFFERS
    _DEC_PASSES = 2          # decoder(quantized) and decoder(encoder output), models_new.py:1603-1612

    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int, n_components: int,
                 encoder_type: str = "recurrent", use_gnn: bool = True, kmeans_loss: float = 0.0,
                 interaction_regularization: float = 0.0, beta: float = 1.0, **kw):
        super().__init__(input_shape, edge_feature_shape, adjacency_matrix, latent_dim, n_components,
                         encoder_type=encoder_type, use_gnn=use_gnn, kmeans_loss=kmeans_loss, **kw)
        self.beta = float(beta)

    def forward_eval(self, x, a, want_loc: bool = True):
        """(enc [B,D], quant [B,D], soft [B,K], idx [B], loc_q, loc_e [B,T,N*F] or None)."""
        x, a = self._prep(x, a)
        T, N, F = self.input_shape
        outs = ([], [], [], [], [], [])
        for s in range(0, x.shape[0], self.max_batch):
            xb, ab = x[s:s + self.max_batch], a[s:s + self.max_batch]
            B = xb.shape[0]
            enc = torch.empty(B, self.latent_dim, device=self.device)
            quant = torch.empty(B, self.latent_dim, device=self.device)
            soft = torch.empty(B, self.n_components, device=self.device)
            idx = torch.empty(B, dtype=torch.int32, device=self.device)
            lq = torch.empty(B, T, N * F, device=self.device) if want_loc else None
            le = torch.empty(B, T, N * F, device=self.device) if want_loc else None
            check(self.L.dof_vqvae_forward_eval(self.handle, ptr(self.state), ptr(xb), ptr(ab), B, ptr(enc), ptr(quant),
                                                ptr(soft), ptr(idx), ptr(lq), ptr(le), _stream()))
            for o, t in zip(outs, (enc, quant, soft, idx, lq, le)):
                o.append(t)
        cat = lambda l: None if l[0] is None else (l[0] if len(l) == 1 else torch.cat(l))
        return tuple(cat(o) for o in outs)

    def __call__(self, x, a, return_losses: bool = False, return_all_outputs: bool = True):
        """Eval-mode ``VQVAEPT.forward(return_all_outputs=True)``: (loc_q, loc_e, quant, soft, enc, None)."""
        enc, quant, soft, idx, lq, le = self.forward_eval(x, a)
        return (lq, le, quant, soft, enc, None) if return_all_outputs else (lq, le)

    def encode(self, x, a):
        return self.forward_eval(x, a, want_loc=False)[0]

    def embed(self, x, a):
        """(embedding = encoder output, soft counts): what embedding_per_video reads for VQ-VAE models
        (model_utils_new.py:598-609)."""
        enc, quant, soft, idx, _, _ = self.forward_eval(x, a, want_loc=False)
        return enc, soft

    def loss_grad(self, x, a, distill: Optional["Distillation"] = None, dropout_masks=None):
        if not self.training_capable:
            raise _lib.DofError("model was created with training=False")
        x, a = self._prep(x, a)
        self._set_dropout(x.shape[0], dropout_masks)
        dc = distill.cfg(x.shape[0]) if distill is not None and distill.lambda_distill > 0.0 else None
        check(self.L.dof_vqvae_loss_grad_distill(self.handle, ptr(self.state), ptr(self.grad), ptr(x), ptr(a), x.shape[0],
                                                 self.beta, self.kmeans_weight, C.byref(dc) if dc is not None else None,
                                                 ptr(self.logs), _stream()))
        return self.logs

    def loss_eval(self, x, a):
        """The validation step of ``fit_VQVAE`` (training.py:1165-1170): eval-mode forward + loss terms, teacher off."""
        if not self.training_capable:
            raise _lib.DofError("model was created with training=False")
        x, a = self._prep(x, a)
        check(self.L.dof_vqvae_loss_eval(self.handle, ptr(self.state), ptr(x), ptr(a), x.shape[0], self.beta, self.kmeans_weight,
                                         ptr(self.logs), _stream()))
        return self.logs

    def adam_step(self, lr: float, clip: float = 0.75, grad_scale: float = 1.0, weight_decay: float = 1e-4, **kw):
        """clip_grad_value_(0.75) + build_optimizer_generic's Adam(lr, weight_decay=1e-4) (losses.py:805-814)."""
        super().adam_step(lr, lr, clip=clip, grad_scale=grad_scale, weight_decay=weight_decay, **kw)

    def logs_dict(self) -> Dict[str, float]:
        v = self.logs.detach().cpu().tolist()
        return {k: v[i] for i, k in enumerate(VQ_LOG_KEYS)}


@dataclass
class ContrastiveAugCfg:
    """``ContrastiveCfg.aug_*`` (model_utils_new.py:173-189)."""
    min_shift: int = 1
    max_shift: int = 6
    p_shift: float = 0.8
    max_rot: float = 30.0
    n_rot: int = 4
    p_rot: float = 0.0
    max_interp: int = 8
    min_interp: int = 3
    p_interp: float = 0.3
    noise_sigma: float = 0.03
    p_noise: float = 0.0


@dataclass
class AugParams:
    """Random decisions of one augmented view (device tensors), the input of ``dof_contrastive_views``."""
    start: torch.Tensor                                   # [B] int32
    rot_pivot: List[int] = field(default_factory=list)
    rot_nodes: List[List[int]] = field(default_factory=list)
    rot_theta: Optional[torch.Tensor] = None              # [R,B] float32 radians
    interp_t0: Optional[torch.Tensor] = None              # [B] int32
    interp_len: Optional[torch.Tensor] = None             # [B] int32 (0 = off)
    noise: Optional[torch.Tensor] = None                  # [B,N,3] float32


class RotationTable:
    """``build_rotation_precomp`` (training.py:2064-2125): triplets (a, b, c) around every node b with >= 2
    neighbours and, per triplet, the nodes reachable from a / from c without passing through b."""

    def __init__(self, edge_index, n_nodes: int):
        adj = [[] for _ in range(n_nodes)]
        for u, v in np.asarray(edge_index).reshape(-1, 2).tolist():
            adj[u].append(v)
            adj[v].append(u)
        self.triplets: List[Tuple[int, int, int]] = []
        for b in range(n_nodes):
            nb = adj[b]
            for i in range(len(nb)):
                for j in range(i + 1, len(nb)):
                    self.triplets.append((nb[i], b, nb[j]))

        def branch(center, side):
            seen, stack = {side}, [side]
            while stack:
                u = stack.pop()
                for v in adj[u]:
                    if v != center and v not in seen:
                        seen.add(v)
                        stack.append(v)
            return list(seen)

        self.branches_a = [branch(b, a) for a, b, c in self.triplets]
        self.branches_c = [branch(b, c) for a, b, c in self.triplets]


class ContrastiveB200(VaDEB200):
    """Stand-in for ``ContrastivePT(encoder_type="recurrent", use_gnn=True, similarity_function="cosine",
    loss_function="nce")``.  ``input_shape`` holds the FULL window length; the encoder sees ``T // 2``."""
    _MODEL = _lib.MODEL_CONTRASTIVE
    _BUFFERS = ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence") + TFM_BUFFERS
    _DEC_PASSES = 0

    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int = 8,
                 encoder_type: str = "recurrent", use_gnn: bool = True, temperature: float = 0.1,
                 similarity_function: str = "cosine", loss_function: str = "nce", beta: float = 0.1, tau: float = 0.1,
                 edge_index=None, edge_index_local=None, max_batch: int = 4096, **kw):
        if similarity_function not in ("cosine", "dot", "euclidean", "edit") or loss_function not in ("nce", "dcl", "hard_dcl", "fc"):
            raise NotImplementedError("deepof_b200 implements the cosine / dot / euclidean / edit similarities with the nce / "
                                      f"dcl / hard_dcl / fc losses (got {similarity_function!r}, {loss_function!r})")
        self.similarity_function = similarity_function
        self.loss_function, self.beta, self.tau = loss_function, float(beta), float(tau)
        Tf, N, F = (int(v) for v in input_shape)
        _, E, Fe = (int(v) for v in edge_feature_shape)
        self.full_time_steps = Tf
        # the two views go through the encoder as ONE batch of 2B windows
        super().__init__((Tf // 2, N, F), (Tf // 2, E, Fe), adjacency_matrix, latent_dim, 1, encoder_type=encoder_type,
                         use_gnn=use_gnn, max_batch=2 * int(max_batch), **kw)
        self.max_windows = int(max_batch)
        self.temperature = float(temperature)
        if edge_index is None:       # sorted (i < j) edges of the adjacency: the reference's edge_columns order
            r, c = np.nonzero(np.triu(self.adjacency_matrix))
            edge_index = np.stack([r, c], 1)
        self.edge_index = np.ascontiguousarray(np.asarray(edge_index, dtype=np.int32).reshape(-1, 2))
        assert self.edge_index.shape[0] == E
        self.rotations = RotationTable(self.edge_index if edge_index_local is None else edge_index_local, N)
        self._x2 = torch.empty(2 * self.max_windows, Tf // 2, N, F, device=self.device)
        self._a2 = torch.empty(2 * self.max_windows, Tf // 2, E, Fe, device=self.device)
        self.z_all = torch.empty(2 * self.max_windows, self.latent_dim, device=self.device)

    def __call__(self, x, a):
        """``ContrastivePT.forward``: half windows -> embeddings [B,D]."""
        x, a = self._prep(x, a)
        outs = []
        for s in range(0, x.shape[0], self.max_batch):
            xb, ab = x[s:s + self.max_batch], a[s:s + self.max_batch]
            enc = torch.empty(xb.shape[0], self.latent_dim, device=self.device)
            check(self.L.dof_encode(self.handle, ptr(self.state), ptr(xb), ptr(ab), xb.shape[0], ptr(enc), _stream()))
            outs.append(enc)
        return outs[0] if len(outs) == 1 else torch.cat(outs)

    def embed(self, x, a):
        """Contrastive models have no soft counts; embedding_per_video uses model(x, a) (model_utils_new.py:590-597)."""
        return self(x, a), None

    # ---- augmentation decisions (training.py:2128-2402): same distributions as the reference, drawn with torch
    def draw_augmentation(self, B: int, cfg: ContrastiveAugCfg, generator: Optional[torch.Generator] = None,
                          host_generator: Optional[torch.Generator] = None) -> AugParams:
        dev, g, hg = self.device, generator, host_generator
        Tf, N = self.full_time_steps, self.input_shape[1]
        half = Tf // 2
        base = (Tf - half) // 2
        rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
        rint = lambda lo, hi, s: torch.randint(lo, hi, s, device=dev, generator=g)
        apply = rnd(B) < cfg.p_shift
        shift = rint(cfg.min_shift, cfg.max_shift + 1, (B,)) * (rint(0, 2, (B,)) * 2 - 1) * apply.long()
        out = AugParams(start=(base + shift).clamp(0, Tf - half).int())
        M = len(self.rotations.triplets)
        if cfg.n_rot > 0 and cfg.max_rot > 0.0 and cfg.p_rot > 0.0 and M > 0:
            app = (rnd(B) < cfg.p_rot).float()
            perm = torch.randperm(M, generator=hg).tolist()      # discrete graph choices stay on the host
            chosen, count = [], [0] * N
            for k in perm:
                b0 = self.rotations.triplets[k][1]
                if count[b0] >= 2:
                    continue
                count[b0] += 1
                chosen.append(k)
                if len(chosen) >= cfg.n_rot:
                    break
            thetas = []
            for k in chosen:
                side_a = bool(torch.rand((), generator=hg) < 0.5)
                nodes = self.rotations.branches_a[k] if side_a else self.rotations.branches_c[k]
                if not nodes:
                    continue
                thetas.append((rnd(B) * 2.0 - 1.0) * (float(cfg.max_rot) * math.pi / 180.0) * app)
                out.rot_pivot.append(self.rotations.triplets[k][1])
                out.rot_nodes.append(list(nodes))
            if thetas:
                out.rot_theta = torch.stack(thetas).contiguous()
        if cfg.max_interp > 0 and cfg.p_interp > 0.0 and half >= 3:
            app = rnd(B) < cfg.p_interp
            L = rint(cfg.min_interp, cfg.max_interp + 1, (B,))
            t0 = torch.minimum(rint(1, half - 1, (B,)), (half - L - 1).clamp_min(1))
            out.interp_t0, out.interp_len = t0.int(), (L * app.long()).int()
        if cfg.noise_sigma > 0.0 and cfg.p_noise > 0.0:
            app = (rnd(B) < cfg.p_noise).float().view(B, 1)
            axis = rint(0, 2, (B, N))
            off = cfg.noise_sigma * torch.randn(B, N, device=dev, generator=g) * app
            ds = cfg.noise_sigma * torch.randn(B, N, device=dev, generator=g) * app
            out.noise = torch.stack([off * (axis == 0).float(), off * (axis == 1).float(), ds], dim=-1).contiguous()
        return out

    def views(self, x_full, prm: AugParams):
        """(x2 [2B,T/2,N,3], a2 [2B,T/2,E,1]): rows 0..B-1 the main view, B..2B-1 the augmented view."""
        x_full = torch.as_tensor(x_full, dtype=torch.float32).to(self.device).contiguous()
        B = x_full.shape[0]
        Tf, N = self.full_time_steps, self.input_shape[1]
        assert x_full.shape[1:] == (Tf, N, 3) and B <= self.max_windows
        i32 = lambda t: None if t is None else torch.as_tensor(t).to(self.device, torch.int32).contiguous()
        f32 = lambda t: None if t is None else torch.as_tensor(t).to(self.device, torch.float32).contiguous()
        start, th, t0, ln, nz = i32(prm.start), f32(prm.rot_theta), i32(prm.interp_t0), i32(prm.interp_len), f32(prm.noise)
        v = DofViewsCfg()
        v.T_full, v.N, v.E = Tf, N, self.edge_index.shape[0]
        v.edges = self.edge_index.ctypes.data_as(C.POINTER(C.c_int))
        v.start = start.data_ptr()
        v.n_rot = 0 if th is None else th.shape[0]
        for k in range(v.n_rot):
            v.rot_pivot[k] = int(prm.rot_pivot[k])
            v.rot_mask[k] = sum(1 << int(n) for n in prm.rot_nodes[k])
        v.rot_theta = None if th is None else th.data_ptr()
        v.interp_t0 = None if t0 is None else t0.data_ptr()
        v.interp_len = None if ln is None else ln.data_ptr()
        v.noise = None if nz is None else nz.data_ptr()
        x2, a2 = self._x2[:2 * B], self._a2[:2 * B]
        check(self.L.dof_contrastive_views(C.byref(v), ptr(x_full), B, ptr(x2), ptr(a2), _stream()))
        self._keep = (start, th, t0, ln, nz)      # keep the device arrays alive until the stream has consumed them
        return x2, a2

    def loss_grad(self, x_full, prm: AugParams, distill: Optional["Distillation"] = None, dropout_masks=None):
        """views + encoder on both + NT-Xent (+ distillation head on the main view) + backward into ``self.grad``;
        returns the device log vector."""
        if not self.training_capable:
            raise _lib.DofError("model was created with training=False")
        x2, a2 = self.views(x_full, prm)
        B = x2.shape[0] // 2
        self._set_dropout(0, dropout_masks, encoder_windows=2 * B)
        kind = {"nce": 0, "dcl": 1, "hard_dcl": 2, "fc": 3}[self.loss_function]
        sim = 0 if self.similarity_function in ("cosine", "dot") else 1
        dc = distill.cfg(B) if distill is not None and distill.lambda_distill > 0.0 else None
        check(self.L.dof_contrastive_loss_grad_distill(self.handle, ptr(self.state), ptr(self.grad), ptr(x2), ptr(a2), B, kind, sim,
                                                       self.temperature, self.tau, self.beta,
                                                       C.byref(dc) if dc is not None else None, ptr(self.logs),
                                                       ptr(self.z_all[:2 * B]), _stream()))
        return self.logs

    def loss_eval(self, x_full, prm: AugParams):
        """The validation step of ``fit_contrastive``: views, eval-mode encoder on both, the loss; teacher off."""
        if not self.training_capable:
            raise _lib.DofError("model was created with training=False")
        x2, a2 = self.views(x_full, prm)
        B = x2.shape[0] // 2
        kind = {"nce": 0, "dcl": 1, "hard_dcl": 2, "fc": 3}[self.loss_function]
        sim = 0 if self.similarity_function in ("cosine", "dot") else 1
        check(self.L.dof_contrastive_loss_eval(self.handle, ptr(self.state), ptr(x2), ptr(a2), B, kind, sim, self.temperature, self.tau,
                                               self.beta, ptr(self.logs), ptr(self.z_all[:2 * B]), _stream()))
        return self.logs

    def main_view(self, x_full):
        """The middle half window and its recomputed edge lengths (training.py:518-525; ``get_q_contrastive``,
        logging.py:83-115): what the distillation head sees."""
        B = x_full.shape[0]
        prm = AugParams(start=torch.full((B,), (self.full_time_steps // 2) // 2, dtype=torch.int32, device=self.device))
        x2, a2 = self.views(x_full, prm)
        return x2[:B], a2[:B]

    def adam_step(self, lr: float, clip: float = 0.75, grad_scale: float = 1.0, weight_decay: float = 1e-4, **kw):
        super().adam_step(lr, lr, clip=clip, grad_scale=grad_scale, weight_decay=weight_decay, **kw)

    def logs_dict(self) -> Dict[str, float]:
        v = self.logs.detach().cpu().tolist()
        return {k: v[i] for i, k in enumerate(CON_LOG_KEYS)}
