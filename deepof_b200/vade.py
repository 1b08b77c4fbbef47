"""Host-side mirror of the reference ``VaDEPT`` model object (recurrent encoder, GNN path)
on top of the deepof_b200 C-ABI.

Mirrors what downstream reference code touches (reference
``deepof/clustering/models_new.py:1794-1976``, ``model_utils_new.py:545-621``):
``window_size``, ``encoder(x, a)``, ``model(x, a) -> (loc, emb, q, kmeans)`` in eval mode,
``state_dict() / load_state_dict()`` with the reference key names and shapes,
``latent_space.{gmm_means, gmm_log_vars, prior}``, and
``str(model.encoder.spatial_gnn_block) == "CensNetConvPT()"``.

torch is used for device memory, streams and RNG only; all arithmetic runs in the CUDA
library.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DofAdamCfg, DofConfig, DofVadeLossCfg, LOG_KEYS, check, lib, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class VadeLossCfg:
    """One phase of the reference ``VadeLoss`` (losses.py:383-457)."""
    pretrain_mode: bool = False
    kl_weight: float = 1.0
    l1_activity_weight: float = 0.1
    kmeans_loss_weight: float = 0.0
    model_kmeans_weight: float = 1.0
    repel_weight: float = 0.0
    repel_length_scale: float = 1.0
    nonempty_weight: float = 2e-2
    nonempty_floor: float = 0.05 / 8
    nonempty_p: int = 2
    tf_cluster_weight: float = 0.0
    reg_cat_clusters_weight: float = 0.0
    temporal_cohesion_weight: float = 0.0
    reg_scatter_weight: float = 0.0
    reg_scatter_beta: float = 1.0
    gmm_logvar_clamp: Tuple[float, float] = (-8.0, 8.0)
    mc_samples: int = 32
    lambda_distill: float = 0.0
    distill_sharpen_T: float = 0.5
    distill_conf_weight: bool = False
    distill_conf_thresh: float = 0.3

    @staticmethod
    def pretrain_defaults(n_components: int, kl_weight: float = 0.0) -> "VadeLossCfg":
        # VaDECfg defaults, reference model_utils_new.py:152-157
        return VadeLossCfg(pretrain_mode=True, kl_weight=kl_weight, kmeans_loss_weight=1.0, repel_weight=0.5,
                           repel_length_scale=0.5, nonempty_weight=2e-2,
                           nonempty_floor=max(1e-4, 0.05 / n_components), nonempty_p=2)

    @staticmethod
    def main_defaults(n_components: int, kl_weight: float = 1.0) -> "VadeLossCfg":
        # VaDECfg / CommonFitCfg defaults, reference model_utils_new.py:66,135-150
        return VadeLossCfg(pretrain_mode=False, kl_weight=kl_weight, kmeans_loss_weight=0.0, repel_weight=0.0,
                           repel_length_scale=1.0, nonempty_weight=2e-2,
                           nonempty_floor=max(1e-4, 0.05 / n_components), nonempty_p=2)

    def to_c(self) -> DofVadeLossCfg:
        return DofVadeLossCfg(
            int(self.pretrain_mode), self.kl_weight, self.l1_activity_weight, self.kmeans_loss_weight,
            self.model_kmeans_weight, self.repel_weight, self.repel_length_scale, self.nonempty_weight,
            self.nonempty_floor, int(self.nonempty_p), self.tf_cluster_weight, self.reg_cat_clusters_weight,
            self.temporal_cohesion_weight, self.reg_scatter_weight, self.reg_scatter_beta,
            self.gmm_logvar_clamp[0], self.gmm_logvar_clamp[1], int(self.mc_samples), self.lambda_distill,
            float(self.distill_sharpen_T or 0.0), int(self.distill_conf_weight), self.distill_conf_thresh)


def state_layout(cfg: DofConfig):
    """[(name, offset, numel, shape, group)] in reference state_dict order."""
    L = lib()
    n = L.dof_state_num_entries(C.byref(cfg))
    if n < 0:
        check(-1)
    out = []
    name = C.create_string_buffer(128)
    off, numel, ndim, grp = C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
    shape = (C.c_int * 4)()
    for i in range(n):
        check(L.dof_state_entry(C.byref(cfg), i, name, C.byref(off), C.byref(numel), C.byref(ndim), shape,
                                C.byref(grp)))
        out.append((name.value.decode(), off.value, numel.value, tuple(shape[: ndim.value]), grp.value))
    return out


def graph_operators(adjacency: np.ndarray):
    """(laplacian, edge_laplacian, incidence) float32 numpy, reference censNetConv_pt.py:160-175."""
    L = lib()
    A = np.ascontiguousarray(np.asarray(adjacency, dtype=np.float64))
    N = A.shape[0]
    ne = C.c_int()
    check(L.dof_graph_operators(A.ctypes.data_as(C.POINTER(C.c_double)), N, 0, None, None, None, C.byref(ne)))
    E = ne.value
    lap = np.zeros((N, N), np.float32)
    elap = np.zeros((E, E), np.float32)
    inc = np.zeros((N, E), np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    check(L.dof_graph_operators(A.ctypes.data_as(C.POINTER(C.c_double)), N, E, fp(lap), fp(elap), fp(inc),
                                C.byref(ne)))
    return lap, elap, inc


class _Named:
    def __init__(self, rep):
        self._rep = rep

    def __repr__(self):
        return self._rep


TFM_BUFFERS = ("encoder.head.2.running_mean", "encoder.head.2.running_var", "encoder.head.2.num_batches_tracked",
               "encoder.head.5.running_mean", "encoder.head.5.running_var", "encoder.head.5.num_batches_tracked")


class VaDEB200:
    """B200-native stand-in for ``VaDEPT(encoder_type="recurrent" | "transformer" | "TCN", use_gnn=True)``."""
    _MODEL = _lib.MODEL_VADE
    _BUFFERS = ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence", "latent_space.prior",
                "latent_space.pretrain") + TFM_BUFFERS
    _DEC_PASSES = 1          # decoder passes per training step (dropout mask layout of the transformer family)

    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int, n_components: int,
                 encoder_type: str = "recurrent", use_gnn: bool = True, kmeans_loss: float = 1.0,
                 interaction_regularization: float = 0.0, device: Optional[int] = None, max_batch: int = 4096,
                 training: bool = True, seed: Optional[int] = None):
        if encoder_type not in _lib.ENCODER_KINDS or not use_gnn:
            raise NotImplementedError("deepof_b200 implements the recurrent, transformer and TCN GNN encoders "
                                      f"(got encoder_type={encoder_type!r}, use_gnn={use_gnn})")
        self.encoder_type = encoder_type
        if not torch.cuda.is_available():
            raise _lib.DofError("deepof_b200 needs a CUDA device (no CPU fallback)")
        self.L = lib()
        T, N, F = (int(v) for v in input_shape)
        T2, E, Fe = (int(v) for v in edge_feature_shape)
        assert T == T2
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.adjacency_matrix = np.asarray(adjacency_matrix, dtype=np.float64)
        lap, elap, inc = graph_operators(self.adjacency_matrix)
        assert inc.shape[1] == E, f"adjacency has {inc.shape[1]} edges, edge_feature_shape says {E}"
        self.cfg = DofConfig(T, N, E, F, Fe, int(latent_dim), int(n_components), self._MODEL,
                             _lib.ENCODER_KINDS[encoder_type])
        self._drop_step = 0
        self._drop_seed = int(seed) if seed is not None else 0
        self.window_size = T
        self.latent_dim, self.n_components = int(latent_dim), int(n_components)
        self.kmeans_weight = float(kmeans_loss)
        self.input_shape, self.edge_feature_shape = (T, N, F), (T, E, Fe)
        self.layout = state_layout(self.cfg)
        self.n_state = int(self.L.dof_state_numel(C.byref(self.cfg)))
        self.state = torch.zeros(self.n_state, dtype=torch.float32, device=self.device)
        self._views = OrderedDict()
        for name, off, numel, shape, grp in self.layout:
            self._views[name] = self.state[off:off + numel].view(shape if len(shape) else ())
        self.max_batch, self.training_capable = int(max_batch), bool(training)
        nbytes = self.L.dof_workspace_bytes(C.byref(self.cfg), self.max_batch, int(training))
        if nbytes == 0:
            check(-1)
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        check(self.L.dof_create(C.byref(self.cfg), self.device.index, self.max_batch, int(training),
                                C.c_void_p(self.workspace.data_ptr()), nbytes, C.byref(h)))
        self.handle = h
        self.grad = torch.zeros_like(self.state) if training else None
        self.adam_m = torch.zeros_like(self.state) if training else None
        self.adam_v = torch.zeros_like(self.state) if training else None
        self.adam_steps = [0, 0, 0, 0]
        self.logs = torch.zeros(_lib.DOF_N_LOGS, dtype=torch.float32, device=self.device)
        self._training = False
        # surfaces poked by reference helper code
        self.encoder = self._Encoder(self)
        if self._MODEL == _lib.MODEL_VADE:
            self.latent_space = self._Latent(self)
        self.reset_parameters(seed)
        with torch.no_grad():
            self._views["encoder.laplacian"].copy_(torch.from_numpy(lap))
            self._views["encoder.edge_laplacian"].copy_(torch.from_numpy(elap))
            self._views["encoder.incidence"].copy_(torch.from_numpy(inc))

    # ---- reference-looking sub-objects
    class _Encoder:
        def __init__(self, m):
            self._m = m
            self.spatial_gnn_block = _Named("CensNetConvPT()")

        def __call__(self, x, a):
            return self._m.forward_eval(x, a, want_loc=False)[0]

    class _Latent:
        def __init__(self, m):
            self._m = m

        @property
        def gmm_means(self):
            return self._m._views["latent_space.gmm_means"]

        @property
        def gmm_log_vars(self):
            return self._m._views["latent_space.gmm_log_vars"]

        @property
        def prior(self):
            return self._m._views["latent_space.prior"]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.L.dof_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- parameters
    def reset_parameters(self, seed: Optional[int] = None):
        """torch-default initialisers of the reference modules."""
        g = torch.Generator(device="cpu")
        if seed is not None:
            g.manual_seed(seed)
        else:
            g.seed()
        D, K = self.latent_dim, self.n_components

        def uni(shape, bound):
            return (torch.rand(shape, generator=g) * 2 - 1) * bound

        for name, off, numel, shape, grp in self.layout:
            v = self._views[name]
            leaf = name.rsplit(".", 1)[-1]
            if name in ("encoder.laplacian", "encoder.edge_laplacian", "encoder.incidence"):
                continue
            if name.endswith("num_batches_tracked") or name.endswith("running_mean"):
                val = torch.zeros(shape)
            elif name.endswith("running_var"):
                val = torch.ones(shape)
            elif name == "latent_space.prior":
                val = torch.full(shape, 1.0 / K)
            elif name == "latent_space.pretrain":
                val = torch.zeros(())
            elif ".gru" in name:      # nn.GRU: U(+-1/sqrt(hidden))
                hidden = shape[0] // 3
                val = uni(shape, 1.0 / math.sqrt(hidden))
            elif ".norm" in name or (name.startswith("encoder.head.") and len(shape) == 1 and name.split(".")[2] in ("2", "5")):
                val = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)        # LayerNorm / BatchNorm affine
            elif ".bn" in name and leaf in ("weight", "bias"):                 # BatchNorm affine of the TCN blocks / decoder
                val = torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
            elif self.encoder_type == "TCN" and (".conv1." in name or ".conv2." in name or ".downsample." in name):
                # nn.init.normal_(std=0.05) weights, zero biases (models_new.py:420-428)
                val = torch.randn(shape, generator=g) * 0.05 if leaf == "weight" else torch.zeros(shape)
            elif self.encoder_type == "TCN" and (name.startswith("encoder.head.") or name.startswith("decoder.fc")):
                # xavier_uniform_ weights, zero biases (models_new.py:604-607, 776-779)
                val = uni(shape, math.sqrt(6.0 / (shape[0] + shape[1]))) if leaf == "weight" else torch.zeros(shape)
            elif self.encoder_type == "transformer" and ("_tf." in name or name.startswith("encoder.head.") or name.startswith("decoder.")):
                # xavier_uniform_ weights, zero biases (models_new.py:868-871, 1085-1089, 1225-1230); embed / prob_decoder keep
                # nn.Linear's default init
                if ".embed." in name or "prob_decoder" in name:
                    fan_in = self._views[name[:-4] + "weight"].shape[1] if leaf == "bias" else shape[1]
                    val = uni(shape, 1.0 / math.sqrt(fan_in))
                elif leaf == "weight":
                    val = uni(shape, math.sqrt(6.0 / (shape[0] + shape[1])))
                else:
                    val = torch.zeros(shape)
            elif "conv1d.weight" in name:   # kaiming_uniform(a=sqrt5) == U(+-1/sqrt(fan_in))
                val = uni(shape, 1.0 / math.sqrt(shape[1] * shape[2]))
            elif "spatial_gnn_block" in name:
                if leaf.endswith("bias"):
                    val = uni(shape, 1.0 / math.sqrt(D))   # fan_in of kernel[2D, D] is size(1) = D
                else:
                    fan_out, fan_in = shape[0], shape[1]
                    val = uni(shape, math.sqrt(6.0 / (fan_in + fan_out)))
            elif name == "vq_layer.codebook":      # uniform_(0, 1), models_new.py:1349-1351
                val = torch.rand(shape, generator=g)
            elif name in ("latent_space.gmm_means", "latent_space.gmm_log_vars"):
                val = torch.randn(shape, generator=g) * math.sqrt(2.0 / (K + D))
            elif leaf == "weight":     # nn.Linear
                val = uni(shape, 1.0 / math.sqrt(shape[1]))
            elif leaf == "bias":
                wshape = self._views[name[:-4] + "weight"].shape
                val = uni(shape, 1.0 / math.sqrt(wshape[1]))
            else:
                raise AssertionError(name)
            with torch.no_grad():
                v.copy_(val.to(self.device))

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, v.detach().clone().to(torch.int64) if k.endswith("num_batches_tracked") else v.detach().clone())
                           for k, v in self._views.items())

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        missing = [k for k in self._views if k not in sd]
        extra = [k for k in sd if k not in self._views]
        if strict and (missing or extra):
            raise KeyError(f"state_dict mismatch: missing={missing} unexpected={extra}")
        with torch.no_grad():
            for k, v in self._views.items():
                if k in sd:
                    t = torch.as_tensor(sd[k]).to(torch.float32)
                    if tuple(t.shape) != tuple(v.shape):
                        raise ValueError(f"shape mismatch for {k}: {tuple(t.shape)} vs {tuple(v.shape)}")
                    v.copy_(t.to(self.device))
        return self

    def parameters(self):
        return [v for (name, *_r, grp), v in zip(self.layout, self._views.values()) if grp > 0]

    def named_parameters(self):
        return [(name, v) for (name, *_r, grp), v in zip(self.layout, self._views.values())
                if name not in self._BUFFERS and not name.endswith(("running_mean", "running_var", "num_batches_tracked"))]

    def to(self, *a, **k):
        return self

    def eval(self):
        self._training = False
        return self

    def train(self, mode: bool = True):
        self._training = bool(mode)
        return self

    def set_pretrain_mode(self, flag: bool):
        with torch.no_grad():
            self._views["latent_space.pretrain"].fill_(1.0 if flag else 0.0)

    # ---- forward
    def _prep(self, x, a):
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device, non_blocking=True).contiguous()
        a = torch.as_tensor(a, dtype=torch.float32).to(self.device, non_blocking=True).contiguous()
        T, N, F = self.input_shape
        _, E, Fe = self.edge_feature_shape
        assert x.shape[1:] == (T, N, F) and a.shape[1:] == (T, E, Fe), (x.shape, a.shape)
        assert x.shape[0] == a.shape[0]
        return x, a

    def forward_eval(self, x, a, want_loc: bool = True):
        """Eval-mode forward: (enc [B,D], emb [B,D], q [B,K], loc [B,T,N*F] or None)."""
        x, a = self._prep(x, a)
        T, N, F = self.input_shape
        outs = ([], [], [], [])
        for s in range(0, x.shape[0], self.max_batch):
            xb, ab = x[s:s + self.max_batch], a[s:s + self.max_batch]
            B = xb.shape[0]
            enc = torch.empty(B, self.latent_dim, device=self.device)
            emb = torch.empty(B, self.latent_dim, device=self.device)
            q = torch.empty(B, self.n_components, device=self.device)
            loc = torch.empty(B, T, N * F, device=self.device) if want_loc else None
            check(self.L.dof_vade_forward_eval(self.handle, ptr(self.state), ptr(xb), ptr(ab), B, ptr(enc),
                                               ptr(emb), ptr(q), ptr(loc), _stream()))
            for o, t in zip(outs, (enc, emb, q, loc)):
                o.append(t)
        cat = lambda l: None if l[0] is None else (l[0] if len(l) == 1 else torch.cat(l))
        return tuple(cat(o) for o in outs)

    def embed(self, x, a):
        """(embedding = z_mean, q): what reference embedding_per_video reads as model(x,a)[1], [2]."""
        x, a = self._prep(x, a)
        embs, qs = [], []
        for s in range(0, x.shape[0], self.max_batch):
            xb, ab = x[s:s + self.max_batch], a[s:s + self.max_batch]
            B = xb.shape[0]
            emb = torch.empty(B, self.latent_dim, device=self.device)
            q = torch.empty(B, self.n_components, device=self.device)
            check(self.L.dof_vade_embed(self.handle, ptr(self.state), ptr(xb), ptr(ab), B, ptr(emb), ptr(q),
                                        _stream()))
            embs.append(emb)
            qs.append(q)
        return (embs[0], qs[0]) if len(embs) == 1 else (torch.cat(embs), torch.cat(qs))

    def __call__(self, x, a, return_gmm_params: bool = False):
        """Eval-mode ``VaDEPT.forward``: (loc, embedding, q, kmeans_loss=0)."""
        enc, emb, q, loc = self.forward_eval(x, a, want_loc=True)
        return loc, emb, q, torch.zeros((), device=self.device)

    # ---- training
    def dropout_mask_bytes(self, B: int, encoder_windows: Optional[int] = None) -> int:
        return int(self.L.dof_dropout_mask_bytes(C.byref(self.cfg), int(encoder_windows or B), int(B if self._DEC_PASSES else 0),
                                                 self._DEC_PASSES))

    def _set_dropout(self, B: int, masks=None, encoder_windows: Optional[int] = None):
        """Transformer family: explicit keep masks (uint8 on the device, reference draw order — see ``dof_set_dropout``)
        for parity runs, else a fresh Philox stream per step (seed advances with every call)."""
        if self.encoder_type != "transformer":
            return
        if masks is not None:
            m = torch.as_tensor(masks).to(self.device, torch.uint8).contiguous()
            need = self.dropout_mask_bytes(B, encoder_windows)
            if m.numel() != need:
                raise ValueError(f"dropout masks: {m.numel()} bytes, the step needs {need}")
            self._drop_keep = m
            check(self.L.dof_set_dropout(self.handle, 0, ptr(m), m.numel()))
        else:
            self._drop_step += 1
            self._drop_keep = None
            seed = (self._drop_seed * 0x9E3779B97F4A7C15 + self._drop_step) & 0xFFFFFFFFFFFFFFFF
            check(self.L.dof_set_dropout(self.handle, seed, None, 0))

    def loss_grad(self, x, a, loss_cfg: VadeLossCfg, eps=None, mc_eps=None, tau_batch=None, class_weight=None,
                  teacher_marginal=None, dropout_masks=None):
        """forward + VadeLoss + backward into ``self.grad``; returns the device log vector."""
        if not self.training_capable:
            raise _lib.DofError("model was created with training=False")
        x, a = self._prep(x, a)
        B = x.shape[0]
        D, K = self.latent_dim, self.n_components
        # eps / mc_eps None: the kernels draw the noise themselves (Philox, keyed per step) — nothing is materialised
        f32 = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32).to(self.device).contiguous()
        eps, mc_eps, tau_batch, class_weight = f32(eps), f32(mc_eps), f32(tau_batch), f32(class_weight)
        if eps is None or (mc_eps is None and not loss_cfg.pretrain_mode):
            self._noise_step = getattr(self, "_noise_step", 0) + 1
            check(self.L.dof_set_noise_seed(self.handle, ((self._drop_seed + 1) * 0xD1B54A32D192ED03 + self._noise_step) & 0xFFFFFFFFFFFFFFFF))
        floor = torch.full((K,), float(loss_cfg.nonempty_floor), device=self.device)
        if teacher_marginal is not None:   # reference losses.py:672-678
            floor = torch.maximum(floor, 0.9 * f32(teacher_marginal))
        c = loss_cfg.to_c()
        self._set_dropout(B, dropout_masks)
        check(self.L.dof_vade_loss_grad(self.handle, ptr(self.state), ptr(self.grad), ptr(x), ptr(a), B, ptr(eps),
                                        ptr(mc_eps), ptr(tau_batch), ptr(class_weight), ptr(floor), C.byref(c),
                                        ptr(self.logs), _stream()))
        return self.logs

    def loss_eval(self, x, a, loss_cfg: VadeLossCfg, mc_eps=None, tau_batch=None, class_weight=None, teacher_marginal=None):
        """The validation step (``validate_one_epoch_indexed``, training.py:190-229): eval-mode forward (z = z_mean, no
        dropout, BatchNorm running statistics) + the criterion's terms; returns the device log vector.  No gradient,
        ``self.grad`` is left untouched."""
        if not self.training_capable:
            raise _lib.DofError("model was created with training=False")
        x, a = self._prep(x, a)
        B = x.shape[0]
        D, K = self.latent_dim, self.n_components
        f32 = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32).to(self.device).contiguous()
        mc_eps, tau_batch, class_weight = f32(mc_eps), f32(tau_batch), f32(class_weight)
        if mc_eps is None and not loss_cfg.pretrain_mode:
            self._noise_step = getattr(self, "_noise_step", 0) + 1
            check(self.L.dof_set_noise_seed(self.handle, ((self._drop_seed + 1) * 0xD1B54A32D192ED03 + self._noise_step) & 0xFFFFFFFFFFFFFFFF))
        floor = torch.full((K,), float(loss_cfg.nonempty_floor), device=self.device)
        if teacher_marginal is not None:
            floor = torch.maximum(floor, 0.9 * f32(teacher_marginal))
        c = loss_cfg.to_c()
        check(self.L.dof_vade_loss_eval(self.handle, ptr(self.state), ptr(x), ptr(a), B, ptr(mc_eps), ptr(tau_batch),
                                        ptr(class_weight), ptr(floor), C.byref(c), ptr(self.logs), _stream()))
        return self.logs

    def initialize_gmm_from_data(self, data_loader, n_samples: int = 10000):
        """``VaDEPT.initialize_gmm_from_data`` (models_new.py:1907-1947): z_mean of the first ``n_samples`` windows in
        loader order (eval mode), a scikit-learn diagonal GaussianMixture (reg_covar 1e-4, numpy's global RNG like the
        reference) on them, its means / log-covariances written into the latent space.  ``data_loader`` yields
        ``(x, a, ...)`` batches.  The GMM fit is the reference's own third-party call, off the step path."""
        from .gmm_init import gmm_from_embeddings
        embs, got = [], 0
        for batch in data_loader:
            z = self.embed(batch[0], batch[1])[0]
            embs.append(z.detach().cpu())
            got += z.shape[0]
            if got >= n_samples:
                break
        means, log_vars = gmm_from_embeddings(torch.cat(embs).numpy()[:n_samples], self.n_components)
        with torch.no_grad():
            self.latent_space.gmm_means.copy_(torch.from_numpy(means).float().to(self.device))
            self.latent_space.gmm_log_vars.copy_(torch.from_numpy(log_vars).float().to(self.device))
        return means, log_vars

    def adam_step(self, lr_base: float, lr_gmm: float, lr_decoder: Optional[float] = None, clip: float = 0.75,
                  grad_scale: float = 1.0, active=(True, True, True), betas=(0.9, 0.999), eps: float = 1e-8,
                  weight_decay: float = 0.0):
        """clip_grad_value_ + Adam.  Groups: encoder+latent heads / decoder / GMM (or VQ codebook)."""
        o = DofAdamCfg()
        lrs = (0.0, lr_base, lr_base if lr_decoder is None else lr_decoder, lr_gmm)
        for g in range(1, 4):
            if active[g - 1]:
                self.adam_steps[g] += 1
            o.lr[g] = lrs[g]
            o.step[g] = max(1, self.adam_steps[g])
            o.active[g] = int(bool(active[g - 1]))
            o.weight_decay[g] = weight_decay
        o.clip_value, o.grad_scale, o.beta1, o.beta2, o.eps = clip, grad_scale, betas[0], betas[1], eps
        check(self.L.dof_clip_adam(self.handle, ptr(self.state), ptr(self.grad), ptr(self.adam_m),
                                   ptr(self.adam_v), C.byref(o), _stream()))

    def logs_dict(self) -> Dict[str, float]:
        v = self.logs.detach().cpu().tolist()
        return {k: v[i] for i, k in enumerate(LOG_KEYS)}

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return OrderedDict((name, self.grad[off:off + numel].view(shape if len(shape) else ()))
                           for name, off, numel, shape, grp in self.layout)

    def debug(self, name: str, dtype=torch.float32):
        n = C.c_int64()
        p = self.L.dof_debug_tensor(self.handle, name.encode(), C.byref(n))
        if not p:
            raise KeyError(name)
        # copy out of the workspace
        off = p - self.workspace.data_ptr()
        raw = self.workspace[off:off + n.value * 4]
        return raw.view(dtype).clone()
