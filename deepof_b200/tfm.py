"""Host-side mirror of the reference ``TFMEncoderPT`` (``deepof/clustering/models_new.py:985-1164``) for INFERENCE:
``encoder(x, a) -> [B, latent_dim]`` with the module in ``eval()`` — the call ``embedding_per_video``
(``model_utils_new.py:545-621``) makes on the transformer model family.  The forward runs in the CUDA library
(``dof_tfm_encode``, ``csrc/tfm.cuh``); torch is device memory only.  The training step of this encoder is not built
yet: ``train()`` raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import DofTfmCfg, DofTfmDecCfg, check, lib, ptr
from .vade import _stream, graph_operators


def tfm_layout(cfg: DofTfmCfg):
    """[(name, offset, numel, shape)] in the reference's state_dict order."""
    L = lib()
    n = L.dof_tfm_num_entries(C.byref(cfg))
    if n < 0:
        check(-1)
    out = []
    name = C.create_string_buffer(128)
    off, numel, ndim = C.c_int64(), C.c_int64(), C.c_int()
    shape = (C.c_int * 4)()
    for i in range(n):
        check(L.dof_tfm_entry(C.byref(cfg), i, name, C.byref(off), C.byref(numel), C.byref(ndim), shape))
        out.append((name.value.decode(), off.value, numel.value, tuple(shape[: ndim.value])))
    return out


class TFMEncoderB200:
    """``TFMEncoderPT(input_shape, edge_feature_shape, adjacency_matrix, latent_dim, use_gnn=True)`` in eval mode."""
    _BUFFERS = ("laplacian", "edge_laplacian", "incidence", "head.2.running_mean", "head.2.running_var",
                "head.5.running_mean", "head.5.running_var")

    def __init__(self, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int, use_gnn: bool = True,
                 num_layers: int = 2, num_heads: int = 4, dff: int = 128, key_dim: Optional[int] = None,
                 device: Optional[int] = None, max_batch: int = 1024, seed: Optional[int] = None):
        if not use_gnn:
            raise NotImplementedError("deepof_b200 implements the GNN path of the transformer encoder only")
        T, N, F = (int(v) for v in input_shape)
        _, E, Fe = (int(v) for v in edge_feature_shape)
        adjacency_matrix = np.asarray(adjacency_matrix)
        assert adjacency_matrix.shape == (N, N), "Adjacency must be NxN and match input nodes."
        if key_dim is None:                                      # models_new.py:1014-1019
            key_dim = max((min(64, N * F) // num_heads) * num_heads, num_heads)
        self.input_shape, self.edge_feature_shape = (T, N, F), (T, E, Fe)
        self.latent_dim, self.key_dim, self.window_size = int(latent_dim), int(key_dim), T
        self.cfg = DofTfmCfg(T, N, E, F, Fe, int(latent_dim), int(key_dim), int(num_heads), int(dff), int(num_layers))
        self.L = lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.max_batch = int(max_batch)
        self.layout = tfm_layout(self.cfg)
        total = self.L.dof_tfm_numel(C.byref(self.cfg))
        host = torch.zeros(total)
        g = torch.Generator().manual_seed(int(seed)) if seed is not None else None
        lap, elap, inc = graph_operators(adjacency_matrix)
        if inc.shape[1] != E:
            raise ValueError(f"edge_feature_shape has {E} edges, the adjacency matrix {inc.shape[1]}")
        for name, off, numel, shape in self.layout:
            v = host[off:off + numel].view(shape)
            if name == "laplacian":
                v.copy_(torch.from_numpy(lap))
            elif name == "edge_laplacian":
                v.copy_(torch.from_numpy(elap))
            elif name == "incidence":
                v.copy_(torch.from_numpy(inc))
            elif name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
                v.fill_(1.0)                                     # LayerNorm / BatchNorm scales, BN variance
            elif len(shape) == 2 and shape[1] > 1 or name.endswith("_kernel"):
                bound = math.sqrt(6.0 / (shape[0] + shape[1]))  # xavier_uniform_
                v.copy_((torch.rand(shape, generator=g) * 2.0 - 1.0) * bound)
            elif name.endswith("_weights"):
                v.copy_((torch.rand(shape, generator=g) * 2.0 - 1.0) * math.sqrt(6.0 / (shape[0] + 1)))
        self.state = host.to(self.device)
        self._views = {name: self.state[off:off + numel].view(shape) for name, off, numel, shape in self.layout}
        self._ws = torch.empty(self.L.dof_tfm_workspace_bytes(C.byref(self.cfg), self.max_batch), dtype=torch.uint8, device=self.device)
        self.training = False

    # ---- torch.nn.Module-like surface ---------------------------------------------------------------------------
    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("the training step of the transformer encoder is not built; eval() only")
        return self

    def to(self, *a, **k):
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.clone() for k, v in self._views.items()}

    def load_state_dict(self, sd, strict: bool = True) -> None:
        """Accepts the reference's ``TFMEncoderPT.state_dict()`` (optionally prefixed with ``encoder.``); the integer
        ``num_batches_tracked`` buffers are ignored."""
        sd = {(k[len("encoder."):] if k.startswith("encoder.") else k): v for k, v in sd.items()}
        missing = [k for k in self._views if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        for k, v in self._views.items():
            if k in sd:
                t = torch.as_tensor(np.asarray(sd[k]) if not torch.is_tensor(sd[k]) else sd[k]).to(torch.float32)
                if tuple(t.shape) != tuple(v.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(v.shape)}")
                v.copy_(t.to(self.device))

    # ---- forward ------------------------------------------------------------------------------------------------
    def encode(self, x, a, return_cores: bool = False):
        T, N, F = self.input_shape
        _, E, Fe = self.edge_feature_shape
        x = torch.as_tensor(x).to(self.device, torch.float32).contiguous()
        a = torch.as_tensor(a).to(self.device, torch.float32).contiguous()
        assert tuple(x.shape[1:]) == (T, N, F), f"Input shape mismatch: got {tuple(x.shape[1:])}, expected {(T, N, F)}"
        assert tuple(a.shape[1:]) == (T, E, Fe) and a.shape[0] == x.shape[0]
        B = x.shape[0]
        out = torch.empty(B, self.latent_dim, device=self.device)
        nodes = torch.empty(B * N, self.key_dim, device=self.device) if return_cores else None
        edges = torch.empty(B * E, self.key_dim, device=self.device) if return_cores else None
        for s in range(0, B, self.max_batch):
            e = min(B, s + self.max_batch)
            check(self.L.dof_tfm_encode(C.byref(self.cfg), ptr(self.state), ptr(x[s:e]), ptr(a[s:e]), e - s, ptr(self._ws),
                                        self._ws.numel(), ptr(out[s:e]), ptr(nodes[s * N:e * N]) if return_cores else None,
                                        ptr(edges[s * E:e * E]) if return_cores else None, _stream()))
        return (out, nodes, edges) if return_cores else out

    __call__ = encode


class TFMDecoderB200:
    """``TFMDecoderPT(output_shape=(T, N*F), latent_dim, num_layers=2, num_heads=8, dff=128)`` in eval mode
    (``models_new.py:1167-1266``): ``decode(z) -> loc [B, T, N*F]``, the mean of the reconstruction distribution."""

    def __init__(self, output_shape, latent_dim: int, num_layers: int = 2, num_heads: int = 8, dff: int = 128,
                 device: Optional[int] = None, max_batch: int = 1024):
        T, Dx = (int(v) for v in output_shape)
        self.output_shape, self.latent_dim, self.max_batch = (T, Dx), int(latent_dim), int(max_batch)
        self.cfg = DofTfmDecCfg(T, Dx, int(latent_dim), int(num_heads), int(dff), int(num_layers))
        self.L = lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        n = self.L.dof_tfm_dec_num_entries(C.byref(self.cfg))
        if n < 0:
            check(-1)
        name = C.create_string_buffer(128)
        off, numel, ndim = C.c_int64(), C.c_int64(), C.c_int()
        shape = (C.c_int * 4)()
        self.layout = []
        for i in range(n):
            check(self.L.dof_tfm_dec_entry(C.byref(self.cfg), i, name, C.byref(off), C.byref(numel), C.byref(ndim), shape))
            self.layout.append((name.value.decode(), off.value, numel.value, tuple(shape[: ndim.value])))
        self.state = torch.zeros(self.L.dof_tfm_dec_numel(C.byref(self.cfg)), device=self.device)
        self._views = {nm: self.state[o:o + ne].view(sh) for nm, o, ne, sh in self.layout}
        for nm, v in self._views.items():
            if nm.endswith("norm1.weight") or nm.endswith("norm2.weight"):
                v.fill_(1.0)
        self._ws = torch.empty(self.L.dof_tfm_dec_workspace_bytes(C.byref(self.cfg), self.max_batch), dtype=torch.uint8, device=self.device)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.clone() for k, v in self._views.items()}

    def load_state_dict(self, sd, strict: bool = True) -> None:
        sd = {(k[len("decoder."):] if k.startswith("decoder.") else k): v for k, v in sd.items()}
        for k, v in self._views.items():
            if k in sd:
                v.copy_(torch.as_tensor(sd[k]).to(self.device, torch.float32).reshape(v.shape))
            elif strict:
                raise KeyError(k)

    def decode(self, z) -> torch.Tensor:
        T, Dx = self.output_shape
        z = torch.as_tensor(z).to(self.device, torch.float32).contiguous()
        B = z.shape[0]
        loc = torch.empty(B, T, Dx, device=self.device)
        for s in range(0, B, self.max_batch):
            e = min(B, s + self.max_batch)
            check(self.L.dof_tfm_decode(C.byref(self.cfg), ptr(self.state), ptr(z[s:e]), e - s, ptr(self._ws), self._ws.numel(),
                                        ptr(loc[s:e]), _stream()))
        return loc

    __call__ = decode


class TFMModelB200:
    """Inference stand-in for the reference models built with ``encoder_type="transformer"`` (``VaDEPT``, ``VQVAEPT``,
    ``ContrastivePT``): the transformer encoder plus the read-out ``embedding_per_video`` needs — VaDE: ``z_mean`` and the
    GMM posterior (``GaussianMixtureLatentPT``, models_new.py:1745-1791); VQ-VAE: encoder output and soft counts
    (``VectorQuantizerPT``, :1358-1423); contrastive: the encoder output.  ``reconstruct`` runs the transformer decoder
    (``TFMDecoderB200``).  The training step is not built: ``__call__`` / ``train`` raise.  ``load_state_dict`` takes the
    reference checkpoint's state_dict; tensors the library does not use are kept on the host so that ``state_dict()``
    round-trips."""

    def __init__(self, model_name: str, input_shape, edge_feature_shape, adjacency_matrix, latent_dim: int,
                 n_components: int = 1, max_batch: int = 1024, device: Optional[int] = None, seed: Optional[int] = None):
        self.model_name = str(model_name).lower()
        if self.model_name not in ("vade", "vqvae", "contrastive"):
            raise ValueError(f"unknown model_name {model_name!r}")
        if self.model_name == "contrastive":                     # ContrastivePT encodes half windows (models_new.py:2013)
            input_shape = (int(input_shape[0]) // 2,) + tuple(input_shape[1:])
            edge_feature_shape = (int(edge_feature_shape[0]) // 2,) + tuple(edge_feature_shape[1:])
        self.encoder = TFMEncoderB200(input_shape, edge_feature_shape, adjacency_matrix, latent_dim, max_batch=max_batch,
                                      device=device, seed=seed)
        self.input_shape, self.edge_feature_shape = self.encoder.input_shape, self.encoder.edge_feature_shape
        self.latent_dim, self.n_components, self.max_batch = int(latent_dim), int(n_components), int(max_batch)
        self.window_size, self.device, self.L = self.encoder.window_size, self.encoder.device, self.encoder.L
        self.training_capable = False
        D, K = self.latent_dim, self.n_components
        if self.model_name == "vade":
            shapes = {"latent_space.gmm_means": (K, D), "latent_space.gmm_log_vars": (K, D), "latent_space.prior": (K,),
                      "latent_space.encoder_mean.weight": (D, D), "latent_space.encoder_mean.bias": (D,),
                      "latent_space.encoder_log_var.weight": (D, D), "latent_space.encoder_log_var.bias": (D,)}
        elif self.model_name == "vqvae":
            shapes = {"vq_layer.codebook": (D, K)}
        else:
            shapes = {}
        self.decoder = None
        if self.model_name != "contrastive":                     # models_new.py:1490-1497: heads 8, dff 128, 2 layers
            T, N, F = self.input_shape
            self.decoder = TFMDecoderB200((T, N * F), latent_dim, device=device, max_batch=max_batch)
        self._head = {k: torch.zeros(s, device=self.device) for k, s in shapes.items()}
        if "latent_space.prior" in self._head:
            self._head["latent_space.prior"].fill_(1.0 / K)
        self._other: Dict[str, torch.Tensor] = {}

    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("the training step of the transformer model family is not built; eval() only")
        return self

    def to(self, *a, **k):
        return self

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {"encoder." + k: v for k, v in self.encoder.state_dict().items()}
        out.update({k: v.clone() for k, v in self._head.items()})
        if self.decoder is not None:
            out.update({"decoder." + k: v for k, v in self.decoder.state_dict().items()})
        out.update(self._other)
        return out

    def load_state_dict(self, sd, strict: bool = True) -> None:
        self.encoder.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=strict)
        for k, v in self._head.items():
            if k in sd:
                v.copy_(torch.as_tensor(sd[k]).to(self.device, torch.float32).reshape(v.shape))
            elif strict:
                raise KeyError(k)
        if self.decoder is not None:
            self.decoder.load_state_dict({k: v for k, v in sd.items() if k.startswith("decoder.")}, strict=strict)
        self._other = {k: torch.as_tensor(v).clone() for k, v in sd.items()
                       if not k.startswith("encoder.") and not k.startswith("decoder.") and k not in self._head}

    def encode(self, x, a):
        return self.encoder(x, a)

    def embed(self, x, a):
        """``(embedding [B,D], soft assignments [B,K] or None)`` exactly as ``embedding_per_video`` reads them from the
        reference models (model_utils_new.py:585-612)."""
        enc = self.encoder(x, a)
        B, D, K = enc.shape[0], self.latent_dim, self.n_components
        h = self._head
        if self.model_name == "vade":
            emb, q = torch.empty(B, D, device=self.device), torch.empty(B, K, device=self.device)
            scratch = torch.empty(3 * B * D, device=self.device)
            check(self.L.dof_latent_eval(ptr(enc), ptr(h["latent_space.encoder_mean.weight"]), ptr(h["latent_space.encoder_mean.bias"]),
                                         ptr(h["latent_space.encoder_log_var.weight"]), ptr(h["latent_space.encoder_log_var.bias"]),
                                         ptr(h["latent_space.gmm_means"]), ptr(h["latent_space.gmm_log_vars"]), ptr(h["latent_space.prior"]),
                                         B, D, K, ptr(emb), ptr(q), ptr(scratch), _stream()))
            return emb, q
        if self.model_name == "vqvae":
            quant, soft = torch.empty(B, D, device=self.device), torch.empty(B, K, device=self.device)
            idx = torch.empty(B, dtype=torch.int32, device=self.device)
            scratch = torch.empty(4 + D * D + K, dtype=torch.float64, device=self.device)
            check(self.L.dof_vq_eval(ptr(enc), ptr(h["vq_layer.codebook"]), B, D, K, ptr(quant), ptr(soft), ptr(idx), ptr(scratch), _stream()))
            return enc, soft
        return enc, None

    def reconstruct(self, x, a) -> torch.Tensor:
        """Mean of the reconstruction distribution ``model(x, a)[0]`` in eval mode: VaDE decodes ``z_mean``, VQ-VAE the
        encoder output (``VQVAEPT.forward`` returns the decode of the quantized latents first; use ``decoder(quant)``)."""
        if self.decoder is None:
            raise NotImplementedError("ContrastivePT has no decoder")
        return self.decoder(self.embed(x, a)[0])

    def __call__(self, *a, **k):
        raise NotImplementedError("only inference read-outs are built for the transformer family: .embed(x, a), "
                                  ".encoder(x, a), .reconstruct(x, a)")
