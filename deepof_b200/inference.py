"""Streaming ``embedding_per_video`` (reference ``deepof/clustering/model_utils_new.py:452-748``).

The reference moves every video's materialised window table to the device and runs ``model(x, a)`` in chunks of
256; here the windows never exist outside one batch: the loader kernel builds each batch from the resident frame
table and the model's eval forward consumes it on the same stream.  Returns, per video, the embeddings ``[Nw, D]``
and the soft counts ``[Nw, K]`` (``None`` for contrastive models, like the reference).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .loader import WindowLoader


def embedding_per_video(model, loader: WindowLoader, batch_size: Optional[int] = None) -> Tuple[List[torch.Tensor], List[Optional[torch.Tensor]]]:
    bs = int(batch_size or getattr(model, "max_windows", None) or model.max_batch)
    bs = min(bs, getattr(model, "max_windows", model.max_batch))
    T, N, F = model.input_shape
    assert (loader.T, loader.N) == (T, N), "loader geometry does not match the model"
    xbuf = torch.empty(bs, T, N, 3, device=loader.device)
    abuf = torch.empty(bs, T, loader.E, 1, device=loader.device)
    embs: List[torch.Tensor] = []
    softs: List[Optional[torch.Tensor]] = []
    for v, nw in enumerate(loader.n_windows_per_video):
        e_parts, q_parts = [], []
        w0 = int(loader.window_offsets[v])
        for s in range(0, nw, bs):
            n = min(bs, nw - s)
            x, a = loader.load(w0 + s, n, xbuf, abuf)
            e, q = model.embed(x, a)
            e_parts.append(e.clone())
            q_parts.append(None if q is None else q.clone())
        D = model.latent_dim
        embs.append(torch.cat(e_parts) if e_parts else torch.empty(0, D, device=loader.device))
        softs.append(None if (not q_parts or q_parts[0] is None) else torch.cat(q_parts))
    return embs, softs
