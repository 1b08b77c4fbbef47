"""Frame-table window loader on top of the deepof_b200 C-ABI (``dof_load_windows``).

Stands in for the reference's CPU chain between per-video pose tables and the window store the
trainer iterates (paths relative to the reference root):

* ``Coordinates.get_graph_dataset(..., preprocess=True)`` -> ``TableDict.preprocess`` -> ``deepof/utils.py``
  ``scale_table`` :2425-2566, ``_pp_*`` :2665-3027, ``rolling_window`` :3354-3377 (centre / align / speed /
  size-normalise / log1p / per-video + global standard scalers / clip + interpolate / windows), and
* ``deepof/clustering/dataset.py`` ``BatchDictDataset`` :183-290 (every window materialised to HDF5) and
  ``_H5BatchIterableDataset.__iter__`` :561-671 (contiguous batches, epoch-seeded shuffle of batch starts,
  ``starts[rank::world]``).

Only the raw frames ``[n_frames, N, 2]`` of each video stay resident in HBM (2.2 GB for 10 M windows of
14 body parts instead of 56 GB of materialised windows); a batch of windows is produced on the fly by one
kernel launch per video segment.  torch holds the device memory; every number is computed by the CUDA
library.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import DofLoaderCfg, check, lib, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def reference_divisors(size: float, edges: np.ndarray, n_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-column size divisors ``scale_table`` applies to the merged (coords | speeds | distances) table
    (``deepof/utils.py:2510-2526``): ``out.loc[:, (bp1, bp2)]`` on the flat column Index selects the two SPEED
    columns ``bp1``, ``bp2`` rather than the distance column, so the speed of node ``n`` ends up divided by
    ``size ** (1 + degree(n))`` and the distances are not divided at all.  Pinned by ``tests/golden/loader_*.npz``."""
    deg = np.zeros(n_nodes, dtype=np.int64)
    for i, j in np.asarray(edges).reshape(-1, 2):
        deg[i] += 1
        deg[j] += 1
    return float(size) ** (1 + deg).astype(np.float64), np.ones(len(edges), dtype=np.float64)


def batch_starts(n: int, batch_size: int, epoch: int, seed: Optional[int], shuffle: bool = True, rank: int = 0,
                 world: int = 1, drop_last: bool = False) -> np.ndarray:
    """Start indices of the contiguous batches one rank visits in one epoch — the reference's
    ``_H5BatchIterableDataset.__iter__`` (``deepof/clustering/dataset.py:576-618``): all ranks shuffle the same
    list with ``default_rng(seed + epoch)`` (epoch counted from 1), truncate it to a multiple of ``world`` and
    take ``starts[rank::world]``."""
    bs = int(batch_size)
    starts = np.arange(0, (n // bs) * bs, bs, dtype=np.int64) if drop_last else np.arange(0, n, bs, dtype=np.int64)
    rng = np.random.default_rng(((seed if seed is not None else 0) + epoch) % (2 ** 32))
    if shuffle:
        rng.shuffle(starts)
    if world > 1:
        starts = starts[:(len(starts) // world) * world]
        starts = starts[rank::world]
    return starts


@dataclass
class VideoConstants:
    """Constants of one video's affine column maps (what the reference's scaler objects hold)."""
    size: float
    speed_div: np.ndarray
    dist_div: np.ndarray
    speed_mean1: float = 0.0
    speed_std1: float = 1.0
    dist_mean1: float = 0.0
    dist_std1: float = 1.0


@dataclass
class GlobalScalers:
    """The groupwise global StandardScalers of ``_pp_fit_global_scaler`` (``deepof/utils.py:2796-2863``)."""
    speed_mean: float = 0.0
    speed_std: float = 1.0
    dist_mean: float = 0.0
    dist_std: float = 1.0
    coord_mean: float = 0.0
    coord_std: float = 1.0


def _safe_std(v: float) -> float:
    return v if v > 0.0 else 1.0   # sklearn's _handle_zeros_in_scale


class WindowLoader:
    """Windows of several videos, concatenated in video order like the reference's window store.

    frames: one ``[n_frames, N, 2]`` float32 array / tensor per video (raw x, y per sorted body part).
    edges:  ``[E, 2]`` node index pairs in the reference's sorted edge order (``data.py:2791-2793``).
    nose / tail_base: the size reference of ``scale_table``.  ``global_scalers=None`` fits the global scalers
    on every row of every video (the reference fits them on a random row sample, ``utils.py:2665-2793`` —
    pass its values to reproduce a given run).
    """

    def __init__(self, frames: Sequence, edges, window_size: int, window_step: int = 1, *, nose: int,
                 tail_base: int, center_node: int = -1, align_node: int = -1, arena_center=(0.0, 0.0),
                 fps: float = 25.0, clip: float = 10.0, global_scalers: Optional[GlobalScalers] = None,
                 video_constants: Optional[List[VideoConstants]] = None, reference_quirks: bool = True,
                 device: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("deepof_b200.WindowLoader needs a CUDA device; there is no CPU path")
        self._lib = lib()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.T, self.step = int(window_size), int(window_step)
        self.edges = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
        self.E = int(self.edges.shape[0])
        self.frames = [torch.as_tensor(np.asarray(f) if not torch.is_tensor(f) else f, dtype=torch.float32)
                       .to(self.device).contiguous() for f in frames]
        self.N = int(self.frames[0].shape[1])
        assert all(f.dim() == 3 and f.shape[1] == self.N and f.shape[2] == 2 for f in self.frames)
        self.nose, self.tail_base = int(nose), int(tail_base)
        self.center_node, self.align_node = int(center_node), int(align_node)
        self.cx, self.cy = float(arena_center[0]), float(arena_center[1])
        self.fps, self.clip = float(fps), float(clip)
        self.reference_quirks = bool(reference_quirks)
        self.n_windows_per_video = [int(self._lib.dof_loader_num_windows(int(f.shape[0]), self.T, self.step))
                                    for f in self.frames]
        self.window_offsets = np.concatenate([[0], np.cumsum(self.n_windows_per_video)]).astype(np.int64)
        self.n_windows = int(self.window_offsets[-1])
        self._keep = []    # host arrays referenced by the ctypes structs
        self.video_constants = video_constants if video_constants is not None else [self._fit_video(v) for v in range(len(self.frames))]
        self.global_scalers = global_scalers if global_scalers is not None else self._fit_global()
        self._cfgs = [self._make_cfg(v, self.video_constants[v], self.global_scalers, self.clip) for v in range(len(self.frames))]

    # ---- affine column maps -----------------------------------------------------------------
    def _make_cfg(self, v: int, vc: VideoConstants, gs: GlobalScalers, clip: float) -> DofLoaderCfg:
        s1, d1 = _safe_std(vc.speed_std1), _safe_std(vc.dist_std1)
        s2, d2, c2 = _safe_std(gs.speed_std), _safe_std(gs.dist_std), _safe_std(gs.coord_std)
        sp_scale = np.ascontiguousarray(1.0 / (np.asarray(vc.speed_div, np.float64) * s1 * s2))
        sp_shift = np.full(self.N, -(vc.speed_mean1 / s1 + gs.speed_mean) / s2, dtype=np.float64)
        d_div = np.ascontiguousarray(np.asarray(vc.dist_div, np.float64))
        d_scale = np.full(self.E, 1.0 / (d1 * d2), dtype=np.float64)
        d_shift = np.full(self.E, -(vc.dist_mean1 / d1 + gs.dist_mean) / d2, dtype=np.float64)
        self._keep += [sp_scale, sp_shift, d_div, d_scale, d_shift]
        return DofLoaderCfg(self.T, self.step, self.N, self.E, self.center_node, self.align_node, self.cx, self.cy,
                            self.fps, clip, 1.0 / (vc.size * c2), -gs.coord_mean / c2, _dptr(sp_scale), _dptr(sp_shift),
                            _dptr(d_div), _dptr(d_scale), _dptr(d_shift), self.edges.ctypes.data_as(C.POINTER(C.c_int)))

    def _moments(self, cfg: DofLoaderCfg, v: int, shift, out: torch.Tensor):
        sh = (C.c_double * 3)(*shift)
        f = self.frames[v]
        check(self._lib.dof_loader_moments(C.byref(cfg), ptr(f), int(f.shape[0]), sh, ptr(out), _stream()))

    @staticmethod
    def _mean_std(acc0: np.ndarray, acc1: np.ndarray, g: int, shift: float) -> Tuple[float, float]:
        """acc0: moments about 0 (mean), acc1: moments about that mean (population std, like sklearn)."""
        n = acc0[3 * g]
        if n <= 0:
            return 0.0, 1.0
        d = acc1[3 * g + 1] / n
        return float(shift + d), float(np.sqrt(max(acc1[3 * g + 2] / n - d * d, 0.0)))

    def _two_pass(self, cfgs, videos) -> List[Tuple[float, float]]:
        """(mean, std) per column group over `videos`, two passes so the variance is centred."""
        acc = torch.zeros(9, dtype=torch.float64, device=self.device)
        for c, v in zip(cfgs, videos):
            self._moments(c, v, (0.0, 0.0, 0.0), acc)
        a0 = acc.cpu().numpy()
        mean = [a0[3 * g + 1] / a0[3 * g] if a0[3 * g] > 0 else 0.0 for g in range(3)]
        acc.zero_()
        for c, v in zip(cfgs, videos):
            self._moments(c, v, mean, acc)
        a1 = acc.cpu().numpy()
        return [self._mean_std(a0, a1, g, mean[g]) for g in range(3)]

    def _fit_video(self, v: int) -> VideoConstants:
        """Per-video constants of ``scale_table`` (``deepof/utils.py:2478-2489, 2547-2563``)."""
        f = self.frames[v]
        nf = int(f.shape[0])
        ln = torch.empty(max(nf, 1), dtype=torch.float64, device=self.device)
        check(self._lib.dof_loader_pair_length(ptr(f), nf, self.N, self.nose, self.tail_base, ptr(ln), _stream()))
        ln = ln[:nf]
        ln = ln[~torch.isnan(ln)]
        if ln.numel():                                   # np.nanmedian: midpoint of the two middle values
            srt = torch.sort(ln).values
            size = float(0.5 * (srt[(srt.numel() - 1) // 2] + srt[srt.numel() // 2]))
        else:
            size = float("nan")
        if not (np.isfinite(size) and size > 0.0):
            size = 1.0
        sd, dd = reference_divisors(size, self.edges, self.N) if self.reference_quirks else \
            (np.full(self.N, size), np.full(self.E, size))
        vc = VideoConstants(size=size, speed_div=sd, dist_div=dd)
        ident = GlobalScalers()
        (_, _), (sm, ss), (dm, ds) = self._two_pass([self._make_cfg(v, vc, ident, 0.0)], [v])
        vc.speed_mean1, vc.speed_std1, vc.dist_mean1, vc.dist_std1 = sm, ss, dm, ds
        return vc

    def _fit_global(self) -> GlobalScalers:
        ident = GlobalScalers()
        vids = list(range(len(self.frames)))
        cfgs = [self._make_cfg(v, self.video_constants[v], ident, 0.0) for v in vids]
        (cm, cs), (sm, ss), (dm, ds) = self._two_pass(cfgs, vids)
        return GlobalScalers(speed_mean=sm, speed_std=ss, dist_mean=dm, dist_std=ds, coord_mean=cm, coord_std=cs)

    # ---- windows ----------------------------------------------------------------------------
    def __len__(self) -> int:
        return self.n_windows

    def load(self, start: int, count: int, out_x: Optional[torch.Tensor] = None, out_a: Optional[torch.Tensor] = None):
        """x [count,T,N,3], a [count,T,E,1] for windows start .. start+count-1 of the concatenated store
        (the slice ``X[s:e], A[s:e]`` of ``dataset.py:630-640``).  Enqueued on the current stream."""
        if start < 0 or count < 0 or start + count > self.n_windows:
            raise IndexError(f"windows [{start}, {start + count}) outside [0, {self.n_windows})")
        x = out_x if out_x is not None else torch.empty(count, self.T, self.N, 3, device=self.device)
        a = out_a if out_a is not None else torch.empty(count, self.T, self.E, 1, device=self.device)
        assert x.is_contiguous() and a.is_contiguous() and x.shape[0] >= count and a.shape[0] >= count
        v = int(np.searchsorted(self.window_offsets, start, side="right") - 1)
        done = 0
        while done < count:
            w0 = start + done - int(self.window_offsets[v])
            n = min(count - done, self.n_windows_per_video[v] - w0)
            if n > 0:
                f = self.frames[v]
                check(self._lib.dof_load_windows(C.byref(self._cfgs[v]), ptr(f), int(f.shape[0]), int(w0), int(n),
                                                 ptr(x[done:]), ptr(a[done:]), _stream()))
                done += n
            v += 1
        return x[:count], a[:count]

    def video_index(self, start: int, count: int) -> np.ndarray:
        """``video_idx[s:e]`` of the reference store (``dataset.py:255-256``)."""
        idx = np.arange(start, start + count)
        return (np.searchsorted(self.window_offsets, idx, side="right") - 1).astype(np.int64)

    def epoch(self, batch_size: int, epoch: int, seed: Optional[int] = None, shuffle: bool = True, rank: int = 0,
              world: int = 1, drop_last: bool = False):
        """Yield ``(x, a, idx)`` batches of one epoch for this rank, like the reference's DataLoader."""
        for s in batch_starts(self.n_windows, batch_size, epoch, seed, shuffle, rank, world, drop_last):
            n = int(min(batch_size, self.n_windows - s))
            x, a = self.load(int(s), n)
            yield x, a, torch.arange(int(s), int(s) + n, device=self.device)
