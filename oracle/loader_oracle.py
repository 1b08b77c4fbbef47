"""CPU oracle for the window loader (frames -> model-ready windows) of mlfpm/deepof.

TEST INFRASTRUCTURE — NOT THE PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of ``bench.py`` may import this module.

A numpy (float64) restatement of what the reference does between a per-video pose table and the
``x[Nw,T,N,3]`` / ``a[Nw,T,E,1]`` arrays its window store holds.
Paths are relative to /root/reference/deepof:

  centre            data.py:1844-1869   (arena: subtract (cx, cy); body part: subtract that node)
  align             data.py:1878-1928 -> utils.py:2097-2142 (mode="all"), rotate utils.py:1298-1319,
                    |v| < 1e-5 -> 0     data.py:1912
  speed             utils.py:3788-3857  rolling_speed(window=3, shift=2, rounds=3) on the RAW coordinates
  edge length       utils.py:863-881    compute_dist on the RAW coordinates
  size-normalise,   utils.py:2425-2566  scale_table: / median |Nose - Tail_base|, log1p(dist),
  per-video scalers                     groupwise StandardScaler for speeds and inner distances
  global scalers    utils.py:2866-2921  _pp_apply_global (groupwise speed / dist_inner / coord)
  clip + interp     utils.py:2990-3004  |z| > 10 -> NaN -> linear interpolation, both directions
  sanitize          utils.py:2577-2583  interpolate + fillna(0)
  windows           utils.py:3354-3377  rolling_window(window_size, window_step)
  layout            data.py:2791-2833, 2877-2880 (sorted nodes: x.. | y.. | speed..), clustering/dataset.py:16-26
  batch starts      clustering/dataset.py:576-618 (contiguous batches, epoch-seeded shuffle, rank sharding)

Parity pinning: ``tests/golden/loader_*.npz`` are produced by the UNMODIFIED reference functions
(``tests/golden/make_loader_golden.py``); ``tests/test_loader_cpu.py`` checks this oracle against
every one of them.  Parity is therefore pinned for the arithmetic; the pandas glue of data.py and
the random row sample of the global-scaler fit are inputs here (see the generator's docstring).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np


@dataclass
class LoaderCfg:
    T: int
    step: int = 1
    center_node: int = -1          # -1: arena centre (cx, cy)
    align_node: int = -1           # -1: no alignment
    cx: float = 0.0
    cy: float = 0.0
    fps: float = 25.0
    size: float = 1.0              # per-video median |Nose - Tail_base|
    speed_mean1: float = 0.0       # per-video groupwise scaler (scale_table)
    speed_std1: float = 1.0
    dist_mean1: float = 0.0
    dist_std1: float = 1.0
    speed_mean2: float = 0.0       # global scalers (_pp_apply_global)
    speed_std2: float = 1.0
    dist_mean2: float = 0.0
    dist_std2: float = 1.0
    coord_mean2: float = 0.0
    coord_std2: float = 1.0
    clip: float = 10.0             # interpolate_normalized; <= 0 disables
    speed_div: Optional[np.ndarray] = None   # [N] per-column size divisor of the speeds (None: size)
    dist_div: Optional[np.ndarray] = None    # [E] per-column size divisor of the edge lengths (None: size)


def reference_divisors(size: float, edges: np.ndarray, n_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """The divisors scale_table ACTUALLY applies on the merged (coords | speeds | dists) table the
    reference hands it (utils.py:2963-2974): the speed loop divides every speed column by `size`
    (utils.py:2516-2518), then `out.loc[:, (bp1, bp2)] = ... / s` (utils.py:2520-2526) addresses — on a
    flat object Index a tuple key is a LIST of labels — the two speed columns bp1 and bp2, not the distance
    column.  Net effect (pinned by tests/golden/loader_*.npz, which record the divisors read back from the
    reference's output): speed of node n is divided by size**(1 + degree(n)); distances are not divided."""
    deg = np.zeros(n_nodes, dtype=np.int64)
    for i, j in np.asarray(edges).reshape(-1, 2):
        deg[i] += 1
        deg[j] += 1
    return float(size) ** (1 + deg).astype(np.float64), np.ones(len(edges), dtype=np.float64)


def _divs(cfg: "LoaderCfg", n_nodes: int, n_edges: int) -> Tuple[np.ndarray, np.ndarray]:
    sd = np.full(n_nodes, cfg.size) if cfg.speed_div is None else np.asarray(cfg.speed_div, dtype=np.float64)
    dd = np.full(n_edges, cfg.size) if cfg.dist_div is None else np.asarray(cfg.dist_div, dtype=np.float64)
    return sd, dd


def aligned_coords(frames: np.ndarray, cfg: LoaderCfg) -> np.ndarray:
    """[F,N,2] raw -> centred, aligned coordinates (before size normalisation)."""
    p = np.asarray(frames, dtype=np.float64)
    if cfg.center_node >= 0:
        c = p - p[:, cfg.center_node:cfg.center_node + 1, :]
    else:
        c = p - np.array([cfg.cx, cfg.cy])
    if cfg.align_node >= 0:
        ang = np.arctan2(c[:, cfg.align_node, 0], c[:, cfg.align_node, 1])   # utils.py:2125
        cs, sn = np.cos(ang)[:, None], np.sin(ang)[:, None]
        x = cs * c[..., 0] - sn * c[..., 1]                                    # utils.py:1313-1317
        y = sn * c[..., 0] + cs * c[..., 1]
        c = np.stack([x, y], -1)
        c[np.abs(c) < 1e-5] = 0.0                                              # data.py:1912
    return c


def raw_speed(frames: np.ndarray, fps: float) -> np.ndarray:
    """[F,N]: rolling_speed(window=3, shift=2, rounds=3) * frame_rate; NaN for the first 4 frames."""
    p = np.asarray(frames, dtype=np.float64)
    F = p.shape[0]
    d = np.full(p.shape[:2], np.nan)
    d[2:] = np.sqrt((((p[2:] / 2.0) - (p[:-2] / 2.0)) ** 2).sum(-1))          # utils.py:3826-3838
    sp = np.full(p.shape[:2], np.nan)
    if F >= 5:
        sp[4:] = (d[2:-2] + d[3:-1] + d[4:]) / 3.0                             # rolling(3).mean()
    return np.round(sp, 3) * fps                                               # utils.py:3848-3855


def raw_dist(frames: np.ndarray, edges: np.ndarray) -> np.ndarray:
    p = np.asarray(frames, dtype=np.float64)
    ab = p[:, edges[:, 0]] - p[:, edges[:, 1]]
    return np.sqrt((ab * ab).sum(-1))                                          # utils.py:877-880


def video_stats(frames: np.ndarray, edges: np.ndarray, cfg: LoaderCfg, nose: int, tail_base: int,
                quirks: bool = True) -> dict:
    """Per-video constants of scale_table: size factor and the groupwise scalers' mean / std
    (population std, NaNs ignored — sklearn StandardScaler).  quirks=True applies reference_divisors()."""
    p = np.asarray(frames, dtype=np.float64)
    size = float(np.nanmedian(np.hypot(p[:, nose, 0] - p[:, tail_base, 0], p[:, nose, 1] - p[:, tail_base, 1])))
    sd, dd = reference_divisors(size, edges, p.shape[1]) if quirks else (size, size)
    sp = raw_speed(p, cfg.fps) / sd
    dn = np.log1p(np.clip(raw_dist(p, edges) / dd, 0.0, None))
    return dict(size=size, speed_mean1=float(np.nanmean(sp)), speed_std1=float(np.nanstd(sp)),
                dist_mean1=float(dn.mean()), dist_std1=float(dn.std()))


def _interp_columns(z: np.ndarray) -> np.ndarray:
    """pandas interpolate(limit_direction="both") column-wise, then fillna(0)."""
    F = z.shape[0]
    idx = np.arange(F)
    flat = z.reshape(F, -1).copy()
    for j in range(flat.shape[1]):
        col = flat[:, j]
        ok = np.isfinite(col)
        if ok.all():
            continue
        flat[:, j] = np.interp(idx, idx[ok], col[ok]) if ok.any() else 0.0
    return flat.reshape(z.shape)


def frame_features(frames: np.ndarray, edges: np.ndarray, cfg: LoaderCfg) -> Tuple[np.ndarray, np.ndarray]:
    """Per-frame standardised features: nodes [F,N,3] = (x, y, speed), edges [F,E]."""
    sd, dd = _divs(cfg, frames.shape[1], len(edges))
    c = aligned_coords(frames, cfg) / cfg.size
    zc = (c - cfg.coord_mean2) / cfg.coord_std2
    sp = raw_speed(frames, cfg.fps) / sd
    zs = ((sp - cfg.speed_mean1) / cfg.speed_std1 - cfg.speed_mean2) / cfg.speed_std2
    dn = np.log1p(np.clip(raw_dist(frames, edges) / dd, 0.0, None))
    zd = ((dn - cfg.dist_mean1) / cfg.dist_std1 - cfg.dist_mean2) / cfg.dist_std2
    if cfg.clip > 0:
        for z in (zc, zs, zd):
            with np.errstate(invalid="ignore"):
                z[np.abs(z) > cfg.clip] = np.nan
    zc, zs, zd = _interp_columns(zc), _interp_columns(zs), _interp_columns(zd)
    return np.concatenate([zc, zs[..., None]], -1), zd


def n_windows(n_frames: int, T: int, step: int) -> int:
    return 0 if n_frames < T else (n_frames - T) // step + 1


def load_windows(frames: np.ndarray, edges: np.ndarray, cfg: LoaderCfg, start: int = 0,
                 count: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """x [B,T,N,3], a [B,T,E,1] (float32) for windows start .. start+count-1."""
    nodes, ed = frame_features(frames, edges, cfg)
    nw = n_windows(frames.shape[0], cfg.T, cfg.step)
    count = nw - start if count is None else count
    assert 0 <= start and start + count <= nw
    f0 = (start + np.arange(count)) * cfg.step
    idx = f0[:, None] + np.arange(cfg.T)[None, :]
    return nodes[idx].astype(np.float32), ed[idx][..., None].astype(np.float32)


def batch_starts(n: int, batch_size: int, epoch: int, seed: Optional[int], shuffle: bool = True, rank: int = 0,
                 world: int = 1, drop_last: bool = False) -> np.ndarray:
    """Contiguous-batch start indices of one epoch for one rank (clustering/dataset.py:576-618).
    `epoch` is 1-based like the reference's per-__iter__ counter."""
    bs = batch_size
    starts = np.arange(0, (n // bs) * bs, bs, dtype=np.int64) if drop_last else np.arange(0, n, bs, dtype=np.int64)
    rng = np.random.default_rng(((seed if seed is not None else 0) + epoch) % (2 ** 32))
    if shuffle:
        rng.shuffle(starts)
    if world > 1:
        starts = starts[:(len(starts) // world) * world]
        starts = starts[rank::world]
    return starts
