"""Golden vectors for the window loader, produced by the UNMODIFIED reference functions.

Run in the build container only (needs /root/reference):

    python tests/golden/make_loader_golden.py

``deepof.utils`` imports here once its non-arithmetic dependencies (matplotlib, cv2, shapely,
sleap_io, segment_anything, h5py, deepof.data ...) are stubbed; every arithmetic step below is the
reference's own function, called on a synthetic single-animal pose table:

  utils.align_trajectories(mode="all")  (utils.py:2097-2142, rotate :1298-1319)
  utils.rolling_speed                   (utils.py:3788-3857)  speeds from the raw coordinates
  utils.compute_dist                    (utils.py:863-881)    edge lengths from the raw coordinates
  utils.scale_table                     (utils.py:2425-2566)  size-normalise, log1p, per-video scalers
  utils._pp_apply_global                (utils.py:2866-2921)  global scalers (sklearn StandardScaler)
  utils._pp_sanitize_numeric            (utils.py:2577-2583)
  utils.rolling_window                  (utils.py:3354-3377)
  clustering.dataset.reorder_and_reshape (dataset.py:16-26)

Glue that lives in ``deepof/data.py`` (not importable here: pandas plumbing around those calls) is
restated inline and cited: arena centring ``data.py:1848-1850``, align column ordering + |v|<1e-5
zeroing ``data.py:1895-1912``, clip / interpolate ``utils.py:2990-3004``, node / edge column order
``data.py:2791-2833,2877-2880``.  Global scalers are fitted on every row of the per-video-scaled
table (the reference fits them on a random row sample, ``utils.py:2665-2793``; the sample choice is
an input of the loader, not arithmetic).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DEEPOF_REFERENCE", "/root/reference")


class _Stub(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    NAMES = ("matplotlib", "cv2", "h5py", "sleap_io", "segment_anything", "shapely", "deepof.data",
             "deepof.data_loading", "deepof.legacy_smote_handling", "natsort", "duckdb", "IPython", "optuna")

    def find_spec(self, name, path, target=None):
        if any(name == n or name.startswith(n + ".") for n in self.NAMES):
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = MagicMock(name=spec.name)
        m.__path__, m.__name__, m.__spec__, m.__loader__ = [], spec.name, spec, self
        return m

    def exec_module(self, module):
        pass


def load_reference():
    # pandas >= 3 returns read-only views from to_numpy(); the reference (pandas 1.x/2.x era) writes into
    # them (utils.py:2534).  Environment shim only: hand back a writable copy, arithmetic untouched.
    _to_numpy = pd.DataFrame.to_numpy

    def to_numpy_writable(self, *a, **k):
        out = _to_numpy(self, *a, **k)
        return out if out.flags.writeable else np.array(out, copy=True)

    pd.DataFrame.to_numpy = to_numpy_writable
    sys.meta_path.insert(0, _Stub())
    pkg = types.ModuleType("deepof")
    pkg.__path__ = [os.path.join(REF, "deepof")]
    sys.modules["deepof"] = pkg
    import deepof.utils as U
    import deepof.clustering.dataset as DS
    return U, DS


BODYPARTS = ["Center", "Left_bhip", "Left_ear", "Left_fhip", "Nose", "Right_bhip", "Right_ear", "Right_fhip",
             "Spine_1", "Spine_2", "Tail_1", "Tail_2", "Tail_base", "Tail_tip"]   # sorted, deepof_14-like


def chain_adjacency(n):
    A = np.zeros((n, n))
    for i in range(n - 1):
        A[i, i + 1] = A[i + 1, i] = 1.0
    if n > 5:
        A[0, 5] = A[5, 0] = 1.0
    return A


def synth_frames(n_frames, n_nodes, seed, outliers=True):
    """Mouse-like random walk in a 400 px arena, fp32."""
    rng = np.random.default_rng(seed)
    centre = np.cumsum(rng.normal(0, 1.5, size=(n_frames, 2)), axis=0) + 200.0
    heading = np.cumsum(rng.normal(0, 0.08, size=n_frames))
    body = rng.normal(0, 18.0, size=(n_nodes, 2))
    body[4] = [0.0, 42.0]      # Nose
    body[12] = [0.0, -38.0]    # Tail_base
    body[0] = [0.0, 0.0]       # Center
    c, s = np.cos(heading), np.sin(heading)
    rot = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)      # [F,2,2]
    pts = np.einsum("fij,nj->fni", rot, body) + centre[:, None, :] + rng.normal(0, 0.4, size=(n_frames, n_nodes, 2))
    if outliers:   # tracking jumps -> |z| > 10 -> clipped and interpolated by the reference
        for f, n in ((37, 3), (38, 3), (90, 9), (5, 1), (n_frames - 2, 7)):
            pts[f, n] += rng.normal(0, 1.0, size=2) * 1500.0
    return pts.astype(np.float32)


def run_case(U, DS, name, n_frames, T, step, seed, align="Center", outliers=True):
    aid = "A"
    names = [f"{aid}_{b}" for b in BODYPARTS]
    N = len(names)
    adj = chain_adjacency(N)
    edges = [(i, j) for i in range(N) for j in range(i + 1, N) if adj[i, j] != 0]
    frames = synth_frames(n_frames, N, seed, outliers)
    fps = 25.0
    cx, cy = 200.0, 200.0
    cols = pd.MultiIndex.from_tuples([(n, ax) for n in names for ax in ("x", "y")])
    raw = pd.DataFrame(frames.reshape(n_frames, 2 * N).astype(np.float64), columns=cols)

    # --- coords: centre on the arena (data.py:1848-1850), align on `align` (data.py:1878-1928)
    tab = raw.copy()
    tab.loc[:, (slice(None), ["x"])] -= cx
    tab.loc[:, (slice(None), ["y"])] -= cy
    align_bp = f"{aid}_{align}"
    align_cols = [(align_bp, "x"), (align_bp, "y")]
    other_cols = [c for c in tab.columns if c[0].startswith(aid) and c[0] != align_bp]
    ordered = align_cols + other_cols
    aligned = U.align_trajectories(np.array(tab[ordered]), mode="all", run_numba=False)
    aligned[np.abs(aligned) < 1e-5] = 0.0
    coords = pd.DataFrame(aligned, columns=pd.MultiIndex.from_tuples(ordered))

    # --- speeds from the raw table (data.py:2725 -> rolling_speed)
    speeds = U.rolling_speed(raw.copy(), frame_rate=fps, deriv=1, typ="coords")
    assert list(speeds.columns) == names

    # --- edge lengths from the raw table (compute_dist on [p_i | p_j])
    dist = {}
    for (i, j) in edges:
        pair = np.concatenate([frames[:, i].astype(np.float64), frames[:, j].astype(np.float64)], axis=1)
        dist[(names[i], names[j])] = np.asarray(U.compute_dist(pair)).reshape(-1)
    dists = pd.DataFrame(dist)

    merged = pd.concat([coords, speeds, dists], axis=1)
    orig_cols = merged.columns

    # --- per-video scaling exactly as _pp_pass2_scale_and_save calls it (utils.py:2963-2974)
    kw = dict(scale="standard", animal_ids=[aid], dist_standardize="groupwise", speed_standardize="groupwise",
              log_distances=True)
    tab_local = U.scale_table(merged, standardize=True, coord_standardize=None, **kw)

    # --- global scalers (legacy dict of sklearn scalers), fitted on all rows
    from sklearn.preprocessing import StandardScaler
    ct = U.infer_column_types(tab_local)
    gs = {}
    gs["speed"] = StandardScaler().fit(tab_local[ct["speeds"]].to_numpy(float).reshape(-1, 1))
    gs["dist_inner"] = StandardScaler().fit(tab_local[ct["inner_dists"]].to_numpy(float).reshape(-1, 1))
    gs["dist_intra"] = None
    gs["coord"] = StandardScaler().fit(tab_local[ct["coords"]].to_numpy(float).reshape(-1, 1))
    tab2 = U._pp_apply_global(tab_local.copy(), speed_standardize="groupwise", dist_standardize="groupwise",
                              coord_standardize="groupwise", global_scaler=gs)

    # --- clip outliers and interpolate (utils.py:2990-3004), sanitize (:3018)
    clip = 10
    scalars = [c for c in ct["scalars"] if c in tab2.columns]
    coord_cols_clip = [c for c in tab2.columns if isinstance(c, tuple) and len(c) == 2 and c[1] in ("x", "y")]
    clip_cols = list(dict.fromkeys(scalars + coord_cols_clip))
    arr = tab2[clip_cols].to_numpy(float)
    n_clipped = int((np.abs(arr) > clip).sum())
    arr[np.abs(arr) > clip] = np.nan
    tab2[clip_cols] = pd.DataFrame(arr, index=tab2.index, columns=clip_cols).interpolate(limit_direction="both")
    tab2 = tab2.reindex(columns=orig_cols)
    final = U._pp_sanitize_numeric(tab2)

    # --- windows + node / edge column order (data.py:2791-2833, 2877-2880) + dataset reshape
    win = U.rolling_window(final.to_numpy(float), T, step)
    node_cols = [(n, "x") for n in names] + [(n, "y") for n in names] + names
    feat = list(final.columns)
    node_idx = [feat.index(c) for c in node_cols]
    edge_idx = [feat.index((names[i], names[j])) for (i, j) in edges]
    x = DS.reorder_and_reshape(np.ascontiguousarray(win[:, :, node_idx])).astype(np.float32)
    a = np.ascontiguousarray(win[:, :, edge_idx])[..., None].astype(np.float32)

    # --- the constants the loader takes, READ BACK from the reference's own tables (nothing assumed):
    # stage 1 of scale_table (standardize=False) gives the per-column divisors, stage 2 the groupwise scalers.
    # NB on a merged (object-Index) table `out.loc[:, (bp1, bp2)]` (utils.py:2525) selects the two SPEED
    # columns bp1, bp2 (a tuple is a list of labels for a flat Index), so the reference divides each speed
    # column by size once more per incident edge and leaves the distances un-normalised.  The divisors below
    # capture whatever the reference did.
    t1 = U.scale_table(merged, standardize=False, coord_standardize=None, **kw)
    size = float(np.nanmedian(np.hypot(merged[(f"{aid}_Nose", "x")] - merged[(f"{aid}_Tail_base", "x")],
                                       merged[(f"{aid}_Nose", "y")] - merged[(f"{aid}_Tail_base", "y")])))
    with np.errstate(invalid="ignore", divide="ignore"):
        speed_div = np.nanmedian(merged[names].to_numpy(float) / t1[names].to_numpy(float), axis=0)
        dist_div = np.nanmedian(dists.to_numpy(float) / np.expm1(t1[list(dists.columns)].to_numpy(float)), axis=0)
        coord_div = np.nanmedian(np.abs(merged[ordered].to_numpy(float) / t1[ordered].to_numpy(float)))
    assert abs(coord_div - size) < 1e-9 * size
    spn = t1[names].to_numpy(float)
    dn = t1[list(dists.columns)].to_numpy(float)
    consts = dict(size=size, speed_mean1=float(np.nanmean(spn)), speed_std1=float(np.nanstd(spn)),
                  dist_mean1=float(dn.mean()), dist_std1=float(dn.std()),
                  speed_mean2=float(gs["speed"].mean_[0]), speed_std2=float(gs["speed"].scale_[0]),
                  dist_mean2=float(gs["dist_inner"].mean_[0]), dist_std2=float(gs["dist_inner"].scale_[0]),
                  coord_mean2=float(gs["coord"].mean_[0]), coord_std2=float(gs["coord"].scale_[0]))
    # the groupwise scalers of stage 2 are exactly (t1 - mean) / std
    chk = (spn - consts["speed_mean1"]) / consts["speed_std1"]
    assert np.nanmax(np.abs(chk - tab_local[names].to_numpy(float))) < 1e-9
    chk = (dn - consts["dist_mean1"]) / consts["dist_std1"]
    assert np.abs(chk - tab_local[list(dists.columns)].to_numpy(float)).max() < 1e-9
    out = os.path.join(HERE, f"loader_{name}.npz")
    np.savez_compressed(out, frames=frames, adjacency=adj, edges=np.asarray(edges, np.int32), T=T, step=step, fps=fps,
                        cx=cx, cy=cy, center_node=-1, align_node=BODYPARTS.index(align), nose=BODYPARTS.index("Nose"),
                        tail_base=BODYPARTS.index("Tail_base"), clip=float(clip), x=x, a=a, n_clipped=n_clipped,
                        speed_div=speed_div, dist_div=dist_div,
                        **{"c_" + k: v for k, v in consts.items()})
    print(f"{name}: frames {frames.shape} -> x {x.shape} a {a.shape}, clipped {n_clipped}, size {size:.3f} "
          f"({os.path.getsize(out) / 1024:.0f} KB)")


if __name__ == "__main__":
    U, DS = load_reference()
    run_case(U, DS, "w25", n_frames=160, T=25, step=1, seed=21)
    run_case(U, DS, "w24s3", n_frames=131, T=24, step=3, seed=22)
    run_case(U, DS, "clean", n_frames=64, T=25, step=1, seed=23, outliers=False)
