"""CPU tests of the loader oracle against the reference-generated goldens, and of the host-side
batch-start logic (reference clustering/dataset.py:576-618)."""
import glob
import os

import numpy as np
import pytest

from oracle import loader_oracle as LO

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loader_*.npz")))


def cfg_from_golden(g, use_stats=None):
    c = LO.LoaderCfg(T=int(g["T"]), step=int(g["step"]), center_node=int(g["center_node"]), align_node=int(g["align_node"]),
                     cx=float(g["cx"]), cy=float(g["cy"]), fps=float(g["fps"]), clip=float(g["clip"]))
    for k in ("size", "speed_mean1", "speed_std1", "dist_mean1", "dist_std1", "speed_mean2", "speed_std2", "dist_mean2",
              "dist_std2", "coord_mean2", "coord_std2"):
        setattr(c, k, float(g["c_" + k]))
    c.speed_div, c.dist_div = g["speed_div"], g["dist_div"]
    if use_stats:
        for k, v in use_stats.items():
            setattr(c, k, v)
    return c


def test_goldens_present():
    assert len(GOLD) >= 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_windows(path):
    g = np.load(path)
    cfg = cfg_from_golden(g)
    x, a = LO.load_windows(g["frames"], g["edges"], cfg)
    assert x.shape == g["x"].shape and a.shape == g["a"].shape
    # float64 restatement vs float64 reference, both cast to fp32: tolerance = fp32 rounding
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(a, g["a"], rtol=0, atol=2e-6)
    # invariants the reference's own tests pin (tests/test_utils.py:199-236, :496-548)
    assert x.shape[0] == (g["frames"].shape[0] - cfg.T) // cfg.step + 1
    if int(g["n_clipped"]) == 0:
        al = int(g["align_node"])
        xa = x[..., al, 0] * cfg.coord_std2 + cfg.coord_mean2      # aligned x of the align node == 0
        assert np.abs(xa).max() < 1e-5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_video_stats_match_reference_scalers(path):
    g = np.load(path)
    cfg = cfg_from_golden(g)
    st = LO.video_stats(g["frames"], g["edges"], cfg, int(g["nose"]), int(g["tail_base"]))
    for k, v in st.items():
        assert abs(v - float(g["c_" + k])) <= 1e-9 * max(1.0, abs(v)), k
    # the divisors the reference applied (read back from its output) are the documented quirk
    sd, dd = LO.reference_divisors(st["size"], g["edges"], g["frames"].shape[1])
    np.testing.assert_allclose(sd, g["speed_div"], rtol=1e-9)
    np.testing.assert_allclose(dd, g["dist_div"], rtol=1e-9)


def test_partial_batches_equal_full_table():
    g = np.load([p for p in GOLD if p.endswith("loader_w25.npz")][0])
    cfg = cfg_from_golden(g)
    x, a = LO.load_windows(g["frames"], g["edges"], cfg)
    xs, as_ = LO.load_windows(g["frames"], g["edges"], cfg, start=17, count=40)
    np.testing.assert_array_equal(xs, x[17:57])
    np.testing.assert_array_equal(as_, a[17:57])


def test_batch_starts_sharding():
    n, bs = 1000, 64
    full = LO.batch_starts(n, bs, epoch=1, seed=7)
    assert sorted(full.tolist()) == list(range(0, n, bs))
    parts = [LO.batch_starts(n, bs, epoch=1, seed=7, rank=r, world=3) for r in range(3)]
    assert all(len(p) == len(full) // 3 for p in parts)
    inter = np.stack(parts, 1).reshape(-1)
    np.testing.assert_array_equal(inter, full[:len(inter)])          # starts[rank::world] of the same shuffle
    assert not np.array_equal(full, LO.batch_starts(n, bs, epoch=2, seed=7))
    np.testing.assert_array_equal(LO.batch_starts(n, bs, epoch=1, seed=None, shuffle=False), np.arange(0, n, bs))
    assert len(LO.batch_starts(n, bs, epoch=1, seed=0, drop_last=True)) == n // bs
