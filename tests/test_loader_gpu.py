"""GPU parity of the window-loader kernel (dof_load_windows & friends, through the C-ABI) against
(1) the goldens the UNMODIFIED reference functions produced (tests/golden/loader_*.npz) and
(2) the CPU oracle (oracle/loader_oracle.py) on fresh seeded frame tables, including the edge cases the
reference tests pin: window count, ragged tails, tables shorter than a window, several videos."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import loader_oracle as LO

pytestmark = pytest.mark.gpu

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loader_*.npz")))
# fp64 kernel vs fp64 reference, both rounded to fp32: one fp32 ulp at |z| <= 10, plus libm differences in
# atan2 / sincos / log1p.  The tolerance is absolute on standardised (unit-variance) features.
ATOL = 4e-6


def _consts(g):
    from deepof_b200 import GlobalScalers, VideoConstants
    vc = VideoConstants(size=float(g["c_size"]), speed_div=g["speed_div"], dist_div=g["dist_div"],
                        speed_mean1=float(g["c_speed_mean1"]), speed_std1=float(g["c_speed_std1"]),
                        dist_mean1=float(g["c_dist_mean1"]), dist_std1=float(g["c_dist_std1"]))
    gs = GlobalScalers(speed_mean=float(g["c_speed_mean2"]), speed_std=float(g["c_speed_std2"]),
                       dist_mean=float(g["c_dist_mean2"]), dist_std=float(g["c_dist_std2"]),
                       coord_mean=float(g["c_coord_mean2"]), coord_std=float(g["c_coord_std2"]))
    return vc, gs


def _loader_from_golden(g, fit=False):
    from deepof_b200 import WindowLoader
    vc, gs = _consts(g)
    return WindowLoader([g["frames"]], g["edges"], int(g["T"]), int(g["step"]), nose=int(g["nose"]),
                        tail_base=int(g["tail_base"]), center_node=int(g["center_node"]), align_node=int(g["align_node"]),
                        arena_center=(float(g["cx"]), float(g["cy"])), fps=float(g["fps"]), clip=float(g["clip"]),
                        video_constants=None if fit else [vc], global_scalers=None if fit else gs)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_kernel_matches_reference_goldens(path):
    g = np.load(path)
    ld = _loader_from_golden(g)
    assert len(ld) == g["x"].shape[0] == (g["frames"].shape[0] - int(g["T"])) // int(g["step"]) + 1
    x, a = ld.load(0, len(ld))
    torch.cuda.synchronize()
    np.testing.assert_allclose(x.cpu().numpy(), g["x"], rtol=0, atol=ATOL)
    np.testing.assert_allclose(a.cpu().numpy(), g["a"], rtol=0, atol=ATOL)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_fitted_scalers_match_reference(path):
    """size factor + groupwise scaler statistics computed on the device == the reference's scalers; the global
    scalers fitted on every row are what the golden generator fitted."""
    g = np.load(path)
    ld = _loader_from_golden(g, fit=True)
    vc, gs = ld.video_constants[0], ld.global_scalers
    ref_vc, ref_gs = _consts(g)
    for k in ("size", "speed_mean1", "speed_std1", "dist_mean1", "dist_std1"):
        assert abs(getattr(vc, k) - getattr(ref_vc, k)) <= 1e-9 * max(1.0, abs(getattr(ref_vc, k))), k
    np.testing.assert_allclose(vc.speed_div, ref_vc.speed_div, rtol=1e-12)
    np.testing.assert_allclose(vc.dist_div, ref_vc.dist_div, rtol=1e-12)
    for k in ("speed_mean", "speed_std", "dist_mean", "dist_std", "coord_mean", "coord_std"):
        assert abs(getattr(gs, k) - getattr(ref_gs, k)) <= 1e-8 * max(1.0, abs(getattr(ref_gs, k))), k
    x, a = ld.load(0, len(ld))
    np.testing.assert_allclose(x.cpu().numpy(), g["x"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(a.cpu().numpy(), g["a"], rtol=0, atol=2e-5)


def _synthetic(n_frames, N, seed, jumps=()):
    rng = np.random.default_rng(seed)
    centre = np.cumsum(rng.normal(0, 1.5, size=(n_frames, 2)), axis=0) + 250.0
    heading = np.cumsum(rng.normal(0, 0.08, size=n_frames))
    body = rng.normal(0, 18.0, size=(N, 2))
    body[1] = [0.0, 40.0]
    body[2] = [0.0, -36.0]
    c, s = np.cos(heading), np.sin(heading)
    rot = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)
    pts = np.einsum("fij,nj->fni", rot, body) + centre[:, None, :] + rng.normal(0, 0.4, size=(n_frames, N, 2))
    for f, n in jumps:
        pts[f, n] += 2000.0
    return pts.astype(np.float32)


def _edges(N):
    e = [(i, i + 1) for i in range(N - 1)] + ([(0, N - 1)] if N > 3 else [])
    return np.asarray(sorted(e), dtype=np.int32)


def _oracle_cfg(ld, v):
    vc, gs = ld.video_constants[v], ld.global_scalers
    return LO.LoaderCfg(T=ld.T, step=ld.step, center_node=ld.center_node, align_node=ld.align_node, cx=ld.cx, cy=ld.cy,
                        fps=ld.fps, size=vc.size, speed_mean1=vc.speed_mean1, speed_std1=vc.speed_std1,
                        dist_mean1=vc.dist_mean1, dist_std1=vc.dist_std1, speed_mean2=gs.speed_mean,
                        speed_std2=gs.speed_std, dist_mean2=gs.dist_mean, dist_std2=gs.dist_std,
                        coord_mean2=gs.coord_mean, coord_std2=gs.coord_std, clip=ld.clip, speed_div=vc.speed_div,
                        dist_div=vc.dist_div)


@pytest.mark.parametrize("N,T,step,center,align,quirks", [
    (14, 25, 1, -1, 0, True),      # cfg1/cfg2 geometry
    (11, 24, 3, 0, 1, True),       # 88-byte rows (unaligned slabs), body-part centring, step 3
    (11, 50, 1, -1, -1, False),    # contrastive window length, no alignment, plain size normalisation
    (5, 7, 2, -1, 3, True),
])
def test_kernel_matches_oracle_multi_video(N, T, step, center, align, quirks):
    from deepof_b200 import WindowLoader
    edges = _edges(N)
    # jumps at the very start / end of a video make the interpolation search leave the CTA tile
    vids = [_synthetic(400, N, 1, jumps=((0, 0), (1, 0), (2, 0), (398, 2), (399, 2), (120, 1), (121, 1), (122, 1))),
            _synthetic(T + 3, N, 2), _synthetic(T - 1, N, 3), _synthetic(333, N, 4, jumps=((50, 3),))]
    ld = WindowLoader(vids, edges, T, step, nose=1, tail_base=2, center_node=center, align_node=align,
                      arena_center=(250.0, 250.0), fps=30.0, reference_quirks=quirks)
    counts = [LO.n_windows(v.shape[0], T, step) for v in vids]
    assert ld.n_windows_per_video == counts and counts[2] == 0
    xs, as_ = [], []
    for v, fr in enumerate(vids):
        if counts[v]:
            x, a = LO.load_windows(fr, edges, _oracle_cfg(ld, v))
            xs.append(x)
            as_.append(a)
    xo, ao = np.concatenate(xs), np.concatenate(as_)
    x, a = ld.load(0, len(ld))
    np.testing.assert_allclose(x.cpu().numpy(), xo, rtol=0, atol=ATOL)
    np.testing.assert_allclose(a.cpu().numpy(), ao, rtol=0, atol=ATOL)
    # ragged slices across video boundaries, counts that are not multiples of the CTA tile (plain-store path)
    for s, n in ((0, 1), (3, 5), (counts[0] - 2, 7), (len(ld) - 9, 9), (17, 130)):
        n = min(n, len(ld) - s)
        xs_, as2 = ld.load(s, n)
        assert torch.equal(xs_, x[s:s + n]) and torch.equal(as2, a[s:s + n])
    assert np.array_equal(ld.video_index(counts[0] - 1, 3), [0, 1, 1])
    with pytest.raises(IndexError):
        ld.load(len(ld) - 1, 2)


def test_unfitted_statistics_equal_oracle_video_stats():
    from deepof_b200 import WindowLoader
    N, T = 14, 25
    edges = _edges(N)
    fr = _synthetic(5000, N, 9, jumps=((77, 4),))
    ld = WindowLoader([fr], edges, T, nose=1, tail_base=2, align_node=0, arena_center=(250.0, 250.0), fps=25.0)
    st = LO.video_stats(fr, edges, LO.LoaderCfg(T=T, fps=25.0), 1, 2)
    vc = ld.video_constants[0]
    for k, v in st.items():
        assert abs(getattr(vc, k) - v) <= 1e-9 * max(1.0, abs(v)), k


def test_full_size_sliding_window_properties():
    """BASELINE cfg2 size: 1M windows from one frame table.  Size-independent properties: window w at step t
    equals window w+1 at step t-1 (they are the same frame), and a strided checksum over all windows equals
    the checksum of the per-frame features counted with their multiplicity."""
    from deepof_b200 import WindowLoader
    N, T, nW, B = 14, 25, 1 << 20, 4096
    edges = _edges(N)
    fr = _synthetic(nW + T - 1, N, 11)
    ld = WindowLoader([fr], edges, T, nose=1, tail_base=2, align_node=0, arena_center=(250.0, 250.0))
    assert len(ld) == nW
    xbuf = torch.empty(B, T, N, 3, device="cuda")
    abuf = torch.empty(B, T, N, 1, device="cuda")
    for s in range(0, nW, B * 16):                      # every 16th batch
        x, a = ld.load(s, B, xbuf, abuf)
        assert torch.equal(x[:-1, 1:], x[1:, :-1]) and torch.equal(a[:-1, 1:], a[1:, :-1])
        assert torch.isfinite(x).all() and torch.isfinite(a).all()
        assert float(x.abs().max()) <= 10.0 + 1e-6
        # per-frame feature sums: frame f of the batch is x[f, 0] for f < B and x[B-1, f-B+1] beyond
        per_frame = torch.cat([x[:, 0].double().sum((1, 2)) + a[:, 0].double().sum((1, 2)),
                               x[-1, 1:].double().sum((1, 2)) + a[-1, 1:].double().sum((1, 2))])
        mult = torch.minimum(torch.minimum(torch.arange(1, B + T, device="cuda"), torch.full((B + T - 1,), T, device="cuda")),
                             torch.arange(B + T - 1, 0, -1, device="cuda")).double()
        chk = float((per_frame * mult).sum())
        got = float(x.double().sum() + a.double().sum())
        assert abs(chk - got) <= 1e-6 * max(1.0, abs(got))


def test_streaming_embedding_per_video_equals_materialised_windows():
    """embedding_per_video over the frame table == model.embed on the materialised windows of each video
    (reference model_utils_new.py:452-748), for a VaDE and a VQ-VAE model."""
    from deepof_b200 import VaDEB200, VQVAEB200, WindowLoader, embedding_per_video
    from oracle import vade_oracle as O
    N, T = 14, 25
    adj = O.default_adjacency(N)
    r, c = np.nonzero(np.triu(adj))
    edges = np.stack([r, c], 1).astype(np.int32)
    vids = [_synthetic(300, N, 5), _synthetic(T + 10, N, 6), _synthetic(20, N, 7)]
    ld = WindowLoader(vids, edges, T, nose=1, tail_base=2, align_node=0, arena_center=(250.0, 250.0))
    for cls, K in ((VaDEB200, 6), (VQVAEB200, 12)):
        m = cls((T, N, 3), (T, len(edges), 1), adj, 8, K, max_batch=128, training=False, seed=2)
        embs, softs = embedding_per_video(m, ld, batch_size=100)
        assert [e.shape[0] for e in embs] == ld.n_windows_per_video and embs[2].shape == (0, 8)
        x, a = ld.load(0, len(ld))
        e_all, q_all = m.embed(x, a)
        assert torch.equal(torch.cat(embs), e_all) and torch.equal(torch.cat([s for s in softs if s is not None]), q_all)
        assert softs[0].shape == (ld.n_windows_per_video[0], K)


def test_streaming_embedding_per_video_transformer_checkpoint():
    """The same streaming path on a transformer-encoder model restored from a reference checkpoint's tensors
    (TFMModelB200): frames -> windows -> transformer encoder -> latent read-out, never materialising the window table."""
    from deepof_b200 import TFMModelB200, WindowLoader, embedding_per_video
    from helpers import load_golden_of, rel_l2
    g = load_golden_of("tfmmodel", "vade")
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    r, c = np.nonzero(np.triu(g["adjacency"]))
    edges = np.stack([r, c], 1).astype(np.int32)
    vids = [_synthetic(150, N, 8), _synthetic(T + 3, N, 9)]
    ld = WindowLoader(vids, edges, T, nose=1, tail_base=2, align_node=0, arena_center=(250.0, 250.0))
    m = TFMModelB200("vade", (T, N, 3), (T, E, 1), g["adjacency"], D, K, max_batch=64)
    m.load_state_dict({k[2:]: g[k] for k in g if k.startswith("p/")})
    embs, softs = embedding_per_video(m, ld, batch_size=50)
    assert [e.shape[0] for e in embs] == ld.n_windows_per_video
    x, a = ld.load(0, len(ld))
    e_all, q_all = m.embed(x, a)
    assert rel_l2(torch.cat(embs).cpu(), e_all.cpu()) < 1e-6 and rel_l2(torch.cat(softs).cpu(), q_all.cpu()) < 1e-6
    assert torch.allclose(q_all.sum(1), torch.ones_like(q_all[:, 0]), atol=1e-5)
