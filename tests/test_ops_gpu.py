"""GPU: single CUDA kernels, called through the C-ABI test hooks, against torch fp64 on the
same device (oracle building blocks where they exist)."""
import ctypes as C

import pytest
import torch

from oracle import vade_oracle as O

pytestmark = pytest.mark.gpu

A_PLAIN, A_SPLIT, A_CONV5, A_TSHIFT = 0, 1, 2, 3


@pytest.fixture(scope="module")
def L():
    from deepof_b200 import _lib
    return _lib.lib()


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rnd(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, device="cuda", generator=g)


@pytest.mark.parametrize("M,N,K,wT", [(1000, 96, 32, 0), (777, 48, 64, 0), (300, 16, 448, 0), (515, 18, 12, 0),
                                      (640, 224, 16, 1), (129, 32, 96, 1), (4096, 42, 32, 0), (50, 6, 6, 1),
                                      # both operands streamed (tc_gemm_big.cuh): K > 128 / weights too large to stay resident
                                      (4096, 256, 256, 0), (2304, 200, 160, 1), (4096, 128, 1792, 0), (3000, 36, 264, 1),
                                      (2500, 768, 256, 0), (5000, 300, 44, 1)])
def test_gemm_rows_plain(L, M, N, K, wT):
    A = rnd(M, K, seed=1)
    W = rnd(N, K, seed=2) if wT == 0 else rnd(K, N, seed=2)
    bias = rnd(N, seed=3)
    Cout = torch.zeros(M, N, device="cuda")
    rc = L.dof_test_gemm_rows(P(A), K, A_PLAIN, 0, 0, 0, P(W), W.shape[1], wT, P(bias), P(Cout), N, M, N, K, 0, 0,
                              None, S())
    assert rc == 0, L.dof_last_error()
    ref = A.double() @ (W.double().t() if wT == 0 else W.double()) + bias.double()
    # the TMEM accumulator TRUNCATES on every tcgen05.mma accumulate (3 per 8 columns of K): the error grows linearly with K
    tol = 2e-6 if K <= 128 else max(4e-6, 1e-8 * K)
    assert rel(Cout, ref) < tol
    # relu + accumulate + mask
    C2 = rnd(M, N, seed=4)
    base = C2.clone()
    mask = rnd(M, N, seed=5)
    rc = L.dof_test_gemm_rows(P(A), K, A_PLAIN, 0, 0, 0, P(W), W.shape[1], wT, P(bias), P(C2), N, M, N, K, 1, 1,
                              P(mask), S())
    assert rc == 0
    ref2 = torch.relu(ref + base.double()) * (mask > 0)
    assert rel(C2, ref2) < tol


def test_gemm_rows_views(L):
    S_, T, Cc, N = 37, 25, 64, 32
    M = S_ * T
    X = rnd(S_, T, Cc, seed=1)
    # conv5 'same' forward == oracle conv
    W = rnd(N, Cc, 5, seed=2)
    out = torch.zeros(M, N, device="cuda")
    rc = L.dof_test_gemm_rows(P(X), Cc, A_CONV5, T, 1, 0, P(W), Cc * 5, 0, None, P(out), N, M, N, Cc * 5, 0, 0, None, S())
    assert rc == 0
    ref = O.conv1d_same_k5(X.double(), W.double()).reshape(M, N)
    assert rel(out, ref) < 2e-6
    # split view: columns [0,2H) u [3H,4H)
    H = 16
    G = rnd(M, 4 * H, seed=3)
    Wi = rnd(3 * H, 20, seed=4)
    dX = torch.zeros(M, 20, device="cuda")
    rc = L.dof_test_gemm_rows(P(G), 4 * H, A_SPLIT, 2 * H, H, 0, P(Wi), 20, 1, None, P(dX), 20, M, 20, 3 * H, 0, 0, None, S())
    assert rc == 0
    Gi = torch.cat([G[:, :2 * H], G[:, 3 * H:]], 1).double()
    assert rel(dX, Gi @ Wi.double()) < 2e-6
    # time-shift view
    for shift in (-1, 1):
        out = torch.zeros(M, N, device="cuda")
        Wt = rnd(N, Cc, seed=5)
        rc = L.dof_test_gemm_rows(P(X), Cc, A_TSHIFT, T, shift, 0, P(Wt), Cc, 0, None, P(out), N, M, N, Cc, 0, 0, None, S())
        assert rc == 0
        Xs = torch.zeros_like(X)
        if shift == -1:
            Xs[:, 1:] = X[:, :-1]
        else:
            Xs[:, :-1] = X[:, 1:]
        assert rel(out, Xs.reshape(M, Cc).double() @ Wt.double().t()) < 2e-6


@pytest.mark.parametrize("M,N,K,oT", [(5000, 96, 32, 0), (3333, 48, 64, 0), (257, 16, 448, 0), (999, 16, 32, 1),
                                      (100000, 32, 15, 0), (64, 18, 6, 0)])
def test_gemm_wgrad(L, M, N, K, oT):
    Pm, Q = rnd(M, N, seed=1), rnd(M, K, seed=2)
    dW = torch.zeros(N, K, device="cuda") if oT == 0 else torch.zeros(K, N, device="cuda")
    db = torch.zeros(N, device="cuda")
    rc = L.dof_test_gemm_wgrad(P(Pm), N, A_PLAIN, 0, 0, P(Q), K, A_PLAIN, 0, 0, P(dW), dW.shape[1], oT, P(db), M, N, K, S())
    assert rc == 0, L.dof_last_error()
    ref = Pm.double().t() @ Q.double()
    assert rel(dW if oT == 0 else dW.t(), ref) < 5e-6
    assert rel(db, Pm.double().sum(0)) < 5e-6


def test_gemm_wgrad_views(L):
    S_, T, Cc, N = 41, 24, 12, 20
    M = S_ * T
    X = rnd(S_, T, Cc, seed=1)
    dY = rnd(M, N, seed=2)
    dW = torch.zeros(N, Cc * 5, device="cuda")
    rc = L.dof_test_gemm_wgrad(P(dY), N, A_PLAIN, 0, 0, P(X), Cc, A_CONV5, T, 1, P(dW), Cc * 5, 0, None, M, N, Cc * 5, S())
    assert rc == 0
    Xd = X.double().requires_grad_(False)
    Wd = torch.zeros(N, Cc, 5, device="cuda", dtype=torch.float64, requires_grad=True)
    (O.conv1d_same_k5(Xd, Wd).reshape(M, N) * dY.double()).sum().backward()
    assert rel(dW, Wd.grad.reshape(N, Cc * 5)) < 5e-6
    # split P + shifted Q (the dW_hh pattern)
    H = 8
    G = rnd(M, 4 * H, seed=3)
    Hs = rnd(S_, T, 2 * H, seed=4)
    for d, shift in ((0, -1), (1, 1)):
        dWh = torch.zeros(3 * H, H, device="cuda")
        dbh = torch.zeros(3 * H, device="cuda")
        Hd = Hs[:, :, d * H:(d + 1) * H]
        rc = L.dof_test_gemm_wgrad(P(G), 4 * H, A_PLAIN, 0, 0, C.c_void_p(Hs.data_ptr() + d * H * 4), 2 * H, A_TSHIFT, T,
                                   shift, P(dWh), H, 0, P(dbh), M, 3 * H, H, S())
        assert rc == 0
        Hsh = torch.zeros_like(Hd)
        if shift == -1:
            Hsh[:, 1:] = Hd[:, :-1]
        else:
            Hsh[:, :-1] = Hd[:, 1:]
        ref = G[:, :3 * H].double().t() @ Hsh.reshape(M, H).double()
        assert rel(dWh, ref) < 5e-6
        assert rel(dbh, G[:, :3 * H].double().sum(0)) < 5e-6


@pytest.mark.parametrize("S_,T,H,I,ldx,off", [(200, 25, 32, 64, 64, 0), (333, 24, 16, 32, 32, 0), (411, 11, 32, 16, 48, 1),
                                              (180, 25, 16, 64, 64, 3), (1000, 5, 32, 48, 48, 0)])
def test_gru_wgrad_merged(L, S_, T, H, I, ldx, off):
    """All four parameter gradients of both directions from one pass over dG (TMA + MN-major tf32 operands)."""
    M = S_ * T
    G = [rnd(M, 4 * H, seed=3 + d) for d in range(2)]
    Xfull = rnd(M, ldx, seed=5)
    X = Xfull[:, :I]
    Hs = rnd(S_, T, 2 * H, seed=6)
    per = 3 * H * I + 3 * H * H + 6 * H
    out = torch.zeros(2 * per + off, device="cuda")[off:]        # off != 0: gradient tensors that are not 16-byte aligned
    rc = L.dof_test_gru_wgrad(P(G[0]), P(G[1]), P(Xfull), ldx, P(Hs), P(out), M, T, I, H, S())
    assert rc == 0, L.dof_last_error()
    torch.cuda.synchronize()
    for d in range(2):
        o = out[d * per:(d + 1) * per]
        dWih, dWhh = o[:3 * H * I].view(3 * H, I), o[3 * H * I:3 * H * I + 3 * H * H].view(3 * H, H)
        dbih, dbhh = o[-6 * H:-3 * H], o[-3 * H:]
        Gd = G[d].double()
        Gi = torch.cat([Gd[:, :2 * H], Gd[:, 3 * H:]], 1)
        Gh = Gd[:, :3 * H]
        Hd = Hs[:, :, d * H:(d + 1) * H].double()
        Hsh = torch.zeros_like(Hd)
        if d == 0:
            Hsh[:, 1:] = Hd[:, :-1]
        else:
            Hsh[:, :-1] = Hd[:, 1:]
        assert rel(dWih, Gi.t() @ X.double()) < 5e-6, d
        assert rel(dWhh, Gh.t() @ Hsh.reshape(M, H)) < 5e-6, d
        assert rel(dbih, Gi.sum(0)) < 5e-6 and rel(dbhh, Gh.sum(0)) < 5e-6, d


@pytest.mark.parametrize("R,W,relu_in", [(1000, 64, 0), (4097, 32, 1), (33, 12, 0), (500, 128, 0), (10, 256, 0)])
def test_layernorm(L, R, W, relu_in):
    x = rnd(R, W, seed=1)
    if relu_in:
        x = torch.relu(x)
    w, b, dy = rnd(W, seed=2), rnd(W, seed=3), rnd(R, W, seed=4)
    y, dx = torch.zeros_like(x), torch.zeros_like(x)
    mu, rs = torch.zeros(R, device="cuda"), torch.zeros(R, device="cuda")
    dw, db = torch.zeros(W, device="cuda"), torch.zeros(W, device="cuda")
    rc = L.dof_test_layernorm(P(x), P(w), P(b), 1e-3, P(y), P(mu), P(rs), P(dy), P(dx), P(dw), P(db), R, W, relu_in, S())
    assert rc == 0, L.dof_last_error()
    xd = x.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = O.layer_norm(xd, wd, bd, 1e-3)
    (yr * dy.double()).sum().backward()
    assert rel(y, yr) < 2e-6
    gx = xd.grad * (x > 0) if relu_in else xd.grad
    assert rel(dx, gx) < 1e-5
    assert rel(dw, wd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5


def _gru_case(L, S_, T, I, H, lens, final_only, seed):
    dev = "cuda"
    X = rnd(S_, T, I, seed=seed)
    k = 1.0 / H ** 0.5
    prm = {}
    for d in ("", "_reverse"):
        prm["weight_ih_l0" + d] = rnd(3 * H, I, seed=seed + 1 + len(d)) * k
        prm["weight_hh_l0" + d] = rnd(3 * H, H, seed=seed + 2 + len(d)) * k
        prm["bias_ih_l0" + d] = rnd(3 * H, seed=seed + 3 + len(d)) * k
        prm["bias_hh_l0" + d] = rnd(3 * H, seed=seed + 4 + len(d)) * k
    len_t = torch.full((S_,), T, dtype=torch.int32, device=dev) if lens is None else lens.to(dev).int()
    # ---- reference (fp64, oracle building block)
    Xd = X.double().requires_grad_(True)
    pd = {k_: v.double().requires_grad_(True) for k_, v in prm.items()}
    out_ref, hn_ref = O.bigru(Xd, len_t.long(), pd, "")
    R1, R2 = rnd(S_, T, 2 * H, seed=seed + 9).double(), rnd(S_, 2 * H, seed=seed + 10).double()
    loss = (hn_ref * R2).sum() if final_only else (out_ref * R1).sum()
    loss.backward()
    # ---- ours
    gi = [(X @ prm["weight_ih_l0" + d].t() + prm["bias_ih_l0" + d]).contiguous() for d in ("", "_reverse")]
    hout = torch.full((S_, T, 2 * H), 7.0, device=dev)
    gt = [torch.zeros(S_, T, 4 * H, device=dev) for _ in range(2)]
    hn = torch.zeros(S_, 2 * H, device=dev)
    rc = L.dof_test_gru_fwd(P(gi[0]), P(gi[1]), T * 3 * H, 3 * H, P(prm["weight_hh_l0"]), P(prm["weight_hh_l0_reverse"]),
                            P(prm["bias_hh_l0"]), P(prm["bias_hh_l0_reverse"]), P(len_t), P(hout), P(gt[0]), P(gt[1]),
                            P(hn), S_, T, H, S())
    assert rc == 0, L.dof_last_error()
    assert rel(hout, out_ref) < 5e-6, ("hout", rel(hout, out_ref))
    assert rel(hn, hn_ref) < 5e-6, ("hn", rel(hn, hn_ref))
    dg = [torch.full((S_, T, 4 * H), 3.0, device=dev) for _ in range(2)]
    dout = None if final_only else R1.float().contiguous()
    dhn = R2.float().contiguous() if final_only else None
    rc = L.dof_test_gru_bwd(P(prm["weight_hh_l0"]), P(prm["weight_hh_l0_reverse"]), P(len_t), P(hout), P(gt[0]), P(gt[1]),
                            P(dout), P(dhn), P(dg[0]), P(dg[1]), S_, T, H, S())
    assert rc == 0, L.dof_last_error()
    dX = torch.zeros(S_, T, I, device=dev, dtype=torch.float64)
    for di, d in enumerate(("", "_reverse")):
        G = dg[di].double()
        dGi = torch.cat([G[..., :2 * H], G[..., 3 * H:]], -1)
        dGh = G[..., :3 * H]
        dX += dGi @ prm["weight_ih_l0" + d].double()
        assert rel(dGi.sum((0, 1)), pd["bias_ih_l0" + d].grad) < 2e-5, ("db_ih", d)
        assert rel(dGh.sum((0, 1)), pd["bias_hh_l0" + d].grad) < 2e-5, ("db_hh", d)
        assert rel(torch.einsum("stg,sti->gi", dGi, X.double()), pd["weight_ih_l0" + d].grad) < 2e-5, ("dW_ih", d)
        Hd = hout[..., di * H:(di + 1) * H].double()
        Hp = torch.zeros_like(Hd)
        if di == 0:
            Hp[:, 1:] = Hd[:, :-1]
        else:
            Hp[:, :-1] = Hd[:, 1:]
        assert rel(torch.einsum("stg,sth->gh", dGh, Hp), pd["weight_hh_l0" + d].grad) < 2e-5, ("dW_hh", d)
    assert rel(dX, Xd.grad) < 2e-5, ("dX", rel(dX, Xd.grad))


@pytest.mark.parametrize("S_,T,I,H", [(300, 25, 32, 32), (130, 25, 64, 16), (77, 24, 12, 12), (50, 24, 24, 6),
                                      (64, 25, 16, 8), (40, 10, 20, 64), (33, 25, 48, 24), (20, 7, 9, 5), (45, 12, 8, 4), (45, 12, 8, 8)])
def test_gru_full_sequences(L, S_, T, I, H):
    _gru_case(L, S_, T, I, H, None, False, seed=10)
    _gru_case(L, S_, T, I, H, None, True, seed=20)


def test_gru_packed_lengths(L):
    S_, T = 97, 25
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(0, T + 1, (S_,), generator=g)
    lens[0], lens[1], lens[2] = 0, 1, T
    _gru_case(L, S_, T, 32, 32, lens, False, seed=30)
    _gru_case(L, S_, T, 64, 16, lens, True, seed=40)


@pytest.mark.parametrize("S_,T,I,H,packed,repeat", [(300, 25, 32, 32, False, False), (1000, 25, 64, 16, True, False),
                                                     (128, 24, 16, 16, False, True), (257, 7, 32, 32, True, False),
                                                     (200, 25, 48, 32, True, False)])
def test_gru_fused_tc_layer(L, S_, T, I, H, packed, repeat):
    """Fused tcgen05 GRU layer (projection + recurrence + gates in one kernel) vs the fp64 oracle GRU, and its saved
    gates vs the SIMT kernel's (they feed the same backward kernel)."""
    dev = "cuda"
    X = rnd(S_, 1 if repeat else T, I, seed=50)
    Xfull = X.expand(S_, T, I).contiguous() if repeat else X
    k = 1.0 / H ** 0.5
    prm = {}
    for d in ("", "_reverse"):
        prm["weight_ih_l0" + d] = rnd(3 * H, I, seed=51 + len(d)) * k
        prm["weight_hh_l0" + d] = rnd(3 * H, H, seed=52 + len(d)) * k
        prm["bias_ih_l0" + d] = rnd(3 * H, seed=53 + len(d)) * k
        prm["bias_hh_l0" + d] = rnd(3 * H, seed=54 + len(d)) * k
    if packed:
        g = torch.Generator().manual_seed(1)
        lens = torch.randint(0, T + 1, (S_,), generator=g)
        lens[0], lens[1], lens[2] = 0, 1, T
        len_t = lens.to(dev).int()
    else:
        len_t = torch.full((S_,), T, dtype=torch.int32, device=dev)
    out_ref, hn_ref = O.bigru(Xfull.double(), len_t.long(), {k_: v.double() for k_, v in prm.items()}, "")
    hout = torch.full((S_, T, 2 * H), 7.0, device=dev)
    gt = [torch.zeros(S_, T, 4 * H, device=dev) for _ in range(2)]
    hn = torch.zeros(S_, 2 * H, device=dev)
    names = ["weight_ih_l0", "weight_ih_l0_reverse", "weight_hh_l0", "weight_hh_l0_reverse", "bias_ih_l0", "bias_ih_l0_reverse",
             "bias_hh_l0", "bias_hh_l0_reverse"]
    w8 = (C.c_void_p * 8)(*[prm[n].data_ptr() for n in names])
    ran = _ran(L, lambda: check_rc(L, L.dof_test_gru_layer_fwd(P(X), 0 if False else (I if repeat else T * I), 0 if repeat else I, w8,
                                                               P(len_t), P(hout), P(gt[0]), P(gt[1]), P(hn), S_, T, H, I, 0, S())))
    assert any(n.startswith("gru_fwd_tc") for n in ran), ran
    print("fused gru", (S_, T, I, H), rel(hout, out_ref), rel(hn, hn_ref))
    assert rel(hout, out_ref) < 5e-6 and rel(hn, hn_ref) < 5e-6
    # saved gates == what the SIMT kernel saves
    gi = [(Xfull @ prm["weight_ih_l0" + d].t() + prm["bias_ih_l0" + d]).contiguous() for d in ("", "_reverse")]
    hout2 = torch.zeros_like(hout)
    gt2 = [torch.zeros(S_, T, 4 * H, device=dev) for _ in range(2)]
    hn2 = torch.zeros_like(hn)
    rc = L.dof_test_gru_fwd(P(gi[0]), P(gi[1]), T * 3 * H, 3 * H, P(prm["weight_hh_l0"]), P(prm["weight_hh_l0_reverse"]),
                            P(prm["bias_hh_l0"]), P(prm["bias_hh_l0_reverse"]), P(len_t), P(hout2), P(gt2[0]), P(gt2[1]),
                            P(hn2), S_, T, H, S())
    assert rc == 0, L.dof_last_error()
    for d in range(2):
        assert rel(gt[d], gt2[d]) < 1e-5, ("gates", d, rel(gt[d], gt2[d]))


@pytest.mark.parametrize("S_,T,I,H,packed,final_only,mask", [(300, 25, 32, 32, False, False, True), (1000, 25, 64, 16, True, True, False),
                                                              (257, 7, 32, 32, True, False, False), (130, 24, 16, 16, True, False, False)])
def test_gru_fused_tc_backward(L, S_, T, I, H, packed, final_only, mask):
    """Fused tcgen05 BPTT (dG + dX) vs the SIMT BPTT kernel fed by the same forward, and dX vs dGi . W_ih in fp64."""
    dev = "cuda"
    X = rnd(S_, T, I, seed=60)
    k = 1.0 / H ** 0.5
    names = ["weight_ih_l0", "weight_ih_l0_reverse", "weight_hh_l0", "weight_hh_l0_reverse", "bias_ih_l0", "bias_ih_l0_reverse",
             "bias_hh_l0", "bias_hh_l0_reverse"]
    prm = {n: rnd(*((3 * H, I) if "weight_ih" in n else (3 * H, H) if "weight_hh" in n else (3 * H,)), seed=61 + i) * k for i, n in enumerate(names)}
    w8 = (C.c_void_p * 8)(*[prm[n].data_ptr() for n in names])
    if packed:
        g = torch.Generator().manual_seed(2)
        lens = torch.randint(0, T + 1, (S_,), generator=g)
        lens[0], lens[1], lens[2] = 0, 1, T
        len_t = lens.to(dev).int()
    else:
        len_t = torch.full((S_,), T, dtype=torch.int32, device=dev)
    Sp = (S_ + 127) // 128 * 128
    hout = torch.zeros(S_, T, 2 * H, device=dev)
    hn = torch.zeros(S_, 2 * H, device=dev)
    gt_rm = [torch.zeros(S_, T, 4 * H, device=dev) for _ in range(2)]
    gt_ti = [torch.zeros(Sp * T * 4 * H, device=dev) for _ in range(2)]
    for tiled, gt in ((0, gt_rm), (1, gt_ti)):
        check_rc(L, L.dof_test_gru_layer_fwd(P(X), T * I, I, w8, P(len_t), P(hout), P(gt[0]), P(gt[1]), P(hn), S_, T, H, I, tiled, S()))
    dout = None if final_only else rnd(S_, T, 2 * H, seed=70)
    dhn = rnd(S_, 2 * H, seed=71) if final_only else None
    # SIMT reference path
    dg_ref = [torch.full((S_, T, 4 * H), 3.0, device=dev) for _ in range(2)]
    rc = L.dof_test_gru_bwd(P(prm["weight_hh_l0"]), P(prm["weight_hh_l0_reverse"]), P(len_t), P(hout), P(gt_rm[0]), P(gt_rm[1]),
                            P(dout), P(dhn), P(dg_ref[0]), P(dg_ref[1]), S_, T, H, S())
    assert rc == 0, L.dof_last_error()
    # fused path
    dg = [torch.full((S_, T, 4 * H), 5.0, device=dev) for _ in range(2)]
    dx = torch.full((S_, T, I), 9.0, device=dev)
    mk = (rnd(S_, T, I, seed=72) if mask else None)
    ran = _ran(L, lambda: check_rc(L, L.dof_test_gru_layer_bwd(w8, P(len_t), P(hout), P(gt_ti[0]), P(gt_ti[1]), P(dout), P(dhn), P(dg[0]),
                                                               P(dg[1]), P(dx), P(mk), S_, T, H, I, S())))
    assert any(n.startswith("gru_bwd_tc") for n in ran), ran
    for d in range(2):
        print("fused bwd dG", (S_, T, I, H), d, rel(dg[d], dg_ref[d]))
        assert rel(dg[d], dg_ref[d]) < 1e-5, (d, rel(dg[d], dg_ref[d]))
    dx_ref = torch.zeros(S_, T, I, device=dev, dtype=torch.float64)
    for di, sfx in enumerate(("", "_reverse")):
        G = dg_ref[di].double()
        dx_ref += torch.cat([G[..., :2 * H], G[..., 3 * H:]], -1) @ prm["weight_ih_l0" + sfx].double()
    if mask:
        dx_ref = dx_ref * (mk > 0)
    print("fused bwd dX", rel(dx, dx_ref))
    assert rel(dx, dx_ref) < 1e-5


@pytest.mark.parametrize("S_,T,I,H,packed,final_only,mask,off", [(300, 25, 32, 32, False, False, True, 0), (1000, 25, 64, 16, True, True, False, 0),
                                                                  (257, 7, 32, 32, True, False, False, 1), (130, 24, 16, 16, True, False, False, 0),
                                                                  (113, 25, 32, 16, False, False, False, 3), (4000, 25, 32, 32, True, False, True, 0)])
def test_gru_fused_backward_with_weight_gradients(L, S_, T, I, H, packed, final_only, mask, off):
    """Second-generation fused backward (gru_bwdw_tc.cuh: BPTT + dX + the four parameter gradients in one kernel, dG never
    in HBM) vs the fp64 autograd of the oracle's GRU fed with the same inputs."""
    dev = "cuda"
    X = rnd(S_, T, I, seed=80)
    k = 1.0 / H ** 0.5
    names = ["weight_ih_l0", "weight_ih_l0_reverse", "weight_hh_l0", "weight_hh_l0_reverse", "bias_ih_l0", "bias_ih_l0_reverse",
             "bias_hh_l0", "bias_hh_l0_reverse"]
    prm = {n: rnd(*((3 * H, I) if "weight_ih" in n else (3 * H, H) if "weight_hh" in n else (3 * H,)), seed=81 + i) * k for i, n in enumerate(names)}
    w8 = (C.c_void_p * 8)(*[prm[n].data_ptr() for n in names])
    if packed:
        g = torch.Generator().manual_seed(3)
        lens = torch.randint(0, T + 1, (S_,), generator=g)
        lens[0], lens[1], lens[2] = 0, 1, T
        len_t = lens.to(dev).int()
    else:
        len_t = torch.full((S_,), T, dtype=torch.int32, device=dev)
    Sp = (S_ + 127) // 128 * 128
    hout = torch.zeros(S_, T, 2 * H, device=dev)
    hn = torch.zeros(S_, 2 * H, device=dev)
    gt_ti = [torch.zeros(Sp * T * 4 * H, device=dev) for _ in range(2)]
    check_rc(L, L.dof_test_gru_layer_fwd(P(X), T * I, I, w8, P(len_t), P(hout), P(gt_ti[0]), P(gt_ti[1]), P(hn), S_, T, H, I, 1, S()))
    dout = None if final_only else rnd(S_, T, 2 * H, seed=90)
    dhn = rnd(S_, 2 * H, seed=91) if final_only else None
    mk = (rnd(S_, T, I, seed=92) if mask else None)
    dx = torch.full((S_, T, I), 9.0, device=dev)
    per = 3 * H * I + 3 * H * H + 6 * H
    out = torch.zeros(2 * per + off, device=dev)[off:]          # off != 0: gradient tensors that are not 16-byte aligned
    ran = _ran(L, lambda: check_rc(L, L.dof_test_gru_layer_bwdw(P(X), w8, P(len_t), P(hout), P(gt_ti[0]), P(gt_ti[1]), P(dout), P(dhn),
                                                                P(dx), P(mk), P(out), S_, T, H, I, S())))
    assert any(n.startswith("gru_bwdw_tc") for n in ran), ran
    torch.cuda.synchronize()
    # fp64 reference: autograd through the oracle's bidirectional GRU
    Xd = X.double().requires_grad_(True)
    pd = {k_: v.double().requires_grad_(True) for k_, v in prm.items()}
    out_ref, hn_ref = O.bigru(Xd, len_t.long(), pd, "")
    loss = (hn_ref * dhn.double()).sum() if final_only else (out_ref * dout.double()).sum()
    loss.backward()
    dx_ref = Xd.grad * (mk > 0) if mask else Xd.grad
    print("fused bwdw dX", (S_, T, I, H), rel(dx, dx_ref))
    assert rel(dx, dx_ref) < 1e-5
    for d, sfx in enumerate(("", "_reverse")):
        o = out[d * per:(d + 1) * per]
        got = {"weight_ih_l0": o[:3 * H * I].view(3 * H, I), "weight_hh_l0": o[3 * H * I:3 * H * I + 3 * H * H].view(3 * H, H),
               "bias_ih_l0": o[-6 * H:-3 * H], "bias_hh_l0": o[-3 * H:]}
        for n, v in got.items():
            e = rel(v, pd[n + sfx].grad)
            print("fused bwdw", n + sfx, e)
            assert e < 1e-5, (n + sfx, e)


def check_rc(L, rc):
    assert rc == 0, L.dof_last_error()


# ---------------------------------------------------------------------------
# tcgen05 (3xTF32) GEMMs: same hooks, shapes large enough to be dispatched to the
# tensor-core kernels; the per-kernel profile proves which kernel ran.
# ---------------------------------------------------------------------------
def _ran(L, fn):
    L.dof_profile_begin()
    fn()
    buf = C.create_string_buffer(4096)
    L.dof_profile_end(buf, 4096)
    return {ln.split()[0] for ln in buf.value.decode().strip().split("\n") if ln}


@pytest.mark.parametrize("M,N,K,wT", [(20000, 96, 32, 0), (20000, 48, 64, 0), (20000, 64, 48, 1), (20000, 32, 96, 1),
                                      (5000, 42, 32, 0), (4097, 16, 8, 0), (3000, 32, 128, 0), (3000, 64, 96, 0), (2048, 8, 4, 0)])
def test_tc_gemm_rows(L, M, N, K, wT):
    L.dof_set_tensor_cores(1)
    A = rnd(M, K, seed=1)
    W = rnd(N, K, seed=2) if wT == 0 else rnd(K, N, seed=2)
    bias = rnd(N, seed=3)
    Cout = torch.zeros(M, N, device="cuda")
    names = _ran(L, lambda: L.dof_test_gemm_rows(P(A), K, A_PLAIN, 0, 0, 0, P(W), W.shape[1], wT, P(bias), P(Cout), N, M,
                                                 N, K, 0, 0, None, S()))
    assert "gemm_rows_tc" in names, names
    ref = A.double() @ (W.double().t() if wT == 0 else W.double()) + bias.double()
    print("tc rows", (M, N, K, wT), rel(Cout, ref))
    assert rel(Cout, ref) < 5e-6
    C2 = rnd(M, N, seed=4)
    base, mask = C2.clone(), rnd(M, N, seed=5)
    rc = L.dof_test_gemm_rows(P(A), K, A_PLAIN, 0, 0, 0, P(W), W.shape[1], wT, P(bias), P(C2), N, M, N, K, 1, 1, P(mask), S())
    assert rc == 0
    assert rel(C2, torch.relu(ref + base.double()) * (mask > 0)) < 5e-6
    # the SIMT kernel must agree
    L.dof_set_tensor_cores(0)
    C3 = torch.zeros(M, N, device="cuda")
    names = _ran(L, lambda: L.dof_test_gemm_rows(P(A), K, A_PLAIN, 0, 0, 0, P(W), W.shape[1], wT, P(bias), P(C3), N, M,
                                                 N, K, 0, 0, None, S()))
    L.dof_set_tensor_cores(1)
    assert "gemm_rows" in names and "gemm_rows_tc" not in names
    assert rel(C3, ref) < 5e-6


def test_tc_gemm_rows_split_view(L):
    L.dof_set_tensor_cores(1)
    M, H, I = 30000, 16, 64
    G = rnd(M, 4 * H, seed=3)
    Wi = rnd(3 * H, I, seed=4)
    dX = torch.zeros(M, I, device="cuda")
    names = _ran(L, lambda: L.dof_test_gemm_rows(P(G), 4 * H, A_SPLIT, 2 * H, H, 0, P(Wi), I, 1, None, P(dX), I, M, I, 3 * H,
                                                 0, 0, None, S()))
    assert "gemm_rows_tc" in names
    Gi = torch.cat([G[:, :2 * H], G[:, 3 * H:]], 1).double()
    assert rel(dX, Gi @ Wi.double()) < 5e-6


@pytest.mark.parametrize("M,N,K,oT", [(50000, 96, 32, 0), (50000, 48, 64, 0), (50000, 48, 16, 0), (20000, 16, 16, 1),
                                      (8192, 128, 64, 0), (4100, 4, 4, 0), (30000, 32, 96, 0)])
def test_tc_gemm_wgrad(L, M, N, K, oT):
    L.dof_set_tensor_cores(1)
    Pm, Q = rnd(M, N, seed=1), rnd(M, K, seed=2)
    dW = torch.zeros(N, K, device="cuda") if oT == 0 else torch.zeros(K, N, device="cuda")
    db = torch.zeros(N, device="cuda")
    names = _ran(L, lambda: L.dof_test_gemm_wgrad(P(Pm), N, A_PLAIN, 0, 0, P(Q), K, A_PLAIN, 0, 0, P(dW), dW.shape[1], oT,
                                                  P(db), M, N, K, S()))
    assert "gemm_wgrad_tc" in names, names
    ref = Pm.double().t() @ Q.double()
    print("tc wgrad", (M, N, K, oT), rel(dW if oT == 0 else dW.t(), ref), rel(db, Pm.double().sum(0)))
    assert rel(dW if oT == 0 else dW.t(), ref) < 1e-5
    assert rel(db, Pm.double().sum(0)) < 1e-5


def test_tc_gemm_wgrad_views(L):
    L.dof_set_tensor_cores(1)
    S_, T, H = 400, 25, 16
    M = S_ * T
    G = rnd(M, 4 * H, seed=3)
    X = rnd(M, 32, seed=5)
    dWi = torch.zeros(3 * H, 32, device="cuda")
    dbi = torch.zeros(3 * H, device="cuda")
    names = _ran(L, lambda: L.dof_test_gemm_wgrad(P(G), 4 * H, A_SPLIT, 2 * H, H, P(X), 32, A_PLAIN, 0, 0, P(dWi), 32, 0,
                                                  P(dbi), M, 3 * H, 32, S()))
    assert "gemm_wgrad_tc" in names
    Gi = torch.cat([G[:, :2 * H], G[:, 3 * H:]], 1).double()
    assert rel(dWi, Gi.t() @ X.double()) < 1e-5 and rel(dbi, Gi.sum(0)) < 1e-5
    Hs = rnd(S_, T, 2 * H, seed=4)
    for d, shift in ((0, -1), (1, 1)):
        dWh = torch.zeros(3 * H, H, device="cuda")
        dbh = torch.zeros(3 * H, device="cuda")
        Hd = Hs[:, :, d * H:(d + 1) * H]
        names = _ran(L, lambda: L.dof_test_gemm_wgrad(P(G), 4 * H, A_PLAIN, 0, 0, C.c_void_p(Hs.data_ptr() + d * H * 4),
                                                      2 * H, A_TSHIFT, T, shift, P(dWh), H, 0, P(dbh), M, 3 * H, H, S()))
        assert "gemm_wgrad_tc" in names
        Hsh = torch.zeros_like(Hd)
        if shift == -1:
            Hsh[:, 1:] = Hd[:, :-1]
        else:
            Hsh[:, :-1] = Hd[:, 1:]
        ref = G[:, :3 * H].double().t() @ Hsh.reshape(M, H).double()
        assert rel(dWh, ref) < 1e-5 and rel(dbh, G[:, :3 * H].double().sum(0)) < 1e-5


@pytest.mark.parametrize("S_,T,dm,heads,causal,pad,rate", [(37, 25, 40, 4, 0, True, 0.1), (20, 25, 64, 8, 1, False, 0.2),
                                                           (9, 12, 32, 4, 0, False, 0.0), (5, 40, 24, 3, 1, True, 0.3)])
def test_tfm_attention_train_forward_backward(L, S_, T, dm, heads, causal, pad, rate):
    """Attention block of the transformer training step: key-padding / causal masks, dropout on the weights with an
    injected keep mask; forward and the recompute-based backward vs autograd in float64."""
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(S_, T, 3 * dm, generator=g).cuda()
    kpad = None
    if pad:
        kp = (torch.rand(S_, T, generator=g) < 0.25)
        kp[:, 0] = False                                   # never all keys padded (the reference yields NaN there)
        kpad = kp.to(torch.uint8).cuda()
    keep = (torch.rand(S_, heads, T, T, generator=g) >= rate).to(torch.uint8).cuda() if rate > 0 else None
    dout = torch.randn(S_, T, dm, generator=g).cuda()
    out = torch.zeros(S_, T, dm, device="cuda")
    dqkv = torch.zeros_like(qkv)
    assert L.dof_test_tfm_attention(P(qkv), P(kpad), P(keep), rate, causal, S_, T, dm, heads, P(out), None, None, S()) == 0, L.dof_last_error()
    assert L.dof_test_tfm_attention(P(qkv), P(kpad), P(keep), rate, causal, S_, T, dm, heads, None, P(dout), P(dqkv), S()) == 0, L.dof_last_error()
    x = qkv.double().requires_grad_(True)
    hd = dm // heads
    q, k, v = (x[..., i * dm:(i + 1) * dm].view(S_, T, heads, hd).transpose(1, 2) for i in range(3))
    sc = q @ k.transpose(-1, -2) / hd ** 0.5
    if pad:
        sc = sc.masked_fill(kpad.bool()[:, None, None, :], float("-inf"))
    if causal:
        sc = sc.masked_fill(torch.triu(torch.ones(T, T, dtype=torch.bool, device="cuda"), 1), float("-inf"))
    p = torch.softmax(sc, -1)
    if keep is not None:
        p = p * keep.double() / (1.0 - rate)
    ref = (p @ v).transpose(1, 2).reshape(S_, T, dm)
    ref.backward(dout.double())
    assert rel(out, ref.detach()) < 2e-6
    assert rel(dqkv, x.grad) < 5e-6


@pytest.mark.parametrize("S_,T,cin,C,dil", [(180, 25, 32, 32, 1), (180, 25, 32, 32, 8), (170, 25, 3, 32, 2), (170, 25, 1, 32, 4),
                                            (200, 25, 64, 64, 4), (176, 24, 24, 64, 8), (40, 12, 32, 32, 2), (7, 5, 8, 16, 1)])
def test_tcn_dilated_causal_conv_vs_torch(L, S_, T, cin, C, dil):
    """The TCN family's Conv1d(k = 4, causal, dilated) as the library runs it — a GEMM over an A_TAPS view of [rows, channels]
    activations, its input gradient and its fused-tap weight gradient — against torch.nn.functional.conv1d + autograd in fp64.
    Sizes with >= 2048 / 4096 rows take the tcgen05 kernels (incl. the 3 / 1-channel inputs padded to pitch 4 and the K = 256
    decoder shape staged in 32-column blocks), the small ones the SIMT kernels.  3xTF32: 1e-5 rel-L2."""
    R = S_ * T
    ldx = (cin + 3) // 4 * 4
    x = rnd(S_, T, cin, seed=1)
    X = torch.zeros(R, ldx, device="cuda")
    X[:, :cin] = x.reshape(R, cin)
    W, b = rnd(C, cin, 4, seed=2) * 0.2, rnd(C, seed=3)
    dA = rnd(S_, T, C, seed=4)
    xd, Wd, bd = x.double().requires_grad_(True), W.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.conv1d(torch.nn.functional.pad(xd.transpose(1, 2), (3 * dil, 0)), Wd, bd, dilation=dil).transpose(1, 2)
    (ref * dA.double()).sum().backward()
    A = torch.zeros(R, C, device="cuda")
    assert L.dof_test_tcn_conv(0, P(X), ldx, cin, T, dil, P(W), P(b), P(A), C, R, None, None, None, S()) == 0, L.dof_last_error()
    assert rel(A, ref.reshape(R, C)) < 1e-5
    dAf = dA.reshape(R, C).contiguous()
    dW, db = torch.zeros_like(W), torch.zeros_like(b)
    assert L.dof_test_tcn_conv(2, P(X), ldx, cin, T, dil, P(W), None, P(dAf), C, R, None, P(dW), P(db), S()) == 0, L.dof_last_error()
    assert rel(dW, Wd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5
    if ldx == cin:
        dX = torch.zeros(R, cin, device="cuda")
        assert L.dof_test_tcn_conv(1, P(X), ldx, cin, T, dil, P(W), None, P(dAf), C, R, P(dX), None, None, S()) == 0, L.dof_last_error()
        assert rel(dX, xd.grad.reshape(R, cin)) < 1e-5
