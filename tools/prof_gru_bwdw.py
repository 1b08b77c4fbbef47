"""Stand-alone launcher of the fused GRU backward + weight-gradient kernel at cfg2 size, with the per-role clock timeline of
CTA (0, 0).  usage: python tools/prof_gru_bwdw.py H I [S]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from deepof_b200 import _lib

H, I = int(sys.argv[1]), int(sys.argv[2])
S_ = int(sys.argv[3]) if len(sys.argv) > 3 else 57344
T = 25
L = _lib.lib()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(S_, T, I, device=dev, generator=g)
k = 1.0 / H ** 0.5
w = [torch.randn(3 * H, I, device=dev, generator=g) * k for _ in range(2)] + [torch.randn(3 * H, H, device=dev, generator=g) * k for _ in range(2)] + \
    [torch.randn(3 * H, device=dev, generator=g) * k for _ in range(4)]
w8 = (C.c_void_p * 8)(*[t.data_ptr() for t in w])
Sp = (S_ + 127) // 128 * 128
hout = torch.empty(S_, T, 2 * H, device=dev)
gt = [torch.empty(Sp * T * 4 * H, device=dev) for _ in range(2)]
hn = torch.empty(S_, 2 * H, device=dev)
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.dof_test_gru_layer_fwd(P(X), T * I, I, w8, None, P(hout), P(gt[0]), P(gt[1]), P(hn), S_, T, H, I, 1, st) == 0, L.dof_last_error()
dout = torch.randn(S_, T, 2 * H, device=dev, generator=g)
dx = torch.empty(S_, T, I, device=dev)
out = torch.zeros(2 * (3 * H * I + 3 * H * H + 6 * H), device=dev)
dbg = torch.zeros(8 * T * 4, dtype=torch.int64, device=dev)
for it in range(3):
    L.dof_test_gru_bwdw_timeline(P(dbg) if it == 2 else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = L.dof_test_gru_layer_bwdw(P(X), w8, None, P(hout), P(gt[0]), P(gt[1]), P(dout), None, P(dx), None, P(out), S_, T, H, I, st)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, L.dof_last_error()
    print("fused gru bwdw H=%d I=%d S=%d: %.3f ms" % (H, I, S_, e0.elapsed_time(e1)))
L.dof_test_gru_bwdw_timeline(None)
d = dbg.view(8, T, 4).cpu()
t0 = int(d[0, 0, 0])
names = ["gate w4 : loop top | dh read | tiles free | arrived", "gate w11: loop top | dh read | tiles free | arrived",
         "MMA A   : loop top | a_full | dh committed | dx committed", "MMA B   : a_full | w committed",
         "load w0 : loop top | XH free | x staged+arrived | dX drained", "load w14: loop top | XH free | x staged+arrived"]
for r, nm in enumerate(names):
    print(nm)
    for s in (0, 1, 2, 3, 10, 11, 24):
        print("   step %2d:" % s, " ".join("%7d" % (int(v) - t0) if int(v) else "      -" for v in d[r, s]))
print("cycles per step (gate w4, loop top to loop top), steps 2..24: %.0f" % ((int(d[0, 24, 0]) - int(d[0, 2, 0])) / 22))
