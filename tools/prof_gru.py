"""Stand-alone launcher of the fused GRU layer kernel at cfg2 size, for ncu.  usage: python tools/prof_gru.py H I [S]"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from deepof_b200 import _lib

H, I = int(sys.argv[1]), int(sys.argv[2])
S_ = int(sys.argv[3]) if len(sys.argv) > 3 else 57344
MODE = sys.argv[4] if len(sys.argv) > 4 else "all"      # all | nogt | nostore
TILED = int(sys.argv[5]) if len(sys.argv) > 5 else 0
T = 25
L = _lib.lib()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(S_, T, I, device=dev, generator=g)
k = 1.0 / H ** 0.5
w = [torch.randn(3 * H, I, device=dev, generator=g) * k for _ in range(2)] + [torch.randn(3 * H, H, device=dev, generator=g) * k for _ in range(2)] + \
    [torch.randn(3 * H, device=dev, generator=g) * k for _ in range(4)]
w8 = (C.c_void_p * 8)(*[t.data_ptr() for t in w])
hout = torch.empty(S_, T, 2 * H, device=dev)
gt = [torch.empty(S_, T, 4 * H, device=dev) for _ in range(2)]
hn = torch.empty(S_, 2 * H, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
dbg = torch.zeros(8 * T * 4, dtype=torch.int64, device=dev)
for it in range(3):
    L.dof_test_gru_bwdw_timeline(P(dbg) if it == 2 else None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = L.dof_test_gru_layer_fwd(P(X), T * I, I, w8, None, None if MODE == "nostore" else P(hout),
                                  P(gt[0]) if MODE == "all" else None, P(gt[1]) if MODE == "all" else None, P(hn), S_, T, H, I, TILED, st)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, L.dof_last_error()
    print("fused gru layer H=%d I=%d S=%d mode=%s tiled=%d: %.3f ms" % (H, I, S_, MODE, TILED, e0.elapsed_time(e1)))
L.dof_test_gru_bwdw_timeline(None)
d = dbg.view(8, T, 4).cpu()
t0 = int(d[0, 0, 0])
names = ["gate w4 : loop top | acc ready | math done | h published", "gate w4 : gates + output rows staged", "MMA     : loop top | h ready | h-part issued | x-part of next step issued",
         "store w13: loop top | rows staged | drained", "prod w0 : loop top | x staged | helped draining"]
for r, nm in enumerate(names):
    print(nm)
    for s in (0, 1, 2, 10, 11, 24):
        print("   step %2d:" % s, " ".join("%7d" % (int(v) - t0) if int(v) else "      -" for v in d[r, s]))
print("cycles per step (gate w4, loop top to loop top), steps 2..24: %.0f" % ((int(d[0, 24, 0]) - int(d[0, 2, 0])) / 22))
