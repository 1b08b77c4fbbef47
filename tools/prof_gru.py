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
for it in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = L.dof_test_gru_layer_fwd(P(X), T * I, I, w8, None, None if MODE == "nostore" else P(hout),
                                  P(gt[0]) if MODE == "all" else None, P(gt[1]) if MODE == "all" else None, P(hn), S_, T, H, I, TILED, st)
    e1.record()
    torch.cuda.synchronize()
    assert rc == 0, L.dof_last_error()
    print("fused gru layer H=%d I=%d S=%d mode=%s tiled=%d: %.3f ms" % (H, I, S_, MODE, TILED, e0.elapsed_time(e1)))
