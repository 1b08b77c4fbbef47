"""Times dof_tfm_encode (transformer encoder, eval forward) at the cfg5 geometry: B windows of 25 x 14 x 3."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from deepof_b200 import TFMEncoderB200
from oracle.vade_oracle import default_adjacency

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, T, D = 14, 25, 16
adj = default_adjacency(N)
E = int(np.count_nonzero(np.triu(adj)))
m = TFMEncoderB200((T, N, 3), (T, E, 1), adj, D, seed=1, max_batch=B)
x = torch.randn(B, T, N, 3, device="cuda")
a = torch.randn(B, T, E, 1, device="cuda")
for _ in range(3):
    m(x, a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    m(x, a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"tfm encode B={B}: {ms:.3f} ms/batch, {B / ms * 1e3:.0f} windows/s")
