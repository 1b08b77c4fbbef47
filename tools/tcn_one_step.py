"""One TCN VaDE training step at a given batch (profiling target: ncu -k regex:... python tools/tcn_one_step.py B steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepof_b200.training import VaDETrainer
from oracle import vade_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T, N, D, K = 25, 14, 16, 8
adj = O.default_adjacency(N)
E = int(np.count_nonzero(np.triu(adj)))
x, a = O.synthetic_windows(256, T, adj, seed=3)
x = x.repeat(B // 256, 1, 1, 1).cuda(); a = a.repeat(B // 256, 1, 1, 1).cuda()
x += 0.01 * torch.randn_like(x)
tr = VaDETrainer((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, seed=3, encoder_type="TCN")
tr.set_phase("main", kl_weight=0.8, lr_base=5e-4, lr_gmm=2e-4)
for i in range(steps):
    tr.train_step_device(x, a)
torch.cuda.synchronize()
print("ok", float(tr.model.logs[0]))
