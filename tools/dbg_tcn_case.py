"""Debug helper: run one tcnstep / tcnvade golden on the GPU and print where the gradient differs from the reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import load_golden_of, sub, rel_l2
from deepof_b200 import VQVAEB200

case = sys.argv[1] if len(sys.argv) > 1 else "vq_cfg3"
g = load_golden_of("tcnstep", case)
T, N, E, D, K, B = (int(v) for v in g["meta"])
m = VQVAEB200((T, N, 3), (T, E, 1), g["adjacency"], D, K, encoder_type="TCN", beta=float(g["beta"]), max_batch=B, training=True, seed=1)
m.load_state_dict(sub(g, "p/"))
m.loss_grad(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
gd = m.grad_dict()
torch.set_printoptions(precision=5, linewidth=200, sci_mode=False)
for k in ("encoder.edge_tcn.blocks.0.conv1.weight", "encoder.edge_tcn.blocks.0.conv1.bias", "encoder.edge_tcn.blocks.0.bn1.weight",
          "encoder.edge_tcn.blocks.0.bn1.bias", "encoder.edge_tcn.blocks.0.downsample.weight", "encoder.edge_tcn.blocks.0.downsample.bias",
          "encoder.edge_tcn.blocks.0.bn2.weight", "encoder.edge_tcn.blocks.1.conv1.weight"):
    r = torch.from_numpy(g["g/" + k]).squeeze()
    o = gd[k].cpu().squeeze()
    print(k, "rel", rel_l2(o, r), "ref norm", float(r.norm()))
    if r.numel() <= 128:
        d = (o - r)
        print("  ref ", r.flatten()[:32])
        print("  diff", d.flatten()[:32])
# per output channel error of conv2.weight of block 0
r = torch.from_numpy(g["g/encoder.edge_tcn.blocks.0.conv2.weight"]); o = gd["encoder.edge_tcn.blocks.0.conv2.weight"].cpu()
print("conv2.weight err per out channel", ((o - r).flatten(1).norm(dim=1) / r.flatten(1).norm(dim=1).clamp_min(1e-9)))
print("conv2.weight err per in channel", ((o - r).transpose(0, 1).flatten(1).norm(dim=1) / r.transpose(0, 1).flatten(1).norm(dim=1).clamp_min(1e-9)))
print("conv2.weight err per tap", ((o - r).permute(2, 0, 1).flatten(1).norm(dim=1) / r.permute(2, 0, 1).flatten(1).norm(dim=1).clamp_min(1e-9)))

# ---- forward intermediates of the edge stack, block 0, vs a CPU fp32 / fp64 evaluation
from oracle import tcn_oracle as TC, vade_oracle as O
p = sub(g, "p/")
a = torch.from_numpy(g["a"])
xe = O.group_reshape(a).reshape(B * E, T, 1).transpose(1, 2)
pre = "encoder.edge_tcn.blocks.0."
for dt in (torch.float32, torch.float64):
    pp = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in p.items()}
    st = {}
    A1 = TC.causal_conv(xe.to(dt), pp[pre + "conv1.weight"], pp[pre + "conv1.bias"], 1)
    Y1 = torch.relu(TC.batch_norm(A1, pp, pre + "bn1.", True, st, (0, 2)))
    A2 = TC.causal_conv(Y1, pp[pre + "conv2.weight"], pp[pre + "conv2.bias"], 1)
    Y2p = TC.batch_norm(A2, pp, pre + "bn2.", True, st, (0, 2))
    gA1 = m.debug("edge_a1")[:B * E * T * 32].view(B * E, T, 32).transpose(1, 2).cpu()
    gA2 = m.debug("edge_a2")[:B * E * T * 32].view(B * E, T, 32).transpose(1, 2).cpu()
    print(dt, "A1 rel", rel_l2(gA1, A1), "A2 rel", rel_l2(gA2, A2))
    mu, var, n = st[pre + "bn2."]
    gy = (gA2.to(dt) - mu.view(1, -1, 1)) / torch.sqrt(var.view(1, -1, 1) + 1e-3) * pp[pre + "bn2.weight"].view(1, -1, 1) + pp[pre + "bn2.bias"].view(1, -1, 1)
    flips = ((gy > 0) != (Y2p > 0))
    print("  relu flips per channel (gpu A2 through the CPU statistics):", flips.sum(dim=(0, 2)).tolist())
    print("  active fraction ch21:", float((Y2p[:, 21] > 0).float().mean()), "min |y2pre| ch21:", float(Y2p[:, 21].abs().min()))
    print("  var ch21", float(var[21]), "mean", float(mu[21]), "A2 ch21 abs err max", float((gA2[:, 21].to(dt) - A2[:, 21]).abs().max()))
