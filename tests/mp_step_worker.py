"""Worker of tests/test_multigpu_gpu.py (one process per GPU under torch.distributed.run, NCCL): three data-parallel
VaDE training steps through VaDETrainer; rank 0 checks

* the all-reduced gradient of step 1 (sum over ranks, 1/world folded into the clip+Adam kernel) against the CPU oracle:
  mean over ranks of the rank-LOCAL gradients (every batch statistic is rank-local, as under the reference's DDP),
* the post-Adam parameters after step 1 against the oracle's clip + Adam on that mean gradient,
* parameters bit-identical on all ranks after 3 steps.

Writes a JSON verdict to argv[1]."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from deepof_b200.training import VaDETrainer
    from oracle import vade_oracle as O
    T, N, D, K, B = 25, 14, 16, 8, 64
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    x, a = O.synthetic_windows(B * world, T, adj, seed=11)
    g = torch.Generator().manual_seed(2)
    eps, mc = torch.randn(world, 3, B, D, generator=g), torch.randn(world, 3, 32, B, D, generator=g)
    tr = VaDETrainer((T, N, 3), (T, E, 1), adj, D, K, max_batch=B, seed=5 + rank, world_size=world, rank=rank)   # broadcast from rank 0
    tr.set_phase("main", kl_weight=0.7, lr_base=5e-4, lr_gmm=2e-4)
    with torch.no_grad():
        tr.model.latent_space.gmm_means.mul_(3.0)
    dist.broadcast(tr.model.state, src=0)
    p0 = {k: v.cpu() for k, v in tr.model.state_dict().items()}
    xs, as_ = x[rank * B:(rank + 1) * B].cuda(), a[rank * B:(rank + 1) * B].cuda()
    res = {}
    for step in range(3):
        tr.train_step_device(xs, as_, eps=eps[rank, step], mc_eps=mc[rank, step])
        if step == 0:
            grad1 = tr.model.grad.clone()                      # all-reduced SUM over ranks
            state1 = tr.model.state.clone()
    torch.cuda.synchronize()
    gathered = [torch.empty_like(tr.model.state) for _ in range(world)]
    dist.all_gather(gathered, tr.model.state)
    if rank == 0:
        res["params_bit_identical_after_3_steps"] = all(torch.equal(gathered[0], t) for t in gathered[1:])
        graph = O.graph_operators(adj)
        ocfg = O.LossCfg.main_defaults(K, 0.7)
        grads = []
        for r in range(world):
            _, gr, _ = O.train_step(x[r * B:(r + 1) * B], a[r * B:(r + 1) * B], p0, graph, D, ocfg, eps=eps[r, 0], mc_eps=mc[r, 0])
            grads.append(gr)
        names = [k for k, v in grads[0].items() if v is not None]
        mean = {k: sum(gr[k] for gr in grads) / world for k in names}
        lay = {k: (off, n, shape) for k, off, n, shape, grp in tr.model.layout}
        got = {k: (grad1[lay[k][0]:lay[k][0] + lay[k][1]].cpu() / world).view(lay[k][2]) for k in names}
        num = sum(float((got[k].double() - mean[k].double()).pow(2).sum()) for k in names)
        den = sum(float(mean[k].double().pow(2).sum()) for k in names)
        res["allreduced_grad_rel_l2_vs_oracle_mean"] = (num / den) ** 0.5
        p1 = {k: v.clone() for k, v in p0.items()}
        O.adam_step(p1, {k: mean.get(k) for k in p1}, {}, 5e-4, 2e-4)
        worst = 0.0
        for k in names:
            off, n, shape = lay[k]
            worst = max(worst, float((state1[off:off + n].cpu().view(shape) - p1[k]).abs().max()))
        res["post_adam_worst_abs_diff"] = worst
        res["world"] = world
        res["allreduce"] = tr.exchange.kind
        res["allreduce_fallback_reason"] = tr.exchange.why
        res["nccl_version"] = ".".join(str(v) for v in torch.cuda.nccl.version())
        with open(out, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
