#!/usr/bin/env python
"""bench.py — pose-windows/sec trained (VaDE / GRU, window 25 x 14 body parts), BASELINE.json cfg2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full VaDE training step (main phase: MC-KL with 32 samples) on one batch of
4096 synthetic windows per GPU: forward + VadeLoss + backward + (all-reduce when N>1) +
clip_grad_value_(0.75) + Adam.  Prints ONE JSON line (rank 0).

  value : windows/s with the inputs already resident in HBM as RAW pose frames of a 1M-window video;
          every step = loader kernel (frames -> standardised windows) + the training step on consecutive
          batches (every step touches ~16 GB of activations >> 126 MB L2).
  e2e   : the same step driven through the public host API (VaDETrainer.train_steps, the host epoch
          loop) from PINNED HOST buffers: every step's H2D copy of its batch (side stream, overlapped
          with the previous step) and D2H read of its loss are inside the timed region.
  roofline / kernels : per-kernel-class CUDA-event timing (library-side events around every
          launch) taken on extra steps right after the timed region.
  cpu_baseline : the CPU oracle (ATen-GRU variant, oracle/vade_oracle.py) on this host's cores
          on a bounded sample (rank 0, N=1 only).
--impl reference runs only that CPU arm (the reference is Python and cannot travel to the GPU
box; its arithmetic is restated in oracle/, pinned to reference-generated golden vectors).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG = dict(T=25, N=14, E=14, F=3, Fe=1, D=16, K=8, batch=4096, pool_windows=1 << 20)
FLOPS_PER_WINDOW = 88.6e6        # useful matmul/conv FLOPs fwd+bwd, cfg2
WORKLOAD = "cfg2: VaDE GRU encoder, 1M synthetic windows (25x14x3 + 25x14x1), latent=16, 8 clusters, batch 4096/GPU, main phase (MC-KL S=32)"
# secondary workloads (python bench.py --workload ...): same contract, not the headline
WORKLOADS = {
    "cfg2": dict(CFG, kind="vade", flops=88.6e6, name=WORKLOAD, metric="pose-windows/sec trained (VaDE, win=25x28)"),
    "vqvae": dict(CFG, K=64, kind="vqvae", flops=92.6e6,
                  name="cfg3r: VQ-VAE with the RECURRENT encoder/decoder (the transformer cfg3 is its own line), 1M synthetic "
                       "windows (25x14x3 + 25x14x1), latent=16, codebook=64, batch 4096/GPU",
                  metric="pose-windows/sec trained (VQ-VAE recurrent, win=25x28)"),
    "cfg3": dict(CFG, K=64, kind="vqvae", encoder="transformer", flops=180.1e6,
                 name="cfg3: VQ-VAE transformer encoder / causal transformer decoder, 1M synthetic windows (25x14x3 + 25x14x1), latent=16, "
                      "codebook=64, batch 4096/GPU, dropout on (in-kernel Philox)",
                 metric="pose-windows/sec trained (VQ-VAE transformer, win=25x28)"),
    "cfg5": dict(CFG, D=64, K=16, kind="vade", encoder="transformer", flops=263.2e6, pool_windows=1 << 21,
                 name="cfg5: VaDE transformer encoder / causal transformer decoder, synthetic windows (25x14x3 + 25x14x1), latent=64, "
                      "16 clusters, batch 4096/GPU, main phase (MC-KL S=32), dropout on (in-kernel Philox)",
                 metric="pose-windows/sec trained (VaDE transformer, win=25x28)"),
    "tcn": dict(CFG, kind="vade", encoder="TCN", flops=279.4e6,
                name="tcn: VaDE TCN encoder / TCN decoder (cfg2 geometry with encoder_type=TCN), 1M synthetic windows (25x14x3 + 25x14x1), "
                     "latent=16, 8 clusters, batch 4096/GPU, main phase (MC-KL S=32)",
                metric="pose-windows/sec trained (VaDE TCN, win=25x28)"),
    "cfg4": dict(T=50, N=22, E=26, F=3, Fe=1, D=16, K=1, batch=4096, pool_windows=1 << 19, kind="contrastive", flops=290.4e6,
                 name="cfg4: contrastive NT-Xent (nce, cosine, tau=0.1), recurrent encoder, 2 animals x 11 body parts (N=22, E=26), "
                      "win=50 (encoder sees 25), latent=16, batch 4096/GPU, reference-default augmentations",
                 metric="pose-windows/sec trained (contrastive NT-Xent, win=50x44)"),
}


def adjacency(n):
    A = np.zeros((n, n))
    for i in range(n - 1):
        A[i, i + 1] = A[i + 1, i] = 1.0
    if n > 5:
        A[0, 5] = A[5, 0] = 1.0
    return A


def adjacency_two_animals():
    """2 x 11 body parts, 12 within-animal edges each + 2 cross-animal edges = 26 (cfg4 sizes)."""
    A = np.zeros((22, 22))
    for o in (0, 11):
        for i in range(10):
            A[o + i, o + i + 1] = A[o + i + 1, o + i] = 1.0
        for i, j in ((0, 5), (2, 8)):
            A[o + i, o + j] = A[o + j, o + i] = 1.0
    for i, j in ((0, 11), (0, 21)):
        A[i, j] = A[j, i] = 1.0
    return A


def synth_pool(n, T, adj, seed, device):
    """Standardised xy/speed ~ N(0,1); edges = standardised log1p distances recomputed from x
    (SURVEY 8d).  Generated on `device` in chunks."""
    g = torch.Generator(device=device).manual_seed(seed)
    N = adj.shape[0]
    rows, cols = np.nonzero(np.triu(adj))
    rows_t, cols_t = torch.as_tensor(rows, device=device), torch.as_tensor(cols, device=device)
    x = torch.empty(n, T, N, 3, device=device)
    a = torch.empty(n, T, len(rows), 1, device=device)
    for s in range(0, n, 65536):
        e = min(n, s + 65536)
        x[s:e] = torch.randn(e - s, T, N, 3, generator=g, device=device)
        d = (x[s:e, :, rows_t, :2] - x[s:e, :, cols_t, :2]).norm(dim=-1)
        a[s:e, ..., 0] = torch.log1p(d)
    a = (a - 0.9) / 0.45   # fixed standardisation constants (mean/std of log1p|N(0,2I)|)
    return x, a


def synth_frames(n_frames, N, seed, device):
    """Raw pose frames [n_frames, N, 2] (pixels) of one synthetic video: a mouse-like rigid body doing a
    random walk with heading drift + per-point tracking noise.  The loader kernel turns them into the
    standardised windows the model trains on."""
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=device, dtype=torch.float64)
    centre = torch.cumsum(rn(2, n_frames) * 1.5, 1).t().contiguous()    # scan along the contiguous axis
    centre = centre - centre.mean(0) + 300.0
    heading = torch.cumsum(rn(n_frames) * 0.08, 0)
    body = rn(N, 2) * 18.0
    body[0] = torch.tensor([0.0, 0.0], dtype=torch.float64)
    body[1] = torch.tensor([0.0, 40.0], dtype=torch.float64)
    body[2] = torch.tensor([0.0, -36.0], dtype=torch.float64)
    c, s = torch.cos(heading), torch.sin(heading)
    px = c[:, None] * body[None, :, 0] - s[:, None] * body[None, :, 1]
    py = s[:, None] * body[None, :, 0] + c[:, None] * body[None, :, 1]
    pts = torch.stack([px, py], -1) + centre[:, None, :] + rn(n_frames, N, 2) * 0.4
    return pts.float().contiguous()


class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed regions (B200_PROFILING.md's clocks line).
    NVML in-process every 10 ms (a 120 ms timed region is too short for `nvidia-smi -lms`, whose first row can
    arrive after the region has ended); `nvidia-smi` loop as the fallback when pynvml cannot initialise."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.h, self.nv = index, [], None, None, None
        self.sm, self.pw, self.mx, self.bits, self.source = [], [], None, 0, None
        self.stop_ev, self.armed = threading.Event(), threading.Event()

    def arm(self):
        """Samples count from here on.  start() (NVML initialisation / the nvidia-smi fork: 100+ ms on a multi-GPU box) runs
        BEFORE the barrier that opens the timed region — done after it on rank 0 only, the other ranks' first timed step waited
        for rank 0 inside the gradient exchange and the max-over-ranks step time carried the delay."""
        self.armed.set()

    def _nvml_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip()]
        if ids and all(v.strip().isdigit() for v in ids) and self.index < len(ids):
            return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self._nvml_index())
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self._nvml_index()), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_ev.is_set():
            if not self.armed.is_set():
                self.armed.wait(0.01)
                continue
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            self.stop_ev.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            if self.armed.is_set():
                self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_ev.set()
            self.th.join(timeout=2)
            nv = self.nv
            masks = [getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)]
            reasons = [n for n, m in zip(self.NAMES, masks) if self.bits & m]
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                    "sm_max_mhz": self.mx, "power_w_max": max(self.pw) if self.pw else None, "samples": len(self.sm),
                    "reasons": reasons, "source": "nvml, 10 ms, value + e2e timed regions"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons, "source": "nvidia-smi -lms 50"}


def cpu_reference_arm(steps, warmup, batch=256, workload="cfg2"):
    """The reference path restated on CPU (oracle; ATen GRU like the reference's nn.GRU, torch SDPA-free transformer):
    forward + loss + backward + clip + Adam on `batch` windows per step.  Imports nothing of the product."""
    from oracle import vade_oracle as O
    from oracle import models_oracle as MO
    from oracle import tfm_oracle as TO
    from oracle import params as OP
    O.USE_ATEN_GRU = True
    c = WORKLOADS[workload]
    kind, enc = c["kind"], c.get("encoder", "recurrent")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    adj = adjacency_two_animals() if kind == "contrastive" else adjacency(c["N"])
    graph = O.graph_operators(adj)
    p = OP.random_params(kind, enc, c["N"], c["E"], c["F"], c["Fe"], c["D"], c["K"], graph, seed=1234 + 2)
    x, a = synth_pool(batch * 4, c["T"], adj, 1234 + 2, "cpu")
    cfg = O.LossCfg.main_defaults(c["K"], 0.8)
    rows, cols = np.nonzero(np.triu(adj))
    ei = torch.from_numpy(np.stack([rows, cols], 1))
    rot = MO.rotation_table(ei.numpy(), c["N"])
    gen = torch.Generator().manual_seed(7)

    def tfm_masks():
        dk = p["encoder.node_tf.embed.weight"].shape[0]
        mk = {}
        for core, S in (("node", batch * c["N"]), ("edge", batch * c["E"])):
            for nm, shp in TO.dropout_mask_shapes(S, c["T"], dk, 4, 2):
                mk[f"{core}.{nm}"] = (torch.rand(shp, generator=gen) >= 0.1).float()
        for pre in ("dec.", "dec1."):
            for nm, shp in TO.decoder_mask_shapes(batch, c["T"], 4 * c["D"], 8, 128, 2):
                mk[nm.replace("dec.", pre)] = (torch.rand(shp, generator=gen) >= 0.2).float()
        return mk

    state = {}
    times = []
    for i in range(warmup + steps):
        s = (i % 4) * batch
        t0 = time.perf_counter()
        if enc == "TCN":
            from oracle import tcn_oracle as TCO
            eps, mc = torch.randn(batch, c["D"], generator=gen), torch.randn(32, batch, c["D"], generator=gen)
            logs, grads, _ = TCO.vade_train_step(x[s:s + batch], a[s:s + batch], p, graph, cfg, eps, mc_eps=mc)
            O.adam_step(p, grads, state, 5e-4, 2e-4)
        elif enc == "transformer":
            mk = tfm_masks()                       # the reference draws its dropout masks inside the step as well
            if kind == "vade":
                eps, mc = torch.randn(batch, c["D"], generator=gen), torch.randn(32, batch, c["D"], generator=gen)
                logs, grads, _ = TO.vade_train_step(x[s:s + batch], a[s:s + batch], p, graph, cfg, mk, eps, mc_eps=mc)
                O.adam_step(p, grads, state, 5e-4, 2e-4)
            else:
                logs, grads, _ = TO.vqvae_train_step(x[s:s + batch], a[s:s + batch], p, graph, mk, 1.0, 0.0)
                MO.adam_step_generic(p, grads, state, 1e-3)
        elif kind == "vade":
            logs, grads, _ = O.train_step(x[s:s + batch], a[s:s + batch], p, graph, c["D"], cfg)
            O.adam_step(p, grads, state, 5e-4, 2e-4)
        elif kind == "vqvae":
            logs, grads, _ = MO.vqvae_train_step(x[s:s + batch], a[s:s + batch], p, graph, c["D"], 1.0, 0.0)
            MO.adam_step_generic(p, grads, state, 1e-3)
        else:
            prm = MO.draw_aug_params(batch, c["T"], c["N"], MO.AugCfg(), rot)
            logs, grads, _ = MO.contrastive_train_step(x[s:s + batch], p, graph, c["D"], ei, prm, 0.1)
            MO.adam_step_generic(p, grads, state, 1e-3)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.median(times))
    return {"value": batch / (ms / 1e3), "ms_per_step": ms, "cores": cores, "batch": batch,
            "sample": f"median of {steps} steps x {batch} windows per step (the reference's best batch-size class, SURVEY 8d) of the {workload} "
                      f"workload (fwd+loss+bwd+clip+Adam) after {warmup} warm-up steps, torch {torch.__version__} CPU, {cores} threads"}


def workload_config(c, B, world, pool_n):
    """The `config` object — identical keys in the b200 and the reference arm."""
    return {"workload": c["name"], "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "pool_windows_per_gpu": pool_n,
            "inputs": "raw pose frames resident in HBM; windows built per step by the loader kernel",
            "l2": "inputs+activations per step (GBs) exceed the 126 MB L2; consecutive batches of the video"}


def pool_size(c, world):
    pool_n = c["pool_windows"] // world
    return max(c["batch"], pool_n // c["batch"] * c["batch"])


def run_workload(c, steps, warmup, world, rank, local, dev, sample_clocks=True, with_kernels=True):
    """Measure one workload on this rank's GPU (all ranks call it together): value, e2e, per-kernel-class timing."""
    import ctypes as C
    import torch.distributed as dist
    from deepof_b200 import _lib
    from deepof_b200.training import VaDETrainer, VQVAETrainer, ContrastiveTrainer
    from deepof_b200 import WindowLoader
    kind, enc = c["kind"], c.get("encoder", "recurrent")
    adj = adjacency_two_animals() if kind == "contrastive" else adjacency(c["N"])
    B, T, D, K = c["batch"], c["T"], c["D"], c["K"]
    shapes = ((T, c["N"], c["F"]), (T, c["E"], c["Fe"]))
    if kind == "vade":
        trainer = VaDETrainer(*shapes, adj, D, K, max_batch=B, seed=1234 + 2, world_size=world, rank=rank, encoder_type=enc)
        trainer.set_phase("main", kl_weight=0.8, lr_base=5e-4, lr_gmm=2e-4)
    elif kind == "vqvae":
        trainer = VQVAETrainer(*shapes, adj, D, K, max_batch=B, seed=1234 + 2, world_size=world, rank=rank, encoder_type=enc)
    else:
        trainer = ContrastiveTrainer(*shapes, adj, D, max_batch=B, seed=1234 + 2, world_size=world, rank=rank, encoder_type=enc)
    xbuf = trainer._xf if kind == "contrastive" else trainer._xs
    abuf = torch.empty(B, T, c["E"], c["Fe"], device=dev) if kind == "contrastive" else trainer._as
    pool_n = pool_size(c, world)
    # one synthetic video per rank, resident in HBM as RAW frames; the loader kernel produces each batch
    rows, cols = np.nonzero(np.triu(adj))
    edges = np.stack([rows, cols], 1).astype(np.int32)
    frames = synth_frames(pool_n + T - 1, c["N"], 1234 + 2 + 1000 * rank, dev)
    loader = WindowLoader([frames], edges, T, 1, nose=1, tail_base=2, center_node=-1, align_node=0,
                          arena_center=(300.0, 300.0), fps=25.0)
    assert len(loader) == pool_n
    nb = pool_n // B
    # pinned host copy of the first batches (materialised windows, the reference's step_fn input) for the e2e leg
    hb = min(nb, 16)
    xh = torch.empty((hb * B, T, c["N"], c["F"]), pin_memory=True)
    ah = torch.empty((hb * B, T, c["E"], c["Fe"]), pin_memory=True)
    for i in range(hb):
        xb, ab = loader.load(i * B, B)
        xh[i * B:(i + 1) * B].copy_(xb); ah[i * B:(i + 1) * B].copy_(ab)
    torch.cuda.synchronize()
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident(n, first):
        for i in range(n):
            s = ((first + i) % nb) * B
            xb, ab = loader.load(s, B, xbuf, abuf)      # frames -> windows on the device
            trainer.train_step_device(xb, ab)

    def run_e2e(n, first):
        # the public host loop: every step copies ITS batch from pinned host memory (on a side stream, overlapped with the
        # previous step's kernels) and reads ITS loss back (non-blocking D2H); one host synchronisation at the end
        bl = []
        for i in range(n):
            s = ((first + i) % hb) * B
            bl.append((xh[s:s + B], ah[s:s + B]))
        return trainer.train_steps(bl)[-1]

    # ---- value: device-resident inputs
    run_resident(warmup, 0)
    sampler = ClockSampler(local) if sample_clocks else None
    if rank == 0 and sampler:
        sampler.start()                  # slow part (NVML init / fork) outside the timed region and before the barrier
    barrier()
    if rank == 0 and sampler:
        sampler.arm()
    l0 = L.dof_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_resident(steps, warmup)
    e1.record()
    barrier()
    launches = L.dof_launch_count() - l0
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world / (ms / 1e3)

    # ---- e2e: pinned host buffers through the public API
    run_e2e(2, 0)
    barrier()
    e0.record()
    loss = run_e2e(steps, 2)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    clocks = sampler.stop() if (rank == 0 and sampler) else None
    h2d = (xh[:B].numel() + (0 if kind == "contrastive" else ah[:B].numel())) * 4   # contrastive recomputes the edges
    e2e = {"value": B * world / (ms_e2e / 1e3), "unit": "windows/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "last_loss": loss}

    # ---- per-kernel-class timing (library-side CUDA events around each launch), 3 extra steps
    kernels, roof = None, None
    if with_kernels:
        psteps = 3
        L.dof_set_concurrency(0)         # per-kernel events need non-overlapping kernels: one stream for this leg only
        if rank == 0:
            L.dof_profile_begin()
        run_resident(psteps, 0)          # every rank steps (the gradient all-reduce is collective); rank 0 records events
        barrier()
        L.dof_set_concurrency(1)
        if rank == 0:
            buf = C.create_string_buffer(16384)
            L.dof_profile_end(buf, 16384)
            kernels = {}
            for ln in buf.value.decode().strip().split("\n"):
                name, cnt, tot, fl, by = ln.split()
                kernels[name] = {"launches_per_step": int(cnt) / psteps, "ms_per_step": float(tot) / psteps,
                                 "flops_per_step": float(fl) / psteps, "bytes_per_step": float(by) / psteps}
            tot_ms = sum(k["ms_per_step"] for k in kernels.values())
            for k in kernels.values():
                k["share"] = k["ms_per_step"] / tot_ms
                if k["flops_per_step"] > 0:
                    k["tflops"] = k["flops_per_step"] / (k["ms_per_step"] / 1e3) / 1e12
                if k["bytes_per_step"] > 0:
                    k["gbs"] = k["bytes_per_step"] / (k["ms_per_step"] / 1e3) / 1e9
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            top = max(kernels, key=lambda n: kernels[n]["ms_per_step"])
            roof = roofline_for(top, kernels[top], peaks, wl="tcn" if enc == "TCN" else None)
            sustained = peaks.get("bf16_tflops_sustained", 1400.0)
            roof["step_useful_tflops"] = c["flops"] * B / (ms / 1e3) / 1e12
            roof["step_tensor_frac"] = roof["step_useful_tflops"] / sustained
    exch = getattr(trainer, "exchange", None)
    allreduce = exch.kind if exch is not None else None
    del trainer, loader, frames, xh, ah, exch
    torch.cuda.empty_cache()
    return dict(value=value, ms=ms, e2e=e2e, launches=int(launches), clocks=clocks, kernels=kernels, roof=roof, pool_n=pool_n,
                allreduce=allreduce)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=CFG["batch"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    c = dict(WORKLOADS[args.workload])
    if args.batch != CFG["batch"] or args.workload == "cfg2":
        c["batch"] = args.batch
    METRIC = c["metric"]

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(max(1, args.steps), max(3, args.warmup), workload=args.workload)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"],
                "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": workload_config(c, c["batch"], args.gpus, pool_size(c, args.gpus)),
                "note": "reference CPU path restated (oracle port; the Python reference cannot travel to the GPU box); each CPU step is a "
                        "bounded sample of the workload, see cpu_baseline.sample",
                "cpu_baseline": {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = run_workload(c, args.steps, warmup, world, rank, local, dev)
    B = c["batch"]

    # short runs of the other BASELINE configs (same contract, fewer steps): driver-visible, not the headline
    secondary = []
    if not args.no_secondary and args.workload == "cfg2":
        for wl in ("cfg3", "cfg4", "cfg5", "vqvae", "tcn"):
            c2 = dict(WORKLOADS[wl])
            try:
                r2 = run_workload(c2, min(args.steps, 5), 3, world, rank, local, dev, sample_clocks=False)
                rf = r2["roof"] or {}
                secondary.append({"workload": c2["name"], "metric": c2["metric"], "value": r2["value"], "unit": "windows/s",
                                  "ms_per_step": r2["ms"], "steps": min(args.steps, 5), "warmup": 3, "e2e": r2["e2e"],
                                  "gpu_launches": r2["launches"], "batch_per_gpu": c2["batch"],
                                  "step_useful_tflops": rf.get("step_useful_tflops"), "step_tensor_frac": rf.get("step_tensor_frac"),
                                  "top_kernel": rf.get("kernel"), "top_kernel_share": rf.get("kernel_share_of_step"),
                                  "kernels_ms": {k: round(v["ms_per_step"], 4) for k, v in (r2["kernels"] or {}).items()}})
            except Exception as ex:      # a secondary workload must never take the headline down
                secondary.append({"workload": c2["name"], "error": f"{type(ex).__name__}: {ex}"[:300]})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_arm(3, 3, workload=args.workload)
        cpu = {"value": r["value"], "unit": "windows/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {"metric": METRIC, "value": m["value"], "unit": "windows/s",
                "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": m["ms"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(c, B, world, m["pool_n"]),
                "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": m["clocks"], "roofline": m["roof"], "kernels": m["kernels"],
                "allreduce": m["allreduce"] and {"kind": m["allreduce"], "note": "peer_memory = barrier + one-shot reduce over NVLink peer "
                                                 "memory + barrier (csrc/peer.cuh) instead of an NCCL call; nccl = torch.distributed.all_reduce"},
                "kernels_note": "per-kernel-class CUDA-event times of 3 extra steps run on ONE stream (kernels do not overlap); the timed "
                                "steps run the node and edge encoder blocks concurrently on two streams, so ms_per_step is below their sum",
                "cpu_baseline": cpu, "secondary": secondary}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TENSOR_CLASSES = ("gru_", "gemm_", "tfm_attn", "ntxent")   # step GEMMs (GRU, attention, FFN, Linear, conv-as-GEMM), NT-Xent


def roofline_for(name, k, peaks, wl=None):
    """Roofline of the dominant kernel class on the roof SURVEY 8(d) assigns to its row: the tensor pipe (useful
    FLOPs / sustained dense bf16 peak) for the GEMM-class kernels — fused GRU layers, tall-skinny GEMMs, attention —
    and HBM bandwidth for everything else.  FLOPs and bytes are ALGORITHMIC (unpadded), counted by the library at launch
    time over the CUDA-event duration of the class; the other roof's fraction is kept next to it (`hbm_view` /
    `tensor_view`).  `traffic` is the DRAM traffic of the class's dominant launch from the committed ncu --set full
    capture (profiles/ncu_traffic.json), per launch."""
    src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    n = max(1.0, k["launches_per_step"])
    tpeak, hpeak = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("hbm_gbs", 6650.0)
    tfrac = (k["tflops"] / tpeak) if k.get("tflops") else None
    hfrac = (k["gbs"] / hpeak) if k.get("gbs") else None
    traffic, tnote = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        key = f"{wl}:{name}" if wl and f"{wl}:{name}" in tj else name        # workload-specific capture of the class's dominant launch
        if key in tj:
            traffic, tnote = tj[key]["dram_bytes_per_launch"], tj[key]["note"]
    except Exception:
        pass
    common = {"kernel": name, "traffic": traffic, "traffic_note": tnote, "launch_ms": k["ms_per_step"] / n,
              "kernel_share_of_step": k["share"],
              "tensor_view": {"achieved_tflops": k.get("tflops"), "peak_tflops": tpeak, "frac": tfrac,
                              "algorithmic_flops_per_launch": k["flops_per_step"] / n},
              "hbm_view": {"achieved_gbs": k.get("gbs"), "peak_gbs": hpeak, "frac": hfrac,
                           "bytes_per_launch": k["bytes_per_step"] / n,
                           "note": "bytes = the kernel's own operand traffic as designed (activations it reads / writes), not SURVEY 8(d)'s "
                                   "compulsory per-window bytes"}}
    tensor_class = any(name.startswith(p) for p in TENSOR_CLASSES)
    if tensor_class and tfrac is not None:
        return dict(common, bound="tensor", achieved=k["tflops"], peak=tpeak, unit="TFLOP/s", frac=tfrac,
                    peak_source=src + ", sustained dense bf16")
    if hfrac is not None:
        return dict(common, bound="hbm", achieved=k["gbs"], peak=hpeak, unit="GB/s", frac=hfrac, peak_source=src)
    if tfrac is not None:
        return dict(common, bound="tensor", achieved=k["tflops"], peak=tpeak, unit="TFLOP/s", frac=tfrac,
                    peak_source=src + ", sustained dense bf16")
    return dict(common, bound="hbm", achieved=None, peak=hpeak, unit="GB/s", frac=None, peak_source=src)


if __name__ == "__main__":
    main()
