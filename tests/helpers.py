"""Shared helpers for the test-suite (golden loading, error metrics)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(f)[len("vade_"):-len(".npz")]
                  for f in glob.glob(os.path.join(GOLDEN_DIR, "vade_*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"vade_{name}.npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    g["dims"] = dict(T=T, N=N, E=E, D=D, K=K, B=B)
    return g


def sub(g, prefix, as_torch=True, dtype=torch.float32):
    out = {}
    for k, v in g.items():
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(np.array(v)).to(dtype) if as_torch and v.dtype.kind == "f" \
                else (torch.from_numpy(np.array(v)) if as_torch else v)
    return out


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).flatten()
    b = torch.as_tensor(b, dtype=torch.float64).flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden_cases_of(kind):
    return sorted(os.path.basename(f)[len(kind) + 1:-len(".npz")]
                  for f in glob.glob(os.path.join(GOLDEN_DIR, f"{kind}_*.npz")))


def load_golden_of(kind, name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{kind}_{name}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}
