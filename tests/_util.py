"""Helpers shared by the tests: golden loading and error metrics."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    inputs, params = {}, {}
    for k in z.files:
        a = z[k]
        if k.startswith("in."):
            inputs[k[3:]] = torch.from_numpy(a.astype(np.float32))
        elif k.startswith("p."):
            params[k[2:]] = torch.from_numpy(a if a.dtype.kind in "iu" else a.astype(np.float32))
    return inputs, params, torch.from_numpy(z["y_ref"])


def rel_fro(y, ref):
    y, ref = y.double(), ref.double()
    return ((y - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def rel_max(y, ref):
    y, ref = y.double(), ref.double()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
