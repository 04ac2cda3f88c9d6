#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Run:  python oracle/make_golden.py
Needs /root/reference (read-only mount).  For every case it
  1. builds the real reference module (vision_transformers/{ViT,pvt,cvt,cswin,xcit}.py),
     eval() mode, parameters / BN statistics / temperature randomised, and all
     inputs+parameters rounded once to fp16-representable values,
  2. runs the reference forward in fp32 on CPU  -> ``y_ref``,
  3. checks the oracle restatement against it (max-abs <= 2e-6 * max|y|-ish),
  4. stores inputs (fp16), parameters (fp16 except integer buffers) and y_ref (fp32).
The committed vectors are what pins the oracle on machines without the reference
(the GPU box), see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PA_REFERENCE", "/root/reference/vision_transformers")

import oracle  # noqa: E402
from oracle.cases import GOLDEN_CASES, build_reference_case, run_oracle_case  # noqa: E402


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not mounted at {REF}")
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, spec in GOLDEN_CASES.items():
        case = build_reference_case(spec, REF)
        y_or = run_oracle_case(spec, case["inputs"], case["params"])
        y_ref = case["y_ref"]
        err = (y_or - y_ref).abs().max().item()
        ref_mag = y_ref.abs().max().item()
        print(f"{name:28s} out{tuple(y_ref.shape)} max|y|={ref_mag:.4f} oracle-vs-reference max-abs={err:.3e}")
        assert err <= 5e-6 * max(1.0, ref_mag), f"oracle restatement disagrees with reference on {name}"
        blob = {}
        for k, v in case["inputs"].items():
            blob["in." + k] = v.numpy().astype(np.float16)
        for k, v in case["params"].items():
            if v.dtype in (torch.int64, torch.int32):
                blob["p." + k] = v.numpy()
            else:
                blob["p." + k] = v.numpy().astype(np.float16)
        blob["y_ref"] = y_ref.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
    print("wrote", len(GOLDEN_CASES), "golden files to", out_dir)


if __name__ == "__main__":
    main()
