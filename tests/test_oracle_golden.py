"""The oracle restatement must reproduce the committed reference outputs (CPU, fp32).

The golden vectors in tests/golden/ were produced by the REAL reference modules
(oracle/make_golden.py, run in the build container); this is what pins the oracle.
"""
import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_CASES, run_oracle_case
from oracle import cswin_window_table
from _util import golden_names, load_golden


def test_all_golden_cases_present():
    assert set(golden_names()) == set(GOLDEN_CASES)


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_matches_golden(name):
    inputs, params, y_ref = load_golden(name)
    y = run_oracle_case(GOLDEN_CASES[name], inputs, params)
    assert y.shape == y_ref.shape
    tol = 5e-6 * max(1.0, y_ref.abs().max().item())
    assert (y - y_ref).abs().max().item() <= tol


@pytest.mark.parametrize("name", ["vit_b2_n197_c128_h2", "xca_b2_n196_c128_h2"])
def test_oracle_fp64_agrees(name):
    """fp64 evaluation of the restatement stays within fp32 round-off of the golden output."""
    inputs, params, y_ref = load_golden(name)
    y = run_oracle_case(GOLDEN_CASES[name], inputs, params, dtype=torch.float64)
    assert (y.float() - y_ref).abs().max().item() <= 5e-6


@pytest.mark.parametrize("reso,idx,split", [(14, 0, 7), (14, 1, 7), (7, -1, 7), (56, 0, 7), (56, 1, 7), (8, 0, 2)])
def test_window_table_is_a_permutation(reso, idx, split):
    """Index path (cswin.py:199-216) is integer work: bit-exact, a bijection of the L tokens."""
    T = cswin_window_table(reso, idx, split)
    assert T.dtype == np.int64
    assert sorted(T.reshape(-1).tolist()) == list(range(reso * reso))
    # closed form: window (i,j), in-window (r,c)
    H_sp, W_sp = {(-1): (reso, reso), 0: (reso, split), 1: (split, reso)}[idx]
    nJ = reso // W_sp
    w, t = 3 % T.shape[0], 5 % T.shape[1]
    i, j, r, c = w // nJ, w % nJ, t // W_sp, t % W_sp
    assert T[w, t] == (i * H_sp + r) * reso + (j * W_sp + c)


def test_window_table_bad_idx():
    with pytest.raises(ValueError):
        cswin_window_table(14, 2, 7)


def test_aten_port_matches_oracle():
    """The CPU baseline port (same ATen op sequence as the reference) agrees with the einsum oracle."""
    from oracle import vit_attention
    from oracle.aten_port import vit_attention_aten
    torch.manual_seed(0)
    x = torch.randn(2, 50, 128)
    wq, bq = torch.randn(384, 128) * 0.1, torch.randn(384) * 0.1
    wp, bp = torch.randn(128, 128) * 0.1, torch.randn(128) * 0.1
    a = vit_attention(x, wq, bq, wp, bp, 2)
    b = vit_attention_aten(x, wq, bq, wp, bp, 2)
    assert (a - b).abs().max().item() < 1e-5


def test_aten_port_matches_oracle_at_baseline_config_1():
    """The CPU arm of bench.py at BASELINE.json configs[0] exactly (B=2, N=197, dim=768, 12 heads) and with the bench's own
    parameter generator: the timed port and the oracle agree to fp32 round-off, so the CPU number is a number for the reference's math."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from oracle import vit_attention
    from oracle.aten_port import vit_attention_aten
    sd = bench.make_cpu_model()
    torch.manual_seed(1)
    x = torch.randn(2, 197, 768).half().float()
    a = vit_attention(x, sd["qkv.weight"], None, sd["proj.weight"], sd["proj.bias"], 12)
    b = vit_attention_aten(x, sd["qkv.weight"], None, sd["proj.weight"], sd["proj.bias"], 12)
    assert (a - b).abs().max().item() <= 2e-6 * max(1.0, a.abs().max().item())
