"""C-ABI checks that need no GPU: the library loads, exports every symbol declared in include/pa_b200.h,
the ctypes table covers the header, and argument validation returns error codes (never crashes)."""
import ctypes as C
import os
import re

import pytest

from pytorch_attention_b200 import _lib as L
from pytorch_attention_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = L.load()
    names = header_functions()
    assert "pa_vit_fwd" in names and "pa_gemm_tn" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pa_b200.h but not exported by libpa_b200.so"
        assert n in L.SYMBOLS, f"{n} declared in pa_b200.h but missing from the ctypes table"
    assert sorted(L.SYMBOLS) == names


def test_version_and_error_string():
    lib = L.load()
    assert lib.pa_version() == build.header_version() == 101
    assert isinstance(lib.pa_last_error(), bytes)


def test_vit_argument_validation_without_gpu():
    lib = L.load()
    a = L.VitArgs()
    a.B, a.N, a.C, a.H = 2, 197, 768, 5          # 768 % 5 != 0  (ViT.py:70 assert)
    assert lib.pa_vit_fwd(C.byref(a), None, 0, None) == L.PA_ERR_BAD_SHAPE
    assert b"divisible" in lib.pa_last_error()
    a.H = 32                                      # head_dim 24: not a multiple of 16
    assert lib.pa_vit_fwd(C.byref(a), None, 0, None) == L.PA_ERR_UNSUPPORTED
    a.H = 4                                       # head_dim 192 (the reference's default, ViT.py:67) is served
    assert lib.pa_vit_workspace_bytes(C.byref(a)) > 0
    a.H = 12
    assert lib.pa_vit_workspace_bytes(C.byref(a)) >= 2 * 197 * 768 * 2 * 4
    assert lib.pa_vit_fwd(C.byref(a), None, 0, None) == L.PA_ERR_NULL   # x/weights NULL
    assert lib.pa_vit_fwd(None, None, 0, None) == L.PA_ERR_NULL


def test_gemm_and_attn_argument_validation_without_gpu():
    lib = L.load()
    g = L.GemmArgs()
    assert lib.pa_gemm_tn(C.byref(g), None) == L.PA_ERR_NULL
    g.A = g.B = g.D = 16
    g.M, g.N, g.K, g.Z = 4, 4, 12, 1
    assert lib.pa_gemm_tn(C.byref(g), None) == L.PA_ERR_BAD_SHAPE      # K % 8
    at = L.AttnArgs()
    at.q = at.kv = at.o = 16
    at.G, at.H, at.n_q, at.n_k, at.scale = 1, 1, 8, 300, -1.0
    assert lib.pa_attn_core(C.byref(at), None) == L.PA_ERR_UNSUPPORTED  # scale must be > 0
    at.scale, at.head_dim = 1.0, 40
    assert lib.pa_attn_core(C.byref(at), None) == L.PA_ERR_UNSUPPORTED  # head_dim 40: not a multiple of 16


def test_variant_argument_validation_without_gpu():
    lib = L.load()
    p = L.PvtArgs()
    p.B, p.N, p.C, p.H, p.Himg, p.Wimg, p.sr = 1, 64, 128, 3, 8, 8, 1
    assert lib.pa_pvt_fwd(C.byref(p), None, 0, None) == L.PA_ERR_BAD_SHAPE          # 128 % 3 (pvt.py:56)
    p.H, p.Himg = 2, 7
    assert lib.pa_pvt_fwd(C.byref(p), None, 0, None) == L.PA_ERR_BAD_SHAPE          # N != H*W
    le = L.LepeArgs()
    le.B, le.L, le.C, le.H, le.resolution, le.idx, le.split_size = 1, 50, 64, 2, 7, 0, 7
    assert lib.pa_cswin_lepe_fwd(C.byref(le), None) == L.PA_ERR_BAD_SHAPE           # L != resolution^2 (cswin.py:110)
    assert b"wrong size" in lib.pa_last_error()
    le.L, le.idx = 49, 2
    assert lib.pa_cswin_lepe_fwd(C.byref(le), None) == L.PA_ERR_UNSUPPORTED         # ERROR MODE (cswin.py:68-70)
    x = L.XcitArgs()
    x.B, x.N, x.C, x.H = 1, 8, 128, 8
    assert lib.pa_xca_fwd(C.byref(x), None, 0, None) == L.PA_ERR_UNSUPPORTED        # head_dim 16 on the XCA path (64 and 32 exist)
    cv = L.CvtArgs()
    assert lib.pa_cvt_fwd(C.byref(cv), None, 0, None) == L.PA_ERR_BAD_SHAPE
    blk = L.CswinBlockArgs()
    blk.B, blk.L, blk.C, blk.H, blk.reso, blk.split_size = 1, 100, 128, 4, 14, 7
    assert lib.pa_cswin_block_attn_fwd(C.byref(blk), None, 0, None) == L.PA_ERR_BAD_SHAPE


def test_python_error_mapping():
    with pytest.raises(AssertionError):
        L.check(L.PA_ERR_BAD_SHAPE)
    with pytest.raises(ValueError):
        L.check(L.PA_ERR_UNSUPPORTED)
    with pytest.raises(L.PaError):
        L.check(L.PA_ERR_CUDA)
