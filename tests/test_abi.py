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


def test_round2_entry_points_validate_arguments_without_gpu():
    """The sibling / block entry points of round 2 reject bad arguments before anything touches a device."""
    lib = L.load()
    # pa_pvt_fwd: segformer's dense reduction mode, cmt's relative_pos, p2t's external key tokens
    p = L.PvtArgs()
    p.B, p.N, p.C, p.H, p.Himg, p.Wimg, p.sr = 1, 64, 128, 2, 8, 8, 2
    p.sr_mode = 3
    assert lib.pa_pvt_workspace_bytes(C.byref(p)) == 0
    assert lib.pa_pvt_fwd(C.byref(p), None, 0, None) == L.PA_ERR_BAD_SHAPE
    p.sr_mode, p.Himg, p.Wimg, p.N = 1, 9, 9, 81
    assert lib.pa_pvt_fwd(C.byref(p), None, 0, None) == L.PA_ERR_BAD_SHAPE          # 9 % sr_ratio 2
    p.Himg, p.Wimg, p.N = 8, 8, 64
    dense = lib.pa_pvt_workspace_bytes(C.byref(p))
    p.sr_mode = 0
    assert dense > lib.pa_pvt_workspace_bytes(C.byref(p)) > 0                        # the patch matrix of the dense reduction
    p.kv_tokens, p.kv_count = 16, 10                                                 # external key tokens need sr == 1
    assert lib.pa_pvt_fwd(C.byref(p), None, 0, None) == L.PA_ERR_BAD_SHAPE
    blk = L.PvtBlockArgs()
    assert lib.pa_pvt_block_attn_workspace_bytes(C.byref(blk)) == 0
    assert lib.pa_pvt_block_attn_fwd(None, None, 0, None) == L.PA_ERR_NULL
    # pa_p2t_fwd: pyramid levels
    t = L.P2tArgs()
    t.attn.B, t.attn.N, t.attn.C, t.attn.H, t.attn.Himg, t.attn.Wimg, t.attn.sr = 1, 196, 128, 2, 14, 14, 1
    t.n_levels = 5
    assert lib.pa_p2t_workspace_bytes(C.byref(t)) == 0
    assert lib.pa_p2t_fwd(C.byref(t), None, 0, None) == L.PA_ERR_BAD_SHAPE
    t.n_levels = 2
    t.pool_h[0], t.pool_w[0], t.pool_h[1], t.pool_w[1] = 14, 14, 20, 2
    assert lib.pa_p2t_fwd(C.byref(t), None, 0, None) == L.PA_ERR_BAD_SHAPE          # pooled size 20 > H = 14
    t.pool_h[1] = 7
    assert lib.pa_p2t_workspace_bytes(C.byref(t)) > 0
    # pa_bvit_fwd: no output projection only when the inner width equals dim and y is fp16
    b = L.BvitArgs()
    b.B, b.N, b.C, b.H, b.dim_head = 1, 50, 64, 2, 64
    assert lib.pa_bvit_fwd(C.byref(b), None, 0, None) == L.PA_ERR_UNSUPPORTED
    b.H = 1
    assert lib.pa_bvit_workspace_bytes(C.byref(b)) > 0
    b.dim_head = 24
    assert lib.pa_bvit_fwd(C.byref(b), None, 0, None) == L.PA_ERR_UNSUPPORTED       # not a multiple of 16
    # pa_xca_block_attn_fwd
    xb = L.XcaBlockArgs()
    assert lib.pa_xca_block_attn_workspace_bytes(C.byref(xb)) == 0
    xb.attn.B, xb.attn.N, xb.attn.C, xb.attn.H = 1, 196, 128, 4
    assert lib.pa_xca_block_attn_workspace_bytes(C.byref(xb)) > 0
    assert lib.pa_xca_block_attn_fwd(C.byref(xb), None, 0, None) == L.PA_ERR_NULL
    # pa_vit_fwd(topk): the workspace grows by the per-row thresholds
    v = L.VitArgs()
    v.B, v.N, v.C, v.H = 2, 197, 768, 12
    plain = lib.pa_vit_workspace_bytes(C.byref(v))
    v.topk = 100
    assert lib.pa_vit_workspace_bytes(C.byref(v)) >= plain + 2 * 197 * 12 * 4
