"""Co-scheduled kernel: forward time vs PA_CS_LAG (how many m-groups the proj tiles trail their rows' qkv tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cosched_bringup import fresh, timed, setenv
for (B, C, H, N) in [(64, 768, 12, 197), (64, 1024, 16, 197)]:
    m, x = fresh(B, C, H, N)
    res = []
    for lag in [1, 2, 4, 6, 9, 13, 18, 25, 50]:
        setenv(PA_VIT_COSCHED=1, PA_CS_LAG=lag)
        res.append((lag, timed(m, x, 300)))
    setenv(PA_CS_LAG=None)
    res.append(("default", timed(m, x, 300)))
    setenv(PA_VIT_COSCHED=0, PA_VIT_FUSED=1)
    res.append(("sequenced", timed(m, x, 300)))
    setenv(PA_VIT_COSCHED=None, PA_VIT_FUSED=None)
    print(f"B={B} C={C}: " + "  ".join(f"{k}:{v:.1f}" for k, v in res), flush=True)
