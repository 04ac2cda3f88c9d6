"""Tiny driver for ncu: a few GEMM configs, 3 launches each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytorch_attention_b200 import ops
torch.manual_seed(0)
cfgs = [(12608, 2304, 768, 256, 1), (12608, 2304, 768, 256, 2), (12608, 2304, 768, 192, 1), (12608, 768, 768, 192, 1),
        (12608, 768, 768, 256, 2), (131072, 512, 512, 256, 2), (12608, 3072, 1024, 256, 2)]
for (M, N, K, bn, cl) in cfgs:
    A = torch.randn(M, K, device="cuda").half()
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    D = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.gemm_tn(A, B, out=D, block_n=bn, cluster=cl)
    torch.cuda.synchronize()
