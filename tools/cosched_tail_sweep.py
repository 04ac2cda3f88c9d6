"""Device time of the co-scheduled ViT kernel against PA_CS_TAIL (m-groups at the end of the proj phase issued as 256 x 64
quarter tiles): CUDA graph of 24 forwards over 8 rotating inputs, min / median of 7 replays; outputs checked bit for bit against
tail = 0.  usage: cosched_tail_sweep.py [tails, comma separated] [B C H N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa
from pytorch_attention_b200 import _lib
_lib.load()
tails = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,3,4,6,8")]
dbgs = [int(v) for v in os.environ.get("SWEEP_CS_DEBUG", "0").split(",")]
qtails = [int(v) for v in os.environ.get("SWEEP_CS_QTAIL", "0").split(",")]
B, C, H, N = [int(v) for v in (sys.argv[2:6] if len(sys.argv) >= 6 else (64, 768, 12, 197))]
torch.manual_seed(0)
m = pa.ViTAttention(C, H).eval().half().cuda()
xs = [torch.randn(B, N, C, device="cuda").half() for _ in range(8)]
ref = None
for rep in range(int(os.environ.get('SWEEP_REPS', '2'))):
    for tail, qt, dbg in [(t, q, d) for t in tails for q in qtails for d in dbgs]:
        os.environ["PA_CS_TAIL"] = str(tail); os.environ["PA_CS_QTAIL"] = str(qt); os.environ["PA_CS_DEBUG"] = str(dbg); _lib.reload_env()
        with torch.no_grad():
            y = m(xs[0]); torch.cuda.synchronize()
            if ref is None: ref = y.clone()
            same = bool(torch.equal(y, ref))
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for i in range(3): m(xs[i])
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for i in range(24): m(xs[i % 8])
            ts = []
            for _ in range(7):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / 24 * 1e3)
        print(f"PA_CS_TAIL={tail:2d} PA_CS_QTAIL={qt:2d} PA_CS_DEBUG={dbg:2d}: {min(ts):6.1f} us (median {sorted(ts)[3]:6.1f})  bit-equal to tail 0: {same}  path {_lib.load().pa_last_vit_path()}", flush=True)
