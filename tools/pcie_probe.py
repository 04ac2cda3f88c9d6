"""What bounds the end-to-end leg of bench.py?  Pure pinned-host copies of one step's input and output (19.4 MB each way),
alone and concurrently, against the same pipeline with the forward in the middle and 2 / 3 / 4 slots in flight."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_attention_b200 as pa

B, N, C = 64, 197, 768
dev = "cuda"
tokens = B * N
nbytes = B * N * C * 2

def wall(fn, n):
    fn(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(n); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6

hx = [torch.randn(B, N, C).half().pin_memory() for _ in range(4)]
hy = [torch.empty(B, N, C, dtype=torch.float16).pin_memory() for _ in range(4)]
dx = [torch.empty(B, N, C, dtype=torch.float16, device=dev) for _ in range(4)]
dy = [torch.empty(B, N, C, dtype=torch.float16, device=dev) for _ in range(4)]
s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

def h2d(n):
    with torch.cuda.stream(s_in):
        for i in range(n): dx[i & 3].copy_(hx[i & 3], non_blocking=True)
def d2h(n):
    with torch.cuda.stream(s_out):
        for i in range(n): hy[i & 3].copy_(dy[i & 3], non_blocking=True)
def both(n):
    for i in range(n):
        with torch.cuda.stream(s_in): dx[i & 3].copy_(hx[i & 3], non_blocking=True)
        with torch.cuda.stream(s_out): hy[i & 3].copy_(dy[i & 3], non_blocking=True)

for name, fn in (("H2D alone", h2d), ("D2H alone", d2h), ("H2D + D2H concurrently", both)):
    us = wall(fn, 200)
    print(f"{name:26s}: {us:7.1f} us per 19.4 MB step  = {nbytes / us / 1e3:5.1f} GB/s per direction -> copy-bound ceiling {tokens / us:6.2f} M tokens/s")

mod = pa.ViTAttention(C, 12).eval().half().cuda()
for slots in (2, 3, 4):
    ev_in = [torch.cuda.Event() for _ in range(slots)]
    ev_cmp = [torch.cuda.Event() for _ in range(slots)]
    ev_out = [torch.cuda.Event() for _ in range(slots)]
    ys = [None] * slots
    def e2e(n):
        with torch.no_grad():
            for i in range(n):
                b = i % slots
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_cmp[b]); dx[b].copy_(hx[b], non_blocking=True); ev_in[b].record(s_in)
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[b]); s_cmp.wait_event(ev_out[b]); ys[b] = mod(dx[b]); ev_cmp[b].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[b]); hy[b].copy_(ys[b], non_blocking=True); ev_out[b].record(s_out)
    us = wall(e2e, 200)
    print(f"pipeline with forward, {slots} slots in flight: {us:7.1f} us per step -> {tokens / us:6.2f} M tokens/s")
# host-side cost of one step's submissions (no GPU wait)
t0 = time.perf_counter()
with torch.no_grad():
    for i in range(200): mod(dx[0])
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host time to submit one forward (python + ctypes + launch): {(t1 - t0) / 200 * 1e6:6.1f} us")
