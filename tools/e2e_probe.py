#!/usr/bin/env python
"""Where does the end-to-end step lose time against the HBM-resident step?  Toggles the pieces of bench.py's e2e loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import asr_model_b200 as ab
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel
B, N, D = 64, 480000, 512
torch.manual_seed(0)
enc = ab.AudioEncoder(80, D, 4, 4, enc=False, compute="bf16").eval()
fe = LogMel(80, 400)
host = synth.white_noise_batch(B, N, seed=1).pin_memory()
dev_in = [torch.empty(B, N, device="cuda") for _ in range(2)]
pooled = [torch.empty(B, D).pin_memory() for _ in range(2)]
out = torch.empty(B, fe.num_frames(N), D, device="cuda", dtype=torch.bfloat16)
cs = torch.cuda.Stream(); ms_ = torch.cuda.current_stream()
ready = [torch.cuda.Event() for _ in range(2)]; freed = [torch.cuda.Event() for _ in range(2)]
def run(n, h2d, pool, reuse_out):
    for k in range(2): freed[k].record(ms_)
    def feed(i):
        k = i & 1
        with torch.cuda.stream(cs):
            cs.wait_event(freed[k])
            if h2d: dev_in[k].copy_(host, non_blocking=True)
            ready[k].record(cs)
    feed(0)
    for i in range(n):
        k = i & 1
        if i + 1 < n: feed(i + 1)
        ms_.wait_event(ready[k])
        h = enc.forward_pcm(dev_in[k], fe, out=out if reuse_out else None)
        if pool: pooled[k].copy_(torch.mean(h, dim=1, dtype=torch.float32), non_blocking=True)
        freed[k].record(ms_)
def t(**kw):
    run(3, **kw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(20, **kw); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
dev_in[0].copy_(host); dev_in[1].copy_(host)
for kw in (dict(h2d=False, pool=False, reuse_out=True), dict(h2d=False, pool=False, reuse_out=False), dict(h2d=False, pool=True, reuse_out=False),
           dict(h2d=True, pool=False, reuse_out=False), dict(h2d=True, pool=True, reuse_out=False)):
    print(kw, "%.3f ms" % t(**kw), flush=True)
