#!/usr/bin/env python
"""Does running the 64-clip step as micro-batches (activations that fit the 126 MB L2) beat one pass?"""
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
pcm = synth.white_noise_batch(B, N, device="cuda")
out = torch.empty(B, fe.num_frames(N), D, device="cuda", dtype=torch.bfloat16)
def run(mb):
    for s in range(0, B, mb):
        enc.forward_pcm(pcm[s:s + mb], fe, out=out[s:s + mb])
def t(mb, n=20):
    for _ in range(3): run(mb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): run(mb)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mb in (64, 32, 37, 16, 64):
    print("micro-batch %2d: %.3f ms per 64 clips" % (mb, t(mb)), flush=True)
