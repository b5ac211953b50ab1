#!/usr/bin/env python
"""Fused forward (PCM -> hidden) at several batch sizes, pre-heated: does the step gain from activations that fit in L2?
`python tools/batch_sweep.py [enc]` prints ms per step and audio-s/s per batch size (30 s clips, config 2 model)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import asr_model_b200 as ab
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel

enc = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
torch.manual_seed(0)
m = ab.AudioEncoder(80, 512, 4, 4, enc=enc, compute="bf16").eval()
fe = LogMel(80, 400)
pcm_all = synth.white_noise_batch(128, 480000, device="cuda")
for B in (128, 64, 32, 16):
    chunks = [pcm_all[i:i + B] for i in range(0, 128, B)]
    outs = [torch.empty(B, 3001, 512, device="cuda", dtype=torch.bfloat16) for _ in chunks]
    def step():
        for c, o in zip(chunks, outs):
            m.forward_pcm(c, fe, out=o)
    t0 = time.time()
    while time.time() - t0 < 2.0:
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"128 clips as {len(chunks)} x {B}: {ms:.3f} ms = {128 * 30 / ms * 1e3:.0f} audio-s/s")
