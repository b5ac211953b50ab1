#!/usr/bin/env python
"""One fused forward (PCM -> hidden) for ncu: `python tools/prof_step.py [B] [enc]`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import asr_model_b200 as ab
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
enc = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
torch.manual_seed(0)
m = ab.AudioEncoder(80, 512, 4, 4, enc=enc, compute="bf16").eval()
fe = LogMel(80, 400)
pcm = synth.white_noise_batch(B, 480000, device="cuda")
for _ in range(2):
    out = m.forward_pcm(pcm, fe)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.float().abs().mean()))
