"""Front end alone for profiler captures: python tools/prof_frontend.py [B] [mels] [n_fft] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mels = int(sys.argv[2]) if len(sys.argv) > 2 else 80
n_fft = int(sys.argv[3]) if len(sys.argv) > 3 else 400
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
pcm = synth.white_noise_batch(B, 480000, device="cuda")
fe = LogMel(mels, n_fft)
out = torch.empty(B, mels, fe.num_frames(480000), device="cuda")
for _ in range(iters):
    fe(pcm, out=out)
torch.cuda.synchronize()
print("ok")
