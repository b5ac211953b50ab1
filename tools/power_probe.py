#!/usr/bin/env python
"""SM clock, power draw and throttle reasons while the 64 x 30 s step runs back to back for a few seconds."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pynvml
import asr_model_b200 as ab
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
torch.manual_seed(0)
enc = ab.AudioEncoder(80, 512, 4, 4, enc=False, compute="bf16").eval()
fe = LogMel(80, 400)
pcm = synth.white_noise_batch(64, 480000, device="cuda")
out = torch.empty(64, fe.num_frames(480000), 512, device="cuda", dtype=torch.bfloat16)
samples, stop = [], threading.Event()
def sampler():
    while not stop.is_set():
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h), pynvml.nvmlDeviceGetTemperature(h, 0)))
        time.sleep(0.05)
for _ in range(5): enc.forward_pcm(pcm, fe, out=out)
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(1500): enc.forward_pcm(pcm, fe, out=out)
e1.record(); torch.cuda.synchronize()
stop.set(); th.join()
print("ms per step over 1500 steps: %.3f" % (e0.elapsed_time(e1) / 1500))
print("power limit W:", pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0, " max SM MHz:", pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
for t, mhz, w, r, temp in samples[::8]:
    print("t=%.2fs  sm %4d MHz  %6.1f W  %2d C  reasons 0x%x" % (t - t0, mhz, w, temp, r))
