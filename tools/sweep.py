#!/usr/bin/env python
"""BASELINE.json configs 3 and 4 (not bench lines): front-end sweep and the wide encoder.
Prints one JSON object per case; CUDA-event timed, inputs resident in HBM (larger than L2)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import asr_model_b200 as ab
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


def timeit(fn, warm=3, steps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "frontend"):
    B, N = 256, 480000
    pcm = synth.white_noise_batch(B, N, device="cuda")
    for mels in (80, 128):
        for n_fft in (400, 1024):
            fe = LogMel(mels, n_fft)
            out = torch.empty(B, mels, fe.num_frames(N), device="cuda")
            ms = timeit(lambda: fe(pcm, out=out))
            byt = B * (4 * N + 4 * mels * fe.num_frames(N))
            print(json.dumps({"config": 3, "case": f"front end {B} x 30 s, {mels} mel, n_fft {n_fft}", "ms": ms,
                              "audio_s_per_s": B * 30 / (ms * 1e-3), "algorithmic_GBps": byt / ms / 1e6,
                              "hbm_roofline_frac": byt / ms / 1e6 / PEAK["hbm_gbs"]}), flush=True)
            del fe, out
    del pcm
if which in ("all", "wide"):
    B, N = 128, 480000
    pcm = synth.white_noise_batch(B, N, device="cuda")
    fe = LogMel(80, 400)
    T = fe.num_frames(N)
    for enc in (False, True):
        torch.manual_seed(0)
        m = ab.AudioEncoder(80, 1024, 16, 24, enc=enc, compute="bf16").eval()
        ms = timeit(lambda: m.forward_pcm(pcm, fe), warm=2, steps=4)
        fl = (2 * 3 * 80 * 1024 + 24 * 12 * 1024 * 1024 + (8 * 1024 * 1024 + 4 * T * 1024 + 8192 * 1024 if enc else 0)) * B * T
        print(json.dumps({"config": 4, "case": f"wide encoder D=1024 H=16 L=24 enc={enc}, {B} x 30 s", "ms": ms,
                          "audio_s_per_s": B * 30 / (ms * 1e-3), "tflops": fl / ms / 1e9,
                          "tensor_roofline_frac": fl / ms / 1e9 / PEAK["bf16_tflops_sustained"]}), flush=True)
        del m
