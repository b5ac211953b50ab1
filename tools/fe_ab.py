"""Front end alone, timed with CUDA events: python tools/fe_ab.py [tag]   -> one JSON line with ms per launch for the four
config-3 cases (256 x 30 s, fp32 [B, M, T] output, pass 1 + floor pass) and for the fused-path shape (64 x 30 s).
Used with alternating copies of the library (tools/fe_ab.sh A.so B.so)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel

tag = sys.argv[1] if len(sys.argv) > 1 else ""
N = 480000
pcm = synth.white_noise_batch(256, N, device="cuda")
res = {"tag": tag}
for mels in (80, 128):
    for n_fft in (400, 1024):
        fe = LogMel(mels, n_fft)
        out = torch.empty(256, mels, fe.num_frames(N), device="cuda")
        for _ in range(5):
            fe(pcm, out=out)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fe(pcm, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        res[f"{mels}x{n_fft}"] = round(best, 4)
        del fe, out
print(json.dumps(res))
