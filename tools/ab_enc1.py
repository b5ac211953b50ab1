"""enc = 1 step (TransformerEncoderLayer on) and the default step, timed with CUDA events after a pre-heat:
python tools/ab_enc1.py [tag] -> one JSON line (ms per step).  Used with alternating copies of the library."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import asr_model_b200 as ab
from asr_model_b200 import synth
from asr_model_b200.frontend import LogMel

tag = sys.argv[1] if len(sys.argv) > 1 else ""
pcm = synth.white_noise_batch(64, 480000, device="cuda")
fe = LogMel(80, 400)
res = {"tag": tag}
for enc in (True, False):
    torch.manual_seed(0)
    m = ab.AudioEncoder(80, 512, 4, 4, "gelu", "AbbyNormal", norm=False, enc=enc, compute="bf16").cuda().eval()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 2.0:                    # pre-heat: the timed steps run power-capped
        for _ in range(20):
            m.forward_pcm(pcm, fe)
        torch.cuda.synchronize()
    best = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            m.forward_pcm(pcm, fe)
        e1.record()
        torch.cuda.synchronize()
        best.append(round(e0.elapsed_time(e1) / 40, 4))
    res["enc1" if enc else "enc0"] = best
    del m
print(json.dumps(res))
