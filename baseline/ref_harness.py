"""Runs the UNMODIFIED reference modules (model.py / essentials.py of sine2pi/ASR-model) for the
baselines bench.py reports: the CPU path on the box's host cores (`bench.py --impl reference`,
`cpu_baseline`) and PyTorch eager on the B200 (`gpu_eager_baseline`).

MEASUREMENT INFRASTRUCTURE, like oracle/: only bench.py imports it.  Nothing here is on the product path.

Where the reference comes from: `/root/reference` in the build container; on the GPU box (which has no
copy) the three files `__graft_entry__.build()` copied to `baseline/_ref/` (git-ignored, shipped by
gpurun -- SURVEY.md section 8c).  The reference has no setup.py / pyproject, so `pip install` does not
apply; its modules are imported as they are, with the four packages this image lacks and the hot path never
touches (pyworld, soundfile, tensorboardX, tensordict) stubbed in `sys.modules`.

Two things differ from calling `extract_features` verbatim, both dictated by BASELINE.json's configs:
  * `extract_features` hard-codes n_fft = 1024 (essentials.py:475); configs 1-2 name n_fft = 400, so the arm
    makes the very call sequence of essentials.py:470-490 (MelSpectrogram built per utterance, clamp, log10,
    per-utterance max - 8, (x + 4) / 4) with that one kwarg changed;
  * the module-global `device` (model.py:13) is pointed at the CPU for the CPU arm, because on a GPU box the
    reference would otherwise move its activations to cuda:0 (model.py:160-161).
"""
from __future__ import annotations

import os
import statistics
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("model.py", "essentials.py", "optimizerc.py")
_MODS = None


def reference_dir():
    for d in (os.environ.get("ASR_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")):
        if d and all(os.path.exists(os.path.join(d, f)) for f in FILES):
            return d
    return None


def stage_reference_copy() -> bool:
    """Build container only: copy the three reference files next to this harness so they travel to the GPU box."""
    src = os.environ.get("ASR_REFERENCE", "/root/reference")
    if not all(os.path.exists(os.path.join(src, f)) for f in FILES):
        return False
    dst = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        data = open(os.path.join(src, f), "rb").read()
        p = os.path.join(dst, f)
        if not os.path.exists(p) or open(p, "rb").read() != data:
            open(p, "wb").write(data)
    return True


def import_reference():
    """(model, essentials) of the unmodified reference, or None when no copy is reachable."""
    global _MODS
    if _MODS is not None:
        return _MODS
    d = reference_dir()
    if d is None:
        return None
    for name, attrs in {"pyworld": [], "soundfile": [], "tensorboardX": ["SummaryWriter"], "tensordict": ["TensorDict"]}.items():
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, type(a, (), {}))
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, d)
    import warnings
    warnings.filterwarnings("ignore")
    import essentials  # noqa
    import model       # noqa
    _MODS = (model, essentials)
    return _MODS


def _point_device(mods, dev):
    import torch
    for m in mods:
        if hasattr(m, "device"):
            m.device = torch.device(dev)


def reference_logmel(wave, n_mels, n_fft, device=None):
    """essentials.py:470-490 for one utterance (transform constructed per call, as the reference does)."""
    import torch
    import torchaudio
    cfg = {"hop_length": 160, "f_min": 50, "f_max": 8000, "n_mels": n_mels, "n_fft": n_fft,
           "sample_rate": 16000, "pad_mode": "constant", "center": True, "power": 2.0,
           "window_fn": torch.hann_window, "mel_scale": "htk", "norm": None, "normalized": False}
    transform = torchaudio.transforms.MelSpectrogram(**cfg)
    if device is not None:
        transform = transform.to(device)
    mel = transform(wave.float())
    log_mel = torch.clamp(mel, min=1e-10).log10()
    log_mel = torch.maximum(log_mel, log_mel.max() - 8.0)
    return (log_mel + 4.0) / 4.0


def cpu_throughput(batch, secs, mels, n_fft, dims, head, layer, enc, repeats, warm, seed=0):
    """audio-s/s of the reference's own CPU path: per-utterance front end + AudioEncoder.eval() forward, fp32,
    all host threads.  Returns (value, threads, per-step times) or None when the reference is unreachable."""
    import torch
    mods = import_reference()
    if mods is None:
        return None
    model, essentials = mods
    _point_device(mods, "cpu")
    from asr_model_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(seed)
    enc_mod = model.AudioEncoder(mels, dims, head, layer, "gelu", "AbbyNormal", norm=False, enc=enc).eval()
    waves = synth.white_noise_batch(batch, secs * 16000)
    times = []
    with torch.no_grad():
        for i in range(warm + repeats):
            t = time.perf_counter()
            feats = torch.stack([reference_logmel(w, mels, n_fft) for w in waves])     # equal lengths: DataCollator pads nothing
            enc_mod(feats)
            if i >= warm:
                times.append(time.perf_counter() - t)
    return batch * secs / statistics.median(times), torch.get_num_threads(), times


def gpu_eager(batch, secs, mels, n_fft, dims, head, layer, enc, steps=5, warm=3, seed=0):
    """The reference's modules in PyTorch eager on cuda:0 at the bench shapes: fp32 with TF32 exactly as model.py:18-25
    sets it, and under bf16 autocast.  The front end is the reference's op sequence with the transform moved to the GPU
    and the batch handed over at once (its natural GPU form; the per-utterance max is kept).  Returns a dict or None."""
    import torch
    mods = import_reference()
    if mods is None or not torch.cuda.is_available():
        return None
    model, essentials = mods
    _point_device(mods, "cuda:0")
    import torchaudio
    from asr_model_b200 import synth
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    enc_mod = model.AudioEncoder(mels, dims, head, layer, "gelu", "AbbyNormal", norm=False, enc=enc).to(dev).eval()
    waves = synth.white_noise_batch(batch, secs * 16000).to(dev)
    cfg = {"hop_length": 160, "f_min": 50, "f_max": 8000, "n_mels": mels, "n_fft": n_fft, "sample_rate": 16000,
           "pad_mode": "constant", "center": True, "power": 2.0, "window_fn": torch.hann_window, "mel_scale": "htk",
           "norm": None, "normalized": False}
    transform = torchaudio.transforms.MelSpectrogram(**cfg).to(dev)

    def front(w):
        log_mel = torch.clamp(transform(w.float()), min=1e-10).log10()
        log_mel = torch.maximum(log_mel, log_mel.amax(dim=(1, 2), keepdim=True) - 8.0)
        return (log_mel + 4.0) / 4.0

    def timed(fn):
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

    out = {"what": "unmodified model.AudioEncoder + torchaudio MelSpectrogram, PyTorch eager on cuda:0",
           "tf32": bool(torch.backends.cuda.matmul.allow_tf32), "batch": batch, "steps": steps}
    with torch.no_grad():
        ms_fe = timed(lambda: front(waves))
        ms32 = timed(lambda: enc_mod(front(waves)))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms16 = timed(lambda: enc_mod(front(waves)))
    out.update({"front_end_ms": ms_fe, "fp32_tf32_ms_per_step": ms32, "bf16_autocast_ms_per_step": ms16,
                "fp32_tf32_audio_s_per_s": batch * secs / (ms32 * 1e-3), "bf16_autocast_audio_s_per_s": batch * secs / (ms16 * 1e-3)})
    del enc_mod, waves, transform
    torch.cuda.empty_cache()
    return out
