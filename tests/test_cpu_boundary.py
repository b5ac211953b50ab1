"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/asrb200.h declares, fails loudly without a B200, and the Python mirror keeps the
reference's state_dict layout.  No compute calls here (no GPU in this container)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "asrb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asrb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = built_lib.load()
    declared = _declared()
    assert sorted(built_lib.SYMBOLS) == declared, "ctypes table drifted from include/asrb200.h"
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(asrb_[a-z0-9_]+)\b", out))
    assert set(declared) <= exported
    assert lib.asrb_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_not_silently(built_lib):
    lib = built_lib.load()
    assert lib.asrb_device_check(0) == -2                      # ASRB_E_DEVICE
    assert b"CUDA" in lib.asrb_last_error() or b"device" in lib.asrb_last_error()
    plan = C.c_void_p()
    win = torch.hann_window(400)
    fb = oracle.melscale_fbanks_htk(201, 80).contiguous()
    rc = lib.asrb_logmel_plan_create(400, 160, 80, win.data_ptr(), fb.data_ptr(), C.byref(plan))
    assert rc == -2 and not plan.value
    from asr_model_b200 import frontend, encoder
    with pytest.raises(built_lib.AsrbError):
        frontend.log_mel(torch.zeros(2, 1600), 80, 400)       # CPU tensor: no fallback
    enc = encoder.AudioEncoder(80, 64, 4, 1, compute="fp32").eval()
    with pytest.raises(built_lib.AsrbError):
        enc(torch.zeros(1, 80, 10))


def test_argument_errors_have_messages(built_lib):
    lib = built_lib.load()
    plan = C.c_void_p()
    win = torch.hann_window(512)
    fb = torch.zeros(257, 80)
    assert lib.asrb_logmel_plan_create(512, 160, 80, win.data_ptr(), fb.data_ptr(), C.byref(plan)) == -1
    assert b"n_fft=512" in lib.asrb_last_error()
    assert lib.asrb_logmel_f32(None, None, 1, 16, 16, None, None, None, 0, None) == -1
    assert lib.asrb_encoder_forward(None, None, 1, 80, 10, None, 0, None, 0, None) == -1
    assert lib.asrb_encoder_workspace_bytes(None, 1, 1) == 0


@pytest.mark.parametrize("enc", [False, True])
def test_module_keeps_reference_state_dict_layout(enc):
    from asr_model_b200.encoder import AudioEncoder
    m = AudioEncoder(80, 64, 4, 2, "gelu", "AbbyNormal", norm=False, enc=enc)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == oracle.encoder_state_dict_spec(80, 64, 2, enc)
    m.load_state_dict(oracle.random_encoder_state_dict(80, 64, 2, enc, seed=1))
    with pytest.raises(NotImplementedError):
        AudioEncoder(80, 64, 4, 2, norm=True)


def test_attention_module_keeps_reference_state_dict_layout():
    from asr_model_b200.attention import AudioAttention
    a = AudioAttention(64, 4)
    sd = oracle.random_attention_state_dict(64, 4)
    assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_host_constants_match_oracle():
    from asr_model_b200 import frontend
    for n_fft, m in ((400, 80), (1024, 128), (400, 128), (1024, 80)):
        assert torch.equal(frontend._fbank(n_fft // 2 + 1, m, 16000, 50.0, 8000.0), oracle.melscale_fbanks_htk(n_fft // 2 + 1, m))
        assert torch.equal(frontend._hann(n_fft), oracle.hann_periodic(n_fft))


def test_extract_features_signature_matches_reference():
    import inspect
    from asr_model_b200.frontend import extract_features
    names = list(inspect.signature(extract_features).parameters)
    assert names[:13] == ["batch", "tokenizer", "spectrogram", "pitch", "waveform", "harmonics", "aperiodics", "phase",
                          "hilbert", "pitch_tokens", "hop_length", "sample_rate", "mels"]   # essentials.py:423-425
    with pytest.raises(NotImplementedError):
        extract_features({"audio": {"array": [0.0]}, "transcription": "x"}, None, pitch=True)
    with pytest.raises(TypeError):
        extract_features({"audio": 3, "transcription": "x"}, None, spectrogram=True)


def test_load_wave_follows_the_reference(monkeypatch):
    """essentials.py:301-319: str path -> soundfile read + peak normalisation; dict -> as is; anything else -> TypeError."""
    import sys
    import types
    import numpy as np
    import torch
    from asr_model_b200.frontend import load_wave
    fake = types.ModuleType("soundfile")
    mono = np.array([0.1, -0.5, 0.25], dtype=np.float32)
    stereo = np.array([[0.1, -0.2], [0.4, 0.1], [-0.8, 0.05]], dtype=np.float32)
    fake.read = lambda path, dtype="float32": ((mono if "mono" in path else stereo).copy(), 16000)
    monkeypatch.setitem(sys.modules, "soundfile", fake)
    w, sr = load_wave("mono.wav")
    assert sr == 16000 and torch.equal(w, torch.from_numpy(mono / np.float32(0.5)))
    w2, _ = load_wave("stereo.wav")                       # per-channel max (not max-abs), transposed: as the reference writes it
    assert w2.shape == (2, 3) and torch.equal(w2, torch.from_numpy((stereo / stereo.max(axis=0)).T.copy()))
    d, sr = load_wave({"array": mono, "sampling_rate": 8000})
    assert sr == 8000 and torch.equal(d, torch.from_numpy(mono))        # no normalisation on the dict branch
    import pytest
    with pytest.raises(TypeError):
        load_wave(3.14)
    monkeypatch.setitem(sys.modules, "soundfile", None)
    with pytest.raises(ImportError):
        load_wave("x.wav")
