"""Log-mel front end: host side of ``asrb_logmel_f32`` behind the reference's API.

``extract_features`` keeps the reference signature and dict layout
(essentials.py:423-425, 512-521) for the spectrogram branch, so
``prepare_datasets.__getitem__`` (essentials.py:1008-1026) can call it unchanged;
``log_mel`` is the batched form: identical to stacking the per-utterance reference
results (per-utterance dynamic-range floor) with ``DataCollator``'s 0.0 right padding
(essentials.py:555-572).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib

SAMPLE_RATE, HOP, F_MIN, F_MAX = 16000, 160, 50.0, 8000.0


def _hann(n_fft: int) -> torch.Tensor:
    # window_fn=torch.hann_window, win_length=n_fft (essentials.py:480)
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)


def _fbank(n_freqs: int, n_mels: int, sample_rate: int, f_min: float, f_max: float) -> torch.Tensor:
    """HTK triangles, norm=None (ta:functional/functional.py:518-587), built on the host
    with the same fp32 torch ops torchaudio uses so the constants are bit-equal."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).contiguous()


class LogMel:
    """A front-end plan: window + banded filterbank resident on one GPU."""

    def __init__(self, n_mels: int = 128, n_fft: int = 1024, hop_length: int = HOP,
                 sample_rate: int = SAMPLE_RATE, f_min: float = F_MIN, f_max: float = F_MAX,
                 device: Optional[torch.device] = None):
        self.lib = _lib.load()
        self.n_mels, self.n_fft, self.hop = n_mels, n_fft, hop_length
        self.device = torch.device(device if device is not None else "cuda")
        win = _hann(n_fft)
        fb = _fbank(n_fft // 2 + 1, n_mels, sample_rate, f_min, f_max)
        self._plan = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.asrb_logmel_plan_create(n_fft, hop_length, n_mels, win.data_ptr(),
                                                        fb.data_ptr(), C.byref(self._plan)),
                       "asrb_logmel_plan_create")
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None and self._plan.value:
                self.lib.asrb_logmel_plan_destroy(self._plan)
                self._plan = None
        except Exception:            # interpreter shutdown
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._plan

    def num_frames(self, n_samples: int) -> int:
        return 1 + n_samples // self.hop

    def __call__(self, wave: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``wave [B, N]`` (or ``[N]``) fp32 on the plan's GPU -> ``[B, n_mels, 1 + N//hop]``."""
        squeeze = wave.dim() == 1
        if squeeze:
            wave = wave.unsqueeze(0)
        if wave.dim() != 2:
            raise ValueError("wave must be [N] or [B, N]")
        if not wave.is_cuda:
            raise _lib.AsrbError("log_mel needs a CUDA tensor: there is no CPU path")
        wave = wave.float()
        if wave.stride(-1) != 1:
            wave = wave.contiguous()
        B, N = wave.shape
        T = self.num_frames(N)
        if out is None:
            out = torch.empty(B, self.n_mels, T, device=wave.device, dtype=torch.float32)
        if lengths is not None:
            lengths = lengths.to(wave.device, torch.int32).contiguous()
        need = self.lib.asrb_logmel_workspace_bytes(self._plan, B, N)
        if self._ws is None or self._ws.numel() < need or self._ws.device != wave.device:
            self._ws = torch.empty(max(need, 256), device=wave.device, dtype=torch.uint8)
        with torch.cuda.device(wave.device):
            _lib.check(self.lib.asrb_logmel_f32(
                self._plan, wave.data_ptr(), B, N, wave.stride(0) if B > 1 else max(N, 1),
                lengths.data_ptr() if lengths is not None else None,
                out.data_ptr(), self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()),
                "asrb_logmel_f32")
        return out[0] if squeeze else out


_PLANS: Dict[Tuple, LogMel] = {}


def _plan_for(n_mels, n_fft, hop, sr, device) -> LogMel:
    key = (n_mels, n_fft, hop, sr, str(device))
    if key not in _PLANS:
        _PLANS[key] = LogMel(n_mels, n_fft, hop, sr, device=device)
    return _PLANS[key]


def log_mel(wave: torch.Tensor, n_mels: int = 128, n_fft: int = 1024, hop_length: int = HOP,
            sample_rate: int = SAMPLE_RATE, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched essentials.py:469-491.  Defaults are the reference's hard-coded values
    (n_fft=1024 essentials.py:475, mels=128 model.py:743)."""
    if not wave.is_cuda:
        raise _lib.AsrbError("log_mel needs a CUDA tensor: there is no CPU path")
    return _plan_for(n_mels, n_fft, hop_length, sample_rate, wave.device)(wave, lengths)


def waveform_feature(wave: torch.Tensor, hop_length: int = HOP, sample_rate: int = SAMPLE_RATE) -> torch.Tensor:
    """The reference's ``waveform`` feature (essentials.py:493-503), batched: ``[B, N]`` (or ``[N]``) ->
    ``[B, 1, target]`` with ``target = int((N / sample_rate) * (sample_rate // hop_length))`` evaluated in
    Python floats exactly like the reference (so e.g. N = 4640 gives 28, not 29)."""
    if not wave.is_cuda:
        raise _lib.AsrbError("waveform_feature needs a CUDA tensor: there is no CPU path")
    lib = _lib.load()
    squeeze = wave.dim() == 1
    if squeeze:
        wave = wave.unsqueeze(0)
    wave = wave.float()
    if wave.stride(-1) != 1:
        wave = wave.contiguous()
    B, N = wave.shape
    assert sample_rate % hop_length == 0                     # exact_div, essentials.py:296-298
    target = int((N / sample_rate) * (sample_rate // hop_length))
    out = torch.empty(B, 1, target, device=wave.device, dtype=torch.float32)
    with torch.cuda.device(wave.device):
        _lib.check(lib.asrb_waveform_pool_f32(wave.data_ptr(), B, N, wave.stride(0) if B > 1 else max(N, 1), target,
                                              out.data_ptr(), _lib.stream_ptr()), "asrb_waveform_pool_f32")
    return out[0] if squeeze else out


def extract_features(batch, tokenizer=None, spectrogram=False, pitch=False, waveform=False,
                     harmonics=False, aperiodics=False, phase=False, hilbert=False, pitch_tokens=False,
                     hop_length=160, sample_rate=16000, mels=128, n_fft=1024, device="cuda"):
    """Per-utterance drop-in for the reference ``extract_features`` (essentials.py:423-521):
    the spectrogram branch and the waveform (average-pooled PCM) branch.  The other branches
    are CPU WORLD-vocoder features outside this path and raise."""
    if pitch or harmonics or aperiodics or phase or hilbert or pitch_tokens:
        raise NotImplementedError("only spectrogram=True / waveform=True are on the accelerated path")
    labels = tokenizer.encode(batch["transcription" if "transcription" in batch else "sentence"]) \
        if tokenizer is not None else None
    audio = batch["audio"]
    if isinstance(audio, dict):                                   # load_wave dict branch, essentials.py:314-316
        wave = torch.as_tensor(audio["array"]).float()
    elif torch.is_tensor(audio):
        wave = audio.float()
    else:
        raise TypeError("Invalid wave_data format.")            # essentials.py:318
    wave = wave.to(device)
    s_tensor = log_mel(wave, mels, n_fft, hop_length, sample_rate) if spectrogram else None
    w_tensor = waveform_feature(wave, hop_length, sample_rate) if waveform else None       # [1, target]
    return {"waveform": w_tensor, "spectrogram": s_tensor, "pitch_tokens": None, "pitch": None,
            "harmonic": None, "aperiodic": None, "labels": labels, "phase": None}
