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


def check_lengths(lengths: Optional[torch.Tensor], batch: int, n_samples: int, device) -> Optional[torch.Tensor]:
    """``lengths`` -> int32 ``[batch]`` on ``device``.  A host tensor is validated here (0 <= len <= n_samples) before
    upload; a device tensor is not synchronised on -- the kernels clamp every length to [0, n_samples] themselves."""
    if lengths is None:
        return None
    lengths = torch.as_tensor(lengths)
    if lengths.numel() != batch:
        raise ValueError(f"lengths has {lengths.numel()} entries for a batch of {batch}")
    if not lengths.is_cuda and lengths.numel() and (int(lengths.min()) < 0 or int(lengths.max()) > n_samples):
        raise ValueError(f"lengths must lie in [0, {n_samples}]")
    return lengths.to(device, torch.int32).contiguous()


def pooled_target(n_samples: int, hop_length: int = HOP, sample_rate: int = SAMPLE_RATE) -> int:
    """Length of the reference's ``waveform`` feature, in Python floats exactly as written there
    (essentials.py:495: e.g. N = 4640 gives 28, not 29)."""
    assert sample_rate % hop_length == 0                     # exact_div, essentials.py:296-298
    return int((n_samples / sample_rate) * (sample_rate // hop_length))


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
                 out: Optional[torch.Tensor] = None, pooled_target: int = 0):
        """``wave [B, N]`` (or ``[N]``) fp32 on the plan's GPU -> ``[B, n_mels, 1 + N//hop]``.
        ``pooled_target > 0``: the same pass also emits the average-pooled ``waveform`` feature ``[B, 1, target]``
        (essentials.py:493-503) and the call returns ``(log_mel, waveform)``."""
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
        lengths = check_lengths(lengths, B, N, wave.device)
        need = self.lib.asrb_logmel_workspace_bytes(self._plan, B, N)
        if self._ws is None or self._ws.numel() < need or self._ws.device != wave.device:
            self._ws = torch.empty(max(need, 256), device=wave.device, dtype=torch.uint8)
        pooled = torch.empty(B, 1, pooled_target, device=wave.device, dtype=torch.float32) if pooled_target > 0 else None
        with torch.cuda.device(wave.device):
            if pooled is None:
                _lib.check(self.lib.asrb_logmel_f32(
                    self._plan, wave.data_ptr(), B, N, wave.stride(0) if B > 1 else max(N, 1),
                    lengths.data_ptr() if lengths is not None else None,
                    out.data_ptr(), self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()),
                    "asrb_logmel_f32")
            else:
                _lib.check(self.lib.asrb_logmel_waveform_f32(
                    self._plan, wave.data_ptr(), B, N, wave.stride(0) if B > 1 else max(N, 1),
                    lengths.data_ptr() if lengths is not None else None,
                    out.data_ptr(), pooled.data_ptr(), pooled_target, self._ws.data_ptr(), self._ws.numel(),
                    _lib.stream_ptr()), "asrb_logmel_waveform_f32")
        if pooled is not None:
            return (out[0], pooled[0]) if squeeze else (out, pooled)
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
    target = pooled_target(N, hop_length, sample_rate)
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
    wave, sample_rate_in = load_wave(batch["audio"], sample_rate)
    wave = wave.to(device)
    target = pooled_target(wave.shape[-1], hop_length, sample_rate) if waveform else 0
    if spectrogram and waveform and wave.dim() == 1 and 0 < target < wave.shape[-1] and target <= 1 + wave.shape[-1] // hop_length:
        # one pass over the PCM for both features (asrb_logmel_waveform_f32)
        s_tensor, w_tensor = _plan_for(mels, n_fft, hop_length, sample_rate, wave.device)(wave, pooled_target=target)
    else:
        s_tensor = log_mel(wave, mels, n_fft, hop_length, sample_rate) if spectrogram else None
        w_tensor = waveform_feature(wave, hop_length, sample_rate) if waveform else None       # [1, target]
    return {"waveform": w_tensor, "spectrogram": s_tensor, "pitch_tokens": None, "pitch": None,
            "harmonic": None, "aperiodic": None, "labels": labels, "phase": None}


def load_wave(audio, sample_rate: int = SAMPLE_RATE):
    """``load_wave`` of the reference (essentials.py:301-319): a ``str`` path is read with ``soundfile`` as float32 and
    PEAK-NORMALISED (mono: ``wp / max|wp|``; multi-channel: per-channel ``wp / wp.max(axis=0)`` when any maximum is
    positive, then transposed to ``[channels, N]``); a dict (the datasets ``audio`` column) is taken as it is; a tensor is
    accepted as a convenience.  Returns ``(waveform fp32, sample_rate)``.  Host-side glue, not on the timed path."""
    if isinstance(audio, str):
        try:
            import soundfile as sf
        except ImportError as e:
            raise ImportError("extract_features was given a file path: reading it needs the `soundfile` package, "
                              "exactly as the reference's load_wave does (essentials.py:302-303)") from e
        wp, sample_rate = sf.read(audio, dtype="float32")
        if wp.ndim > 1:
            abs_max = wp.max(axis=0)                                        # essentials.py:306 (max, not max-abs, as written there)
            wp = wp / abs_max if any(abs_max > 0) else wp
            return torch.from_numpy(wp.T.copy()), sample_rate
        abs_max = max(abs(wp)) if len(wp) else 0.0
        wp = wp / abs_max if abs_max > 0 else wp
        return torch.from_numpy(wp), sample_rate
    if isinstance(audio, dict):                                             # essentials.py:314-316
        return torch.as_tensor(audio["array"]).float(), audio.get("sampling_rate", sample_rate)
    if torch.is_tensor(audio):
        return audio.float(), sample_rate
    raise TypeError("Invalid wave_data format.")                           # essentials.py:318
