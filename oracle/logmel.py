"""Oracle: log-mel front end (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, op for op, what the reference computes in the ``spectrogram`` branch of
``extract_features`` (essentials.py:469-491) through torchaudio's
``MelSpectrogram`` (ta:transforms/_transforms.py:515-631).  The arithmetic lives in
a third-party dependency that is NOT vendored by the reference and is unpinned
there (no requirements file); the installed version it is restated from is
torchaudio 2.11.0+cu128 on torch 2.11.0+cu128.

``dtype=torch.float64`` runs the same formulas in double precision (tie-breaker
for tolerance arguments; the window and filterbank are still the fp32 constants
the reference uses, promoted).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

SAMPLE_RATE = 16000
HOP = 160
F_MIN = 50.0
F_MAX = 8000.0


def hann_periodic(n_fft: int) -> torch.Tensor:
    """Periodic Hann window, ``w[i] = 0.5 - 0.5 cos(2 pi i / n_fft)``.

    Follows ``window_fn=torch.hann_window`` with ``win_length = n_fft``
    (essentials.py:480; ta:transforms/_transforms.py:97-98 builds
    ``window_fn(win_length)``, periodic=True being torch's default).
    """
    return torch.hann_window(n_fft, periodic=True, dtype=torch.float32)


def _hz_to_mel_htk(f: float) -> float:
    # ta:functional/functional.py:425-456 (htk branch)
    return 2595.0 * math.log10(1.0 + f / 700.0)


def melscale_fbanks_htk(n_freqs: int, n_mels: int, sample_rate: int = SAMPLE_RATE,
                        f_min: float = F_MIN, f_max: float = F_MAX) -> torch.Tensor:
    """HTK triangular filterbank ``[n_freqs, n_mels]``, no area normalisation.

    ta:functional/functional.py:518-587 (``melscale_fbanks``) with ``norm=None``,
    ``mel_scale="htk"`` (essentials.py:481-482) and the triangle builder
    ta:functional/functional.py:489-515.  Built with fp32 torch ops exactly like the
    reference so the constants agree to the last bit.
    """
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_htk(f_min), _hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)          # ta:functional.py:459-486
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)       # [n_freqs, n_mels+2]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def power_spectrogram(wave: torch.Tensor, n_fft: int, hop: int = HOP,
                      dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``|STFT|^2`` of one utterance ``[N]`` -> ``[n_fft//2+1, 1+N//hop]``.

    ta:functional/functional.py:112-145 (``spectrogram``): constant (zero) padding
    of n_fft//2 on both sides because ``center=True, pad_mode="constant"``
    (essentials.py:477-478; torch:functional.py:675-680), frames of n_fft at stride
    hop, periodic Hann, one-sided DFT without normalisation, then ``abs().pow(2)``
    (``power=2.0``, essentials.py:479).
    """
    x = wave.to(dtype)
    win = hann_periodic(n_fft).to(dtype)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=win,
                      center=True, pad_mode="constant", normalized=False, onesided=True,
                      return_complex=True)
    return spec.abs().pow(2.0)


def log_mel_utterance(wave: torch.Tensor, n_mels: int = 128, n_fft: int = 1024,
                      hop: int = HOP, sample_rate: int = SAMPLE_RATE,
                      dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """One utterance ``[N]`` fp32 -> normalised log-mel ``[n_mels, 1+N//hop]``.

    essentials.py:469-490:
      mel     = MelSpectrogram(...)(audio.float())      # :486-487
      log_mel = clamp(mel, min=1e-10).log10()           # :488
      log_mel = maximum(log_mel, log_mel.max() - 8.0)   # :489  (max over THIS utterance)
      s       = (log_mel + 4.0) / 4.0                    # :490
    ``MelScale.forward`` is ``(spec^T @ fb)^T`` (ta:transforms/_transforms.py:407-419).
    The reference hard-codes n_fft=1024 (essentials.py:475) and mels=128
    (model.py:743); BASELINE.json sweeps both, same formula.
    """
    spec = power_spectrogram(wave, n_fft, hop, dtype)                    # [F, T]
    fb = melscale_fbanks_htk(n_fft // 2 + 1, n_mels, sample_rate).to(dtype)
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)     # [M, T]
    log_mel = torch.clamp(mel, min=1e-10).log10()
    log_mel = torch.maximum(log_mel, log_mel.max() - 8.0)
    return (log_mel + 4.0) / 4.0


def log_mel_batch(waves: torch.Tensor, n_mels: int = 128, n_fft: int = 1024,
                  hop: int = HOP, lengths: Optional[Sequence[int]] = None,
                  dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Batched equivalent the CUDA path is compared with: ``[B, N]`` -> ``[B, M, T]``.

    The reference extracts features one utterance at a time in the dataset
    (essentials.py:1008-1026) and pads afterwards in ``DataCollator``
    (essentials.py:555-572), so the batched result is the stack of per-utterance
    results (per-utterance max!), right-padded with 0.0 when ``lengths`` makes the
    clips ragged.
    """
    B, N = waves.shape
    feats: List[torch.Tensor] = []
    for b in range(B):
        n = N if lengths is None else int(lengths[b])
        feats.append(log_mel_utterance(waves[b, :n], n_mels, n_fft, hop, dtype=dtype))
    return collate_spectrograms(feats, t_max=1 + N // hop)


def collate_spectrograms(items: List[torch.Tensor], t_max: Optional[int] = None) -> torch.Tensor:
    """``DataCollator`` spectrogram branch (essentials.py:555-572): right-pad every
    ``[M, T_i]`` with 0.0 on the last dim to the longest, then stack."""
    t_max = max(i.shape[-1] for i in items) if t_max is None else t_max
    out = [torch.nn.functional.pad(i, (0, t_max - i.shape[-1]), mode="constant", value=0)
           for i in items]
    return torch.stack(out)


def waveform_feature(wave: torch.Tensor, sample_rate: int = SAMPLE_RATE, hop: int = HOP) -> torch.Tensor:
    """The ``waveform`` branch of ``extract_features`` (essentials.py:493-510) for one utterance ``[N]``:
    ``target = int((N / sr) * (sr // hop))`` (Python float arithmetic, as written in the reference),
    ``adaptive_avg_pool1d`` when the clip is longer than the target -> ``[1, target]``."""
    n = wave.shape[-1]
    target = int((n / sample_rate) * (sample_rate // hop))
    aud = wave.float().unsqueeze(0).unsqueeze(0)
    if n > target:
        w = torch.nn.functional.adaptive_avg_pool1d(aud, target)
    else:
        w = torch.nn.functional.interpolate(aud, size=target, mode="linear", align_corners=False)
    return w.squeeze(0)
