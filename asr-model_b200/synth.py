"""Synthetic utterances (there is no dataset access): the four signal classes of
SURVEY.md section 8d, generated on the CPU from fixed seeds so every consumer (tests,
bench, smoke, the golden-vector script) sees bit-identical PCM.

  W  white noise U(-0.5, 0.5)                      -- flat spectrum, clamp inactive
  H  "speech-like": 59 harmonics of 120 Hz, 1/k^2 amplitudes, 3 Hz raised-cosine
     envelope, peak 0.3, + 1e-5 N(0,1), last 2 s (or last 1/15th) exactly zero
  T  0.9 sin(2 pi 1000 n / 16000)                  -- ~96 % of bins at the clamp floor
  Z  zeros                                          -- every output is exactly -1.5

PCM is fp32 in [-1, 1] like ``load_wave`` produces (essentials.py:301-319).
"""
from __future__ import annotations

import math

import torch

SAMPLE_RATE = 16000


def make_wave(kind: str, n: int, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64)
    if kind == "W":
        return torch.rand(n, generator=g, dtype=torch.float32) - 0.5
    if kind == "Z":
        return torch.zeros(n, dtype=torch.float32)
    if kind == "T":
        return (0.9 * torch.sin(2 * math.pi * 1000.0 * t / SAMPLE_RATE)).float()
    if kind == "H":
        x = torch.zeros(n, dtype=torch.float64)
        for k in range(1, 60):
            x += torch.sin(2 * math.pi * 120.0 * k * t / SAMPLE_RATE + k) / (k * k)
        env = 0.5 - 0.5 * torch.cos(2 * math.pi * 3.0 * t / SAMPLE_RATE)
        x = x * env
        x = 0.3 * x / x.abs().max().clamp_min(1e-12)
        x = x + 1e-5 * torch.randn(n, generator=g, dtype=torch.float64)
        tail = min(2 * SAMPLE_RATE, max(n // 15, 1))
        x[n - tail:] = 0.0
        return x.float()
    if kind == "2":          # the two-tone known-answer signal of SURVEY.md section 8c
        return (0.5 * torch.sin(2 * math.pi * 440.0 * t / SAMPLE_RATE)
                + 0.25 * torch.sin(2 * math.pi * 3000.0 * t / SAMPLE_RATE)).float()
    raise ValueError(f"unknown signal class {kind!r}")


def make_batch(kinds: str, n: int, seed: int = 1234) -> torch.Tensor:
    """One utterance per character of ``kinds`` (e.g. ``"WHTZ"``), ``[B, n]`` fp32."""
    return torch.stack([make_wave(k, n, seed + i) for i, k in enumerate(kinds)])


def white_noise_batch(b: int, n: int, seed: int = 1234, device="cpu") -> torch.Tensor:
    """Throughput workload: class W for the whole batch (generated on ``device``)."""
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.rand(b, n, generator=g, dtype=torch.float32, device=device) - 0.5
