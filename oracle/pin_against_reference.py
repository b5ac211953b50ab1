#!/usr/bin/env python
"""Pin the oracle against the UNMODIFIED reference and write ``tests/golden/``.

Runs only in the build container (needs ``/root/reference``; the GPU box has no
copy).  It imports the reference's own ``model.py`` / ``essentials.py`` with the four
missing non-arithmetic packages stubbed (pyworld, soundfile, tensorboardX,
tensordict -- none is touched by the hot path), then for every case:

  1. runs the reference (its ``extract_features`` for n_fft=1024, the
     ``torchaudio.transforms.MelSpectrogram`` call it makes -- same kwargs -- for
     n_fft=400; its ``AudioEncoder``, ``sinusoids`` and ``attention`` modules),
  2. runs the oracle restatement on the same input,
  3. asserts they agree (bit-exact for the front end, <= 2e-6 for the encoder),
  4. stores the REFERENCE output as a golden fixture.

Usage:  python oracle/pin_against_reference.py            (rewrites tests/golden/)
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("ASR_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    for name, attrs in {"pyworld": [], "soundfile": [], "tensorboardX": ["SummaryWriter"],
                        "tensordict": ["TensorDict"]}.items():
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, type(a, (), {}))
        sys.modules.setdefault(name, mod)
    sys.path.insert(0, REF)
    import essentials  # noqa
    import model       # noqa
    return model, essentials


class _Tok:
    def encode(self, s):
        return [5, 6, 7]


def ref_logmel(essentials, wave, n_mels, n_fft):
    """The reference's own result for one utterance."""
    import torchaudio
    if n_fft == 1024:
        out = essentials.extract_features(
            {"audio": {"array": wave.numpy(), "sampling_rate": 16000}, "transcription": "x"},
            _Tok(), spectrogram=True, hop_length=160, sample_rate=16000, mels=n_mels)
        return out["spectrogram"].cpu()
    # n_fft=400 (BASELINE config): the reference hard-codes 1024 (essentials.py:475), so
    # make the very call it makes (essentials.py:470-490) with that one kwarg changed.
    cfg = {"hop_length": 160, "f_min": 50, "f_max": 8000, "n_mels": n_mels, "n_fft": n_fft,
           "sample_rate": 16000, "pad_mode": "constant", "center": True, "power": 2.0,
           "window_fn": torch.hann_window, "mel_scale": "htk", "norm": None, "normalized": False}
    mel = torchaudio.transforms.MelSpectrogram(**cfg)(wave.float())
    log_mel = torch.clamp(mel, min=1e-10).log10()
    log_mel = torch.maximum(log_mel, log_mel.max() - 8.0)
    return (log_mel + 4.0) / 4.0


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)     # one thread: fixtures independent of the core count
    model, essentials = import_reference()
    import oracle
    from asr_model_b200 import synth

    os.makedirs(GOLD, exist_ok=True)
    report = {"torch": torch.__version__, "cases": {}}

    # ---------------- front end ----------------
    fe = {}
    for n_mels in (80, 128):
        for n_fft in (400, 1024):
            for kind, n in (("2", 16000), ("W", 8000), ("H", 24000), ("T", 8037), ("Z", 4000)):
                wave = synth.make_wave(kind, n)
                ref = ref_logmel(essentials, wave, n_mels, n_fft)
                ours = oracle.log_mel_utterance(wave, n_mels, n_fft)
                d = float((ref - ours).abs().max())
                assert ref.shape == ours.shape == (n_mels, 1 + n // 160), (ref.shape, ours.shape)
                assert d == 0.0, f"oracle != reference for {(n_mels, n_fft, kind)}: {d}"
                o64 = oracle.log_mel_utterance(wave, n_mels, n_fft, dtype=torch.float64)
                key = f"logmel_m{n_mels}_f{n_fft}_{kind}_{n}"
                fe[key] = ref.numpy().astype(np.float32)
                report["cases"][key] = {"oracle_vs_ref_maxabs": d,
                                        "ref_fp32_vs_fp64_maxabs": float((ref.double() - o64).abs().max())}
    for n_fft, n_mels in ((400, 80), (1024, 128), (400, 128)):
        import torchaudio
        fb_ref = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, 50.0, 8000.0, n_mels, 16000,
                                                       norm=None, mel_scale="htk")
        fb = oracle.melscale_fbanks_htk(n_fft // 2 + 1, n_mels)
        assert torch.equal(fb_ref, fb)
        fe[f"fbank_f{n_fft}_m{n_mels}"] = fb.numpy()
    # waveform feature (essentials.py:493-510) through the reference's own extract_features
    for kind, n in (("W", 8000), ("H", 4640), ("2", 16037)):
        wave = synth.make_wave(kind, n)
        out = essentials.extract_features({"audio": {"array": wave.numpy(), "sampling_rate": 16000}, "transcription": "x"},
                                          _Tok(), waveform=True, hop_length=160, sample_rate=16000)["waveform"].cpu()
        ours = oracle.waveform_feature(wave)
        assert out.shape == ours.shape and torch.equal(out, ours), (kind, n, out.shape, ours.shape)
        fe[f"waveform_{kind}_{n}"] = out.numpy()
        report["cases"][f"waveform_{kind}_{n}"] = {"oracle_vs_ref_maxabs": 0.0, "shape": list(out.shape)}
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **fe)

    # ---------------- sinusoids ----------------
    for ctx, dims in ((4, 8), (3001, 512), (50, 64)):
        ref = essentials.sinusoids(ctx, dims).detach().cpu()
        assert torch.equal(ref, oracle.sinusoids(ctx, dims)), (ctx, dims)
    sin = {"sin_4_8": essentials.sinusoids(4, 8).detach().numpy(),
           "sin_3001_512_row3000": essentials.sinusoids(3001, 512).detach().numpy()[3000]}
    np.savez_compressed(os.path.join(GOLD, "sinusoids.npz"), **sin)

    # ---------------- encoder ----------------
    enc_gold = {}
    for name, (mels, D, H, L, B, T, enc, perturb) in {
        "enc_small": (80, 64, 4, 2, 2, 50, False, True),
        "enc_small_tel": (80, 64, 4, 2, 2, 50, True, True),
        "enc_m128_default": (128, 128, 4, 1, 1, 33, False, False),
        "enc_conv2": (1, 64, 4, 1, 2, 40, False, True),
    }.items():
        m_in = mels
        mod = model.AudioEncoder(80 if mels == 1 else mels, D, H, L, "gelu", "AbbyNormal",
                                 norm=False, enc=enc).eval()
        spec = oracle.encoder_state_dict_spec(80 if mels == 1 else mels, D, L, enc)
        ref_spec = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
        assert spec == ref_spec, "state_dict layout drifted from the reference"
        sd = oracle.random_encoder_state_dict(80 if mels == 1 else mels, D, L, enc, seed=11, perturb=perturb)
        mod.load_state_dict(sd)
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, m_in, T, generator=g) * 0.7 + 0.3
        with torch.no_grad():
            ref = mod(x)
        ours = oracle.audio_encoder_forward(sd, x, H)
        d = float((ref - ours).abs().max())
        assert d <= 2e-6, (name, d)
        enc_gold[name + "_x"] = x.numpy()
        enc_gold[name + "_y"] = ref.numpy()
        report["cases"][name] = {"oracle_vs_ref_maxabs": d, "cfg": [mels, D, H, L, B, T, enc, perturb]}
    # parameter count KAT: AudioEncoder(80,512,4,4) has 6 476 288 parameters
    n_par = sum(p.numel() for p in model.AudioEncoder(80, 512, 4, 4, "gelu", "AbbyNormal").parameters())
    assert n_par == 6476288, n_par
    np.savez_compressed(os.path.join(GOLD, "encoder.npz"), **enc_gold)

    # ---------------- attention + rotary (secondary) ----------------
    att_gold = {}
    D, H, T = 64, 4, 37
    att = model.attention(D, H, 1, n_type="rmsnorm").eval()
    sd = oracle.random_attention_state_dict(D, H, seed=3)
    assert {k: tuple(v.shape) for k, v in att.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    att.load_state_dict(sd)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, T, D, generator=g)
    with torch.no_grad():
        ref = torch.cat([att(x[b:b + 1]) for b in range(2)])     # B=1 semantics per utterance
    ours = oracle.attention_forward(sd, x, H)
    d = float((ref - ours).abs().max())
    assert d <= 2e-6, d
    att_gold["att_x"], att_gold["att_y"] = x.numpy(), ref.numpy()
    report["cases"]["attention_rotary"] = {"oracle_vs_ref_maxabs": d, "cfg": [D, H, T]}
    # cross-attention against another sequence (keys / values and k's rotary magnitudes from xa, model.py:259, 306)
    xa = torch.randn(2, 53, D, generator=g)
    with torch.no_grad():
        refx = torch.cat([att(x[b:b + 1], xa=xa[b:b + 1]) for b in range(2)])
    oursx = oracle.attention_forward(sd, x, H, xa=xa)
    dx = float((refx - oursx).abs().max())
    assert dx <= 2e-6, dx
    att_gold["att_xa"], att_gold["att_cross_y"] = xa.numpy(), refx.numpy()
    report["cases"]["attention_rotary_cross"] = {"oracle_vs_ref_maxabs": dx, "cfg": [D, H, T, 53]}
    # residual.mlp (model.py:573-574): shared RMSNorm, tgate, Linear-GELU-Linear -- the reference's own sub-module, called directly
    Dm = 128
    res = model.residual(Dm, 4, 2, "gelu", "rmsnorm").eval()
    msd = oracle.random_mlp_state_dict(Dm, 3, seed=5)
    full = res.state_dict()
    for k, v in msd.items():
        assert tuple(full[k].shape) == tuple(v.shape), k
        full[k] = v
    full["mlp.0.weight"] = full["mlp.5.weight"] = msd["ln.weight"]          # the shared norm appears under three names
    res.load_state_dict(full)
    xm = torch.randn(2, 29, Dm, generator=g)
    with torch.no_grad():
        refm = res.mlp(xm)
    oursm = oracle.residual_mlp_forward(msd, xm)
    dm = float((refm - oursm).abs().max())
    assert dm <= 2e-6, dm
    att_gold["mlp_x"], att_gold["mlp_y"] = xm.numpy(), refm.numpy()
    report["cases"]["residual_mlp"] = {"oracle_vs_ref_maxabs": dm, "cfg": [Dm, 3, 29]}
    np.savez_compressed(os.path.join(GOLD, "attention.npz"), **att_gold)

    with open(os.path.join(GOLD, "PINNED.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
