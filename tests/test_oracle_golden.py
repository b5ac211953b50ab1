"""The oracle replays the fixtures that oracle/pin_against_reference.py recorded from the
unmodified reference (tests/golden/PINNED.json says how close they were when recorded)."""
import numpy as np
import pytest
import torch

import oracle
from asr_model_b200 import synth


def _cases(golden):
    for key in golden["frontend"].files:
        if key.startswith("logmel_"):
            _, m, f, kind, n = key.split("_")
            yield key, int(m[1:]), int(f[1:]), kind, int(n)


def test_frontend_matches_reference_fixtures(golden):
    n = 0
    for key, m, f, kind, length in _cases(golden):
        ref = torch.from_numpy(golden["frontend"][key])
        out = oracle.log_mel_utterance(synth.make_wave(kind, length), m, f)
        assert out.shape == ref.shape == (m, 1 + length // 160)
        assert float((out - ref).abs().max()) <= 2e-6, key     # recorded bit-exact; slack for other CPUs
        n += 1
    assert n == 20


def test_known_answers_from_survey(golden):
    # zeros -> -1.5 everywhere; two-tone statistics; frame counts
    z = oracle.log_mel_utterance(torch.zeros(4000), 80, 400)
    assert torch.all(z == -1.5)
    for n, t in ((16000, 101), (160000, 1001), (480000, 3001)):
        assert 1 + n // 160 == t
    o = oracle.log_mel_utterance(synth.make_wave("2", 16000), 80, 400)
    assert o.shape == (80, 101)
    assert abs(float(o.sum()) - 147.033654) < 2e-2
    assert abs(float(o.max()) - 1.830703) < 1e-4 and abs(float(o.min()) - (-0.169297)) < 1e-4
    assert abs(float(o[0, 0]) - 1.213067) < 1e-4 and abs(float(o[79, 100]) - 0.712725) < 1e-4
    assert int(o[:, 50].argmax()) == 13
    o = oracle.log_mel_utterance(synth.make_wave("2", 16000), 128, 1024)
    assert abs(float(o.max()) - 2.059790) < 1e-4 and abs(float(o[0, 0]) - 1.240752) < 1e-4
    assert int(o[:, 50].argmax()) == 21


def test_filterbank_and_window_kats(golden):
    fb = oracle.melscale_fbanks_htk(201, 80)
    assert int((fb > 0).sum()) == 390 and abs(float(fb.sum()) - 195.239318) < 1e-3
    assert int((fb > 0).sum(0).max()) == 12
    assert torch.equal(fb, torch.from_numpy(golden["frontend"]["fbank_f400_m80"]))
    fb = oracle.melscale_fbanks_htk(513, 128)
    assert int((fb > 0).sum()) == 1005 and int((fb > 0).sum(0).max()) == 20
    fb = oracle.melscale_fbanks_htk(201, 128)
    assert int(((fb > 0).sum(0) == 0).sum()) == 3          # three all-zero filters: keep them
    w = oracle.hann_periodic(400)
    assert w[0] == 0 and w[200] == 1 and abs(float(w.sum()) - 200) < 1e-4
    assert abs(float((w * w).sum()) - 150) < 1e-4


def test_per_utterance_max_not_batch_max():
    waves = synth.make_batch("WHTZ", 8000)
    out = oracle.log_mel_batch(waves, 80, 400)
    for b in range(4):
        assert torch.equal(out[b], oracle.log_mel_utterance(waves[b], 80, 400))
    assert torch.all(out[3] == -1.5)


def test_ragged_batch_pads_with_zero():
    waves = synth.make_batch("WH", 8000)
    out = oracle.log_mel_batch(waves, 80, 400, lengths=[8000, 4810])
    t1 = 1 + 4810 // 160
    assert torch.all(out[1, :, t1:] == 0.0)
    assert torch.equal(out[1, :, :t1], oracle.log_mel_utterance(waves[1, :4810], 80, 400))


def test_sinusoids_kat(golden):
    s = oracle.sinusoids(4, 8)
    assert np.allclose(s.numpy(), golden["sinusoids"]["sin_4_8"], atol=0)
    assert np.allclose(s[1].numpy(), [0.841471, 0.032177, 0.001036, 3.3e-05, 0.540302, 0.999482, 0.999999, 1.0], atol=1e-6)
    r = oracle.sinusoids(3001, 512)[3000]
    assert np.array_equal(r.numpy(), golden["sinusoids"]["sin_3001_512_row3000"])
    assert abs(float(r[255]) - 0.0998334) < 1e-6 and abs(float(r[511]) - 0.9950042) < 1e-6


@pytest.mark.parametrize("name,head", [("enc_small", 4), ("enc_small_tel", 4), ("enc_m128_default", 4), ("enc_conv2", 4)])
def test_encoder_matches_reference_fixtures(golden, name, head):
    import json, os
    pinned = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "PINNED.json")))
    mels, D, H, L, B, T, enc, perturb = pinned["cases"][name]["cfg"]
    sd = oracle.random_encoder_state_dict(80 if mels == 1 else mels, D, L, enc, seed=11, perturb=perturb)
    x = torch.from_numpy(golden["encoder"][name + "_x"])
    y = oracle.audio_encoder_forward(sd, x, H)
    assert y.shape == (B, T, D)
    assert float((y - torch.from_numpy(golden["encoder"][name + "_y"])).abs().max()) <= 1e-5


def test_attention_matches_reference_fixture(golden):
    sd = oracle.random_attention_state_dict(64, 4, seed=3)
    x = torch.from_numpy(golden["attention"]["att_x"])
    y = oracle.attention_forward(sd, x, 4)
    assert float((y - torch.from_numpy(golden["attention"]["att_y"])).abs().max()) <= 1e-5
    # cross-attention (keys / values and k's rotary magnitudes from xa, model.py:259, 306)
    xa = torch.from_numpy(golden["attention"]["att_xa"])
    yx = oracle.attention_forward(sd, x, 4, xa=xa)
    assert float((yx - torch.from_numpy(golden["attention"]["att_cross_y"])).abs().max()) <= 1e-5


def test_residual_mlp_matches_reference_fixture(golden):
    sd = oracle.random_mlp_state_dict(128, 3, seed=5)
    y = oracle.residual_mlp_forward(sd, torch.from_numpy(golden["attention"]["mlp_x"]))
    assert float((y - torch.from_numpy(golden["attention"]["mlp_y"])).abs().max()) <= 1e-5


def test_encoder_param_count():
    spec = oracle.encoder_state_dict_spec(80, 512, 4, False)
    n = sum(int(np.prod(s)) for k, s in spec.items()
            if "running_" not in k and "num_batches" not in k)
    assert n == 6476288


def test_waveform_feature_matches_reference_fixtures(golden):
    n = 0
    for key in golden["frontend"].files:
        if key.startswith("waveform_"):
            _, kind, length = key.split("_")
            ref = torch.from_numpy(golden["frontend"][key])
            out = oracle.waveform_feature(synth.make_wave(kind, int(length)))
            assert out.shape == ref.shape and float((out - ref).abs().max()) <= 1e-7, key
            n += 1
    assert n == 3
    assert oracle.waveform_feature(torch.zeros(4640)).shape == (1, 28)      # the reference's float quirk: 0.29 * 100 -> 28
