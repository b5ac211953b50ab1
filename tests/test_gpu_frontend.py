"""GPU parity of the fused log-mel kernel (through the C ABI) against the CPU oracle and the
reference fixtures.  Tolerance: 1e-4 max-abs in fp32 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import oracle
from asr_model_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ab(built_lib):
    import asr_model_b200 as ab
    assert torch.cuda.is_available()
    return ab


@pytest.mark.parametrize("n_mels,n_fft", [(80, 400), (128, 400), (80, 1024), (128, 1024)])
def test_mixed_signal_classes_match_oracle(ab, n_mels, n_fft):
    waves = synth.make_batch("WHTZ2WH", 16000 + 37)            # ragged tail, all classes in one batch
    ref = oracle.log_mel_batch(waves, n_mels, n_fft)
    out = ab.log_mel(waves.cuda(), n_mels, n_fft).cpu()
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= TOL
    ref64 = oracle.log_mel_batch(waves, n_mels, n_fft, dtype=torch.float64)
    assert float((out.double() - ref64).abs().max()) <= TOL
    assert torch.all(out[3] == -1.5)                          # class Z: exactly -1.5


def test_reference_fixtures(ab, golden):
    n = 0
    for key in golden["frontend"].files:
        if not key.startswith("logmel_"):
            continue
        _, m, f, kind, length = key.split("_")
        ref = torch.from_numpy(golden["frontend"][key])
        out = ab.log_mel(synth.make_wave(kind, int(length)).cuda(), int(m[1:]), int(f[1:])).cpu()
        assert out.shape == ref.shape, key
        assert float((out - ref).abs().max()) <= TOL, key
        n += 1
    assert n == 20


@pytest.mark.parametrize("n_fft", [400, 1024])
def test_ragged_lengths_pad_with_zero_and_use_own_max(ab, n_fft):
    waves = synth.make_batch("WHT", 12000)
    lengths = [12000, 4810, 0]
    ref = oracle.log_mel_batch(waves, 80, n_fft, lengths=lengths)
    out = ab.log_mel(waves.cuda(), 80, n_fft, lengths=torch.tensor(lengths)).cpu()
    assert float((out - ref).abs().max()) <= TOL
    assert torch.all(out[1, :, 1 + 4810 // 160:] == 0.0)
    assert torch.all(out[2, :, 1:] == 0.0) and torch.all(out[2, :, 0] == -1.5)


def test_edge_shapes(ab):
    for n in (0, 1, 159, 160, 161, 399, 400, 5000):
        w = synth.make_wave("W", max(n, 1))[:n]
        ref = oracle.log_mel_utterance(w, 80, 400) if n > 0 else torch.full((80, 1), -1.5)
        out = ab.log_mel(w.cuda(), 80, 400).cpu()
        assert out.shape == (80, 1 + n // 160)
        assert float((out - ref).abs().max()) <= TOL, n
    assert ab.log_mel(torch.zeros(0, 1600, device="cuda"), 80, 400).shape == (0, 80, 11)


def test_unaligned_rows_take_the_scalar_path(ab):
    base = synth.make_batch("WH", 8003)
    dev = base.cuda()
    view = dev[:, 1:8002]                                       # stride 8003, offset 1: no 16-B alignment
    assert not view.is_contiguous()
    ref = oracle.log_mel_batch(base[:, 1:8002].contiguous(), 80, 400)
    from asr_model_b200.frontend import LogMel
    out = LogMel(80, 400)(view).cpu()
    assert float((out - ref).abs().max()) <= TOL


def test_full_size_30s_clips(ab):
    """BASELINE config sizes: 30 s clips. Oracle on 3 of them + size-independent properties."""
    N = 480000
    waves = torch.cat([synth.make_batch("WHT", N), torch.zeros(1, N)])
    out = ab.log_mel(waves.cuda(), 80, 400)
    assert out.shape == (4, 80, 3001)
    ref = oracle.log_mel_batch(waves[:3], 80, 400)
    assert float((out[:3].cpu() - ref).abs().max()) <= TOL
    assert torch.all(out[3] == -1.5)
    # dynamic range: max - min <= 2.0 (8 in log10 / 4) per utterance
    flat = out.flatten(1)
    assert torch.all(flat.max(1).values - flat.min(1).values <= 2.0 + 1e-6)
    # batch independence: an utterance alone == the same utterance inside a batch
    alone = ab.log_mel(waves[1:2].cuda(), 80, 400)
    assert torch.equal(alone[0], out[1])
    # time-shift covariance: shifting by k hops shifts interior frames
    w = waves[0].cuda()
    a = ab.log_mel(w[:160000], 80, 400)
    bshift = ab.log_mel(w[1600:161600], 80, 400)
    inner = (a[:, 20:900] - bshift[:, 10:890]).abs().max()
    assert float(inner) <= 2e-4


@pytest.mark.parametrize("n_fft", [400, 1024])
def test_blocks_walk_several_tiles_across_ragged_utterances(ab, n_fft):
    """More frame tiles than resident blocks (12 x 30 s = 1128 / 2256 tiles), so every block walks several tiles
    and crosses utterance boundaries with its division-free (utterance, tile) counter; lengths ragged, one empty,
    one full.  Utterances are checked against the oracle one by one (own max, 0.0 padding) and against the same
    utterance computed alone (bit-equal)."""
    N = 480000
    kinds = "WHTWHTWHTWHT"
    waves = synth.make_batch(kinds, N)
    lengths = [N, 0, 123457, 479999, 160, 31, 300000, 480000 - 160, 1, 200000, 399, 77777]
    dev = waves.cuda()
    out = ab.log_mel(dev, 80, n_fft, lengths=torch.tensor(lengths))
    assert out.shape == (12, 80, 3001)
    for b in (0, 2, 3, 6, 11):
        ref = oracle.log_mel_batch(waves[b:b + 1], 80, n_fft, lengths=[lengths[b]])
        assert float((out[b:b + 1].cpu() - ref).abs().max()) <= TOL, b
    for b in range(12):
        tb = 1 + lengths[b] // 160
        assert torch.all(out[b, :, tb:] == 0.0), b
        alone = ab.log_mel(dev[b:b + 1], 80, n_fft, lengths=torch.tensor(lengths[b:b + 1]))
        assert torch.equal(alone[0], out[b]), b


def test_extract_features_drop_in(ab):
    class Tok:
        def encode(self, s):
            return [5, 6]
    w = synth.make_wave("H", 16000)
    r = ab.extract_features({"audio": {"array": w.numpy(), "sampling_rate": 16000}, "transcription": "x"},
                            Tok(), spectrogram=True, mels=128)
    assert r["labels"] == [5, 6] and r["spectrogram"].shape == (128, 101) and r["spectrogram"].is_cuda
    assert float((r["spectrogram"].cpu() - oracle.log_mel_utterance(w, 128, 1024)).abs().max()) <= TOL


def test_waveform_feature_matches_reference(ab, golden):
    """the average-pooled PCM stream (essentials.py:493-510)."""
    for key in golden["frontend"].files:
        if key.startswith("waveform_"):
            _, kind, length = key.split("_")
            ref = torch.from_numpy(golden["frontend"][key])
            out = ab.waveform_feature(synth.make_wave(kind, int(length)).cuda()).cpu()
            assert out.shape == ref.shape and float((out - ref).abs().max()) <= 1e-6, key
    waves = synth.make_batch("WH2", 480000)
    out = ab.waveform_feature(waves.cuda()).cpu()
    assert out.shape == (3, 1, 3000)
    for b in range(3):
        assert float((out[b] - oracle.waveform_feature(waves[b])).abs().max()) <= 1e-6
    r = ab.extract_features({"audio": {"array": waves[0].numpy(), "sampling_rate": 16000}, "transcription": "x"},
                            None, spectrogram=True, waveform=True, mels=80, n_fft=400)
    assert r["waveform"].shape == (1, 3000) and r["spectrogram"].shape == (80, 3001)
    # the pooled stream feeds the encoder through conv2 (model.py:152-155)
    sd = oracle.random_encoder_state_dict(80, 128, 1, False, seed=5, perturb=True)
    m = ab.AudioEncoder(80, 128, 4, 1, compute="fp32").eval()
    m.load_state_dict(sd)
    h = m(r["waveform"]).cpu()
    ref_h = oracle.audio_encoder_forward(sd, oracle.waveform_feature(waves[0]), 4)
    assert float((h - ref_h).abs().max()) <= 1e-4


def test_fused_spectrogram_and_waveform_pass_equals_separate_calls(ab):
    """the average-pooled waveform feature comes out of the front-end kernel's own pass over the PCM
    (asrb_logmel_waveform_f32), bit-equal to the two separate calls -- also when the clip length is not a multiple of the
    hop and pooling bins drift away from the staged span."""
    from asr_model_b200.frontend import LogMel, pooled_target
    for n_fft, n in ((400, 48000), (400, 4640 * 7 + 3), (1024, 16000 + 37), (400, 480000)):
        waves = synth.make_batch("WH2", n).cuda()
        fe = LogMel(80, n_fft)
        tg = pooled_target(n)
        mel, pooled = fe(waves, pooled_target=tg)
        assert pooled.shape == (3, 1, tg)
        assert torch.equal(mel, fe(waves))
        assert torch.equal(pooled, ab.waveform_feature(waves)), (n_fft, n)
        for b in range(3):
            assert float((pooled[b].cpu() - oracle.waveform_feature(waves[b].cpu())).abs().max()) <= 1e-6


def test_waveform_interpolation_branch_matches_reference_op(ab):
    """essentials.py:505-506: when the clip is not longer than the target (hop 1) the reference interpolates linearly."""
    for n in (5, 16, 999, 16000):
        w = synth.make_wave("W", n)
        ref = oracle.waveform_feature(w, sample_rate=16000, hop=1)           # target = n * 16000 / 16000 ... = n: current <= target
        out = ab.waveform_feature(w.cuda(), hop_length=1).cpu()
        assert out.shape == ref.shape and float((out - ref).abs().max()) <= 1e-6, n
    x = torch.arange(8, dtype=torch.float32)
    lib = ab.lib.load()
    out = torch.empty(1, 20, device="cuda")
    ab.lib.check(lib.asrb_waveform_pool_f32(x.cuda().data_ptr(), 1, 8, 8, 20, out.data_ptr(), None), "asrb_waveform_pool_f32")
    ref = torch.nn.functional.interpolate(x.view(1, 1, -1), size=20, mode="linear", align_corners=False).view(1, 20)
    assert float((out.cpu() - ref).abs().max()) <= 1e-6


def test_bad_lengths_are_rejected_on_the_host_and_clamped_on_the_device(ab):
    from asr_model_b200.frontend import LogMel
    waves = synth.make_batch("WH", 8000).cuda()
    fe = LogMel(80, 400)
    with pytest.raises(ValueError):
        fe(waves, lengths=torch.tensor([8000, 8001]))
    with pytest.raises(ValueError):
        fe(waves, lengths=torch.tensor([-1, 5]))
    # a device tensor is not synchronised on: the kernels clamp, so an oversized length behaves like the full clip and
    # nothing is written past the tensor (guard rows stay untouched)
    buf = torch.full((4, 80, fe.num_frames(8000)), 7.0, device="cuda")
    fe(waves, lengths=torch.tensor([10 ** 6, -5], device="cuda"), out=buf[1:3])
    assert torch.all(buf[0] == 7.0) and torch.all(buf[3] == 7.0)
    ref = oracle.log_mel_batch(waves.cpu(), 80, 400, lengths=[8000, 0])
    assert float((buf[1:3].cpu() - ref).abs().max()) <= TOL


@pytest.mark.parametrize("n_mels,n_fft,n", [(80, 400, 160 * 95 + 7), (128, 1024, 160 * 33), (80, 400, 5)])
def test_frontend_writes_every_element_and_nothing_else(ab, n_mels, n_fft, n):
    """compute-sanitizer is closed on this pool: NaN-filled outputs with sentinel guard rows on both sides show that
    logmel_kernel + the floor pass write every element of [B, M, T] and not one byte outside it (ragged lengths included)."""
    from asr_model_b200.frontend import LogMel
    fe = LogMel(n_mels, n_fft)
    B, T = 3, fe.num_frames(n)
    waves = synth.make_batch("WHT", n).cuda()
    guard = 4096
    buf = torch.full((B * n_mels * T + 2 * guard,), 12345.0, device="cuda")
    out = buf[guard: guard + B * n_mels * T].view(B, n_mels, T)
    for lengths in (None, torch.tensor([n, n // 2, 0])):
        out.fill_(float("nan"))
        fe(waves, lengths=lengths, out=out)
        torch.cuda.synchronize()
        assert not torch.isnan(out).any()
        assert bool((buf[:guard] == 12345.0).all()) and bool((buf[-guard:] == 12345.0).all())
    # the fused 16-bit channels-last output of the PCM -> hidden path is covered by the encoder's NaN-filled result tensor
    # (tests/test_gpu_encoder.py::test_full_size_64x30s_batch_properties)


@pytest.mark.parametrize("n_fft,hop", [(400, 128), (400, 200), (1024, 256), (400, 50), (1024, 190)])
def test_other_hop_lengths_take_the_generic_path(ab, n_fft, hop):
    """`hop_length` is a free parameter of extract_features (essentials.py:423-425); 160 has a specialised kernel (compile-time
    frame offsets, bank-skewed staging), every other value runs the run-time-hop instantiation -- incl. hops that are not a
    multiple of 4 (no 16-byte staging) and the pooled waveform feature taken in the same pass."""
    from asr_model_b200.frontend import LogMel
    n = 16000 + 123
    waves = synth.make_batch("WH2Z", n)
    ref = oracle.log_mel_batch(waves, 80, n_fft, hop=hop)
    fe = LogMel(80, n_fft, hop)
    out = fe(waves.cuda()).cpu()
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= TOL
    lengths = [n, 7000, 1234, 0]
    refl = oracle.log_mel_batch(waves, 80, n_fft, hop=hop, lengths=lengths)
    assert float((fe(waves.cuda(), lengths=torch.tensor(lengths)).cpu() - refl).abs().max()) <= TOL
    if 16000 % hop == 0:
        from asr_model_b200.frontend import pooled_target
        tg = pooled_target(n, hop)
        if 0 < tg < n and tg <= fe.num_frames(n):
            mel, pooled = fe(waves.cuda(), pooled_target=tg)
            assert torch.equal(mel.cpu(), out)
            for b in range(4):
                assert float((pooled[b].cpu() - oracle.waveform_feature(waves[b], hop=hop)).abs().max()) <= 1e-6
