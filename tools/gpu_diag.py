#!/usr/bin/env python
"""First-contact diagnostics on a B200: runs every building block and prints error norms
instead of stopping at the first assert.  Output goes to gpurun_out/diag.log."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
g.build()
import oracle
import asr_model_b200 as ab
from asr_model_b200 import synth, _lib

print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
lib = _lib.load()


def step(name, fn):
    t = time.time()
    try:
        r = fn()
        torch.cuda.synchronize()
        print(f"[ok ] {name}: {r}  ({time.time() - t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc(limit=3)


def fe(m, f):
    waves = synth.make_batch("WHTZ2", 16037)
    out = ab.log_mel(waves.cuda(), m, f).cpu()
    ref = oracle.log_mel_batch(waves, m, f)
    return [round(float((out[i] - ref[i]).abs().max()), 7) for i in range(5)]


SECTION = sys.argv[1] if len(sys.argv) > 1 else "all"
for m, f in ((80, 400), (128, 400), (80, 1024), (128, 1024)) if SECTION in ("all", "logmel") else ():
    step(f"logmel m{m} f{f} per-class max-abs", lambda m=m, f=f: fe(m, f))

sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_gemm_tc import CASES, _ref


def gemm(B, T, K, N, taps, epi, act):
    gen = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(B, T, K, device="cuda", generator=gen) * 0.5).bfloat16()
    w = (torch.randn(N, taps * K, device="cuda", generator=gen) / (taps * K) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=gen) * 0.1
    n_out = N // 2 if epi in (1, 4) else N
    res = torch.randn(B, T, n_out, device="cuda", generator=gen).bfloat16() if epi in (2, 3, 5) else None
    gamma = 1 + 0.2 * torch.randn(N, device="cuda", generator=gen)
    beta = 0.1 * torch.randn(N, device="cuda", generator=gen)
    dw, dw_args = None, (None, None, 0, 0, None)
    if epi in (4, 5):
        kw = 15 if epi == 4 else 3
        dw_w = torch.randn(kw, n_out, device="cuda", generator=gen) / kw ** 0.5
        dw_b = 0.1 * torch.randn(n_out, device="cuda", generator=gen)
        pos = torch.randn(T, n_out, device="cuda", generator=gen) if epi == 5 else None
        dw = (dw_w, dw_b, kw, 3 if epi == 4 else 4, pos)
        dw_args = (dw_w.data_ptr(), dw_b.data_ptr(), kw, dw[3], pos.data_ptr() if pos is not None else None)
    out = torch.full((B, T, n_out), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.asrb_test_gemm_tc(a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr() if res is not None else None,
                               gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), B, T, K, N, taps, epi, act, *dw_args, None)
    _lib.check(rc, "gemm")
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, res, gamma, beta, taps, epi, act, N, dw)
    err = (out.float() - ref).abs()
    nan = int(torch.isnan(out.float()).sum())
    bad = (err > 2e-2 + 1e-2 * ref.abs()) | torch.isnan(out.float())
    where = ""
    if bool(bad.any()):
        idx = bad.nonzero()
        where = f" bad {int(bad.sum())} first {idx[0].tolist()} last {idx[-1].tolist()} rows {sorted(set(idx[:, 1].tolist()))[:12]}"
    return f"max {float(err.nan_to_num(9e9).max()):.4g} mean {float(err.nan_to_num(0).mean()):.3g} nan {nan} refmax {float(ref.abs().max()):.3g}{where}"


for c in CASES if SECTION in ("all", "gemm") else ():
    step(f"gemm_tc {c}", lambda c=c: gemm(*c))


def enc(D, H, L, tel, compute, B=2, T=300):
    sd = oracle.random_encoder_state_dict(80, D, L, tel, seed=3, perturb=True)
    waves = synth.make_batch("WH", (T - 1) * 160)
    mel = oracle.log_mel_batch(waves, 80, 400)
    ref = oracle.audio_encoder_forward(sd, mel, H)
    m = ab.AudioEncoder(80, D, H, L, enc=tel, compute=compute).eval()
    m.load_state_dict(sd)
    y = m(mel.cuda()).float().cpu()
    err = (y - ref).abs()
    return f"max-abs {float(err.max()):.5g} rel-absmax {float(err.max() / ref.abs().max()):.4g} nan {int(torch.isnan(y).sum())}"


for cfg in ((64, 4, 2, False, "fp32"), (64, 4, 2, True, "fp32"), (128, 4, 2, True, "fp32"), (512, 4, 4, False, "fp32"),
            (128, 4, 1, False, "bf16"), (128, 4, 2, True, "bf16"), (256, 4, 2, True, "bf16"), (512, 4, 4, False, "bf16"),
            (512, 4, 4, True, "bf16")) if SECTION in ("all", "enc") else ():
    step(f"encoder {cfg}", lambda cfg=cfg: enc(*cfg))
