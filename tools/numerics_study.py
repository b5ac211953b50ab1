"""CPU emulation of the bf16 encoder variant's rounding sites.

Every tensor-core GEMM is emulated as fp32 matmul over operands rounded the way the kernel
rounds them; everything else is fp32 like the kernels' epilogues.  A 'policy' names, per layer
and per site, how the operand is held:

    'b'  bf16               'h'  fp16              's'  split bf16 (hi + lo: two MMAs)       'f'  fp32 (exact)

Sites: stem activation a0, stem weight, and per layer  X (k3 conv input), Wc, Y (LN output: point1
operand AND residual), W1, U, W2, and the store of the layer output (next X / final out).

    python tools/numerics_study.py            # prints the table quoted in DESIGN.md
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test infrastructure; this tool is not product code)
from oracle.encoder import fold_weight_norm, sinusoids, BN_EPS, LN_EPS  # noqa: E402
from asr_model_b200 import synth  # noqa: E402


def rnd(x, how):
    if how == "f":
        return x
    if how == "b":
        return x.bfloat16().float()
    if how == "h":
        return x.half().float()
    if how == "s":
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    raise ValueError(how)


def emulate(sd, mel, pol, L, final_store="b"):
    """pol: dict with keys a0, ws and per-layer lists X, Wc, Y, Yres, W1, U, W2 (strings of length L)."""
    x = rnd(mel, pol["a0"])
    x = F.conv1d(x, rnd(sd["conv1.0.weight"], pol["ws"]), sd["conv1.0.bias"], padding=1)
    x = F.gelu(x)                                   # layer 0's leading GELU is fused into the stem
    for i in range(L):
        D = x.shape[1]
        p = f"encoder.{i}."
        xin = rnd(x, pol["X"][i])
        w = fold_weight_norm(sd[p + "1.parametrizations.weight.original0"], sd[p + "1.parametrizations.weight.original1"])
        y = F.conv1d(xin, rnd(w, pol["Wc"][i]), sd[p + "1.bias"], padding=1)
        y = F.layer_norm(y.transpose(1, -1), (D,), sd[p + "2.gamma"], sd[p + "2.beta"], LN_EPS).transpose(1, -1)
        yop = rnd(y, pol["Y"][i])
        yres = rnd(y, pol["Yres"][i])
        q = p + "3."
        g = F.conv1d(yop, rnd(sd[q + "point1.weight"], pol["W1"][i]), sd[q + "point1.bias"])
        g = F.glu(g, dim=1)
        g = F.conv1d(g, sd[q + "depth.weight"], sd[q + "depth.bias"], padding=7, groups=D)
        g = F.batch_norm(g, sd[q + "bn.running_mean"], sd[q + "bn.running_var"], sd[q + "bn.weight"], sd[q + "bn.bias"], training=False, eps=BN_EPS)
        u = rnd(F.silu(g), pol["U"][i])
        z = F.conv1d(u, rnd(sd[q + "point2.weight"], pol["W2"][i]), sd[q + "point2.bias"]) + yres
        z = F.gelu(z)
        z = F.conv1d(z, sd[p + "5.weight"], sd[p + "5.bias"], padding=1, groups=D)
        z = F.gelu(z)
        x = F.gelu(z) if i < L - 1 else z            # next layer's leading GELU is fused here
    x = x.permute(0, 2, 1).contiguous()
    x = x + sinusoids(x.shape[1], x.shape[-1])
    return rnd(x, final_store)


def policy(L, base="b", **over):
    pol = {"a0": base, "ws": base}
    for k in ("X", "Wc", "Y", "Yres", "W1", "U", "W2"):
        pol[k] = [base] * L
    for k, v in over.items():
        if k in ("a0", "ws"):
            pol[k] = v
        else:
            for i, how in v.items():
                pol[k][i if i >= 0 else L + i] = how
    return pol


def report(name, y, ref):
    err = (y - ref).abs()
    tol = 2e-2 + 1e-2 * ref.abs()
    outside = int((err > tol).sum())
    print(f"{name:58s} max-abs {float(err.max()):.4f}  worst err/tol {float((err / tol).max()):.3f}  outside {outside:6d} / {err.numel()}  fro-rel {float(err.norm() / ref.norm()):.5f}")


def main():
    torch.set_num_threads(os.cpu_count())
    for D, L, perturb, kinds, T in ((512, 4, False, "WHT2", 1001), (512, 4, True, "WH", 1001)):
        sd = oracle.random_encoder_state_dict(80, D, L, False, seed=3, perturb=perturb)
        waves = synth.make_batch(kinds, (T - 1) * 160)
        mel = oracle.log_mel_batch(waves, 80, 400)
        ref = oracle.audio_encoder_forward(sd, mel, 4)
        print(f"== D={D} L={L} perturb={perturb} absmax {float(ref.abs().max()):.3f}")
        last = {-1: "s"}
        lastf = {-1: "f"}
        lasth = {-1: "h"}
        cases = {
            "all bf16 (round 1 kernels)": policy(L),
            "all bf16, final store fp32": (policy(L), "f"),
            "stem split": policy(L, a0="s"),
            "last layer: act split (X, Y, U), res fp32": policy(L, X=last, Y=last, U=last, Yres=lastf),
            "last layer: act split + weights split": policy(L, X=last, Y=last, U=last, Yres=lastf, Wc=last, W1=last, W2=last),
            "last layer: only res fp32": policy(L, Yres=lastf),
            "last layer exact (fp32 everything)": policy(L, X=lastf, Y=lastf, U=lastf, Yres=lastf, Wc=lastf, W1=lastf, W2=lastf),
            "last layer fp16 (act + weights), res fp16": policy(L, X=lasth, Y=lasth, U=lasth, Yres=lasth, Wc=lasth, W1=lasth, W2=lasth),
            "last layer fp16, res fp32": policy(L, X=lasth, Y=lasth, U=lasth, Yres=lastf, Wc=lasth, W1=lasth, W2=lasth),
            "all fp16": policy(L, base="h"),
            "all act split, weights bf16": policy(L, a0="s", X={i: "s" for i in range(L)}, Y={i: "s" for i in range(L)}, U={i: "s" for i in range(L)}, Yres={i: "f" for i in range(L)}),
            "last two layers fp16": policy(L, **{k: {-1: "h", -2: "h"} for k in ("X", "Y", "U", "Yres", "Wc", "W1", "W2")}),
            "last layer split act + split W2 only": policy(L, X=last, Y=last, U=last, Yres=lastf, W2=last),
            "last layer: U, W2 split + res fp32 (point2 only)": policy(L, U=last, W2=last, Yres=lastf),
        }
        for name, c in cases.items():
            pol, fs = c if isinstance(c, tuple) else (c, "b")
            report(name, emulate(sd, mel, pol, L, fs), ref)


if __name__ == "__main__":
    main()
