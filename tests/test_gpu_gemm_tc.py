"""The tcgen05/TMEM/TMA implicit-GEMM kernel in isolation (asrb_test_gemm_tc) against a plain
PyTorch fp32 reference of the same op on the same 16-bit-rounded operands (the library's operand format)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


ACTS = {0: lambda x: x, 1: F.gelu, 2: F.relu, 3: F.silu, 4: lambda x: F.gelu(F.gelu(x))}


def _depthwise(x, dw_w, dw_b, kw):
    """x [B, T, C], dw_w [kw][C]: zero-padded depthwise conv along T (per utterance)."""
    C = x.shape[-1]
    y = F.conv1d(x.transpose(1, 2), dw_w.t().reshape(C, 1, kw).contiguous(), dw_b, padding=kw // 2, groups=C)
    return y.transpose(1, 2)


def _ref(a, w, bias, res, gamma, beta, taps, epi, act, N, dw=None):
    B, T, K = a.shape
    af = a.float()
    wf = w.float().view(N, taps, K)
    acc = torch.zeros(B, T, N, device=a.device)
    for tap in range(taps):
        sh = tap - taps // 2
        src = torch.zeros_like(af)
        if sh < 0:
            src[:, -sh:] = af[:, :T + sh]
        elif sh > 0:
            src[:, :T - sh] = af[:, sh:]
        else:
            src = af
        acc += src @ wf[:, tap].t()
    acc = acc + bias
    if epi in (1, 4):                               # GLU with [128 value | 128 gate] per 256 rows
        v = acc.view(B, T, N // 256, 2, 128)
        g = (v[..., 0, :] * torch.sigmoid(v[..., 1, :])).reshape(B, T, N // 2)
        if epi == 1:
            return g
        dw_w, dw_b, kw, act2, pos = dw
        return ACTS[act2](_depthwise(g, dw_w, dw_b, kw))
    if res is not None:
        acc = acc + res.float()
    if epi == 3:
        return F.layer_norm(acc, (N,), gamma, beta, 1e-5)
    if epi == 5:
        dw_w, dw_b, kw, act2, pos = dw
        y = ACTS[act2](_depthwise(ACTS[act](acc), dw_w, dw_b, kw))
        return y + pos if pos is not None else y
    return ACTS[act](acc)


CASES = [
    # B, T, K, N, taps, epi, act
    (1, 128, 64, 128, 1, 0, 0),
    (1, 256, 128, 256, 1, 0, 0),
    (1, 300, 512, 512, 1, 0, 1),
    (2, 1001, 512, 1536, 1, 0, 0),
    (3, 200, 128, 128, 3, 0, 1),        # k3 conv: halo must not leak across utterances
    (2, 333, 512, 512, 3, 3, 0),        # k3 conv + LayerNorm (2 chunks of 256)
    (1, 129, 128, 128, 1, 3, 0),        # LN, one 128 chunk, double-buffered accumulators
    (2, 260, 256, 384, 1, 3, 0),        # LN over 3 chunks of 128
    (2, 260, 256, 256, 1, 3, 0),
    (2, 500, 512, 512, 1, 2, 1),        # residual + GELU
    (1, 4000, 2048, 512, 1, 3, 0),      # FFN2 + residual + LN
    (1, 777, 512, 2048, 1, 0, 2),       # FFN1 + ReLU
    (40, 130, 256, 256, 3, 0, 4),       # many small utterances, persistent loop wraps
    (2, 300, 256, 512, 1, 4, 0),        # GLU -> depthwise-15 -> SiLU (tiles overlap by 14 frames)
    (3, 115, 512, 1024, 1, 4, 0),
    (2, 300, 256, 256, 1, 5, 1),        # residual + GELU -> depthwise-3 -> GELU(GELU)
    (2, 127, 512, 512, 1, 5, 1),
    (5, 1001, 512, 512, 1, 5, 1),
    # channel-major kernels (gemm_tct.cu): tile edges (254 frames per residual tile, 112 per GLU tile), tiny T,
    # one channel tile, the generic and the 1024-wide store paths
    (1, 1, 64, 128, 1, 5, 1),
    (2, 2, 128, 128, 1, 5, 1),
    (1, 254, 64, 128, 1, 5, 1),
    (1, 255, 64, 128, 1, 5, 1),
    (3, 257, 128, 256, 1, 5, 1),
    (1, 600, 256, 1024, 1, 5, 1),
    (1, 1, 64, 256, 1, 4, 0),
    (2, 112, 64, 256, 1, 4, 0),
    (2, 113, 128, 256, 1, 4, 0),
    (1, 225, 64, 256, 1, 4, 0),
    (1, 500, 256, 2048, 1, 4, 0),
]


@pytest.mark.parametrize("B,T,K,N,taps,epi,act", CASES)
def test_gemm_tc_matches_torch(built_lib, B, T, K, N, taps, epi, act):
    lib = built_lib.load()
    op = built_lib.operand_dtype()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T + K + N + taps + epi)
    a = (torch.randn(B, T, K, device="cuda", generator=g) * 0.5).to(op)
    w = (torch.randn(N, taps * K, device="cuda", generator=g) / (taps * K) ** 0.5).to(op)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    n_out = N // 2 if epi in (1, 4) else N
    res = (torch.randn(B, T, n_out, device="cuda", generator=g)).to(op) if epi in (2, 3, 5) and (T % 2 == 0 or epi != 3) else None
    gamma = 1 + 0.2 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    dw = None
    dw_args = (None, None, 0, 0, None)
    if epi in (4, 5):
        kw = 15 if epi == 4 else 3
        dw_w = torch.randn(kw, n_out, device="cuda", generator=g) / kw ** 0.5
        dw_b = 0.1 * torch.randn(n_out, device="cuda", generator=g)
        act2 = 3 if epi == 4 else (4 if T % 2 else 1)
        pos = torch.randn(T, n_out, device="cuda", generator=g) if (epi == 5 and T % 2 == 0) else None
        dw = (dw_w, dw_b, kw, act2, pos)
        dw_args = (dw_w.data_ptr(), dw_b.data_ptr(), kw, act2, pos.data_ptr() if pos is not None else None)
    guard = 64 * n_out                                  # sentinel rows before and after: nothing may be written outside the tensor
    buf = torch.full((B * T * n_out + 2 * guard,), 12345.0, device="cuda", dtype=op)
    out = buf[guard: guard + B * T * n_out].view(B, T, n_out)
    out.fill_(float("nan"))
    rc = lib.asrb_test_gemm_tc(a.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr() if res is not None else None,
                               gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), B, T, K, N, taps, epi, act, *dw_args, None)
    built_lib.check(rc, "asrb_test_gemm_tc")
    torch.cuda.synchronize()
    ref = _ref(a, w, bias, res, gamma, beta, taps, epi, act, N, dw)
    assert not torch.isnan(out.float()).any(), "rows were left unwritten"
    assert bool((buf[:guard] == 12345.0).all()) and bool((buf[-guard:] == 12345.0).all()), "wrote outside the output tensor"
    err = (out.float() - ref).abs()
    # the store rounding dominates (fp16: 2^-11 relative; bf16: 2^-8), then the MUFU-based activations (2.5e-4 |x| each)
    tol = (4e-3 + 2e-3 * ref.abs()) if op == torch.float16 else (2e-2 + 1e-2 * ref.abs())
    print(f"gemm_tc epi={epi}: max err {float(err.max()):.5f}  worst err/tol {float((err / tol).max()):.3f}  mean {float(err.mean()):.6f}")
    assert bool((err <= tol).all()), f"max err {float(err.max())} at {int(err.argmax())}"
    assert float(err.mean()) < (6e-4 if op == torch.float16 else 3e-3)
