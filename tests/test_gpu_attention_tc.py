"""The tcgen05 flash-style attention kernel in isolation (asrb_test_attention_tc) against a plain
PyTorch fp32 softmax attention on the same 16-bit-rounded q, k, v (the library's operand format)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # B, T, D, H
    (1, 128, 128, 1),
    (2, 300, 256, 2),        # head_dim 128, ragged last tile (keys and queries)
    (2, 1001, 512, 4),
    (1, 3001, 512, 4),       # BASELINE frame count
    (2, 257, 256, 4),        # head_dim 64
    (1, 640, 1024, 16),
    (3, 129, 128, 2),
]


@pytest.mark.parametrize("B,T,D,H", CASES)
def test_attention_tc_matches_torch(built_lib, B, T, D, H):
    lib = built_lib.load()
    g = torch.Generator(device="cuda").manual_seed(B * 7 + T + D + H)
    op = built_lib.operand_dtype()
    qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 1.5).to(op)
    qkv[..., :D] *= 2.0                                   # sharper softmax: the running max really moves
    guard = 64 * D                                         # sentinel rows before and after: nothing may be written outside the tensor
    buf = torch.full((B * T * D + 2 * guard,), 12345.0, device="cuda", dtype=op)
    out = buf[guard: guard + B * T * D].view(B, T, D)
    out.fill_(float("nan"))
    built_lib.check(lib.asrb_test_attention_tc(qkv.data_ptr(), out.data_ptr(), B, T, D, H, None), "asrb_test_attention_tc")
    torch.cuda.synchronize()
    assert bool((buf[:guard] == 12345.0).all()) and bool((buf[-guard:] == 12345.0).all()), "wrote outside the output tensor"
    hd = D // H
    q, k, v = [t.float().view(B, T, H, hd).transpose(1, 2) for t in qkv.split(D, dim=-1)]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    ref = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, T, D)
    assert not torch.isnan(out.float()).any(), "rows were left unwritten"
    err = (out.float() - ref).abs()
    print(f"attention_tc B={B} T={T} D={D} H={H}: max {float(err.max()):.4f} mean {float(err.mean()):.5f} refmax {float(ref.abs().max()):.2f}")
    f16 = op == torch.float16                            # P and O are rounded to the operand format
    assert bool((err <= ((4e-3 + 4e-3 * ref.abs()) if f16 else (2e-2 + 2e-2 * ref.abs()))).all()), float(err.max())
    assert float(err.mean()) < (6e-4 if f16 else 3e-3)
