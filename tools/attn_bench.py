"""The tcgen05 attention kernel alone, timed in a steady loop: python tools/attn_bench.py [B] [T] [D] [H] [iters]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_model_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3001
D = int(sys.argv[3]) if len(sys.argv) > 3 else 512
H = int(sys.argv[4]) if len(sys.argv) > 4 else 4
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 200
lib = _lib.load()
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda") * 0.5).to(_lib.operand_dtype())
out = torch.empty(B, T, D, device="cuda", dtype=_lib.operand_dtype())
st = torch.cuda.current_stream().cuda_stream
for _ in range(min(iters, 20)):
    _lib.check(lib.asrb_test_attention_tc(qkv.data_ptr(), out.data_ptr(), B, T, D, H, st), "attention")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    lib.asrb_test_attention_tc(qkv.data_ptr(), out.data_ptr(), B, T, D, H, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"attention B={B} T={T} D={D} H={H}: {ms:.4f} ms  {4.0 * B * T * T * D / ms / 1e9:.1f} TFLOP/s")
