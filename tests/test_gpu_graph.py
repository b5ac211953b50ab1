"""The compute calls are stream-asynchronous and CUDA-graph capturable (include/asrb200.h)."""
import pytest
import torch

import oracle
from asr_model_b200 import synth

pytestmark = pytest.mark.gpu


def test_fused_forward_replays_from_a_cuda_graph(built_lib):
    import asr_model_b200 as ab
    from asr_model_b200.frontend import LogMel
    sd = oracle.random_encoder_state_dict(80, 256, 2, True, seed=1, perturb=True)
    m = ab.AudioEncoder(80, 256, 4, 2, enc=True, compute="bf16").eval()
    m.load_state_dict(sd)
    fe = LogMel(80, 400)
    waves = synth.make_batch("WHT2", 32000).cuda()
    static_in = waves.clone()
    ref = m.forward_pcm(static_in, fe).clone()            # also creates handle + workspace outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            out = m.forward_pcm(static_in, fe)
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    static_in.copy_(torch.flip(waves, dims=[0]))          # new data, same graph
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, torch.flip(ref, dims=[0]))
