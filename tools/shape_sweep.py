import sys, os
sys.path.insert(0, os.getcwd())
import torch, oracle
import asr_model_b200 as ab
from asr_model_b200 import synth
ok=True
for (mels, D, H, L, enc, B, T) in [(80,128,4,1,False,2,37),(80,256,4,2,False,3,300),(80,384,4,1,False,2,513),(128,640,5,1,False,1,700),(80,768,6,1,True,2,260),(80,1024,16,1,False,1,400),(80,512,4,2,True,5,129)]:
    sd = oracle.random_encoder_state_dict(mels, D, L, enc, seed=D, perturb=True)
    x = torch.randn(B, mels, T, generator=torch.Generator().manual_seed(T))
    ref = oracle.audio_encoder_forward(sd, x, H)
    m = ab.AudioEncoder(mels, D, H, L, "gelu", "AbbyNormal", norm=False, enc=enc, compute="bf16").cuda().eval()
    m.load_state_dict(sd)
    y = m(x.cuda()).float().cpu()
    d = (y-ref).abs()
    rel = float(d.max()/ref.abs().max()); outside=float((d > 2e-2+1e-2*ref.abs()).float().mean())
    good = rel <= 2.6e-2 and outside <= 4e-4 and not torch.isnan(y).any()
    ok &= bool(good)
    print((mels,D,H,L,enc,B,T), 'max-abs %.4f rel %.4f outside %.2e'%(float(d.max()),rel,outside), 'OK' if good else 'FAIL', flush=True)
print('SWEEP', 'OK' if ok else 'FAIL')
