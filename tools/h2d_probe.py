import torch, time
N=64*480000
h=torch.empty(N,dtype=torch.float32).pin_memory(); h.uniform_(-0.5,0.5)
d=torch.empty(N,dtype=torch.float32,device='cuda')
def t(fn,n=10):
    fn(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms=t(lambda: d.copy_(h,non_blocking=True))
print('single copy: %.3f ms  %.1f GB/s'%(ms, N*4/ms/1e6))
s=[torch.cuda.Stream() for _ in range(4)]
def split(k):
    def f():
        cur=torch.cuda.current_stream()
        ev=torch.cuda.Event(); ev.record()
        c=N//k
        for i in range(k):
            with torch.cuda.stream(s[i]):
                s[i].wait_event(ev)
                d[i*c:(i+1)*c].copy_(h[i*c:(i+1)*c],non_blocking=True)
            cur.wait_stream(s[i])
    return f
for k in (2,4):
    ms=t(split(k)); print('%d streams: %.3f ms  %.1f GB/s'%(k,ms,N*4/ms/1e6))
import os
print('cpus',os.cpu_count())
try:
    import subprocess; print(subprocess.run(['nvidia-smi','topo','-m'],capture_output=True,text=True).stdout[:1500])
except Exception as e: print(e)
