"""Probe (torchrun, >= 2 GPUs): can a copy-engine cudaMemcpyAsync write through a symmetric-memory MULTICAST address, and how
fast is one multicast push compared with world-1 unicast pushes?  Prints one line per variant on rank 0."""
import ctypes, os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64 * 3001 * 512                     # one rank's block, bf16 elements
t = symm_mem.empty(world * n, dtype=torch.bfloat16, device=dev)
h = symm_mem.rendezvous(t, dist.group.WORLD)
mc = getattr(h, "multicast_ptr", 0)
if rank == 0:
    print("multicast_ptr", hex(mc) if mc else mc, flush=True)
src = torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device=dev)
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
peers = [h.get_buffer(r, (world * n,), torch.bfloat16) for r in range(world)]
st = torch.cuda.Stream()

def unicast():
    with torch.cuda.stream(st):
        for k in range(1, world):
            r = (rank + k) % world
            peers[r][rank * n:(rank + 1) * n].copy_(src, non_blocking=True)
        t[rank * n:(rank + 1) * n].copy_(src, non_blocking=True)

def multicast():
    with torch.cuda.stream(st):
        rc = cudart.cudaMemcpyAsync(ctypes.c_void_p(mc + rank * n * 2), ctypes.c_void_p(src.data_ptr()), n * 2, 3, ctypes.c_void_p(st.cuda_stream))
        assert rc == 0, rc

def check(tag):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ok = all(bool((t[r * n:(r + 1) * n] == float(r + 1)).all()) for r in range(world))
    f = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(f, op=dist.ReduceOp.MIN)
    if rank == 0: print(tag, "correct" if int(f.item()) else "WRONG", flush=True)
    t.zero_(); torch.cuda.synchronize(); dist.barrier()

def timeit(fn, tag):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
    for _ in range(10): fn()
    with torch.cuda.stream(st):
        e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{tag}: {float(ms.item()):.3f} ms per push of {n * 2 / 1e6:.0f} MB to {world - 1} peers", flush=True)

unicast(); check("unicast")
timeit(unicast, "unicast")
if mc:
    try:
        multicast(); check("multicast copy-engine")
        timeit(multicast, "multicast copy-engine")
    except Exception as e:
        if rank == 0: print("multicast copy-engine failed:", repr(e)[:300], flush=True)
dist.destroy_process_group()
