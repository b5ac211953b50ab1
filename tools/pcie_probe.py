"""torchrun probe: host <-> device copy bandwidth per rank with all ranks copying at once, with and without NUMA-local
placement, from ordinary and from symmetric device memory."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from asr_model_b200.sharded import pin_to_local_numa
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
aff = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 1) + 63) // 64)
info = {"rank": rank, "cpu_count": os.cpu_count(), "allowed": len(os.sched_getaffinity(0)), "nvml_affinity_bits": [hex(int(a)) for a in aff]}
try:
    info["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    info["gpu_numa"] = open(f"/sys/bus/pci/devices/{pynvml.nvmlDeviceGetPciInfo(h).busId.decode().lower()[4:] if isinstance(pynvml.nvmlDeviceGetPciInfo(h).busId, bytes) else pynvml.nvmlDeviceGetPciInfo(h).busId.lower()[4:]}/numa_node").read().strip()
except Exception as e:
    info["numa_err"] = repr(e)[:100]

def bw(src, dst, n=5):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(n): dst.copy_(src, non_blocking=True)
        e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    return src.numel() * src.element_size() * n / e0.elapsed_time(e1) / 1e6

nbytes = 197 * 1024 * 1024
def run(tag):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    res = {"h2d": bw(host, d), "d2h": bw(d, host)}
    import torch.distributed._symmetric_memory as symm
    s = symm.empty(nbytes, dtype=torch.uint8, device=dev); symm.rendezvous(s, dist.group.WORLD)
    res["d2h_from_symmetric"] = bw(s, host)
    # both directions at once
    host2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    st2 = torch.cuda.Stream()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(st2): d.copy_(host2, non_blocking=True)
        host.copy_(s, non_blocking=True)
    torch.cuda.synchronize()
    res["duplex_each_dir"] = nbytes * 5 / (time.perf_counter() - t0) / 1e9
    info[tag] = {k: round(v, 1) for k, v in res.items()}

run("default_placement")
info["pinned_to"] = pin_to_local_numa(local)
info["pinned_to"] = len(info["pinned_to"]) if info["pinned_to"] else None
run("numa_local")
for r in range(world):
    dist.barrier()
    if r == rank: print(info, flush=True)
dist.destroy_process_group()
