"""Utterance-sharded multi-GPU driver.  The reference has no parallelism
at all; utterances are independent in eval mode, so the batch is cut into contiguous
per-rank ranges with replicated weights and NO collective on the data path.  The single
exchange is the gather of encoder outputs: each rank's ``[B/G, T, D]`` block is all-gathered
in place into the ``[B, T, D]`` result (NCCL over NVLink), issued per micro-batch on the
communication stream so it overlaps the next micro-batch's compute.

One process per GPU (``torch.distributed``, backend nccl; gloo for the CPU tests of this
host logic).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def pin_to_local_numa(gpu_index: int) -> Optional[List[int]]:
    """Restrict this process to the host cores NVML reports as local to GPU ``gpu_index`` (so pinned buffers allocated
    afterwards are first-touched on that NUMA node and the feeder thread sits next to its PCIe root port).  With one
    process per GPU this keeps 8 ranks from sharing node 0's memory controllers (round 1: 22 GB/s per GPU of H2D at
    N = 8 against 55 GB/s alone).  Returns the core list, or None when NVML / affinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``n`` utterances: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def micro_batches(lo: int, hi: int, size: int) -> List[Tuple[int, int]]:
    return [(s, min(s + size, hi)) for s in range(lo, hi, size)] if size > 0 else [(lo, hi)]


def gather_outputs(local: torch.Tensor, total: int, group=None, out: Optional[torch.Tensor] = None,
                   async_op: bool = False):
    """All-gather per-rank ``[n_r, T, D]`` blocks into ``[total, T, D]`` (rank order).  Equal
    shards use one in-place ``all_gather_into_tensor``; ragged shards fall back to padding to
    the largest shard.  Returns ``(out, work)``; ``work`` is None when ``async_op`` is False."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    T, D = local.shape[1], local.shape[2]
    if out is None:
        out = torch.empty(total, T, D, dtype=local.dtype, device=local.device)
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        lo, hi = shard_range(total, rank, world)
        if local.data_ptr() != out[lo:hi].data_ptr():
            out[lo:hi].copy_(local)
        work = dist.all_gather_into_tensor(out.view(-1), out[lo:hi].reshape(-1), group=group, async_op=async_op)
        return out, work
    big = max(sizes)
    pad = torch.zeros(big, T, D, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]].copy_(local)
    buf = torch.empty(world * big, T, D, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1), pad.view(-1), group=group)
    o = 0
    for r in range(world):
        out[o:o + sizes[r]].copy_(buf[r * big: r * big + sizes[r]])
        o += sizes[r]
    return out, None


class PeerExchange:
    """Gathered ``[total, T, D]`` tensors in symmetric memory (torch.distributed._symmetric_memory: every
    rank maps every peer's copy over NVLink / NVSwitch).  A rank PUSHES its block into each peer's copy with
    ``cudaMemcpyAsync`` on a side stream, i.e. on the copy engines: no SM is taken from the persistent
    tensor-core kernels of the next step, which is what an NCCL send/recv kernel cannot do when every SM
    is occupied (measured at 8 x B200: 6.3 ms per step with NCCL point-to-point against 3.0 ms of compute).
    A device-side barrier on the signal pads closes each step's exchange.  ``slots`` gathered tensors are
    cycled so that a peer running one step ahead never writes into a tensor the consumer may still read.

    When the symmetric allocation has a MULTICAST mapping (NVSwitch / NVLS), a push is ONE copy-engine write through
    the multicast address: the switch replicates it to every GPU, so a rank reads its block from HBM once and sends it
    over its links once instead of world-1 times (at 8 GPUs: 197 MB instead of 1.38 GB of egress and of source reads
    per step).  The write also lands on the sender's own replica, with the bytes that are already there.
    ``multicast=False`` (or no multicast support) keeps the staggered unicast pushes."""

    def __init__(self, total: int, T: int, D: int, dtype, device, group=None, slots: int = 3, multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.shape, self.dtype = (total, T, D), dtype
        self.bufs, self.hdls, self.peers = [], [], []
        for _ in range(slots):
            t = symm_mem.empty(total, T, D, dtype=dtype, device=device)
            h = symm_mem.rendezvous(t, self.group)
            self.bufs.append(t)
            self.hdls.append(h)
            self.peers.append([h.get_buffer(r, self.shape, dtype) if r != self.rank else t for r in range(self.world)])
        self.stream = torch.cuda.Stream(device)
        self.step = 0
        self.mc_ptrs = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.hdls] if multicast and self.world > 2 else []
        self._memcpy = None
        if self.mc_ptrs and all(self.mc_ptrs):
            import ctypes
            try:
                rt = ctypes.CDLL("libcudart.so.12")
                rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
                rt.cudaMemcpyAsync.restype = ctypes.c_int
                self._memcpy = rt.cudaMemcpyAsync
            except OSError:
                self._memcpy = None
        self.row_bytes = T * D * torch.empty(0, dtype=dtype).element_size()

    def next_slot(self) -> int:
        k = self.step % len(self.bufs)
        self.step += 1
        return k

    def push(self, slot: int, lo: int, hi: int):
        """Rows [lo, hi) of this rank's copy were produced on the current stream: send them to every peer."""
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            src = self.bufs[slot][lo:hi]
            if self._memcpy is not None:                        # one write, replicated by the switch
                rc = self._memcpy(self.mc_ptrs[slot] + lo * self.row_bytes, src.data_ptr(), (hi - lo) * self.row_bytes, 3,
                                  self.stream.cuda_stream)
                if rc != 0:
                    raise RuntimeError(f"cudaMemcpyAsync to the multicast mapping failed ({rc})")
                return
            for k in range(1, self.world):                      # staggered so the peers' ingress ports are used evenly
                r = (self.rank + k) % self.world
                self.peers[slot][r][lo:hi].copy_(src, non_blocking=True)

    def close(self, slot: int) -> torch.cuda.Event:
        """All ranks' pushes of this step have landed once the returned event has completed."""
        with torch.cuda.stream(self.stream):
            self.hdls[slot].barrier(channel=slot)
            done = torch.cuda.Event()
            done.record(self.stream)
        return done


class _EventWork:                       # the same .wait() surface as a c10d Work
    def __init__(self, ev):
        self.ev = ev

    def wait(self):
        torch.cuda.current_stream().wait_event(self.ev)


class ShardedEncoder:
    """Runs ``compute(wave_shard[, out=]) -> [n, T, D]`` on this rank's utterances and gathers.

    ``compute`` is the fused hot path (``AudioEncoder.forward_pcm`` bound to a ``LogMel`` plan) in
    production and any shape-preserving function in the gloo CPU tests.  With ``micro`` > 0 the
    shard is processed in micro-batches.  For equal shards every micro-batch is computed straight
    into its slot of the gathered ``[total, T, D]`` tensor and exchanged with grouped point-to-point
    sends/receives (one NCCL group per micro-batch, no staging copies) while the next micro-batch
    computes; ragged shards fall back to a padded all-gather.
    """

    def __init__(self, compute: Callable[..., torch.Tensor], group=None, micro: int = 0, gather: bool = True,
                 shape_of: Optional[Callable[[torch.Tensor], Tuple[int, int, torch.dtype]]] = None,
                 overlap_steps: bool = False, exchange: str = "auto", slots: int = 3, multicast: bool = True):
        """``overlap_steps``: do not wait for a call's exchange before returning; it is waited for at the
        end of the NEXT call (or by ``finish()``), so the gather of step i rides under the compute of
        step i+1.  The local block of the returned tensor is always valid on the current stream; the
        peers' blocks are valid after the next call / ``finish()``."""
        self.compute, self.group, self.micro, self.gather, self.shape_of = compute, group, micro, gather, shape_of
        self.overlap_steps = overlap_steps
        self._inflight = []
        # "peer": copy-engine pushes through symmetric memory (PeerExchange); "nccl": grouped send/recv;
        # "auto": peer on the NCCL backend when shape_of is known and the rendezvous succeeds, else nccl
        self.exchange = exchange
        import os
        self.slots, self.multicast = slots, multicast and os.environ.get("ASRB_MULTICAST", "1") != "0"
        self._peer: Optional[PeerExchange] = None
        self._peer_key = None

    def _peer_exchange(self, total, T, D, dt, device) -> Optional[PeerExchange]:
        if self.exchange == "nccl" or device.type != "cuda" or dist.get_backend(self.group) != "nccl":
            return None
        key = (total, T, D, dt)
        if self._peer is None or self._peer_key != key:
            self.finish()
            peer, err = None, None
            try:
                peer = PeerExchange(total, T, D, dt, device, self.group, slots=self.slots, multicast=self.multicast)
            except Exception as e:                   # no symmetric memory on this system
                err = e
            # every rank must end up on the same protocol: agree on the outcome before anyone uses it
            ok = torch.tensor([1 if peer is not None else 0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if self.exchange == "peer":
                    raise RuntimeError(f"symmetric-memory peer exchange unavailable on at least one rank: {err}")
                self.exchange = "nccl"               # all ranks together: NCCL point-to-point
                return None
            self._peer, self._peer_key = peer, key
        return self._peer

    def finish(self):
        for w in self._inflight:
            w.wait()
        self._inflight = []

    def _run(self, waves, out=None):
        if out is not None and self.shape_of is not None:
            return self.compute(waves, out=out)
        y = self.compute(waves)
        if out is not None:
            out.copy_(y)
            return out
        return y

    def __call__(self, waves: torch.Tensor, total: Optional[int] = None) -> torch.Tensor:
        """``waves`` holds THIS rank's utterances ``[n_r, N]`` (already resident on the rank's
        device); ``total`` is the global batch.  Returns the gathered ``[total, T, D]`` (or the
        local block when ``gather`` is False)."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n_local = waves.shape[0]
        total = n_local * world if total is None else total
        lo, hi = shard_range(total, rank, world)
        assert hi - lo == n_local, f"rank {rank} holds {n_local} utterances, partition says {hi - lo}"
        if world == 1 or not self.gather:
            outs = [self._run(waves[s:e]) for s, e in micro_batches(0, n_local, self.micro)]
            return outs[0] if len(outs) == 1 else torch.cat(outs)
        if total % world != 0:
            local = torch.cat([self._run(waves[s:e]) for s, e in micro_batches(0, n_local, self.micro)])
            return gather_outputs(local, total, self.group)[0]
        if self.shape_of is not None:
            T, D, dt = self.shape_of(waves)
            px = self._peer_exchange(total, T, D, dt, waves.device)
            if px is not None:
                slot = px.next_slot()
                out = px.bufs[slot]
                for s, e in micro_batches(0, n_local, self.micro):
                    self._run(waves[s:e], out=out[lo + s: lo + e])
                    px.push(slot, lo + s, lo + e)
                work = _EventWork(px.close(slot))
                if self.overlap_steps:
                    prev, self._inflight = self._inflight, [work]
                    for w in prev:
                        w.wait()
                else:
                    work.wait()
                return out
        out = None
        pending = []
        for s, e in micro_batches(0, n_local, self.micro):
            if out is None:
                if self.shape_of is not None:
                    T, D, dt = self.shape_of(waves)
                    out = torch.empty(total, T, D, dtype=dt, device=waves.device)
                    self._run(waves[s:e], out=out[lo + s: lo + e])
                else:
                    y = self.compute(waves[s:e])
                    out = torch.empty(total, y.shape[1], y.shape[2], dtype=y.dtype, device=y.device)
                    out[lo + s: lo + e].copy_(y)
            else:
                self._run(waves[s:e], out=out[lo + s: lo + e])
            ops = []
            for r in range(world):                      # every peer's slice of this micro-batch lands in place
                if r == rank:
                    continue
                peer = dist.get_global_rank(self.group, r) if self.group is not None else r
                ops.append(dist.P2POp(dist.isend, out[lo + s: lo + e], peer, self.group))
                ops.append(dist.P2POp(dist.irecv, out[r * n_local + s: r * n_local + e], peer, self.group))
            pending.extend(dist.batch_isend_irecv(ops))
        if self.overlap_steps:
            prev, self._inflight = self._inflight, pending
            for w in prev:
                w.wait()
        else:
            for w in pending:
                w.wait()
        return out
