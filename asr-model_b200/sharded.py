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
                 overlap_steps: bool = False):
        """``overlap_steps``: do not wait for a call's exchange before returning; it is waited for at the
        end of the NEXT call (or by ``finish()``), so the gather of step i rides under the compute of
        step i+1.  The local block of the returned tensor is always valid on the current stream; the
        peers' blocks are valid after the next call / ``finish()``."""
        self.compute, self.group, self.micro, self.gather, self.shape_of = compute, group, micro, gather, shape_of
        self.overlap_steps = overlap_steps
        self._inflight = []

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
