"""Host logic of the utterance-sharded driver on the gloo backend, world_size 2 (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from asr_model_b200.sharded import ShardedEncoder, gather_outputs, micro_batches, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert micro_batches(0, 10, 4) == [(0, 4), (4, 8), (8, 10)] and micro_batches(3, 9, 0) == [(3, 9)]


def _fake_compute(w):            # stands in for forward_pcm: [n, N] -> [n, T=3, D=4], depends on content
    return torch.stack([w[:, :3] * (d + 1) for d in range(4)], dim=-1)


def _worker(rank, world, port, total, micro, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        waves = torch.randn(total, 8, generator=g)
        lo, hi = shard_range(total, rank, world)
        out = ShardedEncoder(_fake_compute, micro=micro)(waves[lo:hi], total=total)
        ok = torch.equal(out, _fake_compute(waves))
        out2, _ = gather_outputs(_fake_compute(waves[lo:hi]), total)
        ok2 = torch.equal(out2, _fake_compute(waves))
        if total % world == 0:                       # shape_of known: on gloo / CPU the peer exchange must step aside
            se = ShardedEncoder(lambda w, out=None: _fake_compute(w) if out is None else out.copy_(_fake_compute(w)),
                                micro=micro, shape_of=lambda w: (3, 4, torch.float32))
            ok2 = ok2 and torch.equal(se(waves[lo:hi], total=total), _fake_compute(waves)) and se._peer is None
        if total % world == 0:                       # deferred wait: step i's exchange finishes during step i+1
            se = ShardedEncoder(_fake_compute, micro=micro, overlap_steps=True)
            o1 = se(waves[lo:hi], total=total)
            o2 = se(waves[lo:hi] * 2, total=total)
            se.finish()
            ok2 = ok2 and torch.equal(o1, _fake_compute(waves)) and torch.equal(o2, _fake_compute(waves * 2))
        q.put((rank, bool(ok), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,micro", [(8, 0), (8, 2), (7, 0), (6, 2)])
def test_two_rank_gather_reassembles_the_batch(total, micro):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, micro, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    assert all(ok and ok2 for _, ok, ok2 in res), res


def test_single_process_without_dist_runs_locally():
    w = torch.randn(5, 8)
    out = ShardedEncoder(_fake_compute, micro=2)(w)
    assert torch.equal(out, _fake_compute(w))
