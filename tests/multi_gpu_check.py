"""Run under torchrun (one rank per GPU): the gathered encoder outputs of the utterance-sharded driver must
equal every rank's own computation of the whole batch, with the copy-engine peer exchange and with the NCCL
point-to-point fallback, with and without the deferred wait.  Prints 'MULTI_GPU_OK <world>' on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import oracle
    import asr_model_b200 as ab
    from asr_model_b200 import synth
    from asr_model_b200.frontend import LogMel
    from asr_model_b200.sharded import ShardedEncoder, shard_range

    per, n = 3, 24000
    total = per * world
    sd = oracle.random_encoder_state_dict(80, 256, 2, False, seed=3, perturb=True)
    enc = ab.AudioEncoder(80, 256, 4, 2, enc=False, compute="bf16").eval()
    enc.load_state_dict(sd)
    fe = LogMel(80, 400, device=dev)
    waves = synth.white_noise_batch(total, n, seed=7).to(dev)            # same on every rank
    ref = enc.forward_pcm(waves, fe).clone()
    lo, hi = shard_range(total, rank, world)
    shape_of = lambda w: (fe.num_frames(w.shape[1]), 256, torch.bfloat16)
    hot = lambda w, out=None: enc.forward_pcm(w, fe, out=out)
    ok = True
    for exchange, mc in (("peer", True), ("peer", False), ("nccl", False)):      # multicast push (world > 2), unicast pushes, NCCL
        for micro in (0, 2):
            se = ShardedEncoder(hot, micro=micro, shape_of=shape_of, exchange=exchange, multicast=mc)
            ok &= bool(torch.equal(se(waves[lo:hi], total=total), ref))
        se = ShardedEncoder(hot, shape_of=shape_of, overlap_steps=True, exchange=exchange, multicast=mc)
        outs = [se(waves[lo:hi] * (1.0 - 0.1 * i), total=total) for i in range(4)]   # more steps than gathered slots
        se.finish()
        torch.cuda.synchronize()
        ok &= bool(torch.equal(outs[-1], enc.forward_pcm(waves * 0.7, fe)))
        ok &= bool(torch.equal(outs[-2], enc.forward_pcm(waves * 0.8, fe)))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_OK" if int(flag.item()) == 1 else "MULTI_GPU_FAIL", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
