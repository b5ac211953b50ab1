#!/usr/bin/env python
"""Headline benchmark: audio-seconds per second of the fused log-mel + AudioEncoder forward.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--enc 0|1]

Workload (BASELINE.json configs[1]): 64 x 30 s synthetic clips per GPU, 16 kHz, n_fft 400,
hop 160, 80 mel -> AudioEncoder(D=512, H=4, L=4), bf16 tensor-core path, random-init weights.
A step = one pass of the hot path (PCM -> hidden states) over one batch.  Prints ONE JSON line
(rank 0).  `value`: inputs resident in HBM, CUDA-event timed, max over ranks.  `e2e`: the
same through the public API from pinned HOST buffers (H2D of the PCM and D2H of the
per-utterance pooled hidden state inside the timed region).  `roofline`: the dominant kernel
(tcgen05 conv-GEMM + LayerNorm) from per-launch CUDA events.  `cpu_baseline`: the oracle
(port of the reference) on this box's host cores on a bounded sample.
Under torchrun (N > 1) each rank runs its own 64-clip shard (weak scaling) and pushes its encoder
outputs into every peer's gathered tensor over NVLink (symmetric memory + copy engines; NCCL
send/recv if that is unavailable); the exchange of step i overlaps the compute of step i+1 and
all of it is inside the timed region.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SECS, SR, HOP, N_FFT, MELS = 30, 16000, 160, 400, 80
DIMS, HEAD, LAYER = 512, 4, 4
PER_GPU_BATCH = 64


def flops_per_frame(enc: bool, T: int) -> float:
    """Algorithmic encoder flops per frame, SURVEY.md 8d."""
    f = 2 * 3 * MELS * DIMS + LAYER * 12 * DIMS * DIMS
    if enc:
        f += 8 * DIMS * DIMS + 4 * T * DIMS + 8192 * DIMS
    return float(f)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 5 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_port_throughput(batch, secs, enc, repeats, warm):
    """The oracle (CPU port of the reference's path: essentials.py:469-490 per utterance +
    AudioEncoder forward, fp32) on the host cores.  Returns (audio-s/s, threads)."""
    import torch
    import oracle
    from asr_model_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    waves = synth.white_noise_batch(batch, secs * SR)
    sd = oracle.random_encoder_state_dict(MELS, DIMS, LAYER, enc, seed=0)
    times = []
    with torch.no_grad():
        for i in range(warm + repeats):
            t = time.perf_counter()
            mel = oracle.log_mel_batch(waves, MELS, N_FFT)
            oracle.audio_encoder_forward(sd, mel, HEAD)
            if i >= warm:
                times.append(time.perf_counter() - t)
    return batch * secs / statistics.median(times), torch.get_num_threads(), times


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch, secs = 4, SECS                       # bounded sample of the 64 x 30 s workload
    val, threads, times = cpu_port_throughput(batch, secs, bool(args.enc), args.steps, args.warmup)
    sample = f"{batch} x {secs} s clips per step (of the {PER_GPU_BATCH} x {SECS} s workload), {args.steps} timed steps"
    print(json.dumps({
        "impl": "reference", "metric": "audio_seconds_per_second", "value": val, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, world):
    return {"workload": f"log-mel (16 kHz, n_fft {N_FFT}, hop {HOP}, {MELS} mel) + AudioEncoder(D={DIMS}, H={HEAD}, L={LAYER}, "
                        f"enc={bool(args.enc)}) forward, {PER_GPU_BATCH} x {SECS} s clips per GPU",
            "global_batch": PER_GPU_BATCH * world, "clip_seconds": SECS, "frames_per_clip": 1 + SECS * SR // HOP,
            "parallelism": f"utterance-sharded x{world}" + (", encoder outputs gathered on every rank over NVLink" if world > 1 else ""),
            "l2": "inputs larger than L2 (123 MB PCM + 197 MB activations per tensor per step)",
            "weights": "random init", "enc": bool(args.enc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--enc", type=int, default=0, help="1 = with the optional TransformerEncoderLayer (model.py:138)")
    ap.add_argument("--micro", type=int, default=0, help="micro-batch inside a step for the gather (N > 1); 0 = whole shard")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    if rank != 0:
        entry.build()
    import asr_model_b200 as ab
    from asr_model_b200 import _lib, synth
    from asr_model_b200.frontend import LogMel
    from asr_model_b200.sharded import ShardedEncoder

    B, N = PER_GPU_BATCH, SECS * SR
    T = 1 + N // HOP
    torch.manual_seed(0)                      # random init = the reference constructors' default init
    enc = ab.AudioEncoder(MELS, DIMS, HEAD, LAYER, "gelu", "AbbyNormal", norm=False, enc=bool(args.enc), compute="bf16").eval()
    fe = LogMel(MELS, N_FFT, HOP, device=dev)
    pcm_host = synth.white_noise_batch(B, N, seed=1234 + rank).pin_memory()
    pcm = pcm_host.to(dev)

    def hot(w, out=None):
        return enc.forward_pcm(w, fe, out=out)

    # N > 1: the exchange of step i overlaps the compute of step i+1 (waited for one step later)
    sharded = ShardedEncoder(hot, micro=args.micro if world > 1 else 0, gather=world > 1,
                             shape_of=lambda w: (fe.num_frames(w.shape[1]), DIMS, torch.bfloat16),
                             overlap_steps=world > 1)

    def step_resident():
        return sharded(pcm, total=B * world)

    def sync():
        sharded.finish()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    sync()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    sync()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    value = B * world * SECS / (ms_step * 1e-3)

    # ---- end to end from pinned host memory through the public API ----
    # Every step: H2D of that step's PCM (pinned -> device, copy stream) + the fused forward + D2H
    # of the time-pooled hidden state.  Double-buffered like a real feeder: step i+1's PCM streams
    # in while step i computes; nothing is reused between steps.
    pooled_host = [torch.empty(B, DIMS, dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_in = [torch.empty_like(pcm) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def feed(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[k])                 # the buffer's previous consumer is done
            dev_in[k].copy_(pcm_host, non_blocking=True)
            ready[k].record(copy_stream)

    def run_e2e(n):
        for k in range(2):
            freed[k].record(main_stream)
        feed(0)
        for i in range(n):
            k = i & 1
            if i + 1 < n:
                feed(i + 1)
            main_stream.wait_event(ready[k])
            h = sharded(dev_in[k], total=B * world)
            lo = rank * B
            pooled_host[k].copy_(torch.mean(h[lo:lo + B], dim=1, dtype=torch.float32), non_blocking=True)
            freed[k].record(main_stream)

    run_e2e(3)
    sync()
    e0.record()
    run_e2e(args.steps)
    e1.record()
    sync()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = B * world * SECS / (float(t2.item()) / args.steps * 1e-3)

    # ---- per-launch CUDA events: roofline of the dominant kernel, launch count ----
    lib = _lib.load()
    sync()
    lib.asrb_profile_begin()
    psteps = min(args.steps, 5)
    for _ in range(psteps):
        hot(pcm)
    recs = _lib.profile_records()
    by = {}
    for tag, ms, fl, byt in recs:
        d = by.setdefault(tag, {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["n"] += 1; d["ms"] += ms; d["flops"] += fl; d["bytes"] += byt
    launches_per_step = len(recs) // psteps
    pk = peaks()
    step_ms_prof = sum(d["ms"] for d in by.values()) / psteps
    kernels = {t: {"launches_per_step": d["n"] // psteps, "ms_per_step": d["ms"] / psteps,
                   "share": d["ms"] / psteps / step_ms_prof if step_ms_prof else None,
                   "tflops": d["flops"] / d["ms"] / 1e9 if d["ms"] and d["flops"] else None,
                   "gbs": d["bytes"] / d["ms"] / 1e6 if d["ms"] else None} for t, d in by.items()}
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["bytes_per_launch"]
    except Exception:
        pass
    # The dominant kernels are the tcgen05 implicit GEMMs: `gemm_tc_kernel<BN, EPI>` (lanes = frames: stem,
    # k3+LayerNorm) and `gemm_tct_kernel<EPI, NO, ACT2>` (lanes = channels: 1x1+GLU+depthwise-15,
    # 1x1+residual+depthwise-3), > 90 % of the step, three of them within 2x of each other.  The roofline entry
    # aggregates their launches (per launch = totals / launches); the per-kernel numbers are in "kernels".
    fam = {t: d for t, d in by.items() if t.startswith("gemm_tc")}
    other = max((t for t in by if t not in fam), key=lambda t: by[t]["ms"], default=None)
    fam_ms = sum(d["ms"] for d in fam.values())
    if fam and fam_ms >= (by[other]["ms"] if other else 0.0):
        n = sum(d["n"] for d in fam.values())
        fl = sum(d["flops"] for d in fam.values())
        tr = sum(traffic[t] * (d["n"]) for t, d in fam.items() if t in traffic)
        tr_n = sum(d["n"] for t, d in fam.items() if t in traffic)
        achieved = fl / fam_ms / 1e9                       # TFLOP/s
        roof = {"kernel": "tcgen05 GEMMs gemm_tc_kernel<BN,EPI> + gemm_tct_kernel<EPI,NO,ACT2> (%d launches per step: %s)" % (n // psteps, ", ".join(sorted(fam))),
                "bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_sustained"],
                "traffic": tr / tr_n if tr_n else None, "traffic_unit": "DRAM bytes per launch (ncu, profiles/traffic.json)",
                "peak_source": pk["src"] + " (sustained bf16, kernel timed inside a long step)",
                "avg_launch_ms": fam_ms / n, "flops_per_launch": fl / n,
                "share_of_step": fam_ms / psteps / step_ms_prof if step_ms_prof else None}
    else:
        dd = by[other]
        achieved = dd["bytes"] / dd["ms"] / 1e6            # GB/s
        roof = {"kernel": other, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic.get(other), "peak_source": pk["src"],
                "avg_launch_ms": dd["ms"] / dd["n"], "bytes_per_launch": dd["bytes"] / dd["n"]}
    enc_flops = flops_per_frame(bool(args.enc), T) * B * T
    whole = {"encoder_algorithmic_tflop_per_step": enc_flops / 1e12,
             "encoder_tensor_roofline_frac_of_step": enc_flops / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"] if world == 1 else None}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, threads, times = cpu_port_throughput(4, 10, bool(args.enc), 5, 2)     # BASELINE config 1: 4 x 10 s
        cpu = {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port",
               "sample": "BASELINE config 1: 4 x 10 s clips, fp32, median of 5 after 2 warm-ups",
               "cpu_count": os.cpu_count()}
    if rank == 0:
        print(json.dumps({
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_val, "unit": "audio-s/s", "h2d_bytes_per_step": B * N * 4, "d2h_bytes_per_step": B * DIMS * 4,
                    "note": "every step: H2D of its fp32 PCM from pinned memory (double-buffered on a copy stream) + fused forward + D2H of the time-pooled hidden state"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels, "whole_step": whole,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
