set -x
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python bench.py --extras 0 --no-cpu-baseline > gpurun_out/r2s_bench_1gpu_same_box.json 2> gpurun_out/r2s_1.err
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/r2s_$n.err | grep '^{' > gpurun_out/r2s_bench_${n}gpu.json
done
python - <<'PY'
import json
for f in ("1gpu_same_box","2gpu","4gpu","8gpu"):
    try:
        d=json.load(open(f"gpurun_out/r2s_bench_{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["result_on_device"]), "strong", d.get("strong_scaling",{}) and round(d["strong_scaling"]["audio_s_per_s"]), [round(x,2) for x in d["per_rank_ms_per_step"]])
    except Exception as e: print(f, "ERR", e)
PY
