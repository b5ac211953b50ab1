# Final-code evidence for profiles/: TAG=r2w bash tools/run_profiles.sh   (short bench, ncu launch list, ncu --set full rows of the step and of the front end)
set -x
python bench.py --steps 2 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/${TAG:-r2w}_bench_short.json 2> gpurun_out/${TAG:-r2w}_bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG:-r2w}_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --extras 0 --no-cpu-baseline > gpurun_out/${TAG:-r2w}_ncu_launches.log 2>&1
python tools/prof_step.py 64 1 > gpurun_out/${TAG:-r2w}_prof_step.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"gemm_tc|attn_tc|logmel_kernel" -s 19 -c 19 -o /tmp/${TAG:-r2w}_full_step_B64_enc1 -f python tools/prof_step.py 64 1 > gpurun_out/${TAG:-r2w}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/${TAG:-r2w}_full_step_B64_enc1.ncu-rep gpurun_out/${TAG:-r2w}_ncu_full_step_B64_enc1_metrics.csv gpurun_out/traffic.json
ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 2 -c 1 -o gpurun_out/${TAG:-r2w}_logmel400 -f python tools/prof_frontend.py 256 80 400 4 > gpurun_out/${TAG:-r2w}_ncu_fe400.log 2>&1
ncu --set full --clock-control none -k regex:logmel_kernel -s 2 -c 1 -o /tmp/${TAG:-r2w}_logmel1024 -f python tools/prof_frontend.py 256 80 1024 4 > gpurun_out/${TAG:-r2w}_ncu_fe1024.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG:-r2w}_logmel400.ncu-rep gpurun_out/${TAG:-r2w}_ncu_logmel_256x30s_nfft400_metrics.csv
python tools/ncu_summary.py /tmp/${TAG:-r2w}_logmel1024.ncu-rep gpurun_out/${TAG:-r2w}_ncu_logmel_256x30s_nfft1024_metrics.csv
du -sh gpurun_out
