#!/usr/bin/env python
"""Selected raw metrics per launch from an .ncu-rep -> CSV (the files under profiles/), optionally the DRAM bytes per
launch keyed by the library's profiler tags (profiles/traffic.json, read by bench.py).

  python tools/ncu_summary.py <report.ncu-rep> <out.csv> [traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
TAGS = [("gemm_tct_kernel<5", "gemm_tc_res_gelu_dw3_gelu"), ("gemm_tct_kernel<4", "gemm_tc_glu_dw15_silu"),
        ("gemm_tc_kernel<256, 3>", "gemm_tc_layernorm"), ("gemm_tc_kernel<256, 0>", "gemm_tc_bias_act"),
        ("logmel_kernel", "logmel_stft_mel"), ("attn_tc_kernel", "attention_tc")]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    if len(sys.argv) > 3:
        kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        acc = {}
        for r in data:
            for pat, tag in TAGS:
                if pat in r[kn].replace("(int)", ""):
                    b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
                    acc.setdefault(tag, []).append(b)
                    break
        json.dump({"source": f"{out} (ncu --set full --clock-control none, one 64 x 30 s step): dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the launches of the step)",
                   "bytes_per_launch": {t: sum(v) / len(v) for t, v in acc.items()}}, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
