#!/bin/bash
# A/B: scalar vs lane-paired dielectric kernel (RLS_PAIRED=0/1): bench line + a light ncu pass
mkdir -p gpurun_out
TAG=${1:-ab}
for P in 0 1; do
  RLS_PAIRED=$P timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --main-only --e2e-steps 1 --e2e-samples 1048576 > gpurun_out/${TAG}_bench_p$P.json 2> gpurun_out/${TAG}_bench_p$P.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_p$P.json"))
print("paired=$P", round(d["value"]/1e9,3), "G/s", d["ms_per_step"], "ms", d.get("arith"))
PY
  RLS_PAIRED=$P timeout 200 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,launch__registers_per_thread,smsp__warps_eligible.avg.per_cycle_active,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_ggx_dielectric -s 4 -c 1 --csv --log-file gpurun_out/${TAG}_ncu_p$P.csv python bench.py --steps 2 --warmup 3 --no-cpu --main-only --e2e-steps 1 --e2e-samples 1048576 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_ncu_p$P.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print("  ", r[h.index("Metric Name")], r[h.index("Metric Value")])
PY
done
