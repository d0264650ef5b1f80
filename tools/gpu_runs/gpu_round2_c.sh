#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02c_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/r02c_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read())
print('headline', d['value']/1e9, d['roofline']['frac'], d['roofline']['by_policy'], 'e2e', d['e2e']['value']/1e9, 'cpu', d.get('cpu_baseline',{}).get('value',0)/1e6, d.get('cpu_baseline',{}).get('cores'))
print('tolerant', d['tolerant']['value']/1e9, d['tolerant']['exact_rerun_fraction'], d['tolerant'].get('flag_mismatches_vs_headline_policy_full_batch'), d['tolerant'].get('parity_vs_oracle',{}).get('flags'))
for k,v in d['other_workloads'].items(): print(k, v['samples_per_s']/1e9, v.get('hbm_frac'), '| tol', v.get('tolerant',{}).get('samples_per_s',0)/1e9, v.get('tolerant',{}).get('hbm_frac'), v.get('tolerant',{}).get('exact_rerun_fraction'))
PY
python bench.py --steps 5 --warmup 3 --workload disney --main-only > gpurun_out/r02c_bench_disney.json 2> gpurun_out/r02c_bench_disney.err; echo "bench disney rc=$?"; tail -3 gpurun_out/r02c_bench_disney.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench_disney.json').read())
print('disney headline', d['value']/1e9, d['roofline']['by_policy'], 'e2e', d['e2e']['value']/1e9, 'cpu', d.get('cpu_baseline',{}).get('value',0)/1e6)
print('tolerant', d['tolerant']['value']/1e9, d['tolerant']['exact_rerun_fraction'], d['tolerant'].get('flag_mismatches_vs_headline_policy_full_batch'), d['tolerant'].get('parity_vs_oracle',{}).get('flags'))
PY
python tools/tol_flag_hunt.py 24 > gpurun_out/r02c_flag_hunt.log 2>&1; echo "hunt rc=$?"; grep -v "mismatches 0" gpurun_out/r02c_flag_hunt.log | tail -8
