#!/bin/bash
# group-of-8 re-run + the two-samples-per-thread skin tolerance kernel: tests, flag hunt, bench
mkdir -p gpurun_out
python -m pytest -q tests/test_tolerant_policy.py tests/test_full_size_parity.py tests/test_gpu_parity.py -m gpu -x > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02i_tests.log
python tools/tol_flag_hunt.py 6 > gpurun_out/r02i_flag_hunt.log 2>&1; echo "hunt rc=$?"; grep -v "mismatches 0" gpurun_out/r02i_flag_hunt.log | tail -3
timeout 900 python bench.py --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 4194304 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02i_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench_n1.json').read())
print('headline', d['value']/1e9, d['roofline']['by_policy'], 'tolerant', d['tolerant']['value']/1e9, d['tolerant']['exact_rerun_fraction'])
for k,v in d['other_workloads'].items(): print(k, v['samples_per_s']/1e9, v.get('hbm_frac'), '| tol', v.get('tolerant',{}).get('samples_per_s',0)/1e9, v.get('tolerant',{}).get('hbm_frac'), v.get('tolerant',{}).get('exact_rerun_fraction'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rerun|_tol" -c 60 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 1048576 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/r02i_launches.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        m=re.search(r'(k_[A-Za-z0-9_]+)',r[hdr.index('Kernel Name')]); key=m.group(1) if m else '?'
        a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=float(r[hdr.index('Metric Value')])
for k,(c,t) in agg.items(): print(f"{k:36s} {c:3d} launches, {t/c/1e3:9.1f} us each")
PY
