#!/bin/bash
# dense input copy for the re-run of rlDisney / dielectric: tests, flag hunt, bench; A/B of the dielectric tol kernel's occupancy cap
mkdir -p gpurun_out
python -m pytest -q tests/test_tolerant_policy.py tests/test_full_size_parity.py -m gpu -x > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02k_tests.log
python tools/tol_flag_hunt.py 6 > gpurun_out/r02k_flag_hunt.log 2>&1; echo "hunt rc=$?"; grep -v "mismatches 0" gpurun_out/r02k_flag_hunt.log | tail -3
for V in default mb4; do
  LIBV=""; [ $V = mb4 ] && LIBV=$PWD/build/librls_b200_mb4.so
  RLS_B200_LIB=$LIBV timeout 600 python bench.py --no-cpu --no-cpp-driver --main-only --steps 20 --e2e-steps 1 --e2e-samples 4194304 > gpurun_out/r02k_bench_$V.json 2> gpurun_out/r02k_bench_$V.err; echo "bench $V rc=$?"; tail -2 gpurun_out/r02k_bench_$V.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_bench_$V.json').read())
print('$V', 'headline', d['value']/1e9, 'tolerant', d['tolerant']['value']/1e9, d['tolerant']['roofline']['frac'], d['tolerant']['exact_rerun_fraction'])
PY
done
timeout 600 python bench.py --no-cpu --no-cpp-driver --main-only --workload disney --steps 5 --e2e-steps 1 --e2e-samples 4194304 > gpurun_out/r02k_bench_disney.json 2> gpurun_out/r02k_bench_disney.err; echo "bench disney rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02k_bench_disney.json').read())
print('disney headline', d['value']/1e9, 'tolerant', d['tolerant']['value']/1e9, d['tolerant']['roofline']['frac'], d['tolerant']['exact_rerun_fraction'], d['tolerant']['flag_mismatches_vs_headline_policy_full_batch'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rerun|_tol" -c 40 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 2 --warmup 2 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 1048576 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/r02k_launches.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        m=re.search(r'(k_[A-Za-z0-9_]+)',r[hdr.index('Kernel Name')]); key=m.group(1) if m else '?'
        a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=float(r[hdr.index('Metric Value')])
for k,(c,t) in agg.items(): print(f"{k:36s} {c:3d} launches, {t/c/1e3:9.1f} us each")
PY
