#!/bin/bash
# band base 1e-4 (+ GTR1 noise term): tolerance-policy tests, flag hunt over 24 x 2^24 samples per unit, bench line, PCIe pattern probe
mkdir -p gpurun_out
python -m pytest -q tests/test_tolerant_policy.py tests/test_full_size_parity.py -m gpu > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02g_tests.log
python tools/tol_flag_hunt.py 24 > gpurun_out/r02g_flag_hunt.log 2>&1; echo "hunt rc=$?"; grep -v "mismatches 0" gpurun_out/r02g_flag_hunt.log | tail -5
timeout 900 python bench.py --no-cpu --no-cpp-driver > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02g_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_n1.json').read())
e=d['e2e']; print('e2e compact', e['value']/1e9, e['pcie_gbs'], e['matches_device_path_on_decoded_frames'], e.get('non_finite_outputs'), '| full', e['full_frames']['value']/1e9, d['e2e_matches_device'])
print('headline', d['value']/1e9, d['roofline']['by_policy'], 'tolerant', d['tolerant']['value']/1e9, d['tolerant']['exact_rerun_fraction'])
for k,v in d['other_workloads'].items(): print(k, v['samples_per_s']/1e9, v.get('hbm_frac'), '| tol', v.get('tolerant',{}).get('samples_per_s',0)/1e9, v.get('tolerant',{}).get('hbm_frac'), v.get('tolerant',{}).get('exact_rerun_fraction'))
PY
python tools/pcie_pattern_probe.py > gpurun_out/r02g_pcie_probe_n1.txt 2>&1; cat gpurun_out/r02g_pcie_probe_n1.txt
