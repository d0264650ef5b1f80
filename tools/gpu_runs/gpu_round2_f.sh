#!/bin/bash
# compact-frame host forms: test + bench line; then the sanitizer passes
mkdir -p gpurun_out
python -m pytest -q tests/test_gpu_parity.py -k "compact_frame or host_buffer" > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02f_tests.log
timeout 900 python bench.py --no-cpu --main-only --no-other-policy --no-cpp-driver > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_n1.json').read())
e=d['e2e']; print('e2e compact', e['value']/1e9, e['pcie_gbs'], e['matches_device_path_on_decoded_frames'], '| full', e['full_frames']['value']/1e9, e['full_frames']['pcie_gbs'], d['e2e_matches_device'])
PY
bash tools/sanitize.sh gpu r02
ls -la gpurun_out | grep "r02f_\|sanitizer"
