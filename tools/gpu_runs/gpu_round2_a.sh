#!/bin/bash
# First GPU pass of round 2: tolerant-policy tests, full GPU suite, bench under both policies, launch list + ncu of the
# tolerance dielectric kernel.
mkdir -p gpurun_out
python -m pytest tests/test_tolerant_policy.py -x -q -m gpu -s > gpurun_out/r02a_tol_tests.log 2>&1; echo "tol tests rc=$?"
tail -3 gpurun_out/r02a_tol_tests.log
python -m pytest tests -x -q -m gpu > gpurun_out/r02a_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/r02a_gpu_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02a_bench_fast.json 2> gpurun_out/r02a_bench_fast.err; echo "bench fast rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu --arith tolerant > gpurun_out/r02a_bench_tol.json 2> gpurun_out/r02a_bench_tol.err; echo "bench tol rc=$?"
cat gpurun_out/r02a_bench_fast.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fast', d['value']/1e9, d['roofline']['frac'], {k:v['samples_per_s']/1e9 for k,v in d.get('other_workloads',{}).items()}, d['e2e']['value']/1e9)"
cat gpurun_out/r02a_bench_tol.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tol', d['value']/1e9, d['roofline']['frac'], {k:v['samples_per_s']/1e9 for k,v in d.get('other_workloads',{}).items()}, d['e2e']['value']/1e9, d['arith'])"
ncu --set full --clock-control none --import-source on -k regex:k_ggx_dielectric_tol -s 2 -c 1 -o gpurun_out/r02a_prof_dielectric_tol \
    python bench.py --steps 2 --warmup 3 --no-cpu --main-only --arith tolerant --e2e-steps 1 --e2e-samples 1048576 > gpurun_out/r02a_ncu.log 2>&1; echo "ncu rc=$?"
