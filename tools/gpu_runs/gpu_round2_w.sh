#!/bin/bash
# last call of the round: the L.h band (rlDisney tolerance kernel changed) -- tolerance tests, device hunts, ncu capture of that kernel, final bench lines
mkdir -p gpurun_out
timeout 120 python -m pytest -q tests/test_tolerant_policy.py -m gpu > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02w_tests.log
timeout 60 python tools/tol_flag_hunt.py 6 > gpurun_out/r02w_flag_hunt.log 2>&1; echo "hunt rc=$?"; grep -vc "mismatches 0" gpurun_out/r02w_flag_hunt.log
timeout 60 python tests/hunts/gpu_stress_parity.py 16 > gpurun_out/r02w_gpu_stress.log 2>&1; echo "stress rc=$?"; tail -1 gpurun_out/r02w_gpu_stress.log
SMALL="--steps 2 --warmup 3 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 1048576"
timeout 200 ncu --set full --clock-control none -k regex:k_disney_sample_eval_pdf_tol -s 3 -c 1 -f -o gpurun_out/r02_prof_disney_tolerant python bench.py $SMALL --workload disney --main-only --no-other-policy --arith tolerant > /dev/null 2>&1
ncu -i gpurun_out/r02_prof_disney_tolerant.ncu-rep --page raw --csv > gpurun_out/r02_prof_disney_tolerant.raw.csv 2>/dev/null; rm -f gpurun_out/r02_prof_disney_tolerant.ncu-rep
timeout 200 python bench.py --workload disney --main-only --steps 5 --warmup 3 > gpurun_out/r02_bench_disney_n1.json 2> gpurun_out/r02_bench_disney_n1.err; echo "disney rc=$?"
timeout 200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
