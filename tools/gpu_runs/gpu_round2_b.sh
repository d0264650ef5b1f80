#!/bin/bash
mkdir -p gpurun_out
python tools/tol_flag_hunt.py 12 > gpurun_out/r02b_flag_hunt.log 2>&1; echo "hunt rc=$?"; tail -4 gpurun_out/r02b_flag_hunt.log
python -m pytest tests -q -m gpu > gpurun_out/r02b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/r02b_gpu_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/r02b_launches_tol.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --main-only --arith tolerant --e2e-steps 1 --e2e-samples 1048576 > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"
