#!/bin/bash
# GPU suite with the full-size parity tests + ncu of the bit-exact re-run kernels of the tolerance policy
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02e_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/r02e_tests.log
SMALL="--steps 2 --warmup 3 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 1048576 --main-only --no-other-policy --arith tolerant"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_ggx_dielectric_rerun -s 3 -c 1 -f -o gpurun_out/r02e_prof_dielectric_rerun \
    python bench.py $SMALL > /dev/null 2>&1
ncu -i gpurun_out/r02e_prof_dielectric_rerun.ncu-rep --page raw --csv > gpurun_out/r02e_prof_dielectric_rerun.raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_prof_dielectric_rerun.ncu-rep --page details > gpurun_out/r02e_prof_dielectric_rerun.details.txt 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:k_disney_sample_eval_pdf_rerun -s 3 -c 1 -f -o gpurun_out/r02e_prof_disney_rerun \
    python bench.py $SMALL --workload disney > /dev/null 2>&1
ncu -i gpurun_out/r02e_prof_disney_rerun.ncu-rep --page details > gpurun_out/r02e_prof_disney_rerun.details.txt 2>/dev/null
rm -f gpurun_out/r02e_prof_disney_rerun.ncu-rep
ls -la gpurun_out | grep r02e_
