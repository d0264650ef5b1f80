#!/bin/bash
# Round evidence run (one GPU): GPU test-suite, smoke, default bench line, reference arm (both workloads), Disney bench
# line, ncu launch list, `ncu --set full` captures of every fused kernel under BOTH arithmetic policies.
# Outputs under gpurun_out/<tag>_*; afterwards, in the build container:
#   python tools/make_traffic.py <tag> gpurun_out/<tag>_prof_*.raw.csv      -> profiles/traffic.json (stamped with SASS hashes)
#   cp gpurun_out/<tag>_* profiles/ (what is to be judged); python tools/make_ncu_summary.py <tag>  -> profiles/<tag>_ncu_summary.md
TAG=${1:-r02}
STAGE=${2:-all}          # all | tests | bench | ncu
mkdir -p gpurun_out
if [ "$STAGE" = all ] || [ "$STAGE" = tests ]; then
  python -m pytest tests -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
  python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
fi
if [ "$STAGE" = all ] || [ "$STAGE" = bench ]; then
  timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_n1.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_n1_reference.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-160 gpurun_out/${TAG}_bench_n1_reference.json
  timeout 900 python bench.py --workload disney --main-only --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_disney_n1.json 2> gpurun_out/${TAG}_bench_disney_n1.err; cut -c1-160 gpurun_out/${TAG}_bench_disney_n1.json
  timeout 600 python bench.py --impl reference --workload disney --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_disney_n1_reference.json 2>> gpurun_out/${TAG}_bench_ref.err; cut -c1-160 gpurun_out/${TAG}_bench_disney_n1_reference.json
fi
if [ "$STAGE" = all ] || [ "$STAGE" = ncu ]; then
  SMALL="--steps 2 --warmup 3 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 1048576"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu --no-cpp-driver --e2e-steps 1 --e2e-samples 4194304 > /dev/null 2>&1
  # the two headline kernels with source (SASS <-> CUDA lines), one policy per process so that the regex is unambiguous
  for P in fast tolerant; do
    K=k_ggx_dielectric; [ $P = tolerant ] && K=k_ggx_dielectric_tol
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_dielectric_$P \
        python bench.py $SMALL --main-only --no-other-policy --arith $P > /dev/null 2>&1
    ncu -i gpurun_out/${TAG}_prof_dielectric_$P.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_dielectric_$P.raw.csv 2>/dev/null
    K=k_disney_sample_eval_pdf; [ $P = tolerant ] && K=k_disney_sample_eval_pdf_tol
    timeout 400 ncu --set full --clock-control none -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_disney_$P \
        python bench.py $SMALL --workload disney --main-only --no-other-policy --arith $P > /dev/null 2>&1
    ncu -i gpurun_out/${TAG}_prof_disney_$P.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_disney_$P.raw.csv 2>/dev/null
    rm -f gpurun_out/${TAG}_prof_disney_$P.ncu-rep
  done
  # the other configs run under both policies inside the default line: anchored names, first timed launch of each
  for K in k_ggx_sample_eval_pdf k_ggx_sample_eval_pdf_tol k_skin_profile k_skin_profile_tol_x2 k_albedo_sweep k_albedo_sweep_tol; do
    timeout 400 ncu --set full --clock-control none -k "regex:^${K}\$" -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$K \
        python bench.py $SMALL > /dev/null 2>&1
    ncu -i gpurun_out/${TAG}_prof_$K.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_$K.raw.csv 2>/dev/null
    rm -f gpurun_out/${TAG}_prof_$K.ncu-rep      # gpurun brings back at most 64 MiB: raw pages only
  done
fi
ls -la gpurun_out | grep ${TAG}_
