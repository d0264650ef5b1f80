#!/bin/bash
# Round-end evidence run (one GPU): smoke, default bench line, reference arm, ncu launch list,
# ncu --set full captures of the five fused kernels.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cut -c1-200 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_n1_reference.json 2> gpurun_out/${TAG}_bench_ref.err; cut -c1-200 gpurun_out/${TAG}_bench_n1_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ggx_dielectric -s 4 -c 1 -f -o gpurun_out/${TAG}_prof_dielectric python bench.py --steps 2 --warmup 3 --no-cpu --main-only > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_prof_dielectric.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_dielectric.raw.csv 2>/dev/null
for k in k_ggx_sample_eval_pdf k_disney_sample k_skin_profile k_albedo_sweep; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --e2e-samples 1048576 > /dev/null 2>&1
  # gpurun brings back at most 64 MiB: keep the raw page of these, the full report of the dielectric kernel only
  ncu -i gpurun_out/${TAG}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_$k.raw.csv 2>/dev/null
  rm -f gpurun_out/${TAG}_prof_$k.ncu-rep
done
ls -la gpurun_out | grep ${TAG}_
