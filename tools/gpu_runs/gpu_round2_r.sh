#!/bin/bash
# A/B: tolerance kernels at 128 threads x 9 CTAs per SM against 256 x 4 (build/librls_b200_b128.so = -DRLS_TOL_BLOCK=128 -DRLS_TOL_MIN_BLOCKS=9)
mkdir -p gpurun_out
for V in default b128; do
  LIBV=""; [ $V = b128 ] && LIBV=$PWD/build/librls_b200_b128.so
  RLS_B200_LIB=$LIBV timeout 600 python bench.py --no-cpu --no-cpp-driver --steps 20 --e2e-steps 1 --e2e-samples 4194304 > gpurun_out/r02r_bench_$V.json 2> gpurun_out/r02r_bench_$V.err
  RLS_B200_LIB=$LIBV timeout 600 python bench.py --no-cpu --no-cpp-driver --main-only --workload disney --steps 10 --e2e-steps 1 --e2e-samples 4194304 > gpurun_out/r02r_bench_disney_$V.json 2> gpurun_out/r02r_bench_disney_$V.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02r_bench_$V.json').read())
print('$V', 'dielectric tol', d['tolerant']['value']/1e9, {k: round(v.get('tolerant',{}).get('samples_per_s',0)/1e9,2) for k,v in d['other_workloads'].items()})
d=json.loads(open('gpurun_out/r02r_bench_disney_$V.json').read())
print('$V', 'disney tol', d['tolerant']['value']/1e9, d['tolerant']['roofline']['frac'])
PY
done
