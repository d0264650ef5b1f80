#!/bin/bash
# 2-GPU box: the single-process multi-GPU driver, its tests, and a torchrun bench line at N = 2
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r02d_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -5 gpurun_out/r02d_multi_tests.log
for P in fast tolerant; do
  for G in 1 2; do
    ./rlshaders_b200/host/rls_driver --gpus $G --policy $P --reps 10 sweep 2>&1 | tee -a gpurun_out/r02d_driver.log
    ./rlshaders_b200/host/rls_driver --gpus $G --policy $P --reps 10 dielectric 26 2>&1 | tee -a gpurun_out/r02d_driver.log
  done
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02d_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_n2.json').read())
print('n2 headline', d['value']/1e9, d['roofline']['by_policy'], 'e2e', d['e2e']['value']/1e9)
print('sweep', d['other_workloads']['albedo_sweep_65536x4096']['samples_per_s']/1e9, d['other_workloads']['albedo_sweep_65536x4096'].get('tolerant',{}).get('samples_per_s',0)/1e9)
print('cpp', json.dumps(d.get('cpp_driver'))[:1500])
PY
