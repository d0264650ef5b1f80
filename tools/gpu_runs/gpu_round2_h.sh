#!/bin/bash
# 8-GPU box: multi-GPU tests, the single-process C++ driver at 1/2/4/8 devices (both policies), torchrun bench lines at
# N = 8 and 2, and the copy-only PCIe probe at N = 2/4/8 (the host-side ceiling of the e2e number).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_multi_gpu.py -q -x > gpurun_out/r02h_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r02h_multi_tests.log
: > gpurun_out/r02h_driver.log
for P in fast tolerant; do
  for G in 1 2 4 8; do
    ./rlshaders_b200/host/rls_driver --gpus $G --policy $P --reps 10 sweep 2>&1 | grep "^{" >> gpurun_out/r02h_driver.log
    ./rlshaders_b200/host/rls_driver --gpus $G --policy $P --reps 10 dielectric 26 2>&1 | grep "^{" >> gpurun_out/r02h_driver.log
  done
  ./rlshaders_b200/host/rls_driver --gpus 8 --policy $P --reps 5 disney 27 2>&1 | grep "^{" >> gpurun_out/r02h_driver.log
done
cut -c1-170 gpurun_out/r02h_driver.log
for N in 8 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/r02h_bench_n$N.err
done
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N tools/pcie_pattern_probe.py > gpurun_out/r02h_pcie_probe_n$N.txt 2>&1; grep '"pattern": "A"' gpurun_out/r02h_pcie_probe_n$N.txt | cut -c1-400
done
python - <<'PY'
import json
for N in (8, 2):
    d=json.loads(open(f'gpurun_out/r02h_bench_n{N}.json').read().strip().splitlines()[-1])
    print(N, 'headline', d['value']/1e9, d['roofline']['by_policy'], 'e2e', d['e2e']['value']/1e9, 'full', d['e2e']['full_frames']['value']/1e9)
    sw=d['other_workloads']['albedo_sweep_65536x4096']; print('  sweep', sw['samples_per_s']/1e9, sw['tolerant']['samples_per_s']/1e9)
    print('  cpp', json.dumps(d.get('cpp_driver'))[:600])
PY
