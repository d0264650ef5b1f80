#!/bin/bash
# 8-GPU box, final kernels: the torchrun bench line at N = 8 and the single-process C++ driver at 8 devices
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err; echo "bench n8 rc=$?"
: > gpurun_out/r02p_driver.log
for P in fast tolerant; do
  for W in "sweep" "dielectric 26" "disney 27" "skin 27"; do ./rlshaders_b200/host/rls_driver --gpus 8 --policy $P --reps 10 $W 2>&1 | grep "^{" >> gpurun_out/r02p_driver.log; done
done
cut -c1-150 gpurun_out/r02p_driver.log
