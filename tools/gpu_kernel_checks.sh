#!/bin/bash
# Runs each group of kernel parity tests in its own process (a device trap in one kernel must not hide the others)
# and leaves the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for grp in linear attention layernorm timestep embed rounding posterior q_sample xstart philox; do
  timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$grp" > gpurun_out/kt_$grp.log 2>&1
  echo "== $grp rc=$? $(tail -1 gpurun_out/kt_$grp.log)"
done
