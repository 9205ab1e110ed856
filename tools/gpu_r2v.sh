#!/bin/bash
# round 2, session v: whole GPU suite with the pruned-tree SC kernel in place; prefetch stage A/B; threshold campaign for that kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r02v_pytest_gpu.txt 2>&1
tail -6 gpurun_out/r02v_pytest_gpu.txt
{
export POLAR_B200_STRICT_TAU=1e-30
for st in 8 4; do for b in 65536 262144; do
  echo -n "stage $st: "; POLAR_B200_SSC_STAGE=$st timeout 120 python tools/list_rate.py 11 1024 0 1 $b 1.5
done; done
timeout 120 python tools/list_rate.py 12 2048 0 1 32768 2.0
} > gpurun_out/r02v_rates.txt 2>&1
cat gpurun_out/r02v_rates.txt
timeout 600 python tools/flip_margins.py 0.5 gpurun_out/r02v_flip_sc.json sc > gpurun_out/r02v_flip_sc.txt 2>&1
cut -c1-330 gpurun_out/r02v_flip_sc.txt
