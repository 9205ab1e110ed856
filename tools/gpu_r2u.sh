#!/bin/bash
# round 2, session u: sc_ssc.cuh with the per-round barrier: warps per SM x batch x barrier on/off
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -x -k "pruned_tree" 2>&1 | tail -5 ) > gpurun_out/r02u_pytest_ssc.txt 2>&1
tail -2 gpurun_out/r02u_pytest_ssc.txt
{
export POLAR_B200_STRICT_TAU=1e-30
for s in 1 0; do for w in 10 8; do for b in 65536 262144; do
  echo -n "sync $s warps $w: "; POLAR_B200_SSC_SYNC=$s POLAR_B200_SSC_WARPS=$w timeout 120 python tools/list_rate.py 11 1024 0 1 $b 1.5
done; done; done
timeout 120 python tools/list_rate.py 9 256 0 1 4096 2.0
timeout 120 python tools/list_rate.py 9 256 0 1 262144 2.0
timeout 120 python tools/list_rate.py 12 2048 0 1 32768 2.0
timeout 120 python tools/list_rate.py 10 512 0 1 131072 2.0
timeout 120 python tools/list_rate.py 8 128 0 1 262144 2.0
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 8 128 0 1 262144 2.0
} > gpurun_out/r02u_rates.txt 2>&1
cat gpurun_out/r02u_rates.txt
