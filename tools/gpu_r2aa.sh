#!/bin/bash
# round 2, session aa: A/B of the register subtree of sc_ssc.cuh out of line (code size / instruction fetch): default, the
# four-entry node out of line, the eight-entry node out of line
mkdir -p gpurun_out
export POLAR_B200_STRICT_TAU=1e-30
for rep in 1 2; do for d in lib lib_o1 lib_o2; do for b in 65536 262144; do
  echo -n "$d: "; POLAR_B200_LIB_DIR=$PWD/polar_b200/$d timeout 120 python tools/list_rate.py 11 1024 0 1 $b 1.5
done; done; done > gpurun_out/r02aa_outline_ab.txt 2>&1
for d in lib lib_o1 lib_o2; do echo -n "$d: "; POLAR_B200_LIB_DIR=$PWD/polar_b200/$d timeout 120 python tools/list_rate.py 9 256 0 1 262144 2.0; done >> gpurun_out/r02aa_outline_ab.txt 2>&1
cat gpurun_out/r02aa_outline_ab.txt
