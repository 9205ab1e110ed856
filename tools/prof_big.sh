#!/bin/bash
# usage: tools/prof_big.sh <tag> <batch> [lib dir]: warp-state / scheduler / instruction sections of one decode launch at a
# large batch (the full set would replay a long kernel ~40 times)
tag=$1; b=$2; d=${3:-lib}
POLAR_B200_LIB_DIR=$PWD/polar_b200/$d ncu --section WarpStateStats --section SchedulerStats --section InstructionStats --section SpeedOfLight --section MemoryWorkloadAnalysis --section SourceCounters --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o gpurun_out/prof_${tag} python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch $b > gpurun_out/ncu_${tag}.log 2>&1
tail -1 gpurun_out/ncu_${tag}.log
