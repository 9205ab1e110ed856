#!/bin/bash
# usage: tools/prof_cfg.sh <config> <tag>: ncu --set full capture of one decode launch of bench config <config> (batch 16384)
cfg=$1; tag=$2
ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o gpurun_out/prof_${cfg}_$tag python bench.py --config $cfg --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 > gpurun_out/ncu_full_${cfg}_$tag.log 2>&1
tail -1 gpurun_out/ncu_full_${cfg}_$tag.log
