#!/bin/bash
# usage: tools/prof2.sh <tag> <batch> [bench args]: ncu --set full of one decode launch at the given batch
tag=$1; b=$2; shift; shift
ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch $b "$@" > gpurun_out/ncu_full_$tag.log 2>&1
tail -1 gpurun_out/ncu_full_$tag.log
