#!/bin/bash
# usage: tools/prof.sh <tag> [extra bench args]  -> gpurun_out/prof_<tag>.ncu-rep + bench line
tag=$1; shift
python bench.py --steps 3 --no-cpu --e2e-steps 1 "$@" | tee gpurun_out/bench_$tag.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['bler'], d['clocks'])"
ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 "$@" > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log
