#!/bin/bash
# usage: tools/ab_dirs.sh <tag> "<lib dirs under polar_b200/>" "<configs>" [reps]: bench several builds of the libraries on the same box
tag=$1; dirs=$2; cfgs=${3:-c4}; reps=${4:-2}
mkdir -p gpurun_out
for c in $cfgs; do for rep in $(seq $reps); do for d in $dirs; do
  echo -n "$c $d: "
  POLAR_B200_LIB_DIR=$PWD/polar_b200/$d python bench.py --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/ab_$tag.err | tee -a gpurun_out/ab_$tag.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], d['clocks']['sm_mhz'])"
done; done; done
