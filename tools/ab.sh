#!/bin/bash
# usage: tools/ab.sh <tag> [configs...]: bench the in-tree build against polar_b200/lib_base (an older build of the same
# libraries) on the same box; one line per (build, config)
tag=$1; shift; cfgs=${@:-c4}
mkdir -p gpurun_out
for c in $cfgs; do for d in lib lib_base lib; do
  echo -n "$c $d: "
  POLAR_B200_LIB_DIR=$PWD/polar_b200/$d python bench.py --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/ab_$tag.err | tee -a gpurun_out/ab_$tag.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], d['clocks']['sm_mhz'])"
done; done
