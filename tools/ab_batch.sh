#!/bin/bash
# usage: tools/ab_batch.sh "<batches>" [bench args]: in-tree build vs polar_b200/lib_base at several batch sizes
for b in $1; do for d in lib lib_base; do
  echo -n "batch $b $d: "
  POLAR_B200_LIB_DIR=$PWD/polar_b200/$d python bench.py --batch $b --steps 5 --warmup 3 --no-cpu --e2e-steps 1 "${@:2}" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['clocks'])"
done; done
