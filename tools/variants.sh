#!/bin/bash
# usage: tools/variants.sh "<variant list>" [bench args] ; prints cw/s per variant (and with L2 persistence off)
vs=$1; shift
for v in $vs; do for p in 1 0; do
  echo -n "variant $v persist $p: "
  POLAR_B200_FAST_VARIANT=$v POLAR_B200_L2_PERSIST=$p python bench.py --steps 3 --no-cpu --e2e-steps 1 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['bler'])"
done; done
