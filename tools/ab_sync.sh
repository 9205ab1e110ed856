#!/bin/bash
# usage: tools/ab_sync.sh "<batches>" "<configs>": POLAR_B200_SYNC=1 (one block per SM, per-round barrier per sub-partition) vs 0
for c in $2; do for b in $1; do for s in 1 0; do
  echo -n "$c batch $b sync $s: "
  POLAR_B200_SYNC=$s python bench.py --config $c --batch $b --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], round(d['e2e']['value']), d['e2e']['matches_device_arm'])"
done; done; done
