#!/bin/bash
# usage: [ENV=..] tools/line.sh <label> [bench args]: one short result line of a bench run
label=$1; shift
echo -n "$label: "; python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], 'e2e', round(d['e2e']['value']), d['e2e']['matches_device_arm'], d['e2e']['pipelined_chunks'], 'kind', d['roofline']['kernel_kind'], d['clocks']['sm_mhz'])"
