#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -s -k "close_decisions or strict_mode or edge_cases or raw_c_abi" 2>&1 | grep -v "^$" | tail -14
for v in 1 0; do echo -n "strict c4 verify=$v: "; POLAR_B200_VERIFY=$v timeout 300 python bench.py --mode strict --config c4 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2j.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['modes'])"; done
echo -n "strict c5: "; timeout 300 python bench.py --mode strict --config c5 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2j.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['modes'])"
tail -3 gpurun_out/bench_r2j.err
timeout 1500 python tools/flip_margins.py 1.0 gpurun_out/flip_margins_r2j.json 2>&1 | tail -6 | cut -c1-700
