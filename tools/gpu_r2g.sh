#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "strict_mode or edge_cases or fused_sweep or golden or reference_main_prints" 2>&1 | tail -5
timeout 1500 python tools/flip_margins.py 1.0 gpurun_out/flip_margins_r2g.json 2>&1 | tail -8 | cut -c1-1500
for c in c4 c5; do echo -n "strict $c: "; timeout 300 python bench.py --mode strict --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2g.err | tee -a gpurun_out/bench_r2g_strict.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['e2e_sweep']['value']), d['bler'], d['modes'])"; done
tail -3 gpurun_out/bench_r2g.err
