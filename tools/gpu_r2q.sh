#!/bin/bash
timeout 1500 python -m pytest tests -q -m gpu -x -k "block_per_codeword or strict_mode or edge_cases or golden" 2>&1 | tail -4
python tools/prof_exact.py 11 1024 16 32 831 1.0
POLAR_B200_EXACT_THREADS=256 POLAR_B200_EXACT_BPS=3 python tools/prof_exact.py 11 1024 16 32 296 1.0
python tools/prof_exact.py 11 1024 16 4 296 1.0
python tools/prof_exact.py 9 256 16 32 296 1.0
for c in c4 c3 c5; do echo -n "strict $c: "; timeout 300 python bench.py --mode strict --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2q.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['modes']['strict_flagged_per_step'])"; done
