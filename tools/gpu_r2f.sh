#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "block_per_codeword or strict_mode or edge_cases or fused_sweep or cpp_multi" 2>&1 | tail -5
for bps in 2 3 4; do echo -n "bps=$bps: "; POLAR_B200_EXACT_BPS=$bps python tools/prof_exact.py 11 1024 16 32 888 1.0; done
python tools/prof_exact.py 11 1024 0 1 296 1.0
python tools/prof_exact.py 11 1024 16 4 296 1.0
for c in c4 c3 c5 c2 c1; do echo -n "strict $c: "; timeout 300 python bench.py --mode strict --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2f.err | tee -a gpurun_out/bench_r2f_strict.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['e2e_sweep']['value']), d['bler'], d['modes'])"; done
tail -3 gpurun_out/bench_r2f.err
echo -n "wc host buffer c4: "; timeout 300 python bench.py --host-alloc wc --steps 3 --warmup 3 --no-cpu --e2e-steps 3 2>>gpurun_out/bench_r2f.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['e2e'])"
echo -n "minsum: "; for c in c4 c3 c2 c5; do timeout 300 python bench.py --mode minsum --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2f.err | tee -a gpurun_out/bench_r2f_minsum.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], d['bler_per_ebno'])"; done
tail -3 gpurun_out/bench_r2f.err
