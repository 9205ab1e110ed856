#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_r2d.txt 2>&1; tail -8 gpurun_out/pytest_gpu_r2d.txt
for big in 32 64 128; do for bps in 1 2 3; do echo -n "big=$big bps=$bps: "; POLAR_B200_EXACT_BIG=$big POLAR_B200_EXACT_BPS=$bps python tools/prof_exact.py 11 1024 16 32 888 1.0; done; done
python tools/prof_exact.py 11 1024 0 1 296 1.0
python tools/prof_exact.py 11 1024 16 4 296 1.0
python tools/prof_exact.py 9 256 16 32 296 1.0
for bps in 1 2 3; do for c in c4 c5; do
  echo -n "strict bps=$bps $c: "; POLAR_B200_EXACT_BPS=$bps timeout 300 python bench.py --mode strict --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2d.err | tee -a gpurun_out/bench_r2d_strict.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['bler'], d['modes'])"
done; done
for c in c3 c2; do echo -n "strict $c: "; timeout 300 python bench.py --mode strict --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2d.err | tee -a gpurun_out/bench_r2d_strict.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['bler'], d['modes'])"; done
tail -3 gpurun_out/bench_r2d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scl_exact -s 2 -c 1 -f -o gpurun_out/prof_exact_r2d python tools/prof_exact.py > gpurun_out/ncu_exact_r2d.log 2>&1; tail -2 gpurun_out/ncu_exact_r2d.log
