#!/bin/bash
mkdir -p gpurun_out
for mode in fp32 strict; do for c in c4 c3 c2 c5; do
  echo -n "$mode $c: "; timeout 300 python bench.py --mode $mode --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2c.err | tee -a gpurun_out/bench_r2c_$mode.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['bler'], d['modes'])"
done; done
tail -3 gpurun_out/bench_r2c.err
python tools/prof_exact.py
python tools/prof_exact.py 11 1024 0 1 296 1.0
python tools/prof_exact.py 9 256 16 32 296 1.0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scl_exact -s 2 -c 1 -f -o gpurun_out/prof_exact_r2c python tools/prof_exact.py > gpurun_out/ncu_exact_r2c.log 2>&1; tail -2 gpurun_out/ncu_exact_r2c.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scl_fast -s 3 -c 1 -f -o gpurun_out/prof_c4_r2c python bench.py --mode fp32 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 > gpurun_out/ncu_c4_r2c.log 2>&1; tail -2 gpurun_out/ncu_c4_r2c.log
ls -la gpurun_out | tail -8
