#!/bin/bash
# round 2, GPU session B: fixed-point metrics + block-per-codeword double kernel: tests, calibration, bench
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_r2b.txt 2>&1; tail -15 gpurun_out/pytest_gpu_r2b.txt
timeout 900 python tools/margin_calib.py 1.0 gpurun_out/margin_calib_r2b.json > gpurun_out/margin_calib_r2b.log 2>&1; tail -c 9000 gpurun_out/margin_calib_r2b.log
for mode in fp32 strict; do for c in c4 c3 c2 c5; do
  echo -n "$mode $c: "; timeout 300 python bench.py --mode $mode --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2b.err | tee -a gpurun_out/bench_r2b_$mode.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['bler'], d['modes'])"
done; done
tail -5 gpurun_out/bench_r2b.err
echo -n "c2 variant 49: "; POLAR_B200_FAST_VARIANT=49 timeout 300 python bench.py --mode fp32 --config c2 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2b.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'])"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b_full.json 2>>gpurun_out/bench_r2b.err; cat gpurun_out/bench_r2b_full.json
