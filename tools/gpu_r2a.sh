#!/bin/bash
# round 2, GPU session A: tests, margin calibration, bench in both modes, A/B of the round-1 experiments and the on-chip variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
( time timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_r2a.txt 2>&1; tail -8 gpurun_out/pytest_gpu_r2a.txt
timeout 900 python tools/margin_calib.py 1.0 gpurun_out/margin_calib_r2a.json > gpurun_out/margin_calib_r2a.log 2>&1; tail -c 6000 gpurun_out/margin_calib_r2a.log
for mode in fp32 strict; do for c in c4 c3 c2 c5; do
  echo -n "$mode $c: "; POLAR_B200_MODE=$mode timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2a.err | tee -a gpurun_out/bench_r2a_$mode.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['bler'])"
done; done
# on-chip variants (layer 3 in tensor memory, 8 warps/SM) against the defaults
for v in "c4 51" "c3 52" "c2 53"; do set -- $v
  echo -n "variant $2 $1: "; POLAR_B200_MODE=fp32 POLAR_B200_FAST_VARIANT=$2 timeout 300 python bench.py --config $1 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2a.err | tee -a gpurun_out/bench_r2a_onchip.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], d['roofline']['kernel_kind'])"
done
for v in 49 50; do echo -n "variant $v c2: "; POLAR_B200_MODE=fp32 POLAR_B200_FAST_VARIANT=$v timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2a.err | tee -a gpurun_out/bench_r2a_onchip.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'], d['roofline']['kernel_kind'])"; done
# round-1 macro experiments
for d in lib_tm2 lib_top2 lib_ps2; do [ -d polar_b200/$d ] && POLAR_B200_MODE=fp32 tools/ab_dirs.sh r2a_exp "lib $d" "c4" 1; done
ls gpurun_out | tail -30
