#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for d in lib lib_out lib_roll lib_outroll; do
  echo -n "$d c4 fp32: "; POLAR_B200_LIB_DIR=$PWD/polar_b200/$d timeout 300 python bench.py --mode fp32 --config c4 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2m.err | tee -a gpurun_out/ab_r2m.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'])"
done; done
for d in lib lib_out; do echo -n "$d c5 fp32: "; POLAR_B200_LIB_DIR=$PWD/polar_b200/$d timeout 300 python bench.py --mode fp32 --config c5 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2m.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['bler'])"; done
POLAR_B200_LIB_DIR=$PWD/polar_b200/lib_out timeout 600 python -m pytest tests -q -m gpu -x -k "strict_mode or matches_oracle_on_awgn or edge_cases" 2>&1 | tail -3
tail -3 gpurun_out/bench_r2m.err
