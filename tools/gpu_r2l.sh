#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for d in lib lib_nov; do for mode in fp32 strict; do
  echo -n "$d $mode: "; POLAR_B200_LIB_DIR=$PWD/polar_b200/$d timeout 300 python bench.py --mode $mode --config c4 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>>gpurun_out/bench_r2l.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['modes']['strict_flagged_per_step'])"
done; done; done
tail -3 gpurun_out/bench_r2l.err
