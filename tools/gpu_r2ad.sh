#!/bin/bash
# round 2, session ad: FP32 mode on the pruned-tree kernel (ties handed to the leaf-by-leaf fp32 kernel): whole GPU suite, bench c1 / c2
tag=${1:-r02ad}
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -8 gpurun_out/${tag}_pytest_gpu.txt
for c in c1 c2; do python bench.py --config $c --steps 10 --warmup 3 --cpu-seconds 4 >> gpurun_out/${tag}_bench_c1_c2.json 2>> gpurun_out/${tag}_bench.err; done
python - <<'PY'
import json
for l in open("gpurun_out/r02ad_bench_c1_c2.json"):
    d = json.loads(l); print(d["config"]["workload"][:40], round(d["value"]), d["modes"], d["roofline"]["kernel"], d["parity"])
PY
tail -3 gpurun_out/${tag}_bench.err
