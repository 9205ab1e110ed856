#!/bin/bash
# round 2, session ac (2 GPUs): the multi-GPU sweep tests with the pruned-tree SC kernel in place, and bench.py --config c2 on 2 GPUs
tag=${1:-r02ac}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -k "cpp_multi or fused_sweep or bler_sweep" 2>&1 | tail -4 ) | tee gpurun_out/${tag}_pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --config c2 --steps 10 --warmup 3 --no-cpu 2>gpurun_out/${tag}_bench_c2_2gpu.err | grep '^{' | tee gpurun_out/${tag}_bench_c2_2gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sweep', round(d['e2e_sweep']['value']), d['e2e_sweep']['collective'], d['modes'])"
tail -2 gpurun_out/${tag}_bench_c2_2gpu.err
