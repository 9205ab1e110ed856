#!/bin/bash
# multi-GPU session: usage tools/gpu_r2_multi.sh <ngpus> <tag>
n=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/${tag}_topo_${n}gpu.txt
timeout 600 python -m pytest tests -q -m gpu -x -k "cpp_multi or fused_sweep" 2>&1 | tail -3
for host in pinned wc; do
  echo "== bench --gpus $n host buffer $host"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 5 --warmup 3 --host-alloc $host 2>gpurun_out/${tag}_bench_${n}gpu_$host.err | grep '^{' | tee gpurun_out/${tag}_bench_${n}gpu_$host.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e']['h2d_copy_gbs_per_gpu_all_ranks_at_once'], 'sweep', round(d['e2e_sweep']['value']), d['e2e_sweep']['collective'], d['modes'])"
  tail -2 gpurun_out/${tag}_bench_${n}gpu_$host.err
done
echo "== fp32 mode"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --steps 5 --warmup 3 --mode fp32 2>>gpurun_out/${tag}_bench_${n}gpu_fp32.err | grep '^{' | tee gpurun_out/${tag}_bench_${n}gpu_fp32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sweep', round(d['e2e_sweep']['value']))"
echo "== C++ sweep over all visible devices (PolarCode::bler_sweep, ncclCommInitAll)"
python tools/bler_curve.py 262144 1024 > gpurun_out/${tag}_bler_curve_${n}gpu.json 2> gpurun_out/${tag}_bler_curve_${n}gpu.err; tail -c 400 gpurun_out/${tag}_bler_curve_${n}gpu.json; tail -3 gpurun_out/${tag}_bler_curve_${n}gpu.err
grep -o '"N2048_K1024_crc[0-9]*_seconds": [0-9.]*' gpurun_out/${tag}_bler_curve_${n}gpu.json
