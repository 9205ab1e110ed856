#!/bin/bash
# round 2, session aj: warps per SM chosen to fill the rounds evenly (small / odd batches): rates and the pruned-tree tests
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -k "pruned_tree" 2>&1 | tail -3 ) | tee gpurun_out/r02aj_pytest.txt
{
export POLAR_B200_STRICT_TAU=1e-30
timeout 120 python tools/list_rate.py 9 256 0 1 4096 2.0
POLAR_B200_SSC_WARPS=16 timeout 120 python tools/list_rate.py 9 256 0 1 4096 2.0
timeout 120 python tools/list_rate.py 11 1024 0 1 4096 2.0
POLAR_B200_SSC_WARPS=10 timeout 120 python tools/list_rate.py 11 1024 0 1 4096 2.0
timeout 120 python tools/list_rate.py 11 1024 0 1 49152 2.0
POLAR_B200_SSC_WARPS=10 timeout 120 python tools/list_rate.py 11 1024 0 1 49152 2.0
timeout 120 python tools/list_rate.py 11 1024 0 1 65536 2.0
} 2>&1 | tee gpurun_out/r02aj_rates.txt
unset POLAR_B200_STRICT_TAU
python bench.py --config c1 --steps 10 --warmup 3 --cpu-seconds 4 > gpurun_out/r02aj_bench_c1.json 2>/dev/null; cut -c1-160 gpurun_out/r02aj_bench_c1.json
