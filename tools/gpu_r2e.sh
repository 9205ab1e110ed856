#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) > gpurun_out/pytest_gpu_r2e.txt 2>&1; tail -8 gpurun_out/pytest_gpu_r2e.txt
for nt in 64 128 256; do for bps in 3 4; do echo -n "threads=$nt bps=$bps: "; POLAR_B200_EXACT_THREADS=$nt POLAR_B200_EXACT_BPS=$bps python tools/prof_exact.py 11 1024 16 32 888 1.0; done; done
POLAR_B200_EXACT_THREADS=128 python tools/prof_exact.py 11 1024 0 1 296 1.0
timeout 1500 python tools/flip_margins.py 1.0 gpurun_out/flip_margins_r2e.json 2>&1 | tail -8 | cut -c1-1500
