#!/bin/bash
# round 2, session s: the broader tau campaign (other codes / block lengths / list sizes than BASELINE.json's)
mkdir -p gpurun_out
timeout 800 python tools/flip_margins.py 8 gpurun_out/r02s_flip_more.json more > gpurun_out/r02s_flip_more.txt 2>&1
tail -3 gpurun_out/r02s_flip_more.txt | cut -c1-300
