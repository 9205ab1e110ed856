#!/bin/bash
# round 2, session s: the whole GPU suite again (no -x) and the broader tau campaign (other codes / list sizes)
mkdir -p gpurun_out
FLIP_SKIP=3 timeout 700 python tools/flip_margins.py 0.5 gpurun_out/r02s_flip_more2.json more > gpurun_out/r02s_flip_more2.txt 2>&1
tail -3 gpurun_out/r02s_flip_more2.txt | cut -c1-400
