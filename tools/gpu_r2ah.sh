#!/bin/bash
# round 2, session ah: BLER curve again with the final tree (list 1 now through the pruned-tree kernel and its fused count)
mkdir -p gpurun_out
timeout 420 python tools/bler_curve.py 262144 1024 > gpurun_out/r02ah_bler_curve.json 2> gpurun_out/r02ah_bler_curve.err
tail -c 600 gpurun_out/r02ah_bler_curve.json; tail -2 gpurun_out/r02ah_bler_curve.err
