#!/bin/bash
# round 2, session ai: ncu --set full of the pruned-tree kernel on C1 (N=512, batch 4096), summarised on the box
tag=r02ai
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sc_ssc -s 3 -c 1 -f -o /tmp/prof_${tag}_c1 python bench.py --config c1 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_full_c1.log 2>&1
python tools/ncu_summary.py /tmp/prof_${tag}_c1.ncu-rep 20 > gpurun_out/${tag}_c1_ncu_summary.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_${tag}_c1.ncu-rep 4096 c1 $tag > gpurun_out/${tag}_c1_traffic.txt 2>&1
cp profiles/ncu_traffic.json gpurun_out/${tag}_ncu_traffic.json
head -36 gpurun_out/${tag}_c1_ncu_summary.txt; cat gpurun_out/${tag}_c1_traffic.txt
