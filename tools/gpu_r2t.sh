#!/bin/bash
# round 2, session t: the pruned-tree SC kernel (sc_ssc.cuh): its tests, rates against the leaf-by-leaf kernel, one ncu capture
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -x -k "pruned_tree" 2>&1 | tail -25 ) > gpurun_out/r02t_pytest_ssc.txt 2>&1
tail -5 gpurun_out/r02t_pytest_ssc.txt
{
export POLAR_B200_STRICT_TAU=1e-30
for w in 10 8 6; do echo "warps $w:"; POLAR_B200_SSC_WARPS=$w timeout 120 python tools/list_rate.py 11 1024 0 1 65536 1.5; done
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 11 1024 0 1 65536 1.5
timeout 120 python tools/list_rate.py 11 1024 0 1 262144 1.5
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 11 1024 0 1 262144 1.5
timeout 120 python tools/list_rate.py 9 256 0 1 4096 2.0
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 9 256 0 1 4096 2.0
timeout 120 python tools/list_rate.py 9 256 0 1 262144 2.0
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 9 256 0 1 262144 2.0
timeout 120 python tools/list_rate.py 12 2048 0 1 32768 2.0
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 12 2048 0 1 32768 2.0
timeout 120 python tools/list_rate.py 10 512 0 1 131072 2.0
POLAR_B200_SSC=0 timeout 120 python tools/list_rate.py 10 512 0 1 131072 2.0
} > gpurun_out/r02t_rates.txt 2>&1
cat gpurun_out/r02t_rates.txt
POLAR_B200_STRICT_TAU=1e-30 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sc_ssc -s 2 -c 1 -f -o /tmp/prof_ssc python tools/list_rate.py 11 1024 0 1 65536 1.5 > gpurun_out/r02t_ncu.log 2>&1
python tools/ncu_summary.py /tmp/prof_ssc.ncu-rep 40 > gpurun_out/r02t_ssc_ncu_summary.txt 2>&1
python tools/ncu_lines.py /tmp/prof_ssc.ncu-rep 50 > gpurun_out/r02t_ssc_by_source_line.txt 2>&1
python tools/ncu_stall_lines.py /tmp/prof_ssc.ncu-rep 40 > gpurun_out/r02t_ssc_stalls_by_line.txt 2>&1
head -40 gpurun_out/r02t_ssc_ncu_summary.txt
