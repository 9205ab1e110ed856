#!/bin/bash
# round 2, session w: records for the pruned-tree SC kernel (list size 1): bench lines of c1 / c2 (+ the default c4 as a sanity
# check), launch list and ncu --set full of c2 summarised on the box, compute-sanitizer, threshold campaign at scale
tag=${1:-r02w}
mkdir -p gpurun_out
for c in c1 c2; do python bench.py --config $c --steps 10 --warmup 3 --cpu-seconds 4 >> gpurun_out/${tag}_bench_c1_c2.json 2>> gpurun_out/${tag}_bench.err; done; cut -c1-900 gpurun_out/${tag}_bench_c1_c2.json
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_c4_sanity.json 2>> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench_c4_sanity.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_c2.csv python bench.py --config c2 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_bench_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sc_ssc -s 3 -c 1 -f -o /tmp/prof_${tag}_c2 python bench.py --config c2 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_full_c2.log 2>&1
python tools/ncu_summary.py /tmp/prof_${tag}_c2.ncu-rep 40 > gpurun_out/${tag}_c2_ncu_summary.txt 2>&1
python tools/ncu_lines.py /tmp/prof_${tag}_c2.ncu-rep 60 > gpurun_out/${tag}_c2_by_source_line.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_${tag}_c2.ncu-rep 65536 c2 $tag > gpurun_out/${tag}_c2_traffic.txt 2>&1
cp profiles/ncu_traffic.json gpurun_out/${tag}_ncu_traffic.json
head -30 gpurun_out/${tag}_c2_ncu_summary.txt
( compute-sanitizer --tool memcheck python tools/sanitize_probe.py 2>&1 | tail -25; compute-sanitizer --tool racecheck python tools/sanitize_probe.py 2>&1 | tail -8 ) > gpurun_out/${tag}_compute_sanitizer.txt 2>&1; tail -12 gpurun_out/${tag}_compute_sanitizer.txt
timeout 900 python tools/flip_margins.py 10 gpurun_out/${tag}_flip_sc.json sc > gpurun_out/${tag}_flip_sc.txt 2>&1
cut -c1-250 gpurun_out/${tag}_flip_sc.txt
