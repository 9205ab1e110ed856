#!/bin/bash
# ncu --set full capture of the decode kernel for the other BASELINE configs; summarised on the box (the reports
# are ~24 MB each and gpurun_out/ is capped at 64 MiB). usage: tools/prof_others.sh <tag> [configs...]
tag=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o /tmp/prof_${tag}_$c python bench.py --config $c --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_full_${tag}_$c.log 2>&1
  tail -1 gpurun_out/ncu_full_${tag}_$c.log
  python tools/ncu_summary.py /tmp/prof_${tag}_$c.ncu-rep 40 > gpurun_out/${tag}_${c}_ncu_summary.txt 2>&1
  python tools/ncu_funcs.py /tmp/prof_${tag}_$c.ncu-rep > gpurun_out/${tag}_${c}_by_function.txt 2>&1
  python tools/ncu_lines.py /tmp/prof_${tag}_$c.ncu-rep 60 > gpurun_out/${tag}_${c}_by_source_line.txt 2>&1
  python tools/ncu_traffic.py /tmp/prof_${tag}_$c.ncu-rep 65536 $c $tag > gpurun_out/${tag}_${c}_traffic.txt 2>&1
done
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_${tag}.json
ls -la gpurun_out/
