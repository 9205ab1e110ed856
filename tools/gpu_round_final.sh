#!/bin/bash
# Final GPU session of a round: tests, smoke, bench (all configs + reference arm), ncu launch list + full capture
# summarised on the box. Usage: tools/gpu_round_final.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -6 gpurun_out/pytest_gpu_$tag.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_ref_$tag.json
for c in c2 c3 c5; do python bench.py --config $c --steps 5 --warmup 3 --cpu-seconds 4 >> gpurun_out/bench_other_$tag.json 2>> gpurun_out/bench_$tag.err; done; cat gpurun_out/bench_other_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_$tag.log 2>&1
tail -9 gpurun_out/launches_$tag.csv
ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o /tmp/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 > gpurun_out/ncu_full_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_$tag.ncu-rep 45 > gpurun_out/${tag}_c4_ncu_summary.txt 2>&1
python tools/ncu_funcs.py /tmp/prof_$tag.ncu-rep > gpurun_out/${tag}_c4_by_function.txt 2>&1
python tools/ncu_lines.py /tmp/prof_$tag.ncu-rep 60 > gpurun_out/${tag}_c4_by_source_line.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_$tag.ncu-rep 16384 c4 $tag > gpurun_out/${tag}_c4_traffic.txt 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic_$tag.json
cp /tmp/prof_$tag.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/ | tail -20
