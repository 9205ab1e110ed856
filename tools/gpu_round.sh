#!/bin/bash
# One GPU session: tests, smoke, bench, ncu launch list + full capture. Usage: tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.txt; tail -5 gpurun_out/pytest_gpu_$tag.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_ref_$tag.json
for c in c2 c3 c5; do python bench.py --config $c --steps 5 --warmup 3 --cpu-seconds 5 >> gpurun_out/bench_other_$tag.json 2>> gpurun_out/bench_$tag.err; done; cat gpurun_out/bench_other_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_$tag.log 2>&1
tail -12 gpurun_out/launches_$tag.csv
ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_probe.py 2>&1 | tail -3 > gpurun_out/sanitizer_$tag.txt; timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_probe.py 2>&1 | tail -2 >> gpurun_out/sanitizer_$tag.txt; cat gpurun_out/sanitizer_$tag.txt
python tools/bler_curve.py $tag 262144 > gpurun_out/bler_curve_$tag.log 2>&1; tail -3 gpurun_out/bler_curve_$tag.log
python tools/mismatch_rate.py > gpurun_out/mismatch_$tag.txt 2>&1; cat gpurun_out/mismatch_$tag.txt
oracle/_ref/polar_b200_main | grep -v "^Running" > gpurun_out/main_table_$tag.txt 2>&1; cat gpurun_out/main_table_$tag.txt
