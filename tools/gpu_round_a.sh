#!/bin/bash
# Trimmed GPU session: tests, smoke, bench (all configs), ncu launch list + full capture, sanitizer. Usage: tools/gpu_round_a.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_$tag.txt 2>&1; tail -8 gpurun_out/pytest_gpu_$tag.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
for c in c2 c3 c5; do python bench.py --config $c --steps 5 --warmup 3 --cpu-seconds 4 >> gpurun_out/bench_other_$tag.json 2>> gpurun_out/bench_$tag.err; done; cat gpurun_out/bench_other_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_bench_$tag.log 2>&1
tail -12 gpurun_out/launches_$tag.csv
ncu --set full --clock-control none --import-source on -k regex:scl_ -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/
timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_probe.py 2>&1 | tail -3 > gpurun_out/sanitizer_$tag.txt; timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_probe.py 2>&1 | tail -2 >> gpurun_out/sanitizer_$tag.txt; cat gpurun_out/sanitizer_$tag.txt
