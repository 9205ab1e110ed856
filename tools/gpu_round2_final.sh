#!/bin/bash
# Final GPU session of round 2: tests, smoke, bench (all configs, both arms, min-sum line), ncu launch list + full captures
# summarised on the box, sanitizer. Usage: tools/gpu_round2_final.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -6 gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json | cut -c1-3000; tail -3 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench_reference_arm.json
for c in c1 c2 c3 c5; do python bench.py --config $c --steps 10 --warmup 3 --cpu-seconds 4 >> gpurun_out/${tag}_bench_other_configs.json 2>> gpurun_out/${tag}_bench.err; done; cut -c1-700 gpurun_out/${tag}_bench_other_configs.json
python bench.py --mode fp32 --steps 10 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_fp32_mode.json 2>> gpurun_out/${tag}_bench.err
for c in c4 c3 c2 c5; do python bench.py --mode minsum --config $c --steps 10 --warmup 3 --no-cpu >> gpurun_out/${tag}_bench_minsum_mode.json 2>> gpurun_out/${tag}_bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -14 gpurun_out/${tag}_launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:scl_fast -s 3 -c 1 -f -o /tmp/prof_$tag python bench.py --mode fp32 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --batch 16384 > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_$tag.ncu-rep 45 > gpurun_out/${tag}_c4_ncu_summary.txt 2>&1
python tools/ncu_funcs.py /tmp/prof_$tag.ncu-rep > gpurun_out/${tag}_c4_by_function.txt 2>&1
python tools/ncu_lines.py /tmp/prof_$tag.ncu-rep 60 > gpurun_out/${tag}_c4_by_source_line.txt 2>&1
python tools/ncu_traffic.py /tmp/prof_$tag.ncu-rep 16384 c4 $tag > gpurun_out/${tag}_c4_traffic.txt 2>&1
for c in c2 c3 c5; do
  ncu --set full --clock-control none --import-source on -k regex:scl_fast -s 3 -c 1 -f -o /tmp/prof_${tag}_$c python bench.py --mode fp32 --config $c --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${tag}_ncu_full_$c.log 2>&1
  python tools/ncu_summary.py /tmp/prof_${tag}_$c.ncu-rep 20 > gpurun_out/${tag}_${c}_ncu_summary.txt 2>&1
  python tools/ncu_traffic.py /tmp/prof_${tag}_$c.ncu-rep 65536 $c $tag >> gpurun_out/${tag}_c4_traffic.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:scl_exact -s 2 -c 1 -f -o /tmp/prof_${tag}_exact python tools/prof_exact.py 11 1024 16 32 831 1.0 > gpurun_out/${tag}_ncu_exact.log 2>&1
python tools/ncu_summary.py /tmp/prof_${tag}_exact.ncu-rep 25 > gpurun_out/${tag}_second_pass_kernel_ncu_summary.txt 2>&1
cp profiles/ncu_traffic.json gpurun_out/${tag}_ncu_traffic.json
( compute-sanitizer --tool memcheck python tools/sanitize_probe.py 2>&1 | tail -25; compute-sanitizer --tool racecheck python tools/sanitize_probe.py 2>&1 | tail -8 ) > gpurun_out/${tag}_compute_sanitizer.txt 2>&1; tail -12 gpurun_out/${tag}_compute_sanitizer.txt
python tools/bler_curve.py > gpurun_out/${tag}_bler_curve.json 2> gpurun_out/${tag}_bler_curve.err; tail -c 600 gpurun_out/${tag}_bler_curve.json; tail -2 gpurun_out/${tag}_bler_curve.err
ls -la gpurun_out/ | tail -30
