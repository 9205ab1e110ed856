#!/bin/bash
# round 2, session ab: final check of the committed tree: whole GPU suite, smoke(), bench.py default line and the reference arm
tag=${1:-r02ab}
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -6 gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-1500 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench.err; cut -c1-600 gpurun_out/${tag}_bench_reference_arm.json
