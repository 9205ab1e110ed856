#!/bin/bash
# round 2, session z: the pruned-tree SC tests with the extreme-rate codes added
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -k "pruned_tree" 2>&1 | tail -12 ) > gpurun_out/r02z_pytest_ssc.txt 2>&1
cat gpurun_out/r02z_pytest_ssc.txt
