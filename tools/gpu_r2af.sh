#!/bin/bash
# round 2, session af: the pruned-tree tests again (fp32-mode assertions corrected)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -q -m gpu -k "pruned_tree or fp32_kernels_alone or raw_c_abi" 2>&1 | tail -12 ) | tee gpurun_out/r02af_pytest.txt
