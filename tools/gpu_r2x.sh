#!/bin/bash
# round 2, session x: final check of the tree as committed: whole GPU suite, smoke(), racecheck of the pruned-tree SC kernel in full
tag=${1:-r02x}
mkdir -p gpurun_out
( time python -m pytest tests -q -m gpu 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_gpu.txt 2>&1; tail -6 gpurun_out/${tag}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cat > /tmp/probe_ssc.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
for (n, K, crc, B) in [(11, 1024, 16, 21), (9, 256, 0, 40), (12, 2048, 0, 9), (8, 128, 8, 50)]:
    port, pc = Port(n, K, 0.32, crc), PolarCode(n, K, 0.32, crc)
    _, llr = awgn_llrs(port, B, 1.5, seed=3)
    ok = np.array_equal(pc.decode_batch(llr, 1), port.decode_batch(llr, 1))
    print(n, K, crc, B, "kernel kind", pc.info(6), "ok" if ok else "MISMATCH")
PY
( compute-sanitizer --tool racecheck python /tmp/probe_ssc.py 2>&1 | tail -40; compute-sanitizer --tool memcheck python /tmp/probe_ssc.py 2>&1 | tail -8; compute-sanitizer --tool synccheck python /tmp/probe_ssc.py 2>&1 | tail -6 ) > gpurun_out/${tag}_sanitizer_ssc.txt 2>&1
cat gpurun_out/${tag}_sanitizer_ssc.txt
python bench.py --config c2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | cut -c1-200
