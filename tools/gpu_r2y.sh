#!/bin/bash
# round 2, session y: compute-sanitizer --tool synccheck on the pruned-tree SC kernel and, for comparison, on the other kernels
mkdir -p gpurun_out
cat > /tmp/probe_a.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
which = sys.argv[1]
n, K, crc, B = 9, 256, 0, 40
port, pc = Port(n, K, 0.32, crc), PolarCode(n, K, 0.32, crc)
_, llr = awgn_llrs(port, B, 1.5, seed=3)
d = torch.from_numpy(llr).cuda()
if which == "ssc":
    pc.set_strict_tau(1e-30)
    out = pc.decode_device(d, 1, mode="strict")
elif which == "fast1":
    out = pc.decode_device(d, 1, mode="fp32")
elif which == "fast32":
    out = pc.decode_device(d, 32, mode="fp32")
elif which == "exact":
    pc.set_strict_tau(1.0)
    out = pc.decode_device(d, 4, mode="strict")
torch.cuda.synchronize()
print(which, "kernel kind", pc.info(6), "done")
PY
for w in ssc fast1 fast32 exact; do
  echo "=== $w"; compute-sanitizer --tool synccheck python /tmp/probe_a.py $w 2>&1 | grep -v "^=========     " | tail -14
done > gpurun_out/r02y_synccheck.txt 2>&1
cat gpurun_out/r02y_synccheck.txt | cut -c1-220
