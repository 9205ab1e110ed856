#!/bin/bash
# round 2, session ae: which tie / extreme rows differ from the oracle in fp32 mode, per kernel
mkdir -p gpurun_out
cat > /tmp/probe_ties.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from oracle_lib import Port
from polar_b200 import PolarCode
for (n, K, crc) in [(9, 256, 0), (11, 1024, 16)]:
    port, pc = Port(n, K, 0.32, crc), PolarCode(n, K, 0.32, crc)
    N = 1 << n
    rng = np.random.default_rng(5)
    rows = [np.zeros(N), np.full(N, 3.0), np.full(N, -3.0), np.full(N, 1000.0), np.full(N, -1000.0),
            rng.integers(-2, 3, N).astype(np.float64), rng.integers(-1, 2, N) * 39.5,
            np.where(rng.random(N) < 0.5, 0.0, rng.normal(2, 2, N))]
    llr = np.stack(rows).astype(np.float32)
    want = port.decode_batch(llr, 1)
    for env in ({}, {"POLAR_B200_SSC": "0"}, {"POLAR_B200_FORCE_GENERIC": "1"}):
        for k in ("POLAR_B200_SSC", "POLAR_B200_FORCE_GENERIC"):
            os.environ.pop(k, None)
        os.environ.update(env)
        got = pc.decode_batch(llr, 1, mode="fp32")
        bad = np.where((got != want).any(1))[0].tolist()
        print(n, K, crc, env, "kernel kind", pc.info(6), "rows differing from the oracle:", bad, [int((got[b] != want[b]).sum()) for b in bad])
PY
python /tmp/probe_ties.py 2>&1 | tail -8 | tee gpurun_out/r02ae_ties_fp32.txt
