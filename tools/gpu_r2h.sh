#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "8192 or strict_mode or edge_cases" 2>&1 | tail -5
for cfg in "256 3" "128 4" "128 6" "128 8" "96 6" "192 4"; do set -- $cfg; echo -n "threads=$1 bps=$2: "; POLAR_B200_EXACT_THREADS=$1 POLAR_B200_EXACT_BPS=$2 python tools/prof_exact.py 11 1024 16 32 831 1.0; done
for cfg in "256 3" "128 6"; do set -- $cfg; echo -n "strict c4 threads=$1 bps=$2: "; POLAR_B200_EXACT_THREADS=$1 POLAR_B200_EXACT_BPS=$2 timeout 300 python bench.py --mode strict --config c4 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2h.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['modes'])"; done
python - <<'PY'
import sys, os, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
for (n, K, crc, L, B) in [(13, 4096, 16, 32, 8192), (13, 4096, 16, 4, 16384), (13, 4096, 0, 1, 32768)]:
    port, pc = Port(n, K, 0.32, crc), PolarCode(n, K, 0.32, crc)
    _, llr = awgn_llrs(port, 64, 2.0, 5)
    want = port.decode_batch(llr, L, nthreads=os.cpu_count())
    got = pc.decode_batch(llr, L, mode="fp32")
    print("N=8192 L=%d parity on 64 codewords: %d differ, kernel kind %d" % (L, int((got != want).any(1).sum()), pc.info(6)))
    d = torch.from_numpy(np.tile(llr, (B // 64, 1))).cuda()
    for mode in ("fp32", "strict"):
        for _ in range(2): pc.decode_device(d, L, mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3): pc.decode_device(d, L, mode=mode)
        e1.record(); torch.cuda.synchronize()
        print("  %s: %.0f codewords/s (kind %d)" % (mode, B * 3 / (e0.elapsed_time(e1) * 1e-3), pc.info(6)))
PY
