#!/bin/bash
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 ) 2>&1 | tail -12
for v in 1 0; do echo -n "strict c4 verify=$v: "; POLAR_B200_VERIFY=$v timeout 300 python bench.py --mode strict --config c4 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2k.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), round(d['e2e_sweep']['value']), d['modes'])"; done
echo -n "strict c5: "; timeout 300 python bench.py --mode strict --config c5 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_r2k.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['modes'])"
for mr in 4 6 8; do echo -n "e2e chunk cap $mr rounds: "; POLAR_B200_HOST_MAX_ROUNDS=$mr timeout 300 python bench.py --mode strict --config c4 --steps 3 --warmup 3 --no-cpu --e2e-steps 3 2>>gpurun_out/bench_r2k.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), d['e2e']['pipelined_chunks'])"; done
tail -3 gpurun_out/bench_r2k.err
python -c "
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, torch
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
port, pc = Port(11,1024,0.32,16), PolarCode(11,1024,0.32,16)
_, llr = awgn_llrs(port, 65536, 1.0, 3)
d = torch.from_numpy(llr).cuda()
o = pc.decode_device(d, 32, mode='strict'); torch.cuda.synchronize()
g32, g64, cw = pc.verify_gaps()
print('1.0 dB, tau 1e-5: records', pc.last_recorded, 'second pass', pc.last_flagged, 'max |gap64-gap32| %.3g' % np.abs(g64-g32).max(), 'records where double disagrees:', int((g64 < 0).sum()))
"
