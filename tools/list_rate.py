"""Device-resident decode rate for one (n, K, crc, L, batch): usage list_rate.py n K crc L B [ebno]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from polar_b200 import PolarCode
n, K, crc, L, B = (int(x) for x in sys.argv[1:6])
eb = float(sys.argv[6]) if len(sys.argv) > 6 else 2.0
pc = PolarCode(n, K, 0.32, crc)
llr, truth = pc.synthesize(B, [eb], seed=11)
out = pc.decode_device(llr, L)
for _ in range(2):
    pc.decode_device(llr, L, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
e0.record()
for _ in range(steps):
    pc.decode_device(llr, L, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
nerr = int((out != truth).any(dim=1).sum().item())
print("N=%d K=%d crc=%d L=%d B=%d: %.3f ms, %.0f codewords/s, kernel kind %d, BLER %.4g" % (1 << n, K, crc, L, B, ms, B / ms * 1e3, pc.info(6), nerr / B))
