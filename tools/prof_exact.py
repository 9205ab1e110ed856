"""Time the block-per-codeword double kernel (scl_exact.cuh) on a few hundred codewords (developer tool, needs a GPU).
usage: python tools/prof_exact.py [n K crc L count ebno]   (under ncu: -k regex:scl_exact)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["POLAR_B200_F64_EXACT_KERNEL"] = "1"
import numpy as np, torch
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
a = sys.argv[1:]
n, K, crc, L, B, eb = (int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), float(a[5])) if len(a) >= 6 else (11, 1024, 16, 32, 296, 1.0)
port, pc = Port(n, K, 0.32, crc), PolarCode(n, K, 0.32, crc)
_, llr = awgn_llrs(port, B, eb, 99)
d = torch.from_numpy(llr).cuda()
for _ in range(2):
    out = pc.decode_device(d, L, mode="f64")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = pc.decode_device(d, L, mode="f64"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("exact kernel: n=%d L=%d, %d codewords in %.3f ms = %.3f ms per codeword per block (kernel kind %d, %d blocks)" % (
    n, L, B, ms, ms / max(1, -(-B // pc.info(3))), pc.info(6), pc.info(3)), flush=True)
