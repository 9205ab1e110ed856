"""Quick GPU parity + timing probe (developer tool; the real tests are under tests/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode

cfgs = [(3,4,0,1,64,1.0),(4,8,2,2,64,0.0),(5,16,4,4,64,0.0),(6,32,8,8,64,0.0),(7,64,8,32,32,1.0),(7,64,0,3,64,0.0),
        (6,20,3,5,48,-1.0),(8,128,16,16,32,1.0),(9,256,0,1,256,2.0),(9,256,16,32,64,1.5),(9,256,0,32,64,2.0),
        (11,1024,0,1,256,2.0),(11,1024,16,4,128,1.5),(11,1024,16,32,64,1.25),(11,1024,0,2,64,1.5),(11,1024,0,8,64,1.5)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    cfgs = cfgs[:8]
allok = True
for (n,K,crc,L,B,eb) in cfgs:
    port = Port(n,K,0.32,crc)
    pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, B, eb, 1000+n+L)
    want = port.decode_batch(llr, L, nthreads=8)
    t=time.time(); got = pc.decode_batch(llr, L); dt=time.time()-t
    bad = int((got!=want).any(1).sum())
    allok &= bad == 0
    print(f"n={n} K={K} crc={crc} L={L} B={B}: mismatch {bad}, blkerr ref {int((want!=info).any(1).sum())} gpu {int((got!=info).any(1).sum())}  ({dt*1e3:.1f} ms)", flush=True)
print("ALL OK" if allok else "MISMATCHES")
# timing at size
for (n,K,crc,L,B) in [(11,1024,16,32,4096),(11,1024,0,1,65536),(11,1024,16,4,16384),(9,256,0,32,16384)]:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, min(B,2048), 2.0, 5)
    reps = B // llr.shape[0]
    d = torch.from_numpy(llr).cuda().repeat(reps,1).contiguous()
    out = pc.decode_device(d, L); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); out = pc.decode_device(d, L); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    from polar_b200 import unpack_bits
    got = unpack_bits(out.cpu().numpy().view(np.uint32), K)[:info.shape[0]]
    print(f"n={n} K={K} crc={crc} L={L} B={d.shape[0]}: {ms:.2f} ms -> {d.shape[0]/ms*1e3:.0f} cw/s; blkerr {int((got!=info).any(1).sum())}/{info.shape[0]}", flush=True)
