"""Measure how often the fp32 CUDA decoder's K decoded info bits differ from the double-precision
oracle on identical float LLRs (developer tool; prints one line per setting)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode

settings = [(11,1024,0,1,1.0,4000),(11,1024,0,1,2.0,4000),(11,1024,0,4,1.0,2000),(11,1024,16,4,1.0,2000),
            (11,1024,0,32,1.0,1000),(11,1024,16,32,1.0,1000),(11,1024,16,32,1.5,1000),(11,1024,16,32,2.0,1000),
            (9,256,0,32,1.0,4000),(9,256,16,32,1.0,4000)]
if len(sys.argv) > 1:
    scale = float(sys.argv[1]); settings = [s[:5] + (int(s[5]*scale),) for s in settings]
for (n,K,crc,L,eb,B) in settings:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, B, eb, 31337+L+int(eb*100))
    t=time.time(); want = port.decode_batch(llr, L, nthreads=os.cpu_count()); tc=time.time()-t
    wantf = port.decode_batch(llr, L, nthreads=os.cpu_count(), precision=1)
    got = pc.decode_batch(llr, L)
    mm = (got!=want).any(1); mmf = (wantf!=want).any(1)
    be_o = (want!=info).any(1); be_g=(got!=info).any(1)
    print(f"n={n} K={K} crc={crc} L={L} EbN0={eb}: GPU-vs-double mismatch {int(mm.sum())}/{B}; CPU-float-vs-double {int(mmf.sum())}/{B}; "
          f"block errors double {int(be_o.sum())} gpu {int(be_g.sum())}; of the mismatching cw, oracle was wrong in {int((mm&be_o).sum())}, gpu wrong in {int((mm&be_g).sum())}  (oracle {B/tc:.0f} cw/s)", flush=True)
