import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
for (n,K,crc,L,B) in [(11,1024,16,32,8192),(11,1024,0,1,32768),(11,1024,16,4,16384)]:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, 1024, 1.5, 3)
    big = np.tile(llr.astype(np.float64), (B//1024, 1))
    pc.decode_batch_f64(big[:1024], L)
    t=time.time(); out = pc.decode_batch_f64(big, L); dt=time.time()-t
    print(f"f64 mode n={n} K={K} crc={crc} L={L}: {B/dt:.0f} cw/s end to end (B={B}); block errors {(out[:1024]!=info).any(1).sum()}/1024")
