import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from conftest import load_edge
from polar_b200 import PolarCode
for (n,K,crc) in [(9,256,16),(7,64,8)]:
    e = load_edge(n,K,crc); pc = PolarCode(n,K,0.32,crc)
    for L in (1,2,4,32):
        g = pc.decode_batch(e["llr"], L)
        print(n, L, "gpu vs golden:", [i for i in range(len(g)) if not np.array_equal(g[i], e[L][i])])
