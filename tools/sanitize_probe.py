"""Small decodes through every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle_lib import Port, awgn_llrs
from polar_b200 import PolarCode
ok = True
for (n,K,crc,L,B) in [(9,256,16,32,6),(9,256,16,4,19),(9,256,0,16,5),(11,1024,16,32,3),(11,1024,16,8,9),(7,64,8,3,21),(9,256,0,1,40),(8,128,8,32,4),(11,1024,16,1,21),(12,2048,0,1,9),(8,128,8,1,50)]:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, B, 1.5, 5)
    want = port.decode_batch(llr, L)
    got = pc.decode_batch(llr, L)
    g64 = pc.decode_batch_f64(llr.astype(np.float64), L)
    l2, t2 = pc.synthesize(8, [2.0], 3)
    same = np.array_equal(got, want) and np.array_equal(g64, want)
    ok &= same
    print(n, K, crc, L, B, "kernel kind", pc.info(6), "ok" if same else "MISMATCH", flush=True)
# wide-list kernel (lists 33..127, fp32 and f64) and the probability-domain decoder
from oracle_lib import awgn_probs
for (n,K,crc,L,B) in [(9,256,16,40,3),(7,64,8,100,5),(9,256,0,64,2)]:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, B, 1.0, 6)
    want = port.decode_batch(llr, L)
    same = np.array_equal(pc.decode_batch(llr, L), want) and np.array_equal(pc.decode_batch_f64(llr.astype(np.float64), L), want)
    ok &= same
    print(n, K, crc, L, B, "kernel kind", pc.info(6), "ok" if same else "MISMATCH", flush=True)
for (n,K,crc,L,B) in [(9,256,16,8,4),(7,64,8,48,5),(8,100,7,127,2)]:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, p1, p0 = awgn_probs(port, B, 1.0, 7)
    same = np.array_equal(pc.decode_p1_batch(p1, p0, L), port.decode_p1_batch(p1, p0, L))
    ok &= same
    print(n, K, crc, L, B, "p1 kernel kind", pc.info(6), "ok" if same else "MISMATCH", flush=True)
# strict mode with a huge threshold (most codewords take the second pass: block-per-codeword double kernel in list mode,
# then the literal-formula kernel's list mode), the min-sum build, the fused-count sweep
for (n,K,crc,L,B) in [(9,256,16,32,12),(11,1024,16,4,9),(9,256,0,1,40),(11,1024,16,32,4)]:
    port = Port(n,K,0.32,crc); pc = PolarCode(n,K,0.32,crc)
    info, llr = awgn_llrs(port, B, 1.0, 8)
    want = port.decode_batch(llr, L)
    pc.set_strict_tau(1e-3)
    got = pc.decode_batch(llr, L, mode="strict")
    nf = pc.last_flagged
    ms = pc.decode_batch(llr, L, mode="minsum")
    same = np.array_equal(got, want) and np.array_equal(ms, port.decode_batch(llr, L, minsum_only=2))
    c = pc.bler_sweep_device([1.0, 2.0], [L], 64, seed=3, mode="strict")
    ok &= same
    print(n, K, crc, L, B, "strict: second pass on", nf, "of", B, "; sweep counts", c.tolist(), "ok" if same else "MISMATCH", flush=True)
print("ALL OK" if ok else "FAIL")
