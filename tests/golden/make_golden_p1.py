"""Golden fixtures of the probability-domain decoder, generated from the UNMODIFIED reference
(RefPolarCode::decode_scl_p1, PolarCode.cpp:110-128) through oracle/_ref/libpolar_ref.so.

    python tests/golden/make_golden_p1.py        (build container only: needs `make -C oracle ref`)

Outputs (committed): p1_{name}.npz -- float64 likelihoods p1 = P(y|1), p0 = P(y|0) per codeword, the list
size, and the info bits the reference returned; rows [0, n_edge) are the edge inputs of
oracle_lib.edge_probs (exact ties, exact zeros / ones, all-zero = the sigma == 0 branch, denormal scale),
the rest BPSK/AWGN likelihoods built as in the reference's commented-out harness lines (:749-751).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, awgn_probs, edge_probs  # noqa: E402

# name, n, K, crc, L, B (AWGN rows), Eb/N0
CASES = [
    ("n9_K256_crc16_L8", 9, 256, 16, 8, 16, 1.0),
    ("n9_K256_crc0_L1", 9, 256, 0, 1, 16, 2.0),
    ("n9_K256_crc16_L32", 9, 256, 16, 32, 12, 1.0),
    ("n7_K64_crc8_L3", 7, 64, 8, 3, 32, 0.0),
    ("n7_K64_crc8_L100", 7, 64, 8, 100, 16, 0.0),
    ("n11_K1024_crc16_L4", 11, 1024, 16, 4, 4, 1.5),
    ("n5_K16_crc4_L16", 5, 16, 4, 16, 32, 0.0),
]


def main():
    for (name, n, K, crc, L, B, eb) in CASES:
        ref = Ref(n, K, 0.32, crc)
        info, p1, p0 = awgn_probs(ref, B, eb, seed=0x9901 + 100 * n + L)
        e1, e0 = edge_probs(1 << n)
        p1 = np.concatenate([e1, p1]); p0 = np.concatenate([e0, p0])
        dec = np.stack([ref.decode_p1_one(p1[b], p0[b], L) for b in range(len(p1))])
        np.savez_compressed(os.path.join(HERE, "p1_%s.npz" % name), n=n, K=K, crc=crc, L=L, ebno=eb, n_edge=len(e1),
                            p1=p1, p0=p0, info=np.packbits(info, axis=-1), decoded=np.packbits(dec, axis=-1))
        print(name, "block errors on the AWGN rows", int((dec[len(e1):] != info).any(1).sum()), "of", B)


if __name__ == "__main__":
    main()
