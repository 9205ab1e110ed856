"""More golden decode fixtures from the UNMODIFIED reference (oracle/_ref/libpolar_ref.so): list sizes above 32
(the reference accepts up to 127, PolarCode.cpp:497-605) and a block length above 2^13. Same file format as
make_golden.py (decode_{name}.npz), so the port and GPU golden tests pick them up.

    python tests/golden/make_golden_wide.py        (build container only: needs `make -C oracle ref`)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Ref, awgn_llrs  # noqa: E402

# name, n, K, crc, L, B, ebno
DECODES = [
    ("w_n9_K256_crc16_L100", 9, 256, 16, 100, 12, 1.0),
    ("w_n7_K64_crc8_L127", 7, 64, 8, 127, 32, 0.0),
    ("w_n10_K512_crc0_L33", 10, 512, 0, 33, 8, 1.0),
    ("w_n14_K8192_crc16_L4", 14, 8192, 16, 4, 2, 2.0),
]


def main():
    for (name, n, K, crc, L, B, eb) in DECODES:
        ref = Ref(n, K, 0.32, crc)
        info, llr = awgn_llrs(ref, B, eb, seed=0x71DE + n * 100 + L)
        dec = ref.decode_batch(llr, L, nthreads=8)
        np.savez_compressed(os.path.join(HERE, "decode_%s.npz" % name), n=n, K=K, crc=crc, L=L, ebno=eb,
                            llr=llr, info=np.packbits(info, axis=-1), decoded=np.packbits(dec, axis=-1))
        print(name, "block errors", int((dec != info).any(1).sum()), "of", B)


if __name__ == "__main__":
    main()
