"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (where /root/reference exists and `make -C oracle ref` has produced
oracle/_ref/libpolar_ref.so):

    python tests/golden/make_golden.py

Outputs (committed):
    construction_n{n}_K{K}_crc{crc}.npz  frozen mask, reliability order, parity matrix, bit-reversal
                                          table read out of RefPolarCode (PolarCode.h:43-46)
    decode_{name}.npz                     float32 LLR batches + the info bits RefPolarCode::decode_scl_llr
                                          returned for them (LLRs widened to double), incl. edge cases
    ref_main_table.txt                    stdout table of the reference's own main.cpp (progress lines removed)

The reference has no golden vectors of its own; these are the pins.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import REF_MAIN, Ref, awgn_llrs  # noqa: E402

CONSTRUCTIONS = [(9, 256, 0), (9, 256, 16), (11, 1024, 0), (11, 1024, 16), (5, 16, 4), (7, 64, 8), (10, 512, 8)]

# name, n, K, crc, L, B, ebno
DECODES = [
    ("c1_n9_K256_crc0_L1", 9, 256, 0, 1, 64, 2.0),
    ("c2_n11_K1024_crc0_L1", 11, 1024, 0, 1, 16, 2.0),
    ("c3_n11_K1024_crc16_L4", 11, 1024, 16, 4, 16, 1.5),
    ("c4_n11_K1024_crc16_L32", 11, 1024, 16, 32, 12, 1.25),
    ("c5_n9_K256_crc0_L32", 9, 256, 0, 32, 32, 2.0),
    ("c5b_n9_K256_crc16_L32", 9, 256, 16, 32, 32, 1.5),
    ("x_n7_K64_crc8_L3", 7, 64, 8, 3, 64, 0.5),
    ("x_n10_K512_crc8_L8", 10, 512, 8, 8, 24, 1.5),
    ("x_n5_K16_crc4_L16", 5, 16, 4, 16, 64, 0.0),
]


def edge_llrs(code, rng):
    """adversarial inputs of SURVEY.md section 4: exact zeros, the +-40 f-rule boundary, softplus
    overflow (|LLR| > 709.78), small integers (many exact metric ties), mixed huge/small."""
    N = code.N
    rows = []
    rows.append(np.zeros(N))
    rows.append(np.full(N, 1000.0))
    rows.append(np.full(N, -1000.0))
    rows.append(rng.choice([-1000.0, 1000.0], N))
    rows.append(rng.choice([-40.0, 40.0], N))
    rows.append(rng.choice([-39.999996, 39.999996, 40.0, -40.0], N))
    rows.append(rng.integers(-3, 4, N).astype(np.float64))
    rows.append(rng.integers(-1, 2, N).astype(np.float64))
    rows.append(rng.choice([-2.0, 2.0], N))
    rows.append(np.where(rng.random(N) < 0.1, 0.0, rng.normal(2.0, 2.0, N)))
    rows.append(rng.normal(0.0, 300.0, N))
    rows.append(rng.normal(30.0, 20.0, N))
    rows.append(np.full(N, 5.0))
    rows.append(np.full(N, -5.0))
    return np.asarray(rows, np.float32)


def main():
    for (n, K, crc) in CONSTRUCTIONS:
        c = Ref(n, K, 0.32, crc).construction()
        np.savez_compressed(os.path.join(HERE, "construction_n%d_K%d_crc%d.npz" % (n, K, crc)),
                            frozen=np.packbits(c["frozen"]), order=c["order"],
                            crc_matrix=np.packbits(c["crc_matrix"], axis=-1) if crc else np.zeros((0, 0), np.uint8),
                            bitrev=c["bitrev"])
    for (name, n, K, crc, L, B, eb) in DECODES:
        ref = Ref(n, K, 0.32, crc)
        info, llr = awgn_llrs(ref, B, eb, seed=0x601D + n * 100 + L)
        dec = ref.decode_batch(llr, L, nthreads=8)
        np.savez_compressed(os.path.join(HERE, "decode_%s.npz" % name), n=n, K=K, crc=crc, L=L, ebno=eb,
                            llr=llr, info=np.packbits(info, axis=-1), decoded=np.packbits(dec, axis=-1))
        print(name, "block errors", int((dec != info).any(1).sum()), "of", B)
    # edge cases: small code so that every list size is cheap; lists 1, 2, 4, 32
    rng = np.random.default_rng(0xED6E)
    for (n, K, crc) in [(9, 256, 16), (7, 64, 8)]:
        ref = Ref(n, K, 0.32, crc)
        llr = edge_llrs(ref, rng)
        out = {}
        for L in (1, 2, 4, 32):
            out["decoded_L%d" % L] = np.packbits(ref.decode_batch(llr, L, nthreads=8), axis=-1)
        np.savez_compressed(os.path.join(HERE, "edge_n%d_K%d_crc%d.npz" % (n, K, crc)), n=n, K=K, crc=crc, llr=llr, **out)
    if os.path.exists(REF_MAIN):
        txt = subprocess.run([REF_MAIN], capture_output=True, text=True, check=True).stdout
        rows = [ln for ln in txt.splitlines() if ln.strip() and not ln.startswith("Running iteration")]
        with open(os.path.join(HERE, "ref_main_table.txt"), "w") as f:
            f.write("\n".join(rows) + "\n")
        print("\n".join(rows))


if __name__ == "__main__":
    main()
