"""BLER vs Eb/N0 on the GPU (device front end + decoder), next to the oracle on a subsample and the
values read off the reference's results/polar_performance.jpeg. Writes gpurun_out/bler_curve_<tag>.json (copy it to profiles/)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from polar_b200 import PolarCode, bler, unpack_bits
from oracle_lib import Port
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
ebno = [1.0, 1.25, 1.5, 1.75, 2.0, 2.25, 2.5]
res = {"codewords_per_point": total, "ebno_db": ebno, "curves": {}}
for (name, n, K, crc, lists) in [("N2048_K1024_crc0", 11, 1024, 0, [1, 4, 32]), ("N2048_K1024_crc16", 11, 1024, 16, [4, 32])]:
    pc = PolarCode(n, K, 0.32, crc)
    t = time.time(); c = bler.sweep_counts_device(pc, lists, ebno, total, seed=0xB1E5); dt = time.time() - t
    for il, L in enumerate(lists):
        res["curves"]["%s_L%d" % (name, L)] = {"bler": (c[il, :, 0] / c[il, :, 1]).tolist(), "errors": c[il, :, 0].tolist()}
    res["curves"][name + "_seconds"] = dt
    # oracle on a subsample of the same device-generated codewords (1.5 dB, largest list)
    llr, truth = pc.synthesize(1024, [1.5], 0xB1E5 + 2)
    want = Port(n, K, 0.32, crc).decode_batch(llr.cpu().numpy(), lists[-1], nthreads=os.cpu_count())
    got = unpack_bits(pc.decode_device(llr, lists[-1]).cpu().numpy().view(np.uint32), K)
    res["curves"][name + "_oracle_check"] = {"codewords": 1024, "list": lists[-1], "ebno_db": 1.5,
                                            "mismatching_codewords": int((got != want).any(1).sum())}
res["readme_figure_readoffs"] = {"L1_2.0dB": 5e-2, "L32_crc0_2.0dB": 1.5e-3, "L32_crc16_2.0dB": 3e-5}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bler_curve_%s.json" % tag), "w"), indent=1)
print(json.dumps(res, indent=1))
