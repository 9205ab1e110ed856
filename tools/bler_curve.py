"""BLER vs Eb/N0 (developer tool, needs a GPU): the C++ multi-GPU sweep (PolarCode::bler_sweep: device front end + decoder +
fused block-error count, counters reduced with ncclAllReduce) with 95 % Wilson intervals, next to the unmodified CPU
reference on a sample of the same device-generated codewords per point (same codewords through both decoders), and the
values read off the reference's results/polar_performance.jpeg.
usage: python tools/bler_curve.py [codewords per point] [reference sample per point] > profiles/rNN_bler_curve.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from polar_b200 import PolarCode, unpack_bits
from oracle_lib import Port, Ref, have_ref
from bench import wilson

total = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
sample = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ebno = [1.0, 1.25, 1.5, 1.75, 2.0, 2.25, 2.5]
res = {"codewords_per_point": total, "reference_sample_per_point": sample, "ebno_db": ebno, "mode": "strict", "curves": {}}
for (name, n, K, crc, lists) in [("N2048_K1024_crc0", 11, 1024, 0, [1, 4, 32]), ("N2048_K1024_crc16", 11, 1024, 16, [4, 32])]:
    pc = PolarCode(n, K, 0.32, crc)
    cpu = (Ref if have_ref() else Port)(n, K, 0.32, crc)
    t = time.time(); _, c = pc.bler_sweep(ebno, lists, total * len(ebno), seed=0xB1E5); dt = time.time() - t
    for il, L in enumerate(lists):
        pts = []
        for ie, eb in enumerate(ebno):
            e, r = int(c[il, ie, 0]), int(c[il, ie, 1])
            pts.append({"ebno_db": eb, "n": r, "err_gpu": e, "bler_gpu": e / r, "ci95": wilson(e, r)})
        res["curves"]["%s_L%d" % (name, L)] = pts
    res["curves"][name + "_seconds"] = dt
    # the reference on a sample of device-generated codewords per point, largest list: same codewords through both decoders
    L = lists[-1]
    for ie, eb in enumerate(ebno):
        llr, truth = pc.synthesize(sample, [eb], 0xB1E5 + 100 + ie)
        info = unpack_bits(truth.cpu().numpy().view(np.uint32), K)
        want = cpu.decode_batch(llr.cpu().numpy(), L, nthreads=os.cpu_count())
        got = unpack_bits(pc.decode_device(llr, L).cpu().numpy().view(np.uint32), K)
        er, eg = int((want != info).any(1).sum()), int((got != info).any(1).sum())
        res["curves"]["%s_L%d" % (name, L)][ie].update({
            "sample": sample, "err_ref_sample": er, "err_gpu_same_sample": eg, "ci95_ref_sample": wilson(er, sample),
            "codewords_decoded_differently": int((got != want).any(1).sum())})
res["readme_figure_readoffs"] = {"L1_2.0dB": 5e-2, "L32_crc0_2.0dB": 1.5e-3, "L32_crc16_2.0dB": 3e-5}
print(json.dumps(res, indent=1))
