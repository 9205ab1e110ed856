"""Where does fp32 rounding actually change a decoded word? (developer tool, needs a GPU.) Decodes millions of codewords at
1.0 dB with the fp32 kernels (recording decision margins) and with the GPU's double-precision mode -- which equals the
CPU reference on every comparison made so far -- and prints the margins of the codewords that differ, next to the share
of codewords each threshold tau would flag. This is the evidence behind POLAR_B200_DEFAULT_STRICT_TAU.
usage: python tools/flip_margins.py [scale] [out.json] [more | sc]   (more: twelve other codes / lists; sc: list size 1 on sc_ssc.cuh)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from polar_b200 import PolarCode

settings = [(11, 1024, 16, 32, 1.0, 2 ** 21), (11, 1024, 16, 4, 1.0, 2 ** 22), (11, 1024, 0, 1, 1.0, 2 ** 22),
            (9, 256, 16, 32, 1.0, 2 ** 22), (11, 1024, 16, 32, 2.0, 2 ** 20)]
if len(sys.argv) > 3 and sys.argv[3] == "more":     # other block lengths, list sizes and codes than BASELINE.json's
    settings = [(9, 256, 0, 32, 1.0, 2 ** 21), (10, 512, 16, 32, 1.0, 2 ** 20), (12, 2048, 16, 32, 1.0, 2 ** 19),
                (13, 2048, 16, 32, 1.0, 2 ** 17), (8, 128, 8, 32, 1.0, 2 ** 21), (11, 1024, 0, 32, 1.0, 2 ** 20),
                (11, 1024, 16, 2, 1.0, 2 ** 21), (11, 1024, 16, 8, 1.0, 2 ** 20), (11, 1024, 16, 16, 1.0, 2 ** 20),
                (11, 1024, 16, 24, 1.0, 2 ** 19), (11, 512, 16, 32, 0.0, 2 ** 19), (11, 1536, 16, 32, 3.0, 2 ** 19)]
SC = len(sys.argv) > 3 and sys.argv[3] == "sc"
if SC:      # list size 1 on the pruned tree (sc_ssc.cuh): the first arm is that kernel without a second pass (tau -> 0)
    settings = [(11, 1024, 0, 1, 1.0, 2 ** 23), (11, 1024, 16, 1, 2.0, 2 ** 23), (9, 256, 0, 1, 1.0, 2 ** 24), (10, 300, 8, 1, 0.0, 2 ** 23),
                (12, 2048, 16, 1, 1.5, 2 ** 22), (8, 128, 8, 1, 1.0, 2 ** 24), (11, 1536, 16, 1, 3.0, 2 ** 23), (11, 512, 16, 1, 0.0, 2 ** 23)]
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
settings = settings[int(os.environ.get("FLIP_SKIP", "0")):]
CH = 65536
rows = []
for (n, K, crc, L, eb, total) in settings:
    total = max(CH, int(total * scale) // CH * CH)
    pc = PolarCode(n, K, 0.32, crc)
    margin = torch.empty(CH, dtype=torch.float32, device="cuda")
    flips, hist, nflip_err, strict_differ, strict_flagged = [], np.zeros(8, np.int64), 0, 0, 0
    taus = [1e-7, 3e-7, 1e-6, 3e-6, 1e-5, 3e-5, 1e-4, 3e-4]
    t0 = time.time()
    for first in range(0, total, CH):
        llr, truth = pc.synthesize(CH, [eb], seed=991, first_index=first)
        if SC:
            pc.set_strict_tau(1e-30)
            o32 = pc.decode_device(llr, L, mode="strict", margin=margin).clone()
            assert pc.info(6) == 500
            pc.set_strict_tau(1e-5)
        else:
            o32 = pc.decode_device(llr, L, mode="fp32", margin=margin)
        o64 = pc.decode_device(llr, L, mode="f64")
        ost = pc.decode_device(llr, L, mode="strict")
        strict_flagged += pc.last_flagged
        strict_differ += int((ost != o64).any(dim=1).sum().item())
        diff = (o32 != o64).any(dim=1)
        m = margin.clone()
        for i, t in enumerate(taus):
            hist[i] += int((m < t).sum().item())
        if diff.any():
            flips += m[diff].cpu().tolist()
            nflip_err += int(((o32 != truth).any(dim=1) & (o64 != truth).any(dim=1) & diff).sum().item())
    row = dict(n=n, K=K, crc=crc, L=L, ebno=eb, codewords=total, differ=len(flips), margins_of_differing=sorted(flips),
               differing_that_are_block_errors_in_both=nflip_err, strict_mode_differ=strict_differ,
               strict_mode_second_pass_share=strict_flagged / total, flag_share={"%g" % t: hist[i] / total for i, t in enumerate(taus)},
               seconds=time.time() - t0)
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 2:
    json.dump(rows, open(sys.argv[2], "w"), indent=1)
