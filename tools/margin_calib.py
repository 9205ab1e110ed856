"""Calibrate STRICT mode's margin threshold (developer tool, needs a GPU): for every setting decode the same float LLRs
with the fp32 kernels (recording every codeword's smallest decision margin) and with the double-precision CPU reference,
then print the margins of the codewords that differ, the share of codewords each threshold would flag, and the cost /
result of STRICT mode itself. Usage: python tools/margin_calib.py [scale] [out.json]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle_lib import Port, Ref, awgn_llrs, have_ref
from polar_b200 import PolarCode, unpack_bits

settings = [(11, 1024, 16, 32, 1.0, 16000), (11, 1024, 16, 4, 1.0, 32000), (11, 1024, 0, 1, 1.0, 64000),
            (9, 256, 16, 32, 1.0, 32000), (11, 1024, 16, 32, 1.5, 8000), (11, 1024, 0, 32, 1.0, 8000), (9, 256, 0, 8, 1.0, 32000)]
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
TAUS = [1e-7, 3e-7, 1e-6, 3e-6, 1e-5, 3e-5, 1e-4, 3e-4, 1e-3, 3e-3, 1e-2]
rows = []
for (n, K, crc, L, eb, B) in settings:
    B = max(256, int(B * scale))
    cpu = (Ref if have_ref() else Port)(n, K, 0.32, crc)
    pc = PolarCode(n, K, 0.32, crc)
    info, llr = awgn_llrs(cpu, B, eb, 4242 + L + int(eb * 100) + n)
    t = time.time(); want = cpu.decode_batch(llr, L, nthreads=os.cpu_count()); tc = time.time() - t
    d_llr = torch.from_numpy(llr).cuda()
    margin = torch.empty(B, dtype=torch.float32, device="cuda")
    got = unpack_bits(pc.decode_device(d_llr, L, mode="fp32", margin=margin).cpu().numpy().view(np.uint32), K)
    m = margin.cpu().numpy()
    mm = (got != want).any(1)
    strict = unpack_bits(pc.decode_device(d_llr, L, mode="strict").cpu().numpy().view(np.uint32), K)
    nflag = pc.last_flagged
    mm_strict = (strict != want).any(1)
    host_strict = pc.decode_batch(llr, L, mode="strict")
    mm_host = (host_strict != want).any(1)
    dbl = pc.decode_batch_double(llr.astype(np.float64), L, mode="strict")
    mm_dbl = (dbl != want).any(1)
    # timing at a full batch (the LLRs tiled)
    reps = max(1, 65536 // B)
    big = d_llr.repeat(reps, 1).contiguous()
    out = torch.empty((big.shape[0], pc.KW), dtype=torch.int32, device="cuda")
    tim, flagged_at = {}, {}
    for mode, tau in (("fp32", None), ("strict", 1e-6), ("strict", 3e-6), ("strict", 1e-5), ("strict", 3e-5)):
        if tau is not None:
            pc.set_strict_tau(tau)
        for _ in range(2):
            pc.decode_device(big, L, out=out, mode=mode)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            pc.decode_device(big, L, out=out, mode=mode)
        e1.record(); torch.cuda.synchronize()
        key = mode if tau is None else "strict@%g" % tau
        tim[key] = big.shape[0] * 3 / (e0.elapsed_time(e1) * 1e-3)
        if tau is not None:
            flagged_at[key] = pc.last_flagged
    flagged_big = flagged_at
    row = dict(n=n, K=K, crc=crc, L=L, ebno=eb, B=B, cpu=type(cpu).__name__, cpu_cw_per_s=B / tc,
               mismatch_fp32=int(mm.sum()), margins_of_mismatches=sorted(float(x) for x in m[mm]),
               mismatch_strict_device=int(mm_strict.sum()), mismatch_strict_host=int(mm_host.sum()),
               mismatch_strict_double_entry=int(mm_dbl.sum()), flagged_at_default_tau=int(nflag),
               flag_share={"%g" % t: float((m < t).mean()) for t in TAUS},
               margin_quantiles={q: float(np.quantile(m[np.isfinite(m)], q)) for q in (0.001, 0.01, 0.1, 0.5)} if np.isfinite(m).any() else {},
               cw_per_s=tim, flagged_in_timing_batch=flagged_big,
               timing_batch=int(big.shape[0]), block_errors_ref=int((want != info).any(1).sum()))
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 2:
    json.dump(rows, open(sys.argv[2], "w"), indent=1)
