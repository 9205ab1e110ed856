"""Stall mix per SASS region (bucket of instructions). usage: python tools/ncu_regions.py rep [bucket]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 250
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
names = ["stall_wait", "stall_long_sb", "stall_short_sb", "stall_no_inst", "stall_not_selected", "stall_selected", "stall_math", "stall_mio", "stall_branch_resolving", "stall_lg", "stall_barrier", "stall_dispatch"]
idx = [hdr.index(n) for n in names]
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
print("instr", tot, "samples", ts)
print("region        inst%  samp%  cpi* | " + " ".join("%7s" % n.replace("stall_", "")[:7] for n in names))
for b in range(0, len(data), bucket):
    seg = data[b:b + bucket]
    i = sum(int(r[ia]) for r in seg); s = sum(int(r[isamp]) for r in seg)
    if s == 0: continue
    v = [sum(int(r[k]) for r in seg) for k in idx]
    print("%5d-%5d  %6.2f %6.2f %5.2f | " % (b, b + len(seg), 100 * i / tot, 100 * s / ts, (s / ts) / max(i / tot, 1e-9)) + " ".join("%7.1f" % (100 * x / s) for x in v))
v = [sum(int(r[k]) for r in data) for k in idx]
print("all                              | " + " ".join("%7.1f" % (100 * x / ts) for x in v))
