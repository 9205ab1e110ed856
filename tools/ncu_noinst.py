"""Where the instruction-fetch stalls are: stall_no_inst samples per SASS region and the top instructions.
usage: python tools/ncu_noinst.py rep [bucket]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 200
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ia, isamp, ino, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_no_inst"), hdr.index("Source")
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data); tn = sum(int(r[ino]) for r in data)
print("instr", tot, "samples", ts, "no_inst samples", tn, "(%.1f%% of samples)" % (100 * tn / ts))
print("region       inst%  samp%  no_inst% (of all no_inst)  no_inst/samples")
for b in range(0, len(data), bucket):
    seg = data[b:b + bucket]
    i = sum(int(r[ia]) for r in seg); s = sum(int(r[isamp]) for r in seg); n = sum(int(r[ino]) for r in seg)
    print("%5d-%5d  %6.2f %6.2f %6.2f   %5.2f" % (b, b + len(seg), 100 * i / tot, 100 * s / ts, 100 * n / tn, n / max(s, 1)))
top = sorted(range(len(data)), key=lambda i: -int(data[i][ino]))[:30]
print("top instructions by no_inst samples: index, %of no_inst, executions share, sass")
for i in sorted(top):
    r = data[i]; print("%5d %6.2f%% %6.3f%%  %s" % (i, 100 * int(r[ino]) / tn, 100 * int(r[ia]) / tot, r[isrc].strip()[:80]))
