"""Aggregate instructions executed / stall samples per CUDA source line (needs -lineinfo).
usage: python tools/ncu_lines.py rep [ntop]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict(); cur_file = None; hdr = None; cur_key = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr):
        if r[0].strip():
            cur_key = (cur_file, int(r[0]), r[1].strip()[:90]); agg.setdefault(cur_key, [0, 0, 0])
        if r[2].strip() and cur_key:   # a SASS row
            try:
                agg[cur_key][0] += int(r[ia]); agg[cur_key][1] += int(r[isamp]); agg[cur_key][2] += 1
            except ValueError: pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total inst", tot, "samples", tots, "sass", sum(v[2] for v in agg.values()))
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]
for (f, ln, src), (n, s, k) in sorted(top, key=lambda kv: (kv[0][0], kv[0][1])):
    print("%-14s %4d %-90s %6.2f%% inst %6.2f%% samp %4d sass" % (f, ln, src, 100*n/tot, 100*s/max(tots,1), k))
