"""Per CUDA source line: share of a given stall reason's samples. usage: ncu_stall_lines.py rep stall_long_sb [ntop]"""
import csv, io, subprocess, sys, collections
rep, stall = sys.argv[1], sys.argv[2]; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict(); cur_file = None; hdr = None; cur_key = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ist = hdr.index(stall); continue
    if hdr and len(r) == len(hdr):
        if r[0].strip():
            cur_key = (cur_file, int(r[0]), r[1].strip()[:100]); agg.setdefault(cur_key, [0, 0, 0])
        if r[2].strip() and cur_key:
            try:
                agg[cur_key][0] += int(r[ia]); agg[cur_key][1] += int(r[isamp]); agg[cur_key][2] += int(r[ist])
            except ValueError: pass
tots = sum(v[1] for v in agg.values()); tst = sum(v[2] for v in agg.values())
print("samples", tots, stall, tst, "= %.1f%% of all samples" % (100 * tst / max(tots, 1)))
for (f, ln, src), (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:ntop]:
    print("%-14s %4d %-100s %6.2f%% of %s" % (f, ln, src, 100 * k / max(tst, 1), stall))
