"""SASS index ranges -> source lines (needs -lineinfo). usage: python tools/ncu_map.py rep [bucket]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 250
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# rows: file header lines, then per source line rows followed by its SASS rows (Address column filled)
hdr = None; cur = None; cur_file = None; sass = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; iaddr = hdr.index("Address"); continue
    if hdr and len(r) == len(hdr):
        if r[0].strip(): cur = (cur_file, int(r[0]))
        if r[iaddr].strip().startswith("0x") and cur: sass[int(r[iaddr], 16)] = cur
addrs = sorted(sass)
for b in range(0, len(addrs), bucket):
    c = collections.Counter(sass[a] for a in addrs[b:b + bucket])
    print("%5d-%5d " % (b, b + bucket) + ", ".join("%s:%d x%d" % (f, l, n) for (f, l), n in c.most_common(6)))
