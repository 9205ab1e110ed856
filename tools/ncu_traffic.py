"""Extract per-launch DRAM traffic of the decode kernel from an ncu --set full capture and record it
(per codeword) in profiles/ncu_traffic.json, which bench.py reports as roofline.traffic.
usage: python tools/ncu_traffic.py <rep> <codewords in the captured launch> <config> <tag>"""
import csv, io, json, os, subprocess, sys
rep, batch, config, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name); v = float(r[i]); u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {}
d[config] = {"dram_bytes_per_codeword": (rd + wr) / batch, "dram_read_bytes": rd, "dram_write_bytes": wr,
             "captured_batch": batch, "kernel": r[hdr.index("Kernel Name")], "source": os.path.basename(rep), "tag": tag}
json.dump(d, open(out, "w"), indent=1, sort_keys=True)
print(config, d[config])
