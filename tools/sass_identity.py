"""Compare the device code of two builds kernel by kernel (cuobjdump -sass, addresses and encodings stripped,
anonymous-namespace hashes normalised). Used to prove that a change hidden behind a default-off macro leaves the
default binary untouched when no GPU is at hand.
usage: python tools/sass_identity.py <old .o/.so> <new .o/.so>"""
import hashlib, re, subprocess, sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "ANON", m.group(1))
            out[cur] = []
            continue
        if cur and not re.match(r"^\s*/\* 0x", line) and ".headerflags" not in line:
            out[cur].append(re.sub(r"/\*[0-9a-f]*\*/", "", line))
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in out.items()}


old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
same = [k for k in old if old[k] == new.get(k)]
print("kernels: old %d, new %d, identical %d" % (len(old), len(new), len(same)))
for k in old:
    if k not in new:
        print("  missing in new:", k[-90:])
    elif old[k] != new[k]:
        print("  DIFFERENT:", k[-90:])
for k in new:
    if k not in old:
        print("  only in new:", k[-90:])
sys.exit(0 if len(same) == len(old) else 1)
