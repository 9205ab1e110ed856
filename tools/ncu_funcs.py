"""SASS size and dynamic instruction share per device function (needs -lineinfo). usage: ncu_funcs.py rep"""
import csv, io, subprocess, collections, re, sys, os
rep=sys.argv[1]
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict(); cur_file=None; hdr=None; cur_key=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur_file=r[1].split("/")[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; ia=hdr.index("Instructions Executed"); isamp=hdr.index("# Samples"); continue
    if hdr and len(r)==len(hdr):
        if r[0].strip(): cur_key=(cur_file,int(r[0])); agg.setdefault(cur_key,[0,0,0])
        if r[2].strip() and cur_key:
            try: agg[cur_key][0]+=int(r[ia]); agg[cur_key][1]+=1; agg[cur_key][2]+=int(r[isamp])
            except ValueError: pass
def func_table(path):
    src=open(path).read().splitlines(); funcs=[]
    for i,l in enumerate(src,1):
        m=re.match(r'\s*(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__).*?\b(\w+)\(', l)
        if m: funcs.append((i,m.group(1)))
    return funcs
tables={"scl_fast.cuh":func_table(os.path.join(ROOT,"polar_b200/csrc/scl_fast.cuh")),"polar_b200.cu":func_table(os.path.join(ROOT,"polar_b200/csrc/polar_b200.cu"))}
def fn(f,line):
    name='?'
    for (s,n) in tables.get(f,[]):
        if line>=s: name=n
        else: break
    return name
bysass=collections.Counter(); byinst=collections.Counter(); bysamp=collections.Counter()
for (f,ln),(n,k,sm) in agg.items():
    key = fn(f,ln) if f in tables else f
    bysass[key]+=k; byinst[key]+=n; bysamp[key]+=sm
tot=sum(byinst.values()); ts=sum(bysass.values()); tsm=sum(bysamp.values())
print("%-28s %6s %7s %7s %7s"%("function","sass","sass%","inst%","stall%"))
for k,v in byinst.most_common(): print("%-28s %6d %6.1f%% %6.1f%% %6.1f%%"%(k,bysass[k],100*bysass[k]/ts,100*v/tot,100*bysamp[k]/tsm))
print("total sass",ts)
