"""Instruction/sample share per named line range. usage: python scripts/ncu_regions.py rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur=None; hdr=None; L=[]
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr and r[0]!="":
        d=dict(zip(hdr,r))
        try: L.append((cur,int(r[0]),int(d["Instructions Executed"]),int(d["# Samples"])))
        except: pass
import re
src={}
for f in ("oponn_b200/csrc/orx_device.cuh","oponn_b200/csrc/orx_game.cuh"):
    src[f.split("/")[-1]]=open(f).read().split("\n")
# region = enclosing function name (crude: last line matching '__device__ ... name(' at or before)
def func_of(f,ln):
    s=src.get(f)
    if not s: return f
    for i in range(min(ln,len(s))-1,-1,-1):
        m=re.search(r'__device__[^;]*?(\w+)\s*\(',s[i])
        if m and '{' not in s[i].split('(')[0]: return f.split('.')[0][4:]+":"+m.group(1)
    return f
agg={}
for f,ln,ie,sm in L:
    k=func_of(f,ln); a=agg.setdefault(k,[0,0]); a[0]+=ie; a[1]+=sm
ti=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][0]): print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/ts:5.1f}% samples  {k}")
