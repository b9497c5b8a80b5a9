"""Static SASS instruction count per source function (where the code size goes). usage: python scripts/ncu_codesize.py rep"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
src={}
for f in ("oponn_b200/csrc/orx_device.cuh","oponn_b200/csrc/orx_game.cuh"):
    src[f.split("/")[-1]]=open(f).read().split("\n")
def func_of(f,ln):
    s=src.get(f)
    if not s: return f
    for i in range(min(ln,len(s))-1,-1,-1):
        m=re.search(r'__device__[^;]*?(\w+)\s*\(',s[i])
        if m and '{' not in s[i].split('(')[0]: return f.split('.')[0][4:]+":"+m.group(1)
    return f
cur=None; hdr=None; curline=None; cnt={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None: continue
    if r[0]!="": curline=(cur,int(r[0])); continue
    if curline and len(r)>2 and r[2].startswith("0x"):
        k=func_of(*curline); cnt[k]=cnt.get(k,0)+1
tot=sum(cnt.values()); print("total sass", tot)
for k,v in sorted(cnt.items(), key=lambda x:-x[1])[:40]: print(f"{v:6d} {100*v/tot:5.1f}%  {k}")
