"""Per-function instruction totals split by how often the SASS instruction runs per playout. usage: rep playouts"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]; per=int(sys.argv[2])
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
cur=None; hdr=None; curline=None; lo={}; hi={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; ie=hdr.index("Instructions Executed"); continue
    if hdr is None: continue
    if r[0]!="": curline=(cur,int(r[0])); continue
    if curline and len(r)>2 and r[2].startswith("0x"):
        try: n=int(r[ie])/per
        except: continue
        k=func_of(*curline)
        (lo if n<5000 else hi)[k]=(lo if n<5000 else hi).get(k,0)+n
print("low-frequency (<5000 runs/playout) total", int(sum(lo.values())), " high", int(sum(hi.values())))
for k,v in sorted(lo.items(), key=lambda x:-x[1])[:25]: print(f"{int(v):8d}  {k}")
