"""Top source lines of an ncu report by instructions executed / stall samples.
usage: python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] != "":
        d = dict(zip(hdr, r))
        try:
            ie = int(d["Instructions Executed"]); sm = int(d["# Samples"])
        except Exception: continue
        lines.append((cur_file, int(r[0]), ie, sm, r[1].strip()[:110]))
tot_i = sum(l[2] for l in lines); tot_s = sum(l[3] for l in lines)
print(f"total inst {tot_i:.3e} samples {tot_s}")
print("--- by instructions executed")
for f, ln, ie, sm, src in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{100*ie/tot_i:5.1f}% i {100*sm/max(tot_s,1):5.1f}% s  {f}:{ln}  {src}")
print("--- by samples")
for f, ln, ie, sm, src in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{100*ie/tot_i:5.1f}% i {100*sm/max(tot_s,1):5.1f}% s  {f}:{ln}  {src}")
