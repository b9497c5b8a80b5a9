"""Hot-SASS summary of an ncu report (SASS page): how many instructions / bytes of code carry 50 / 90 / 99 % of the executed warp
instructions, the opcode mix, and the contiguous hot regions.  usage: python scripts/ncu_sass_hot.py rep.ncu-rep > profiles/..._sass_hot.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][0], rows[0][1])
hdr = rows[1]; ia = hdr.index("Address"); isrc = hdr.index("Source"); ie = hdr.index("Instructions Executed"); ism = hdr.index("# Samples")
ins = []
for r in rows[2:]:
    if len(r) <= ie or not r[ia].startswith("0x"): continue
    ins.append((int(r[ia], 16), r[isrc].strip(), int(r[ie]), int(r[ism])))
base = ins[0][0]; tot = sum(i[2] for i in ins); n = len(ins)
print(f"static SASS instructions {n} = {16 * n / 1024:.1f} KB; executed warp instructions in this capture {tot:.4e}")
order = sorted(ins, key=lambda x: -x[2]); acc = 0; marks = {0.5: None, 0.9: None, 0.99: None}
for k, i in enumerate(order, 1):
    acc += i[2]
    for m in marks:
        if marks[m] is None and acc >= m * tot: marks[m] = k
for m, k in marks.items(): print(f"  {int(m * 100)} % of executed instructions come from the {k} hottest SASS instructions ({16 * k / 1024:.1f} KB)")
thr = order[marks[0.9] - 1][2]
# contiguous hot regions (instructions at or above the 90 % threshold, gaps of up to 8 instructions bridged)
regions = []; cur = None
for idx, i in enumerate(ins):
    if i[2] >= thr:
        if cur and idx - cur[1] <= 8: cur[1] = idx
        else:
            cur = [idx, idx]; regions.append(cur)
print(f"hot regions (>= {thr} executions, gaps <= 8 bridged): {len(regions)} regions")
for a, b in regions:
    e = sum(i[2] for i in ins[a:b + 1])
    print(f"  +0x{ins[a][0] - base:05x} .. +0x{ins[b][0] - base:05x}  {b - a + 1:5d} instr  {100 * e / tot:5.1f} % of executed")
mix = {}
for i in ins:
    op = i[1].split()[0] if not i[1].startswith("@") else i[1].split()[1]
    op = op.split(".")[0]; mix[op] = mix.get(op, 0) + i[2]
print("opcode mix (share of executed warp instructions):")
for op, e in sorted(mix.items(), key=lambda x: -x[1])[:24]: print(f"  {100 * e / tot:5.1f} %  {op}")
print("40 hottest SASS instructions:")
for i in order[:40]: print(f"  {100 * i[2] / tot:5.2f} % exec {i[3]:6d} samples  +0x{i[0] - base:05x}  {i[1]}")
