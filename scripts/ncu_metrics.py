# usage: python scripts/ncu_metrics.py <report.ncu-rep> <like.json> <out.json> [note]
# Dumps from the raw page of an ncu report the same counters an earlier summary (like.json) holds.
import csv, io, json, subprocess, sys
rep, like, out = sys.argv[1:4]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, vals = rows[0], rows[1], rows[2]
got = {h: (v, u) for h, u, v in zip(head, units, vals)}
keys = [k for k in json.load(open(like)) if not k.startswith("_")]
d = {"_note": note, "_units": {}}
for k in keys:
    if k in got:
        v, u = got[k]
        d["_units"][k] = u
        try:
            d[k] = float(v.replace(",", ""))
        except ValueError:
            d[k] = v
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
