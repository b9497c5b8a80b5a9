"""SASS instruction counts of the lane-per-game kernel: per out-of-line function, and per inlined source function inside
the kernel body (from nvdisasm line info).  usage: python scripts/sass_sizes.py [object-or-so]  (default: compiles orx_api.cu)"""
import os, re, subprocess, sys, tempfile
from collections import Counter
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tmp = tempfile.mkdtemp()
obj = sys.argv[1] if len(sys.argv) > 1 else None
if obj is None:
    obj = os.path.join(tmp, "orx_api.o")
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-c", "-o", obj,
                           os.path.join(ROOT, "oponn_b200", "csrc", "orx_api.cu")])
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
lines = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text._ZN3orx20rollout_lanes_kernel")][0]
cur_fn, cur = "kernel", None
byfn, byline = Counter(), Counter()
for l in lines[start:]:
    m = re.match(r"^(\$?_ZN[^:]+):$", l.strip())
    if m:
        if "$" in m.group(1): cur_fn = m.group(1).split("$")[-1]
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l): byfn[cur_fn] += 1; byline[(cur_fn,) + (cur or ("?", 0))] += 1
    if l.startswith("//-------") and "rollout_lanes" not in l and byfn: break
src = open(os.path.join(ROOT, "oponn_b200", "csrc", "orx_lanes.cuh")).read().split("\n")
def fn_of(ln):
    for i in range(min(ln, len(src)) - 1, -1, -1):
        m = re.search(r"(?:LN_M|LN_COLD|__device__ __forceinline__|__device__ __noinline__|LN_NOINLINE|__global__) [\w\s\*&<>]*?(\w+)\s*\(", src[i])
        if m and not src[i].strip().startswith("//"): return m.group(1)
    return "?"
print("total", sum(byfn.values()), "SASS instructions =", sum(byfn.values()) * 16 // 1024, "KB")
for k, v in byfn.most_common(): print(f"{v:6d} {k}")
agg = Counter()
for (fn, f, ln), v in byline.items():
    if fn == "kernel": agg[fn_of(ln) if f == "orx_lanes.cuh" else f] += v
print("kernel body by inlined source function:", agg.most_common(24))
