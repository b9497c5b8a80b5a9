"""Builds variant libraries oponn_b200/variants/liborx_<name>.so in parallel for on-GPU A/B timing (scripts/run_variants.sh).
usage: python scripts/build_variants.py name:DEFINES:ptxas_min_blocks[:nvvm_min_blocks[:pad]] ...   (DEFINES comma separated)"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
os.makedirs(os.path.join(ROOT, "oponn_b200", "variants"), exist_ok=True)
def one(spec):
    f = spec.split(":") + ["", "", "", ""]
    name, defs, mb, nv, pad = f[0], f[1].replace(",", " "), f[2] or "9", f[3], f[4]
    env = dict(os.environ, ORX_BUILD_OUT=os.path.join(ROOT, "oponn_b200", "variants", f"liborx_{name}.so"), ORX_DEFINES=defs, ORX_PTXAS_MIN_BLOCKS=mb)
    if nv: env["ORX_NVVM_MIN_BLOCKS"] = nv
    if pad: env["ORX_PAD"] = pad
    r = subprocess.run([sys.executable, "-m", "oponn_b200.build", "--force", "-v"], cwd=ROOT, env=env, capture_output=True, text=True)
    info = [l for l in r.stderr.splitlines() if "rollout_kernelILi2ELb0ELb0E" in l or "rollout_lanes" in l]
    regs = ""
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry function '_ZN3orx14rollout_kernelILi2ELb0ELb0E" in l:
            regs = " | ".join(x.strip() for x in lines[i + 2:i + 4])
    return f"{name}: rc={r.returncode} {regs}"
with ThreadPoolExecutor(8) as ex:
    for line in ex.map(one, sys.argv[1:]): print(line, flush=True)
