"""Builds ``liborx.so`` (the C-ABI rollout engine of include/orx.h) in-tree with nvcc for sm_100a.

``python -m oponn_b200.build`` or ``oponn_b200.build.build()``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
LIB = os.path.join(HERE, "liborx.so")
SOURCES = [os.path.join(CSRC, "orx_api.cu"), os.path.join(CSRC, "orx_tree.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("orx_device.cuh", "orx_game.cuh")] + \
    [os.path.join(INCLUDE, f) for f in ("orx.h", "orx_state.h", "orx_tree.h")] + [os.path.join(HERE, "ptx_uniform.py")]

#: resident blocks per SM that ptxas budgets registers for (9 x 128 threads -> 56 registers).  The source carries a
#: loose bound (ORX_MIN_BLOCKS 5) for the front end; see ptx_uniform.transform.
PTXAS_MIN_BLOCKS = int(os.environ.get("ORX_PTXAS_MIN_BLOCKS", "9"))

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: liborx.so cannot be built (there is no CPU build of this engine)")
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, uniform: bool = True) -> str:
    """nvcc -> liborx.so.  With ``uniform`` (default) the build runs nvcc's own steps (taken from ``--dryrun``) and
    applies :mod:`oponn_b200.ptx_uniform` to the PTX between cicc and ptxas: provably warp-uniform conditional
    branches of the rollout kernels become ``bra.uni`` (see that module for why)."""
    if not (force or is_stale()):
        return LIB
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    if os.environ.get("ORX_PAD"):                   # tuning knob: code alignment of the rollout kernels
        flags = flags + [f"-DORX_PAD={int(os.environ['ORX_PAD'])}"]
    if os.environ.get("ORX_NVVM_MIN_BLOCKS"):      # tuning knob: the loose bound the front end sees
        flags = flags + [f"-DORX_MIN_BLOCKS={int(os.environ['ORX_NVVM_MIN_BLOCKS'])}"]
    if not uniform:
        cmd = [nvcc()] + flags + ["-o", LIB] + SOURCES
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return LIB
    keep = os.path.join(HERE, "build")
    os.makedirs(keep, exist_ok=True)
    cmd = [nvcc()] + flags + ["--dryrun", "--keep", "--keep-dir", keep, "-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.join(HERE, ".."))
    if r.returncode != 0:
        raise RuntimeError("nvcc --dryrun failed:\n" + r.stdout + r.stderr)
    steps = [ln[3:] for ln in r.stderr.splitlines() if ln.startswith("#$ ")]
    script = ["set -e"]
    import re
    for st in steps:
        if re.match(r"^[A-Za-z_][A-Za-z0-9_]*=", st):
            if st.startswith("CICC_PATH="):
                script.append(st)
            continue                      # nvcc's own variable dump: not needed to replay the commands
        if st.startswith("rm "):
            st = "rm -f " + st[3:]
        script.append(st)
        if "/cicc" in st.split(" ")[0] or st.lstrip().startswith('"$CICC_PATH/cicc"'):
            ptx = st.rsplit('-o "', 1)[1].rstrip('" ')
            script.append(f'"{sys.executable}" "{os.path.join(HERE, "ptx_uniform.py")}" "{ptx}" "{ptx}" {PTXAS_MIN_BLOCKS}')
    sh = os.path.join(keep, "build.sh")
    open(sh, "w").write("\n".join(script) + "\n")
    r = subprocess.run(["bash", sh], capture_output=True, text=True, cwd=os.path.join(HERE, ".."))
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("liborx.so build failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, uniform="--no-uniform" not in sys.argv))
