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
LIB = os.environ.get("ORX_BUILD_OUT") or os.path.join(HERE, "liborx.so")   # ORX_BUILD_OUT: variant builds
SOURCES = [os.path.join(CSRC, "orx_api.cu"), os.path.join(CSRC, "orx_tree.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("orx_device.cuh", "orx_game.cuh", "orx_lanes.cuh", "orx_devmap.h", "orx_mapimage.h")] + \
    [os.path.join(INCLUDE, f) for f in ("orx.h", "orx_state.h", "orx_tree.h")] + [os.path.join(HERE, "ptx_uniform.py")]

#: resident blocks per SM that ptxas budgets registers for (11 x 128 threads -> 40 registers; round 2 sweep).  The source carries a
#: loose bound (ORX_MIN_BLOCKS 4) for the front end; see ptx_uniform.transform.
PTXAS_MIN_BLOCKS = int(os.environ.get("ORX_PTXAS_MIN_BLOCKS", "11"))

#: the same for the other instantiations (replay, modelled opponents, maps above 64 territories)
PTXAS_MIN_BLOCKS_OTHER = int(os.environ.get("ORX_PTXAS_MIN_BLOCKS_OTHER", "9"))

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: liborx.so cannot be built (there is no CPU build of this engine)")
    return exe


def source_hash() -> str:
    """content hash of everything liborx.so is built from (sources, headers, the PTX pass, the build knobs)"""
    import hashlib
    h = hashlib.sha256()
    for d in sorted(DEPS + [os.path.abspath(__file__)]):
        h.update(os.path.basename(d).encode()); h.update(open(d, "rb").read())
    for k in ("ORX_PTXAS_MIN_BLOCKS", "ORX_PTXAS_MIN_BLOCKS_OTHER", "ORX_DEFINES", "ORX_PAD", "ORX_NVVM_MIN_BLOCKS"):
        h.update(f"{k}={os.environ.get(k, '')};".encode())
    return h.hexdigest()


def is_stale() -> bool:
    """the library is rebuilt when the hash of its sources differs from the one recorded at its last build (mtimes
    are arbitrary after a checkout, and the prebuilt .so travels to the GPU box)"""
    stamp = LIB + ".srchash"
    if not os.path.exists(LIB) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False, uniform: bool = True) -> str:
    """nvcc -> liborx.so.  With ``uniform`` (default) the build runs nvcc's own steps (taken from ``--dryrun``) and
    applies :mod:`oponn_b200.ptx_uniform` to the PTX between cicc and ptxas: provably warp-uniform conditional
    branches of the rollout kernels become ``bra.uni`` (see that module for why)."""
    if not (force or is_stale()):
        return LIB
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    for d in os.environ.get("ORX_DEFINES", "").split():   # tuning knob: extra -D flags (variant builds, scripts/gpu_variants.sh)
        flags = flags + ["-D" + d]
    pad = int(os.environ.get("ORX_PAD", "2"))        # code alignment of the rollout kernels (fetch alignment: +-1 %)
    if pad:
        flags = flags + [f"-DORX_PAD={pad}"]
    if os.environ.get("ORX_NVVM_MIN_BLOCKS"):      # tuning knob: the loose bound the front end sees
        flags = flags + [f"-DORX_MIN_BLOCKS={int(os.environ['ORX_NVVM_MIN_BLOCKS'])}"]
    if not uniform:
        cmd = [nvcc()] + flags + ["-o", LIB] + SOURCES
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        open(LIB + ".srchash", "w").write(source_hash())
        return LIB
    # nvcc's intermediates (preprocessed CUDA headers, cudafe output) never live inside the tree: a scratch directory
    # outside the repository, removed when the build ends
    import tempfile
    keep = tempfile.mkdtemp(prefix="orx_build_")
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
            script.append(f'"{sys.executable}" "{os.path.join(HERE, "ptx_uniform.py")}" "{ptx}" "{ptx}" {PTXAS_MIN_BLOCKS} {PTXAS_MIN_BLOCKS_OTHER} "{LIB}.ptxpass.json"')
    sh = os.path.join(keep, "build.sh")
    open(sh, "w").write("\n".join(script) + "\n")
    r = subprocess.run(["bash", sh], capture_output=True, text=True, cwd=os.path.join(HERE, ".."))
    if verbose:
        sys.stderr.write(r.stderr)
    shutil.rmtree(keep, ignore_errors=True)
    if r.returncode != 0:
        raise RuntimeError("liborx.so build failed:\n" + r.stdout + r.stderr)
    open(LIB + ".srchash", "w").write(source_hash())
    return LIB


#: the same sources compiled WITHOUT the PTX pass: the check library of tests/test_gpu_ptx_guard.py (never loaded by the product)
LIB_NOUNIFORM = os.path.join(HERE, "liborx_nouniform.so")


def build_nouniform(force: bool = False) -> str:
    """plain nvcc build (no bra.uni rewriting, source launch bounds): its records must equal the shipped library's"""
    global LIB
    stamp = LIB_NOUNIFORM + ".srchash"
    if not force and os.path.exists(LIB_NOUNIFORM) and os.path.exists(stamp) and open(stamp).read().strip() == source_hash():
        return LIB_NOUNIFORM
    keep = LIB
    try:
        LIB = LIB_NOUNIFORM
        build(force=True, uniform=False)
    finally:
        LIB = keep
    return LIB_NOUNIFORM


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, uniform="--no-uniform" not in sys.argv))
