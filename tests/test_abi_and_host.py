"""The C-ABI library loads and exports what include/orx.h declares; host-side mirrors of the record layout.
No compute calls here (there is no GPU on the CPU test box)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oponn_b200 import build as orx_build
from oponn_b200 import engine
from oponn_b200.maps import classic42, synth200
from oponn_b200.state import (CARD_WILD, GAME_RESULT_DTYPE, LEAF_RESULT_DTYPE, PHASE_ATTACK, ATTACK_EXECUTE, BoardState, Rules,
                              leaf_layout, pack_leaves)


@pytest.fixture(scope="module")
def lib():
    orx_build.build()
    return engine.load_library()


def header_functions(name="orx.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"static inline[^{]*\{[^}]*\}", "", src)
    return sorted(set(re.findall(r"\b(orx_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from oponn_b200 import tree
    assert header_functions() == sorted(engine.EXPORTS)
    assert header_functions("orx_tree.h") == sorted(tree.TREE_EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions() + header_functions("orx_tree.h"):
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", engine.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (orx_[a-z_]+)", out))
    assert set(header_functions()) | set(header_functions("orx_tree.h")) <= exported
    # plain C ABI: no C++ mangled orx entry points leak out as the interface
    assert not [s for s in exported if s.startswith("_Z")]


def build_c_demo(out_dir) -> str:
    """examples/orx_demo.c compiled as C against include/ and linked with liborx.so"""
    exe = os.path.join(str(out_dir), "orx_demo")
    libdir = os.path.dirname(engine.lib_path())
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "orx_demo.c"), "-o", exe, "-L", libdir, "-lorx", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_are_plain_c_and_a_c_program_links(lib, tmp_path):
    """the boundary is a C ABI: the headers compile as C11 and a C caller links against the shared library
    (orx_tree.h too: compiled here through a one-line translation unit)"""
    exe = build_c_demo(tmp_path)
    assert os.path.exists(exe)
    tu = tmp_path / "tree_h.c"
    tu.write_text('#include "orx_tree.h"\nint main(void) { OrxTreeOptions o; orx_tree_default_options(&o); return o.leaves_in_flight < 1; }\n')
    libdir = os.path.dirname(engine.lib_path())
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(tu), "-o",
                        str(tmp_path / "tree_h"), "-L", libdir, "-lorx", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not __import__("torch").cuda.is_available():   # without a GPU the C caller gets the loud failure too
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stderr
    # orx_tree_default_options is host-only: the tree half of the library runs anywhere
    assert subprocess.run([str(tmp_path / "tree_h")]).returncode == 0


def test_library_is_cuda_not_a_cpu_build():
    """the product contains sm_100a device code and no oracle symbols"""
    out = subprocess.run(["cuobjdump", "-lelf", engine.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    syms = subprocess.run(["nm", "-D", engine.lib_path()], capture_output=True, text=True).stdout
    assert "orc_" not in syms


def test_engine_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.OrxError) as e:
        engine.RolloutEngine(classic42(6), Rules())
    assert e.value.code == engine.ORX_E_CUDA
    assert "no CPU path" in str(e.value)


def test_missing_library_is_an_import_error(monkeypatch, tmp_path):
    monkeypatch.setenv("ORX_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(engine, "_lib", None)
    with pytest.raises(ImportError):
        engine.load_library()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "oponn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "oponn_oracle" not in text and "liboponn_oracle" not in text, f


def test_struct_sizes_match_the_header():
    assert GAME_RESULT_DTYPE.itemsize == 48 and LEAF_RESULT_DTYPE.itemsize == 160
    # compile a probe against the real header
    probe = r'''
    #include <stdio.h>
    #include "orx.h"
    int main(void) { OrxLeafLayout l = orx_leaf_layout(42, 6), m = orx_leaf_layout(200, 8);
      printf("%zu %zu %zu %d %d %d %d %d %d\n", sizeof(OrxLeafHeader), sizeof(OrxGameResult), sizeof(OrxLeafResult),
             l.off_owner, l.off_armies, l.off_ncards, l.off_cards, l.stride, m.stride); return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "p.c"), "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "p"), os.path.join(td, "p.c")])
        vals = [int(x) for x in subprocess.check_output([os.path.join(td, "p")]).split()]
    L, M = leaf_layout(42, 6), leaf_layout(200, 8)
    assert vals == [32, 48, 160, L.off_owner, L.off_armies, L.off_ncards, L.off_cards, L.stride, M.stride]


def test_leaf_pack_round_trip():
    m = classic42(6)
    rng = np.random.default_rng(0)
    bs = BoardState(owner=[int(x) for x in rng.integers(0, 6, 42)], armies=[int(x) for x in rng.integers(1, 99, 42)],
                    currentPlayer=3, currentPhase=PHASE_ATTACK, turnCount=17, playerNumberOfCards=[1, 2, 3, 0, 4, 5],
                    numberOfPlaceableArmies=12, hasInvadedACountry=True, attackState=ATTACK_EXECUTE, attackingCC=5, defendingCC=9,
                    defendingPlayer=2, attackUntilDead=False, cardProgressionPos=7, hand=[3, CARD_WILD, 40])
    buf = bs.pack(m)
    back = BoardState.unpack(m, buf)
    assert back.pack(m).tobytes() == buf.tobytes()
    assert back.hand == [3, CARD_WILD, 40] and len(back.deck) == 41 and back.deck.count(CARD_WILD) == 1
    assert pack_leaves(m, [bs, back]).shape == (2, leaf_layout(42, 6).stride)


def test_maps_are_well_formed():
    for m in (classic42(6), synth200()):
        n = m.n_territories
        for t, nb in enumerate(m.adjacency):
            assert len(set(nb)) == len(nb)
            for x in nb:
                assert t in m.adjacency[x]           # undirected
        # connected
        seen, todo = {0}, [0]
        while todo:
            for x in m.adjacency[todo.pop()]:
                if x not in seen:
                    seen.add(x); todo.append(x)
        assert len(seen) == n
    c = classic42(6)
    assert (c.n_territories, c.n_continents) == (42, 6)
    assert [c.continent.count(i) for i in range(6)] == [9, 4, 7, 6, 12, 4]
    assert sum(len(a) for a in c.adjacency) // 2 == 83
    s = synth200()
    assert (s.n_territories, s.n_continents, s.n_players) == (200, 10, 8)


def test_oponn_txt_settings(tmp_path):
    """Oponn.readSettings (Oponn.java:464-533): same keys, same parsing quirks"""
    from oponn_b200.settings import read_settings, search_options
    p = tmp_path / "Oponn.txt"
    p.write_text("# comment\nlux_agent_modelling=false\niterations = 5 000\ncores=n\ndebug=true\nnonsense\n")
    s = read_settings(str(p), available_cores=16)
    assert (s.opponent_modelling, s.iterations, s.cores, s.debugging) == (False, 5000, 16, True)
    p.write_text("lux_agent_modelling=no\niterations=many\ncores=4\n")
    s = read_settings(str(p), available_cores=16)
    assert (s.opponent_modelling, s.iterations, s.cores, s.debugging) == (True, 1000, 4, False)
    assert search_options(s) == {"iterations": 250, "playouts_per_leaf": 4, "modelling_turn_limit": 2}
    assert read_settings(str(tmp_path / "missing.txt"), available_cores=8).iterations == 1000
    if os.path.exists("/root/reference/Oponn.txt"):           # the reference's own example file (not present on the GPU box)
        s = read_settings("/root/reference/Oponn.txt", available_cores=8)
        assert (s.opponent_modelling, s.iterations, s.cores, s.debugging) == (True, 1000, 8, False)
