"""The PTX warp-uniformity pass (oponn_b200/ptx_uniform.py) on hand-written PTX: it must mark exactly the branches
whose guards are provably the same in every lane, and never one that depends on the lane."""
import os

from conftest import ROOT
from oponn_b200 import ptx_uniform as pu

PTX = r'''
.version 8.8
.target sm_100a
.address_size 64

.visible .entry rollout_kernel_demo(
	.param .u64 p0
)
{
	.reg .pred 	%p<16>;
	.reg .b32 	%r<64>;
	.reg .b64 	%rd<8>;

	ld.param.u64 	%rd1, [p0];
	mov.u32 	%r1, %laneid;
	ld.global.u32 	%r2, [%rd1];
	setp.eq.s32 	%p1, %r2, 0;
	@%p1 bra 	$L_A;
	add.s32 	%r2, %r2, 1;
$L_A:
	setp.lt.s32 	%p2, %r1, 5;
	@%p2 bra 	$L_B;
	mov.u32 	%r3, 7;
	bra.uni 	$L_C;
$L_B:
	mov.u32 	%r3, 9;
$L_C:
	setp.eq.s32 	%p3, %r3, 7;
	@%p3 bra 	$L_D;
	add.s32 	%r2, %r2, 2;
$L_D:
	setp.ne.s32 	%p4, %r1, 0;
	vote.sync.ballot.b32 	%r4, %p4, -1;
	setp.eq.s32 	%p5, %r4, 0;
	@%p5 bra 	$L_E;
	add.s32 	%r2, %r2, 3;
$L_E:
	ld.global.u32 	%r8, [%rd1+4];
	shfl.sync.idx.b32	%r5|%p6, %r1, %r8, 31, -1;
	setp.eq.s32 	%p7, %r5, 3;
	@%p7 bra 	$L_F;
	add.s32 	%r2, %r2, 4;
$L_F:
	shfl.sync.idx.b32	%r6|%p8, %r8, %r1, 31, -1;
	setp.eq.s32 	%p9, %r6, 3;
	@%p9 bra 	$L_G;
	add.s32 	%r2, %r2, 5;
$L_G:
	mov.u32 	%r7, %r8;
	mov.u32 	%r7, %r1;
	mov.u32 	%r7, %r8;
	setp.eq.s32 	%p10, %r7, 3;
	@%p10 bra 	$L_H;
	add.s32 	%r2, %r2, 6;
$L_H:
	st.global.u32 	[%rd1], %r2;
	ret;
}
'''


def marked(text: str):
    out = {}
    for ln in text.split("\n"):
        ln = ln.strip()
        if ln.startswith("@%p") and "bra" in ln:
            out[ln.split()[0][1:]] = "bra.uni" in ln
    return out


def test_marks_only_provably_uniform_branches():
    got = marked(pu.transform(PTX, only="rollout_kernel"))
    assert got["%p1"] is True      # guard from a load at a uniform address
    assert got["%p2"] is False     # guard reads %laneid
    assert got["%p3"] is False     # value defined on both arms of a divergent branch (control dependence)
    assert got["%p5"] is True      # ballot of a lane-dependent predicate is warp-uniform
    assert got["%p7"] is True      # shuffle from a uniform source lane
    assert got["%p9"] is False     # shuffle from a lane-dependent source lane
    assert got["%p10"] is True     # flow sensitive: the last definition reaching the test is uniform


def test_other_functions_are_left_alone():
    assert "bra.uni \t$L_A" not in pu.transform(PTX, only="some_other_kernel")
    assert pu.transform(PTX, only="some_other_kernel").count("@%p1 bra \t$L_A") == 1


def test_min_blocks_rewrite():
    ptx = PTX.replace(")\n{", ")\n.maxntid 128, 1, 1\n.minnctapersm 5\n{", 1)
    out = pu.transform(ptx, only="rollout_kernel", min_blocks=9)
    assert ".minnctapersm 9" in out and ".minnctapersm 5" not in out


def test_built_library_used_the_pass():
    """The build records what the pass did (liborx.so.ptxpass.json).  A parser miss would show as a collapse of these
    counts: the classic all-SimRandom kernel has ~735 conditional branches of which ~490 are provably warp-uniform
    (479 of 718 in round 1), and the library as a whole carries thousands of bra.uni."""
    import json
    import pytest
    from oponn_b200 import build as b
    info = b.LIB + ".ptxpass.json"
    if not os.path.exists(info):
        pytest.skip("liborx.so was not built with the PTX pass on this machine")
    d = json.load(open(info))
    main = [v for k, v in d["kernels"].items() if pu.MAIN_KERNEL in k]
    assert len(main) == 1
    assert 600 <= main[0]["conditional"] <= 900
    assert 0.60 <= main[0]["uniform"] / main[0]["conditional"] <= 0.75
    assert d["bra_uni"] > 1500
    assert len(d["kernels"]) == 8            # every instantiation went through the pass
    assert all(v["uniform"] > 0.5 * v["conditional"] for v in d["kernels"].values())


def test_min_blocks_only_the_main_kernel_gets_the_tight_budget():
    ptx = PTX.replace(")\n{", ")\n.maxntid 128, 1, 1\n.minnctapersm 5\n{", 1)
    out = pu.transform(ptx, only="rollout_kernel", min_blocks=11, min_blocks_other=9)
    assert ".minnctapersm 9" in out     # the test kernel is not the MAIN_KERNEL instantiation
