"""Warp-uniformity pass over nvcc's PTX: marks provably warp-uniform conditional branches ``bra.uni``.

Why: the rollout kernel is one warp = one game with warp-uniform control flow, but most branch conditions come
from values the front end cannot prove uniform (shared-memory loads, ballot / shuffle / redux results).  ptxas then
wraps every such branch in BSSY/BSYNC and guards every ``*.sync`` instruction with a BRA.DIV check: 27 % of the
executed instructions in profile r1b were convergence bookkeeping.  ``bra.uni`` is PTX's way to promise that all
active threads of the warp take the branch together; this pass derives that promise conservatively.

Analysis (per function; a forward data-flow over the CFG, iterated with control dependence to a fixed point):
  * lane-variant seeds: %laneid, %tid, %lanemask_*, atomics, call results, values loaded from call return slots,
    loads from local (thread-private) or generic address space;
  * a destination is variant if any source register, its guard predicate or (for loads) its address is variant;
  * warp-collectives make uniform values whatever their inputs: vote.sync, redux.sync, and shfl.sync.idx whose
    source-lane / clamp / mask operands are uniform;
  * control dependence: every register written inside the span of a divergent branch (forward: branch..join,
    backward: loop body) is variant.
A conditional ``bra`` whose guard is not variant becomes ``bra.uni``.  Only functions whose name matches ``only``
are touched.
"""
from __future__ import annotations

import re
import sys
from typing import Dict, List, Set, Tuple

DEBUG = False
TRACE = None
REG = re.compile(r"%[a-z]+\d+")
VARIANT_SREGS = ("%laneid", "%tid.", "%lanemask_")
NO_DEST = ("st.", "bra", "bar.", "barrier.", "red.", "ret", "exit", "membar", "fence", "trap", "brkpt", "prefetch", "call")


def split_functions(lines: List[str]) -> List[Tuple[int, int, str]]:
    out, i = [], 0
    while i < len(lines):
        m = re.match(r"^(?:\.visible\s+|\.weak\s+)?\.(entry|func)\b", lines[i])
        if m:
            # find the body: first line that is exactly "{" after the header, to the matching "}" at column 0
            j = i
            while j < len(lines) and lines[j].strip() != "{":
                if lines[j].strip().endswith(";") and "(" not in lines[j]:
                    break  # a prototype
                j += 1
            if j < len(lines) and lines[j].strip() == "{":
                k = j + 1
                depth = 1
                while k < len(lines) and depth:
                    s = lines[k].strip()
                    if s == "{":
                        depth += 1
                    elif s == "}":
                        depth -= 1
                    k += 1
                name = re.search(r"([_A-Za-z0-9$]+)\s*\(", " ".join(lines[i:j + 1]))
                out.append((j + 1, k - 1, name.group(1) if name else ""))
                i = k
                continue
        i += 1
    return out


class Ins:
    __slots__ = ("idx", "guard", "op", "dests", "srcs", "addr", "target", "text")


def parse(lines: List[str], lo: int, hi: int):
    ins: List[Ins] = []
    labels: Dict[str, int] = {}
    for idx in range(lo, hi):
        s = lines[idx].strip()
        if not s or s.startswith("//") or s.startswith("."):
            continue
        m = re.match(r"^(\$?[A-Za-z_][\w$]*):$", s)
        if m:
            labels[m.group(1)] = idx
            continue
        if s in ("{", "}"):
            continue
        if s.startswith("{") and s.endswith("}") and ";" in s:
            # an inline-asm block written on one line: "{ .reg .pred q; setp ...; @q st ...; }".  Registers declared
            # inside are private names; treat every statement separately, conservatively (any %-register counts).
            for stmt in s[1:-1].split(";"):
                stmt = stmt.strip()
                if not stmt or stmt.startswith("."):
                    continue
                i = Ins()
                i.idx, i.guard, i.op, i.text, i.target, i.addr = idx, None, "asmblock", s, None, []
                regs = REG.findall(stmt)
                g = re.match(r"^@!?(\S+)\s+(.*)$", stmt)
                body2 = g.group(2) if g else stmt
                opn = body2.split(None, 1)[0]
                if any(opn.startswith(p_) for p_ in NO_DEST):
                    i.dests, i.srcs = [], regs
                else:
                    ops2 = split_operands(body2.split(None, 1)[1] if " " in body2 else "")
                    i.dests = REG.findall(ops2[0]) if ops2 else []
                    i.srcs = regs          # every register of the block may feed the result (private predicates)
                ins.append(i)
            continue
        if not s.endswith(";"):
            continue
        body = s[:-1].strip()
        guard = None
        g = re.match(r"^@!?(%p\d+)\s+(.*)$", body)
        if g:
            guard, body = g.group(1), g.group(2)
        parts = body.split(None, 1)
        op = parts[0]
        rest = parts[1] if len(parts) > 1 else ""
        i = Ins()
        i.idx, i.guard, i.op, i.text, i.target, i.addr = idx, guard, op, s, None, []
        ops = split_operands(rest)
        if op.startswith("bra"):
            i.target = rest.strip()
            i.dests, i.srcs = [], []
        elif any(op.startswith(p) for p in NO_DEST):
            i.dests, i.srcs = [], REG.findall(rest)
        else:
            d = ops[0] if ops else ""
            i.dests = REG.findall(d)
            i.srcs = [r for o in ops[1:] for r in REG.findall(o)]
            if op.startswith("ld.") or op.startswith("ldu.") or op.startswith("atom."):
                i.addr = REG.findall(ops[1]) if len(ops) > 1 else []
        ins.append(i)
    return ins, labels


def split_operands(rest: str) -> List[str]:
    out, depth, cur = [], 0, ""
    for ch in rest:
        if ch in "{[(":
            depth += 1
        elif ch in "}])":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def analyse(lines: List[str], lo: int, hi: int) -> Tuple[List[int], int]:
    """returns (line indices of conditional branches that are provably uniform, number of conditional branches).

    Flow-sensitive (PTX registers are reused across live ranges after phi elimination): a forward "may be
    lane-variant" data-flow over the CFG, re-run with every block that is control dependent on a divergent
    branch (between the branch and its immediate post-dominator) forcing its definitions variant."""
    ins, labels = parse(lines, lo, hi)
    if not ins:
        return [], 0
    for i in ins:
        if i.op.startswith("brx"):
            return [], 0  # indirect branches: not handled, touch nothing
    # ---- basic blocks
    label_at = {idx: name for name, idx in labels.items()}
    label_lines = sorted(labels.values())
    starts = {0}
    for n, i in enumerate(ins):
        if i.op.startswith("bra") or i.op in ("ret", "exit"):
            if n + 1 < len(ins):
                starts.add(n + 1)
    import bisect
    line_of = [i.idx for i in ins]
    label_to_ins = {}
    for name, lidx in labels.items():
        n = bisect.bisect_left(line_of, lidx)   # first instruction after the label line
        label_to_ins[name] = n
        if n < len(ins):
            starts.add(n)
    starts = sorted(starts)
    block_of = [0] * len(ins)
    blocks = []
    for b, st in enumerate(starts):
        en = starts[b + 1] if b + 1 < len(starts) else len(ins)
        blocks.append((st, en))
        for n in range(st, en):
            block_of[n] = b
    nb = len(blocks)
    EXIT = nb
    succ = [[] for _ in range(nb + 1)]
    for b, (st, en) in enumerate(blocks):
        last = ins[en - 1]
        if last.op.startswith("bra"):
            t = label_to_ins.get(last.target)
            if t is None:
                return [], 0
            succ[b].append(block_of[t] if t < len(ins) else EXIT)
            if last.guard is not None:
                succ[b].append(b + 1 if b + 1 < nb else EXIT)
        elif last.op in ("ret", "exit"):
            succ[b].append(EXIT)
        else:
            succ[b].append(b + 1 if b + 1 < nb else EXIT)
    pred = [[] for _ in range(nb + 1)]
    for b in range(nb):
        for t in succ[b]:
            pred[t].append(b)
    # ---- immediate post-dominators (Cooper-Harvey-Kennedy on the reverse graph)
    order, seen = [], [False] * (nb + 1)
    stack = [(EXIT, 0)]
    seen[EXIT] = True
    while stack:
        v, k = stack.pop()
        if k < len(pred[v]):
            stack.append((v, k + 1))
            w = pred[v][k]
            if not seen[w]:
                seen[w] = True
                stack.append((w, 0))
        else:
            order.append(v)
    rpo = list(reversed(order))           # EXIT first
    num = {v: n for n, v in enumerate(rpo)}
    ipdom = {EXIT: EXIT}
    changed = True
    while changed:
        changed = False
        for v in rpo[1:]:
            cand = None
            for s_ in succ[v]:
                if s_ in ipdom:
                    if cand is None:
                        cand = s_
                    else:
                        a, c = cand, s_
                        while a != c:
                            while num[a] > num[c]:
                                a = ipdom[a]
                            while num[c] > num[a]:
                                c = ipdom[c]
                        cand = a
            if cand is not None and ipdom.get(v) != cand:
                ipdom[v] = cand; changed = True

    def region(b: int) -> Set[int]:
        """blocks control dependent on the branch ending block b: reachable from its successors before ipdom(b)"""
        stop = ipdom.get(b, EXIT)
        out, todo = set(), [s_ for s_ in succ[b]]
        while todo:
            v = todo.pop()
            if v == stop or v == EXIT or v in out:
                continue
            out.add(v)
            todo.extend(succ[v])
        return out

    forced: Set[int] = set()
    div_blocks: Set[int] = set()
    # registers as bit positions: the variant sets of the data-flow are Python ints
    regno: Dict[str, int] = {}

    def bits(regs) -> int:
        m = 0
        for r in regs:
            k = regno.get(r)
            if k is None:
                k = regno[r] = len(regno)
            m |= 1 << k
        return m

    KIND_SEED, KIND_UNI, KIND_SHFL, KIND_DATA = 0, 1, 2, 3
    pre = []   # per instruction: (kind, dest mask, source mask, guard mask, is conditional branch)
    for i in ins:
        op, t = i.op, i.text
        generic_or_local_load = op.startswith("ld.") and not any(
            op.startswith(p_) for p_ in ("ld.global", "ld.shared", "ld.param", "ld.const", "ld.volatile.shared", "ld.volatile.global"))
        if any(s_ in t for s_ in VARIANT_SREGS) or op.startswith("atom.") or op.startswith("activemask") or generic_or_local_load or \
                (op.startswith("ld.param") and "retval" in t) or (op.startswith("shfl.sync") and not op.startswith("shfl.sync.idx")):
            kind, src = KIND_SEED, 0   # thread-private (local) and generic loads may differ per lane even at one address
        elif op.startswith("vote.sync") or op.startswith("redux.sync"):
            kind, src = KIND_UNI, 0
        elif op.startswith("shfl.sync.idx"):
            ops = split_operands(t[:-1].split(None, 1)[1] if i.guard is None else t[:-1].split(None, 2)[2])
            kind, src = KIND_SHFL, bits([r for o in ops[2:] for r in REG.findall(o)])   # lane, clamp, membermask
        else:
            kind, src = KIND_DATA, bits(i.srcs)
        pre.append((kind, bits(i.dests), src, bits([i.guard]) if i.guard else 0, i.op.startswith("bra") and i.guard is not None))
    while True:
        # ---- forward data-flow: IN[b] = union OUT[pred]
        IN = [0] * nb
        OUT = [None] * nb
        from collections import deque
        work = deque(range(nb))
        inwork = [True] * nb
        guard_var: Dict[int, bool] = {}
        while work:
            b = work.popleft()
            inwork[b] = False
            cur = IN[b]
            st, en = blocks[b]
            fb = b in forced
            for n in range(st, en):
                kind, dmask, smask, gmask, isbr = pre[n]
                if isbr:
                    guard_var[n] = bool(cur & gmask)
                if not dmask:
                    continue
                v = fb or kind == KIND_SEED or (kind >= KIND_SHFL and bool(cur & smask)) or bool(cur & gmask)
                if TRACE is not None and v:
                    i = ins[n]
                    c = "forced" if fb else next((r for r in i.srcs if (cur >> regno[r]) & 1), None) or \
                        (i.guard if gmask and cur & gmask else "seed")
                    TRACE.setdefault(n, (c, i.text, i.idx))
                if v:
                    cur |= dmask
                elif not gmask:
                    cur &= ~dmask
                # a predicated write with a uniform guard keeps the old value where the guard is false: unchanged
            if OUT[b] is None or cur != OUT[b]:
                OUT[b] = cur
                for t_ in succ[b]:
                    if t_ != EXIT and (cur & ~IN[t_]):
                        IN[t_] |= cur
                        if not inwork[t_]:
                            work.append(t_); inwork[t_] = True
        new_div = {block_of[n] for n, v in guard_var.items() if v}
        if new_div <= div_blocks:
            uni = [ins[n].idx for n, v in guard_var.items() if not v and ins[n].op == "bra"]
            return sorted(uni), len(guard_var)
        for b in sorted(new_div - div_blocks):
            r = region(b)
            if DEBUG:
                st, en = blocks[b]
                print(f"  divergent branch @line {ins[en-1].idx}: {ins[en-1].text}  region {len(r)} blocks", file=sys.stderr)
            forced |= r
        div_blocks |= new_div


#: the tuned all-SimRandom / counter-stream / <= 64-territory instantiation (gets ``min_blocks``; the others ``min_blocks_other``)
MAIN_KERNEL = "rollout_kernelILi2ELb0ELb0E"


def transform(ptx: str, only: str = "rollout_kernel", verbose: bool = False, min_blocks: int = 0, min_blocks_other: int = 0,
              stats_out: list = None) -> str:
    """``min_blocks`` > 0 also rewrites ``.minnctapersm`` of the matching kernels: the front end is compiled with a
    loose ``__launch_bounds__`` (it produces markedly worse code under a tight register budget: 105k vs 158k
    playouts/s measured) and only ptxas is given the tight one.  ``min_blocks_other`` (default: the same) is the budget
    of the instantiations other than MAIN_KERNEL (replay, modelled opponents, large maps: bigger code, fewer blocks fit)."""
    lines = ptx.split("\n")
    stats = []
    if min_blocks > 0:
        cur = ""
        for n, ln in enumerate(lines):
            m = re.match(r"^(?:\.visible\s+|\.weak\s+)?\.(entry|func)\s+(?:\([^)]*\)\s*)?([_A-Za-z0-9$]+)", ln)
            if m:
                cur = m.group(2)
            if only in cur and ln.startswith(".minnctapersm"):
                mb = min_blocks if (MAIN_KERNEL in cur or not min_blocks_other) else min_blocks_other
                lines[n] = f".minnctapersm {mb}"
    for lo, hi, name in split_functions(lines):
        if only not in name:
            continue
        uni, total = analyse(lines, lo, hi)
        for idx in uni:
            lines[idx] = re.sub(r"\bbra\b(?!\.uni)", "bra.uni", lines[idx], count=1)
        stats.append((name, len(uni), total))
    if verbose:
        for name, u, t in stats:
            print(f"ptx_uniform: {name}: {u} of {t} conditional branches marked uniform", file=sys.stderr)
    if stats_out is not None:
        stats_out.extend(stats)
    return "\n".join(lines)


if __name__ == "__main__":
    # usage: ptx_uniform.py in.ptx out.ptx [min_blocks [min_blocks_other [stats.json]]]
    src, dst = sys.argv[1], sys.argv[2]
    mb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    mbo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    st = []
    text = transform(open(src).read(), verbose=True, min_blocks=mb, min_blocks_other=mbo, stats_out=st)   # read fully before (possibly) overwriting in place
    open(dst, "w").write(text)
    if len(sys.argv) > 5:
        import json
        json.dump({"min_blocks": mb, "min_blocks_other": mbo, "bra_uni": text.count("bra.uni"),
                   "kernels": {n: {"uniform": u, "conditional": t} for n, u, t in st}}, open(sys.argv[5], "w"), indent=1)
