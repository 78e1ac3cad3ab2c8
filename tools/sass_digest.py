#!/usr/bin/env python
"""Instruction-class histogram of one kernel's SASS (cuobjdump), whole kernel and hottest loop.

    python tools/sass_digest.py [--lib PATH] [--kernel nb2_kernelILb0] [--md OUT.md]

The "loop" is the backward-branch region with the most instructions (for nb2_kernel: the list-step loop with its 8
unrolled pair evaluations).  Counts are STATIC instructions; they say what the compiler emitted (FFMA2 vs FFMA, MUFU,
LDS, ...), not how often each executes.
"""
import argparse
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "openmm-atmmetaforce-plugin_b200", "libatm_b200.so")

CLASSES = [
    ("FFMA2", r"^FFMA2"), ("FMUL2", r"^FMUL2"), ("FADD2", r"^FADD2"),
    ("FFMA", r"^FFMA"), ("FMUL", r"^FMUL"), ("FADD", r"^FADD"),
    ("MUFU", r"^MUFU"), ("FSEL/FSETP/FMNMX", r"^(FSEL|FSETP|FMNMX|FSET)"),
    ("F2I/I2F/F2F", r"^(F2I|I2F|F2F|FRND|F2FP)"), ("DADD/DMUL/DFMA", r"^D(ADD|MUL|FMA|SETP)"),
    ("int ALU (IADD3/LOP3/SHF/IMAD/LEA/ISETP/SEL/PLOP3/PRMT/MOV)", r"^(IADD|LOP|SHF|IMAD|LEA|ISETP|SEL|PLOP3|PRMT|MOV|IABS|POPC|FLO|VIADD|UIADD|ULOP|USHF|UMOV|UIMAD|ULEA|UISETP|USEL|UPLOP|R2UR|S2R|S2UR|CS2R|VOTE|R2P|P2R|BMSK|SGXT|UPRMT|UFLO|ULDC|LDC|LDCU)"),
    ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDGSTS (cp.async)", r"^LDGSTS"),
    ("RED/ATOM", r"^(RED|ATOM)"), ("UBLKCP/SYNCS (TMA bulk + mbarrier)", r"^(UBLKCP|SYNCS)"),
    ("SHFL", r"^SHFL"), ("BRA/BSSY/BSYNC/EXIT/WARPSYNC", r"^(BRA|BSSY|BSYNC|EXIT|WARPSYNC|CALL|RET|BREAK|NANOSLEEP|YIELD|BAR|DEPBAR|LDGDEPBAR|ACQBULK|MEMBAR|ERRBAR|FENCE|CCTL|ELECT|NOP)"),
]


def disassemble(lib, kernel):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            txt = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
            funcs[cur].append((int(m.group(1), 16), txt))
    names = [f for f in funcs if kernel in f]
    if not names:
        raise SystemExit(f"no kernel matching {kernel!r}; have: {sorted(funcs)[:40]}")
    return names[0], funcs[names[0]]


def histogram(instrs):
    h = collections.Counter()
    for _, t in instrs:
        op = t.split()[0]
        for name, pat in CLASSES:
            if re.match(pat, op):
                h[name] += 1
                break
        else:
            h["other: " + op.split(".")[0]] += 1
    return h


def hottest_loop(instrs):
    """Innermost backward-branch region (no other backward branch inside) with the most fp32 arithmetic."""
    loops = []
    for addr, t in instrs:
        m = re.search(r"0x([0-9a-f]+)\s*$", t) if t.startswith("BRA") else None
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    best, best_fp = [], -1
    for lo, hi in loops:
        if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops):
            continue
        body = [(a, x) for a, x in instrs if lo <= a <= hi]
        fp = sum(1 for _, x in body if re.match(r"^F(FMA|MUL|ADD)", x))
        if fp > best_fp:
            best, best_fp = body, fp
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=LIB)
    ap.add_argument("--kernel", default="nb2_kernelILb0")
    ap.add_argument("--md", default=None)
    a = ap.parse_args()
    name, instrs = disassemble(a.lib, a.kernel)
    loop = hottest_loop(instrs)
    hk, hl = histogram(instrs), histogram(loop)
    lines = [f"# SASS digest of `{name}`", "",
             f"`cuobjdump -sass {os.path.basename(a.lib)}`; static instruction counts.  Whole kernel: {len(instrs)} instructions; "
             f"largest backward-branch region (the inner loop): {len(loop)} instructions.", "",
             "| class | whole kernel | inner loop |", "|---|---:|---:|"]
    keys = [n for n, _ in CLASSES] + sorted(k for k in hk if k.startswith("other"))
    for k in keys:
        if hk.get(k) or hl.get(k):
            lines.append(f"| {k} | {hk.get(k, 0)} | {hl.get(k, 0)} |")
    txt = "\n".join(lines) + "\n"
    print(txt)
    if a.md:
        with open(a.md, "w") as f:
            f.write(txt)


if __name__ == "__main__":
    main()
