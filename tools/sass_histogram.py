#!/usr/bin/env python
"""Opcode histogram per kernel of liblmm.so (cuobjdump -sass): the SASS proof of what each kernel runs on --
DMMA.8x8x4 (FP64 tensor core), UBLKCP (cp.async.bulk, the TMA bulk-copy engine), SYNCS (mbarrier), LDGSTS (cp.async),
ACQBULK / PREEXIT (programmatic dependent launch), DFMA / DADD / DMUL (FP64 pipe), MUFU (RCP64H / RSQ64H seeds), UTCIMMA (tcgen05.mma kind::i8),
UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM allocation), LDTM (tcgen05.ld).

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "linearmixingmodels.jl_b200", "liblmm.so")
INTEREST = ["DMMA", "UBLKCP", "SYNCS", "LDGSTS", "ACQBULK", "PREEXIT", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL",
            "NANOSLEEP", "ATOM", "RED", "UTMALDG", "UTCHMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Za-z0-9_]+)*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op.split(".")[0]] += 1
            if op.startswith(("DMMA", "UBLKCP", "SYNCS", "MUFU", "LDGSTS")):
                kernels[cur][op] += 1
    demangled = {}
    try:
        names = list(kernels)
        d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(names, d))
    except Exception:
        pass
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): opcode counts per kernel")
    print("# columns: total instructions | " + " ".join(INTEREST))
    for k, c in kernels.items():
        name = demangled.get(k, k)
        name = re.sub(r"\(.*\)$", "", name)
        total = sum(v for op, v in c.items() if "." not in op)
        cols = " ".join(f"{op}={c.get(op, 0)}" for op in INTEREST if c.get(op, 0))
        detail = " ".join(f"{op}={v}" for op, v in sorted(c.items()) if "." in op)
        print(f"{name}\n    total={total} {cols}\n    {detail}")


if __name__ == "__main__":
    main()
