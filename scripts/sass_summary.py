#!/usr/bin/env python3
"""Per-kernel SASS mnemonic census of the built library (the "Blackwell tell" of B200_PROFILING.md): how many tcgen05 MMAs
(UTCHMMA / UTCQMMA ...), TMEM loads (LDTM), TMA loads / stores (UTMALDG / UTMASTG), tensor-memory barriers (UTCBAR), legacy
warp-level MMAs (HMMA) and mbarrier waits (SYNCS) each kernel of lib/libparakeet_trt.so contains.
usage: scripts/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt   (needs cuobjdump + cu++filt from the CUDA toolkit)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "trt-asr-engine_b200", "lib", "libparakeet_trt.so")
COLS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCCP", "HMMA", "LDSM", "SYNCS", "BAR"]


def strip_params(name: str) -> str:
    """demangled kernel name without its parameter list (template arguments stay)"""
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0 and i > 0 and name[i - 1] not in "< ,:":
            return name[:i]
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for c in COLS:
                if op == c or op.startswith(c + "."):
                    counts[cur][c] += 1
    names = subprocess.run(["cu++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines() if order else []
    print(f"# {os.path.relpath(LIB, ROOT)}  arch={','.join(sorted(arch))}  kernels={len(order)}")
    print("# instr = SASS instructions in the kernel; the other columns count instructions whose opcode starts with the mnemonic")
    print("# " + " ".join(f"{c:>8}" for c in ["instr"] + COLS) + "  kernel")
    tot = collections.Counter()
    rows = []
    for mangled, name in zip(order, names):
        c = counts[mangled]
        short = strip_params(name)
        rows.append((short, c))
        tot.update(c)
    for short, c in sorted(rows, key=lambda r: (-(r[1]["UTCHMMA"] + r[1]["UTCQMMA"]), -r[1]["HMMA"], r[0])):
        print("  " + " ".join(f"{c[k]:>8}" for k in ["_total"] + COLS) + "  " + short)
    print("# total")
    print("  " + " ".join(f"{tot[k]:>8}" for k in ["_total"] + COLS))


if __name__ == "__main__":
    main()
