"""Instruction counts per kernel from `cuobjdump -sass libmambo_b200.so` (no GPU needed): the SASS mnemonics that show which
hardware paths a kernel uses -- DMMA (FP64 tensor pipe), UBLKCP (TMA bulk copy), SYNCS (mbarrier), LDGSTS (cp.async), ATOM/RED,
UCGABAR (cluster barrier), plus the DFMA / LDG / STG / LDS totals.  usage: sass_counts.py [pattern ...] > profiles/..."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "quantummambo.jl_b200" / "csrc" / "libmambo_b200.so"
KEYS = ["DMMA", "DFMA", "UBLKCP", "SYNCS", "LDGSTS", "UCGABAR", "RED", "ATOM", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main(patterns):
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            base = op.split(".")[0]
            for k in KEYS:
                if base == k or (k in ("UCGABAR", "RED", "ATOM", "SYNCS", "UBLKCP") and base.startswith(k)):
                    counts[cur][k] += 1
    names = demangle(order)
    print("| kernel | instr | " + " | ".join(KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for fn in order:
        name = names[fn]
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
        if patterns and not any(p in short for p in patterns):
            continue
        c = counts[fn]
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) for k in KEYS) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
