"""Opcode histogram of every kernel in lib/libhi_b200.so (cuobjdump -sass): the SASS-level proof of what the hot kernels use -
UTCHMMA (tcgen05.mma), UTMALDG / UTMASTG (TMA tensor load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
LDGSTS (cp.async), SYNCS (mbarrier), MUFU.EX2, ACQBULK / UBLKCP ...  Runs in the build container (no GPU needed).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "hydrainfer_b200" / "lib" / "libhi_b200.so"
KEY = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH", "LDTM", "STTM", "LDGSTS", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "HFMA2", "FMNMX3",
       "LDG", "STG", "LDS", "STS", "ATOM", "RED", "SHFL", "BAR", "ACQBULK", "LDL", "STL", "ELECT", "USETMAXREG", "ERRBAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur[op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode histogram per kernel of lib/libhi_b200.so (sm_100a, `cuobjdump -sass`)\n")
    print("Counts are static instruction counts.  `UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `UTMALDG` / `UTMASTG` = TMA tensor load / store, `UBLKCP` = 1-D bulk copy (cp.async.bulk),")
    print("`LDTM` / `STTM` = tcgen05.ld / st, `LDGSTS` = cp.async, `SYNCS` = mbarrier ops, `LDL` / `STL` = local-memory (spill) traffic.\n")
    print("| kernel | instr | " + " | ".join(KEY) + " |")
    print("|---|---|" + "---|" * len(KEY))
    for (mangled, cnt), name in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", name).replace("void hi::", "")
        total = sum(cnt.values())
        cells = []
        for k in KEY:
            n = sum(v for op, v in cnt.items() if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")))
            cells.append(str(n) if n else "")
        print(f"| `{short}` | {total} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
