#!/usr/bin/env python
"""SASS instruction census of the built extension -> profiles/sass/CENSUS.md (runs on CPU: cuobjdump only).

Counts, per kernel, the mnemonics that prove which hardware path is used (profiling guide: UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, STAS = st.async, SYNCS = mbarrier, LDGMC = multimem) and --
because a pointer that loses its address space silently turns LDS/STS into generic LD/ST -- shared vs generic
memory instructions."""
import glob
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = OrderedDict([
    ("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
    ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("STAS (st.async)", r"\bSTAS"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"\bUCGABAR"), ("ELECT", r"\bELECT"),
    ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("LDGMC (multimem)", r"\bLDGMC"), ("LDS+STS (shared)", r"\b(LDS|STS)\b"),
    ("generic LD/ST", r"\b(LD|ST)\.E\b"), ("HMMA (legacy)", r"\bHMMA"),
])


def census():
    """{kernel name: {"n": instructions, column: count, ...}} of the built extension (``None`` if it is not built)."""
    so = sorted(glob.glob(os.path.join(ROOT, "gossipy_b200", "_C*.so")))
    if not so:
        return None
    out = subprocess.run(["cuobjdump", "-sass", so[0]], capture_output=True, text=True, check=True).stdout
    mangled = list(OrderedDict.fromkeys(re.findall(r"^\s*Function : (\S+)", out, flags=re.M)))
    names = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
    demangled = dict(zip(mangled, names))
    kernels, cur = OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = demangled.get(m.group(1), m.group(1)).strip()
            cur = kernels.setdefault(re.sub(r"\(.*", "", name), {"n": 0, **{c: 0 for c in COLS}})
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        cur["n"] += 1
        for c, pat in COLS.items():
            if re.search(pat, ins):
                cur[c] += 1
    return kernels


def main():
    kernels = census()
    if kernels is None:
        sys.exit("extension not built")
    path = os.path.join(ROOT, "profiles", "sass", "CENSUS.md")
    with open(path, "w") as f:
        f.write("# SASS instruction census of gossipy_b200/_C.so (cuobjdump -sass, sm_100a; `python tools/sass_census.py`)\n\n")
        f.write("| kernel | instrs | " + " | ".join(COLS) + " |\n|" + "---|" * (len(COLS) + 2) + "\n")
        for k, v in sorted(kernels.items(), key=lambda kv: -sum(kv[1][c] for c in list(COLS)[:11])):
            f.write("| `%s` | %d | " % (k, v["n"]) + " | ".join(str(v[c]) for c in COLS) + " |\n")
    print("wrote", path, "(%d kernels)" % len(kernels))


if __name__ == "__main__":
    main()
