#!/usr/bin/env python
"""Register / spill / shared-memory table of every kernel from the build's `-Xptxas -v` logs -> profiles/PTXAS.md.

Runs on CPU after `python -c "import __graft_entry__ as g; g.build(force=True)"` (the logs live next to the objects
under build/gossipy_b200_C/).  The profiling guide asks for this check before GPU time is spent: spills in a hot loop
and a register count that halves the resident CTAs are visible here."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    logs = sorted(glob.glob(os.path.join(ROOT, "build", "gossipy_b200_C", "*.cu.o.log")))
    if not logs:
        sys.exit("no build logs (run __graft_entry__.build(force=True) first)")
    rows = []
    for log in logs:
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                             r"(\d+) bytes spill loads\n.*Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", text):
            rows.append((os.path.basename(log)[:-len(".o.log")],) + m.groups())
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    path = os.path.join(ROOT, "profiles", "PTXAS.md")
    with open(path, "w") as f:
        f.write("# ptxas resource usage of every kernel (sm_100a, `-Xptxas -v`; `python tools/ptxas_report.py`)\n\n"
                "Static shared memory only (the tcgen05 kernels request their dynamic shared memory at launch).\n\n"
                "| file | kernel | registers | stack B | spill stores B | spill loads B | barriers | static smem B |\n|---|---|---|---|---|---|---|---|\n")
        for r, name in sorted(zip(rows, names), key=lambda x: (x[0][0], x[1])):
            src, _, stack, sst, sld, regs, bars, smem = r
            f.write("| %s | `%s` | %s | %s | %s | %s | %s | %s |\n" % (src, re.sub(r"\(.*", "", name), regs, stack, sst, sld, bars or 0, smem or 0))
    print("wrote %s (%d kernels)" % (path, len(rows)))


if __name__ == "__main__":
    main()
