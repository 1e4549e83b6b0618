#!/usr/bin/env python
"""One kernel's SASS listing out of the built extension (cuobjdump -sass, encodings stripped), runs on CPU.

    python tools/sass_dump.py 'mlp1_train_tc4_kernel<8, true, false, false>' profiles/sass/mlp1_train_tc8_fp32eq_full.sass \\
        --title "gb::mlp1_train_tc4_kernel<8, X3=true, SC=false, MOM=false> (the headline kernel, 'tc8')"
"""
import argparse
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel", help="substring of the demangled kernel name")
    ap.add_argument("out")
    ap.add_argument("--title", default=None)
    a = ap.parse_args()
    so = sorted(glob.glob(os.path.join(ROOT, "gossipy_b200", "_C*.so")))
    if not so:
        sys.exit("extension not built")
    text = subprocess.run(["cuobjdump", "-sass", so[0]], capture_output=True, text=True, check=True).stdout
    chunks = re.split(r"^\s*Function : ", text, flags=re.M)[1:]
    hits = []
    for ch in chunks:
        mangled = ch.split("\n", 1)[0].strip()
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        if a.kernel in name:
            hits.append((name, ch))
    if len(hits) != 1:
        sys.exit("%d kernels match %r: %s" % (len(hits), a.kernel, [h[0][:90] for h in hits]))
    name, body = hits[0]
    lines = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?;)", line)
        if m:
            lines.append("/*%s*/ %s" % (m.group(1), m.group(2).rstrip()))
    with open(a.out, "w") as f:
        f.write("// %s  (cuobjdump -sass, sm_100a, encodings stripped)\n" % (a.title or re.sub(r"\(.*", "", name)))
        f.write("\n".join(lines) + "\n")
    print("wrote %s: %d instructions of %s" % (a.out, len(lines), re.sub(r"\(.*", "", name)))


if __name__ == "__main__":
    main()
