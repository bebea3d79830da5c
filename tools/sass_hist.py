#!/usr/bin/env python3
"""SASS opcode histogram of the hot kernels (cuobjdump -sass on the built objects): which Blackwell instructions each one uses.
    python tools/sass_hist.py > profiles/r02/sass_opcodes_final.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sys.argv[1:] or [os.path.join(ROOT, "taseg_b200", "build", f) for f in ("conv_tc.o", "conv_wgrad_tc.o", "conv.o", "bn.o")]
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "UCGABAR_ARV", "LDS", "STS", "LDG", "STG", "ATOMG", "RED",
       "SHFL", "BAR", "FENCE", "MEMBAR", "CCTL", "NANOSLEEP", "HMMA", "FFMA"]
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
print("| kernel | SASS instructions | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for obj in objs:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, hist = None, collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            continue
        m = pat.match(line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    for fn, h in sorted(hist.items()):
        tot = sum(h.values())
        if tot < 400:
            continue
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*\)$", "", name.replace("void tsg::", ""))
        print("| `%s` | %d | " % (name, tot) + " | ".join(str(h.get(k, 0)) for k in KEY) + " |")
