"""Opcode histogram per kernel of libpcp_b200.so (cuobjdump -sass): the mnemonics that prove which hardware units a kernel
uses - UTCHMMA / UTCCP (tcgen05.mma / .cp), LDTM / STTM (tcgen05.ld / .st), UBLKCP / UTMALDG / UTMASTG (bulk and tensor TMA
copies), LDGSTS (cp.async), ATOM / RED / ATOMS (global / shared atomics), REDUX, MATCH, SHFL - next to the instruction count.
    python tools/sass_summary.py [path/to/lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                           "practical-collab-perception_b200", "libpcp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "LDGSTS",
         "SYNCS", "ATOMS", "ATOMG", "ATOM", "RED", "REDUX", "MATCH", "SHFL", "LDG", "STG", "LDS", "STS", "BAR", "FFMA", "MUFU"]
kern = None
hist = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("LDG", "STG", "LDS", "STS", "ATOM", "RED") and op == w):
                hist[kern][w] += 1
                break
        else:
            for w in WATCH:
                if op.startswith(w):
                    hist[kern][w] += 1
                    break
print(f"# {os.path.basename(lib)}: SASS opcode counts per kernel (static, cuobjdump -sass; sm_100a)")
for k, c in hist.items():
    parts = [f"{w}={c[w]}" for w in WATCH if c[w]]
    print(f"{k[:110]}\n    instructions={c['_total']}  " + "  ".join(parts))
