#!/usr/bin/env python
"""Opcode histogram per kernel of the built library (cuobjdump -sass): python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.txt
Shows which kernels carry tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), TMA (UTMALDG / UTMASTG), mma.sync (HMMA), MUFU ..."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "vats_multimodal_lm_b200", "csrc", "libvats_attn.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
fn, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*", "", fn)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        base = op.split(".")[0]
        key = base
        if base in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU", "LDGSTS", "LDSM", "SYNCS", "UBLKCP",
                    "UTMAPF", "UTCCP", "FFMA2", "FADD2", "FMNMX3", "REDUX", "SHFL", "BAR", "STG", "LDG", "STS", "LDS", "ATOMG", "RED"):
            key = ".".join(op.split(".")[:2]) if base in ("MUFU", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "SYNCS") else base
        hist[fn][key] += 1
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "MUFU", "SYNCS", "FFMA2", "FADD2", "FMNMX3")
print(f"# {os.path.basename(lib)}: SASS opcode histogram per kernel (cuobjdump -sass, sm_100a)")
for fn, h in hist.items():
    total = sum(h.values())
    print(f"\n{fn}   [{total} instructions]")
    marks = {k: v for k, v in h.items() if k.split(".")[0] in KEY}
    print("  tensor / TMA / sync : " + (", ".join(f"{k} {v}" for k, v in sorted(marks.items())) or "-"))
    rest = [(k, v) for k, v in h.most_common(14) if k not in marks]
    print("  most frequent       : " + ", ".join(f"{k} {v}" for k, v in rest))
