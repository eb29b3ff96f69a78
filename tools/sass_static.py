#!/usr/bin/env python
"""Static SASS census of the shipped library: per kernel, how often the sm_100a-specific opcodes occur.

    python tools/sass_static.py [video_processor_b200/libvpb200.so] > profiles/<tag>_sass_hist.txt

Counts come from `cuobjdump -sass` (static occurrences in the binary, not executed counts; tools/sass_hist.py
gives executed counts from an ncu report).  UBLKCP = cp.async.bulk (1-D TMA), UTMALDG / UTMASTG = tensor-map
TMA load / store, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed f32x2 math, VABSDIFF4, IDP = byte SIMD."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "video_processor_b200/libvpb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2", "FFMA", "VABSDIFF4", "IDP", "PRMT",
         "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "DFMA", "DADD", "F2F", "I2F", "F2I", "MUFU", "ATOMG", "ATOMS", "RED"]
fn, per, regs = None, collections.OrderedDict(), {}
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        per[fn]["_total"] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w:
                per[fn][w] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


print(f"# static SASS opcode census of {lib} (cuobjdump -sass); columns: occurrences in the kernel's code")
cols = [w for w in WATCH if any(c[w] for c in per.values())]
print("kernel".ljust(44) + "total".rjust(7) + "".join(w.rjust(10) for w in cols))
tot = collections.Counter()
for fn, c in sorted(per.items(), key=lambda kv: demangle(kv[0])):
    print(demangle(fn)[:43].ljust(44) + str(c["_total"]).rjust(7) + "".join(str(c[w]).rjust(10) for w in cols))
    tot.update(c)
print("ALL".ljust(44) + str(tot["_total"]).rjust(7) + "".join(str(tot[w]).rjust(10) for w in cols))
