#!/usr/bin/env python
"""profiles/ptxas_sass_<tag>.md: registers / spills of every kernel in the built
library (cuobjdump -res-usage) and the SASS mnemonics that show which hardware
paths the tensor-core kernels use (cuobjdump -sass).  No GPU needed.
Usage: scripts/make_sass_facts.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = os.path.join(ROOT, "recur_b200", "librecur_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return [re.sub(r"\(.*", "", d).replace("void ", "") for d in out.splitlines()]


res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
rows = []
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        f = dict(kv.split(":") for kv in line.split() if ":" in kv)
        rows.append((cur, int(f.get("REG", 0)), int(f.get("STACK", 0)), int(f.get("SHARED", 0))))
        cur = None
names = demangle([r[0] for r in rows])

sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS",
         "UCGABAR", "MEMBAR", "HMMA", "FFMA", "F2FP", "LDCU", "UTCATOM", "UTMAPF", "UBLKPF", "FENCE"]
per = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if fn and m:
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                per[fn][op.split(".")[0] + ("." + op.split(".")[1] if w in ("UTCHMMA", "UBLKCP", "UTMALDG") and "." in op else "")] += 1
sass_names = dict(zip(per.keys(), demangle(list(per.keys()))))

lines = ["# ptxas and SASS facts, round %s\n" % tag,
         "From the built `recur_b200/librecur_b200.so` (`cuobjdump -res-usage` / `-sass`; "
         "`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`).  No GPU needed; "
         "regenerate with `scripts/make_sass_facts.py %s`.\n" % tag,
         "## Registers per thread, stack (spill) bytes, static shared memory\n",
         "| kernel | registers | stack bytes | static smem bytes |", "|---|---:|---:|---:|"]
for (mangled, reg, stack, shared), name in sorted(zip(rows, names), key=lambda x: -x[0][1]):
    lines.append("| `%s` | %d | %d | %d |" % (name, reg, stack, shared))
lines += ["", "## SASS mnemonics of the tensor-core / bulk-copy kernels\n",
          "`UTCHMMA` = tcgen05.mma kind::f16 (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, "
          "`UTMALDG` = TMA tensor load, `UBLKCP` = cp.async.bulk (1-D bulk copy, both directions), "
          "`SYNCS` = mbarrier operations, `UCGABAR` = cluster barrier.\n",
          "| kernel | mnemonic counts |", "|---|---|"]
for fn, cnt in per.items():
    if any(k.startswith(("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM")) for k in cnt):
        keep = ", ".join("%s x%d" % (k, v) for k, v in sorted(cnt.items())
                         if not k.startswith(("FFMA", "F2FP", "LDCU", "MEMBAR", "FENCE")))
        lines.append("| `%s` | %s |" % (sass_names[fn], keep))
out = os.path.join(ROOT, "profiles", "ptxas_sass_%s.md" % tag)
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
