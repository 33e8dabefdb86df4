"""profiles/r02_sass_mnemonics.md: per-kernel counts of the SASS mnemonics that prove which hardware paths the library uses
(cuobjdump -sass of the in-tree libmotionclone_b200.so; runs in the build container, no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "motionclone_b200", "libmotionclone_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda names: subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")  # noqa: E731
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "HMMA", "MUFU.EX2", "FADD2", "FFMA2"]
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for c in cols:
        if c == "HMMA":
            if op.startswith("HMMA"):
                counts[cur][c] += 1
        elif op.startswith(c):
            counts[cur][c] += 1
names = [n.strip() for n in demangle(order)]
rows = []
for mangled, name in zip(order, names):
    short = re.sub(r"\(.*", "", name)
    if not any(counts[mangled][c] for c in cols):
        continue
    rows.append((short, [counts[mangled][c] for c in cols]))
rows.sort()
out = ["# SASS mnemonic evidence, round 2 (`cuobjdump -sass motionclone_b200/libmotionclone_b200.so`, sm_100a; `scripts/sass_table.py`)", "",
       "`UTCHMMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st (TMEM), `UTMALDG` = cp.async.bulk.tensor (tensor-map TMA), "
       "`UBLKCP` = cp.async.bulk (1-D TMA), `HMMA` = mma.sync, `MUFU.EX2` = ex2.approx, `FADD2`/`FFMA2` = packed fp32 pairs.", "",
       "| kernel | " + " | ".join(cols) + " |", "|---|" + "---|" * len(cols)]
for short, vals in rows:
    out.append(f"| `{short}` | " + " | ".join(str(v) if v else "" for v in vals) + " |")
open(os.path.join(root, "profiles", "r02_sass_mnemonics.md"), "w").write("\n".join(out) + "\n")
print(len(rows), "kernels")
