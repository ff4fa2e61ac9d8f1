"""Per-kernel SASS instruction histogram of libb200clip.so (cuobjdump -sass): the mnemonics that prove a
Blackwell-native kernel (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA, UTCBAR = tcgen05.commit)
and the ones that would betray a legacy path (HMMA = mma.sync).  usage: python tools/sass_histogram.py > profiles/rNN_sass_histogram.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "construction_clip_b200" / "libb200clip.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "ATOMS", "RED", "STG", "LDG", "STS", "LDS"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
hist = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = hist.setdefault(re.sub(r"\(.*", "", name).replace("void ", ""), collections.Counter())
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
        if base in KEYS:
            cur[base] += 1
print("| kernel | instr | " + " | ".join(KEYS) + " |\n|---|---:|" + "---:|" * len(KEYS))
tot = collections.Counter()
for k, c in hist.items():
    if c["total"] < 50:
        continue
    tot.update(c)
    print(f"| `{k[:70]}` | {c['total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in KEYS) + " |")
print(f"| **all kernels** | {tot['total']} | " + " | ".join(str(tot[x]) if tot[x] else "" for x in KEYS) + " |")
