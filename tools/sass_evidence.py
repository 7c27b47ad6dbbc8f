"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths
(UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMA* = TMA loads / stores / reduce-adds, UBLKCP = bulk
copy, DMMA = FP64 tensor cores).  Usage: python tools/sass_evidence.py > profiles/rN_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "modegpt_b200" / "libmodegpt_b200.so"
PAT = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|"
                 r"UBLKCP[.\w]*|UTCBAR[.\w]*|DMMA[.\w]*|HMMA[.\w]*|SYNCS[.\w]*)")

sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
per = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        per[name] = collections.Counter()
        continue
    if name:
        for tok in PAT.findall(line):
            per[name][tok.split(".")[0] + ("." + ".".join(tok.split(".")[1:3]) if "." in tok else "")] += 1
print(f"# cuobjdump -sass {LIB.name} (sm_100a): tensor-core / TMA mnemonics per kernel")
for k, c in per.items():
    keep = {m: n for m, n in c.items() if not m.startswith("SYNCS")}
    if keep:
        print(f"{k}\n    " + ", ".join(f"{m} x{n}" for m, n in sorted(keep.items())))
