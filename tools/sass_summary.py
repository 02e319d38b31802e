"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): UTC*MMA =
tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCCP = tcgen05.cp, UTMALDG / UTMASTG / UTMAREDG / UBLKCP = TMA (REDG: reduce store), HMMA = legacy
mma.sync.   python tools/sass_summary.py > profiles/rNN_sass_summary.txt   (runs here: cuobjdump needs no GPU)"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "labelanything_b200" / "liblabelanything_b200.so"
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTMAPF|HMMA|SYNCS|MUFU)\b")

out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
kernels: "OrderedDict[str, Counter]" = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*\)$", "", name)
        cur = kernels.setdefault(name, Counter())
        continue
    if cur is not None:
        for mn in PAT.findall(line.split("/*")[1] if line.count("/*") >= 2 else line):
            cur[mn] += 1
total = Counter()
print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  ({len(kernels)} kernels)")
for name, c in kernels.items():
    total.update(c)
    tens = {k: v for k, v in c.items() if k not in ("SYNCS", "MUFU")}
    if tens:
        print(f"{name[:110]:110s} " + " ".join(f"{k}={v}" for k, v in sorted(tens.items())))
print("# total: " + " ".join(f"{k}={v}" for k, v in sorted(total.items())))
print(f"# kernels with tcgen05.mma: {sum(1 for c in kernels.values() if any(k.endswith('MMA') and k.startswith('UTC') for k in c))}"
      f"; with legacy HMMA: {sum(1 for c in kernels.values() if c.get('HMMA'))}")
