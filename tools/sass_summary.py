#!/usr/bin/env python
"""Count the Blackwell-native SASS mnemonics per kernel of libcdc_b200.so (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UTMASTG / UBLKCP; HMMA would be the legacy mma.sync path).
Usage: python tools/sass_summary.py [lib.so] > profiles/sass_summary_r2.txt   (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "conditional-diffusion-model-for-compression_b200", "libcdc_b200.so")
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "HMMA", "HGMMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cnt, size, name = collections.defaultdict(collections.Counter), collections.Counter(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void cdc::", "").replace("cdc::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        size[name] += 1
        op = m.group(1).split(".")[0]
        for k in MN:
            if op.startswith(k):
                cnt[name][k] += 1
print(f"# {os.path.basename(lib)}: SASS mnemonics per kernel (instructions = 16-byte SASS words)")
print(f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in MN))
tot = collections.Counter()
for k in sorted(size, key=lambda n: (-cnt[n]["UTCHMMA"], n)):
    print(f"{k[:78]:78s} {size[k]:6d} " + " ".join(f"{cnt[k][m]:8d}" for m in MN))
    tot.update(cnt[k])
print(f"{'TOTAL':78s} {sum(size.values()):6d} " + " ".join(f"{tot[m]:8d}" for m in MN))
