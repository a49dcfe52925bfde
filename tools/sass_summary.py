"""Per-kernel SASS evidence of the shipped library: counts of the Blackwell-native mnemonics
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and of
the legacy tensor path (HMMA) in every kernel of libgdl_b200.so.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "iccv2025-gdl_b200", "gdl_b200", "libgdl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "REDUX", "MEMBAR.ALL.GPU"]
cnt = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cnt[cur]["_instr"] += 1
    if op.startswith("UTCHMMA"):
        cnt[cur]["UTCHMMA"] += 1
        if ".2CTA" in op:
            cnt[cur]["UTCHMMA.2CTA"] += 1
    elif op.startswith("HMMA"):
        cnt[cur]["HMMA"] += 1
    else:
        for p in pats[2:]:
            if op.startswith(p):
                cnt[cur][p] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
print("SASS summary of iccv2025-gdl_b200/gdl_b200/libgdl_b200.so (cuobjdump -sass; sm_100a only)")
print("columns: instructions | UTCHMMA (of which .2CTA) | LDTM | UTMALDG | UTMASTG | UTCBAR | SYNCS | LDGSTS | HMMA (legacy)\n")
tot = collections.Counter()
for (k, c), name in zip(cnt.items(), demangle):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    print("%-86s %6d | %4d (%3d) | %3d | %3d | %3d | %3d | %3d | %3d | %d" % (
        name[:86], c["_instr"], c["UTCHMMA"], c["UTCHMMA.2CTA"], c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UTCBAR"],
        c["SYNCS"], c["LDGSTS"], c["HMMA"]))
    tot.update(c)
print("\nTOTAL over %d kernels: UTCHMMA %d (2CTA %d), LDTM %d, UTMALDG %d, UTMASTG %d, UTCBAR %d, HMMA %d" % (
    len(cnt), tot["UTCHMMA"], tot["UTCHMMA.2CTA"], tot["LDTM"], tot["UTMALDG"], tot["UTMASTG"], tot["UTCBAR"], tot["HMMA"]))
