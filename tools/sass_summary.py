#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libwann_b200.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus registers.  Writes profiles/r02_sass_summary.txt."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "wann_b200", "libwann_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\n\s*REG:(\d+)", res))
pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "UTCATOMSWS", "HMNMX2", "F2FP", "STS.128", "LDS.128", "BAR.SYNC"]
out = ["SASS summary of %s (sm_100a), built from the sources of this commit" % os.path.relpath(so, ROOT),
       "mnemonic counts per kernel (static instruction counts; UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,",
       "UBLKCP = cp.async.bulk (TMA engine, 1-D), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops)", ""]
cur, counts = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = dict.fromkeys(pats, 0); counts[cur]["lines"] = 0
        continue
    if cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line):
        counts[cur]["lines"] += 1
        for p in pats:
            if re.search(r"\b" + re.escape(p) + (r"\b" if "." not in p else ""), line):
                counts[cur][p] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
for (name, c), d in zip(counts.items(), dem or list(counts)):
    out.append(d)
    out.append("    instructions %d, registers %s" % (c["lines"], regs.get(name, "?")))
    out.append("    " + "  ".join("%s=%d" % (p, c[p]) for p in pats if c[p]))
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
