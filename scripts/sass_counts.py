"""profiles/<tag>_sass_counts.txt: per kernel of libavdf_sm100.so, how many of the Blackwell-specific SASS instructions it
contains (cuobjdump -sass): UTCHMMA / UTCHMMA.2CTA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), UCGABAR (barrier.cluster), HMMA (mma.sync), MUFU, plus the instruction total.
    python scripts/sass_counts.py profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_visual_deepfake_detection_b200", "csrc", "libavdf_sm100.so")
PATS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UCGABAR", "HMMA", "MUFU", "SYNCS", "ELECT", "ACQBULK"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, total, name = collections.OrderedDict(), {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*$", "", name).replace("avdf::", "")
        counts[name] = collections.Counter(); total[name] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        total[name] += 1
        for p in PATS:
            if op == p or op.startswith(p + "."):
                if p == "UTCHMMA" and ".2CTA" in op:
                    continue
                counts[name][p] += 1
                break
rows = []
for k, c in counts.items():
    if not any(c.values()) and total[k] < 400:
        continue
    rows.append("%-78s %6d instr  " % (k[:78], total[k]) + "  ".join("%s=%d" % (p, c[p]) for p in PATS if c[p]))
text = "SASS instruction counts per kernel of libavdf_sm100.so (cuobjdump -sass; sm_100a)\n" + "\n".join(rows) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
