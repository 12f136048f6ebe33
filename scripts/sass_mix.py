"""Instruction mix of the stage kernels in the built library (development aid): python scripts/sass_mix.py <pattern>"""
import collections
import re
import subprocess
import sys

pat = sys.argv[1] if len(sys.argv) > 1 else "stage_"
txt = subprocess.run(["cuobjdump", "-sass", "seigen_b200/libseigen_b200.so"], capture_output=True, text=True).stdout
name, mix = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        mix[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        mix[name][m.group(1)] += 1
keys = ["DFMA", "DMUL", "DADD", "LDS", "LD", "LDG", "STG", "ST", "LDC", "IMAD", "LEA", "BRA", "BAR", "SYNCS"]
print("%-72s %6s " % ("kernel", "total") + " ".join("%5s" % k for k in keys))
for n, c in mix.items():
    if pat in n:
        print("%-72s %6d " % (n[-72:], sum(c.values())) + " ".join("%5d" % c[k] for k in keys))
