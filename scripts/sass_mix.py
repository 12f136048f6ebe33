"""Instruction mix of the stage kernels in the built library, from cuobjdump -sass (no GPU needed):

    python scripts/sass_mix.py [pattern] [--md profiles/r02_sass_mix.md]

Columns: UBLKCP = bulk-TMA copies (cp.async.bulk), SYNCS = mbarrier operations, DFMA/DMUL/DADD = FP64 pipe,
LDS = shared loads, LD = generic loads (facet neighbours: shared or global), LDG/STG = global, BAR = CTA barriers.
What must be true for the design in DESIGN.md: every stage kernel has UBLKCP + SYNCS (tiles arrive by TMA on
mbarriers), the reference-element matrices appear as DFMA immediates (no LDC / LDG of matrix entries), and there is
no tensor-core instruction (no FP64 tcgen05 path exists; DMMA evaluated and rejected).
"""
import collections
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
pat = args[0] if args else "stage_"
md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
txt = subprocess.run(["cuobjdump", "-sass", "seigen_b200/libseigen_b200.so"], capture_output=True, text=True).stdout
name, mix = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        mix[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        mix[name][m.group(1)] += 1
names = [n for n in mix if pat in n]
dem = dict(zip(names, subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines())) if names else {}
keys = ["UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DMMA", "LDS", "LD", "LDG", "STG", "ST", "LDC", "BAR", "ATOMG", "RED"]
rows = []
for n in names:
    c = mix[n]
    d = dem[n].replace("void sg::", "").replace("(sg::StageParams)", "")
    tc = sum(v for k, v in c.items() if k.startswith(("UTC", "HMMA", "DMMA", "IMMA", "QMMA")))
    rows.append((d, sum(c.values()), [c[k] for k in keys], tc))
rows.sort()
if md:
    with open(md, "w") as f:
        f.write("# SASS instruction mix of the stage kernels (cuobjdump -sass seigen_b200/libseigen_b200.so)\n\n")
        f.write(__doc__.split("Columns:")[1].strip().replace("\n", " ") + "\n\n")
        f.write("Template arguments: `stage_f_kernel<D, P, TILE, SPLIT, MINB, NS, AXPY, AXS, SYM>`, "
                "`stage_g_kernel<D, P, TILE, SPLIT, MINB, NS, AXPY, AXS, XREG, SYM>`; SYM = 1 rows only "
                "(symmetric stress storage, the default).\n\n")
        f.write("| kernel | total | " + " | ".join(keys) + " | tensor-core |\n|---|---:|" + "---:|" * (len(keys) + 1) + "\n")
        for d, tot, vals, tc in rows:
            if not d.rstrip(">").rstrip().endswith("true"):
                continue
            f.write(f"| `{d}` | {tot} | " + " | ".join(str(v) for v in vals) + f" | {tc} |\n")
    print("wrote", md)
else:
    print("%-80s %6s " % ("kernel", "total") + " ".join("%6s" % k for k in keys))
    for d, tot, vals, tc in rows:
        print("%-80s %6d " % (d[-80:], tot) + " ".join("%6d" % v for v in vals))
