"""Counts the SASS mnemonics that prove tcgen05 / TMEM / TMA use (and the absence of local-memory traffic) per kernel of the
built library: python tools/sass_summary.py [tag]  ->  profiles/<tag>_sass_summary.md   (host only: cuobjdump)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "musediffusion_b200", "libmusediff_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UTCHMMA", ".2CTA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "FMNMX3", "F2FP", "LDL", "STL"]
rows = []
for name, body in zip(names, re.split(r"Function : \S+", sass)[1:]):
    ops = [m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, re.M)]
    c = collections.Counter()
    for o in ops:
        for col in COLS:
            if col == ".2CTA":
                c[col] += o.startswith("UTCHMMA") and ".2CTA" in o
            elif o.startswith(col):
                c[col] += 1
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("md::", "")
    rows.append((short, len(ops), c))
out = ["# SASS evidence (%s): `cuobjdump -sass musediffusion_b200/libmusediff_b200.so` (nvcc 12.9, sm_100a, -lineinfo; `tools/sass_summary.py`)\n" % tag,
       "tcgen05 MMA = `UTCHMMA` (`.2CTA` = cta_group::2), TMEM load / store = `LDTM` / `STTM`, TMA load / store = `UTMALDG` / `UTMASTG`, "
       "tcgen05.commit = `UTCBAR`, mbarrier = `SYNCS`.",
       "`LDL` / `STL` = local-memory traffic (must be 0 in the hot loops; the few that remain belong to the bounded-wait trap path that calls printf).\n",
       "| kernel | SASS instr | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
for short, n, c in rows:
    out.append("| `%s` | %d | %s |" % (short, n, " | ".join(str(int(c[k])) for k in COLS)))
open(os.path.join(ROOT, "profiles", "%s_sass_summary.md" % tag), "w").write("\n".join(out) + "\n")
print("%d kernels" % len(rows))
