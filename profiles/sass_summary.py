"""Per-kernel instruction-count table from the built library (no GPU needed):
    python profiles/sass_summary.py > profiles/r2_sass_summary.txt
Columns: DMMA (fp64 tensor cores; tcgen05 has no fp64 kind, so DMMA is the Blackwell fp64 tensor path), UTMALDG (TMA
tensor loads), SYNCS (mbarrier), UCGABAR / cluster barriers, DFMA, LDS / LDG, ATOM / RED."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "flgp_b200/libflgp_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = [("DMMA", r"\bDMMA\."), ("UTMALDG", r"\bUTMALDG"), ("SYNCS(mbar)", r"\bSYNCS\."), ("CGA-BAR", r"UCGABAR|BAR\.ARV.*CGA|\bCGABAR"),
        ("DFMA", r"\bDFMA\b"), ("DADD/DMUL", r"\bD(ADD|MUL)\b"), ("LDS", r"\bLDS(\.|\b)"), ("LDG", r"\bLDG\."), ("STG", r"\bSTG\."),
        ("ATOM/RED", r"\b(ATOM|ATOMG|RED)\."), ("SHFL", r"\bSHFL\."), ("BAR.SYNC", r"\bBAR\.SYNC")]
rows = []
name, cnt = None, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, cnt))
        name, cnt = m.group(1), collections.Counter()
        continue
    if name:
        for key, p in pats:
            if re.search(p, line):
                cnt[key] += 1
if name:
    rows.append((name, cnt))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("%-78s" % "kernel" + "".join("%12s" % k for k, _ in pats))
for (nm, c), d in sorted(zip(rows, dem), key=lambda t: t[1]):
    short = d.replace("(anonymous namespace)", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")
    short = re.sub(r"\(.*", "", short).replace("void ", "").replace("flgp::", "").replace("::", "")
    print("%-78s" % short[:78] + "".join("%12d" % c[k] for k, _ in pats))
