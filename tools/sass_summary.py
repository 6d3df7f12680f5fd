"""Per-kernel count of the Blackwell tensor-core / TMA / TMEM instructions in the SHIPPED library's SASS (no GPU needed):

    python tools/sass_summary.py [--lib cusrl_b200/lib/libcusrl_b200.so] > profiles/r02_sass_summary.md

`cuobjdump -sass` of libcusrl_b200.so, one row per kernel that contains at least one of the mnemonics
(/opt/skills/guides/B200_PROFILING.md lists what proves what): UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA bulk tensor
load / store (.MULTICAST = one load feeding every CTA of a cluster), UTMAPF = TMA prefetch into L2, LDTM = tcgen05.ld (TMEM ->
registers), UTCBAR = tcgen05.commit onto an mbarrier, UTCATOMSWS = TMEM allocation, SYNCS = mbarrier operations."""
import argparse
import re
import subprocess
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=str(ROOT / "cusrl_b200" / "lib" / "libcusrl_b200.so"))
args = ap.parse_args()

sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLUMNS = OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG(?![.\w]*MULTICAST)"),
    ("UTMALDG.MULTICAST", r"\bUTMALDG[.\w]*MULTICAST"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF"), ("LDTM", r"\bLDTM"),
    ("UTCBAR", r"\bUTCBAR"), ("UTCATOMSWS", r"\bUTCATOMSWS"), ("SYNCS", r"\bSYNCS")])
rows = []
total = Counter()
kernels = 0
for i, body in enumerate(re.split(r"\n\s*Function : ", sass)[1:]):
    kernels += 1
    counts = {k: len(re.findall(rx, body)) for k, rx in COLUMNS.items()}
    if not any(counts[k] for k in COLUMNS if k != "SYNCS"):
        continue
    name = names[i] if i < len(names) else body.split("\n", 1)[0]
    name = re.sub(r"\(.*", "", name).replace("cusrl_b200::", "").replace("void ", "")
    rows.append((name, counts))
    total.update(counts)
print("# SASS evidence: tcgen05 / TMA / TMEM instructions per kernel of the shipped `libcusrl_b200.so`\n")
print("    python tools/sass_summary.py      # cuobjdump -sass, sm_100a cubin only; no GPU needed\n")
print(f"{kernels} kernels in the library, {len(rows)} of them issue tensor-core, TMA or TMEM instructions "
      f"(the others are the HBM-bound SIMT kernels: scans, reductions, gathers, losses, optimizer).\n")
print("| kernel | " + " | ".join(COLUMNS) + " |")
print("|---|" + "---:|" * len(COLUMNS))
for name, counts in sorted(rows):
    print(f"| `{name}` | " + " | ".join(str(counts[k]) if counts[k] else "" for k in COLUMNS) + " |")
print("| **total** | " + " | ".join(str(total[k]) for k in COLUMNS) + " |")
