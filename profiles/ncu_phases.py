"""Split the SASS of one kernel at its BAR.SYNC instructions and report, per phase, the share of stall samples, the instruction
mix and the dominant stall reasons (SASS order follows program order in these straight-line kernels).
usage: python profiles/ncu_phases.py <report.ncu-rep> <kernel-regex>"""
import collections, csv, io, re, subprocess, sys

src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
d = rows[2:]
def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
phase = 0
agg = collections.defaultdict(collections.Counter)
ST = ("wait", "long_sb", "short_sb", "mio", "barrier", "no_inst", "lg", "math", "selected")
seen = set()
for r in d:
    if r[ix["Address"]] in seen:  # a report with several captures lists the kernel once per capture: keep the first
        break
    seen.add(r[ix["Address"]])
    s = r[ix["Source"]]
    a = agg[phase]
    a["samples"] += f(r, "# Samples")
    a["inst"] += f(r, "Instructions Executed")
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", s)
    op = m.group(2).split(".")[0] if m else "?"
    if op in ("DFMA", "DMUL", "DADD"):
        a["fp64"] += f(r, "Instructions Executed")
    if op in ("LDS", "STS"):
        a["smem"] += f(r, "Instructions Executed")
    if op in ("LDG", "STG", "LDGSTS"):
        a["glob"] += f(r, "Instructions Executed")
    for k in ST:
        a[k] += f(r, "stall_" + k)
    if "BAR.SYNC" in s:
        phase += 1
tot = sum(a["samples"] for a in agg.values())
for p, a in agg.items():
    n = max(a["samples"], 1)
    print(f"phase {p}: {a['samples'] / tot * 100:5.1f}% of samples | inst {a['inst'] / 1e6:7.1f}M fp64 {a['fp64'] / 1e6:7.1f}M smem {a['smem'] / 1e6:6.1f}M glob {a['glob'] / 1e6:5.1f}M | "
          + " ".join(f"{k} {a[k] / n * 100:4.1f}" for k in ST))
