"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python profiles/summarize_launches.py gpurun_out/launches.csv [--last-step] > profiles/<name>.md
--last-step: only the launches of the last complete train step of the capture (everything after the second-to-last
accx::adam_flat_kernel up to and including the last one)."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, ns))
if "--last-step" in sys.argv:
    ends = [i for i, (n, _) in enumerate(rows) if "adam_flat_kernel" in n]
    if len(ends) >= 2:
        rows = rows[ends[-2] + 1:ends[-1] + 1]
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot / 1e6:.2f} ms total (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{n[:90]}` | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |")
