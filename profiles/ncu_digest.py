"""Digest of an .ncu-rep (run where ncu is installed): per captured launch -- name, duration, DRAM bytes, issue-slot
utilisation, registers, executed warp instructions, opcode histogram and stall-reason totals (source page).
    python profiles/ncu_digest.py report.ncu-rep [max_launches]"""
import csv
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
lim = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
ix = {n: i for i, n in enumerate(h)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for n, r in enumerate(rows[2:2 + lim]):
    print(f"==== launch {n}: {r[ix['Kernel Name']][:100]}  grid {r[ix.get('launch__grid_size', 0)]}")
    for k in KEYS:
        if k in ix:
            print(f"   {k:62s} {r[ix[k]]:>16s} {rows[1][ix[k]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tables, cur = [], None
for line in src.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        tables.append(cur)
    elif cur is not None:
        cur.append(line)
for n, tab in enumerate(tables[:lim]):
    sr = list(csv.reader(tab))
    try:
        hi = next(i for i, r in enumerate(sr) if "Source" in r and "Address" in r)
    except StopIteration:
        continue
    sh = sr[hi]
    sx = {k: i for i, k in enumerate(sh)}
    body = [r for r in sr[hi + 1:] if len(r) == len(sh)]
    E, S = sx["Instructions Executed"], sx["# Samples"]
    tot = sum(int(r[E] or 0) for r in body) or 1
    c = Counter()
    for r in body:
        t = [o for o in r[sx["Source"]].split() if not o.startswith("@")]
        c[t[0].split(".")[0] if t else "?"] += int(r[E] or 0)
    stall = [k for k in sh if k.startswith("stall_") and "Not Issued" not in k]
    d = sorted(((sum(int(r[sx[k]] or 0) for r in body), k[6:]) for k in stall), reverse=True)[:6]
    print(f"---- launch {n}: {sr[0][1][:80] if len(sr[0]) > 1 else ''}  warp instructions {tot}")
    print("   opcodes: " + "  ".join(f"{o} {100 * v / tot:.1f}%" for o, v in c.most_common(14)))
    print("   stalls:  " + "  ".join(f"{k} {v}" for v, k in d))
