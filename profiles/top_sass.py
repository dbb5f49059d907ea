"""top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (stdin or file)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
body = rows[hi + 1:]
S, A = ix["# Samples"], ix["Source"]
stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[S] or 0) for r in body)
print("total samples", tot, "instructions", len(body), "executed", sum(int(r[ix["Instructions Executed"]] or 0) for r in body))
top = sorted(range(len(body)), key=lambda i: -int(body[i][S] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for i in sorted(top):
    r = body[i]
    st = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stall), reverse=True)[:3]
    print(f"{i:5d} {int(r[S]):6d} {100*int(r[S])/tot:5.1f}% exec={r[ix['Instructions Executed']]:>8s} {r[A].strip()[:70]:70s} {st}")
