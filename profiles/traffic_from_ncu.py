"""Mean DRAM traffic per launch of each accx kernel from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of one bench step ->
profiles/traffic.json (read by bench.py for roofline.traffic).
    python profiles/traffic_from_ncu.py gpurun_out/dram_r01.csv"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# C++ kernel name -> C ABI entry point that launches it
ABI = [("pw_fwd_tc_kernel", "accx_pw_fwd_tc_res"), ("pw_wgrad_tc_kernel", "accx_pw_wgrad_tc"),
       ("dw3x3_tiled_kernel", "accx_dw3x3"), ("bn_bwd_apply_kernel", "accx_bn_bwd_apply"),
       ("bn_bwd_reduce_kernel", "accx_bn_bwd_reduce"), ("se_bwd_apply_kernel", "accx_se_bwd_apply"),
       ("se_apply_kernel", "accx_se_apply"), ("hanc_unpool_bnred_kernel", "accx_hanc_unpool_bnred")]

lines = [l for l in open(sys.argv[1], newline="") if not l.startswith("==")]
per = defaultdict(lambda: defaultdict(float))     # launch id -> metric -> bytes
name = {}
for r in csv.DictReader(lines):
    m = r.get("Metric Name", "")
    if m.startswith("dram__bytes_"):
        per[r["ID"]][m] += float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1)
        name[r["ID"]] = r["Kernel Name"]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i, d in per.items():
    k = re.sub(r"^void ", "", name[i])
    for pat, abi in ABI:
        if pat in k:
            a = agg[abi]
            a[0] += 1
            a[1] += d.get("dram__bytes_read.sum", 0.0)
            a[2] += d.get("dram__bytes_write.sum", 0.0)
out = {k: {"launches": a[0], "dram_read_bytes_per_launch": a[1] / a[0], "dram_write_bytes_per_launch": a[2] / a[0],
           "dram_bytes_per_launch": (a[1] + a[2]) / a[0]} for k, a in agg.items()}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
