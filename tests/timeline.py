"""Timeline of one CUDA-graph replay of the bench step (CUPTI through torch.profiler): per-kernel start / end /
stream, GPU busy vs idle time, concurrency, the longest kernels and the largest gaps.  Diagnostic tool, not a test:
    python tests/timeline.py [--batch 16] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    args = ap.parse_args()
    import accx
    from accx.train import TrainStep
    from torch.profiler import ProfilerActivity, profile

    dev = torch.device("cuda", 0)
    torch.manual_seed(2)
    model = accx.ACC_UNet(3, 1, 32, compute_dtype=torch.bfloat16).to(dev).train()
    model.last_activation = None
    step = TrainStep(model, lr=1e-3, graph=True)
    g = torch.Generator().manual_seed(100)
    x = torch.randn(args.batch, 3, args.hw, args.hw, generator=g).to(dev)
    m = (torch.rand(args.batch, 1, args.hw, args.hw, generator=g) > 0.5).float().to(dev)
    for _ in range(6):
        step(x, m)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(x, m)
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            evs.append((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", -1)
                        if hasattr(e, "device_resource_id") else -1))
    if not evs:
        # older/newer event API: fall back to the chrome trace
        path = args.out + ".trace.json"
        prof.export_chrome_trace(path)
        tr = json.load(open(path))
        for e in tr["traceEvents"]:
            if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e:
                evs.append((e["ts"], e["ts"] + e["dur"], e["name"], e.get("args", {}).get("stream", -1)))
    evs.sort()
    t0 = evs[0][0]
    t1 = max(e[1] for e in evs)
    # union of busy intervals and concurrency-weighted time
    pts = []
    for s, e, _, _ in evs:
        pts.append((s, 1))
        pts.append((e, -1))
    pts.sort()
    busy, cur, last, conc = 0.0, 0, t0, {}
    for t, d in pts:
        if cur > 0:
            busy += t - last
        conc[cur] = conc.get(cur, 0.0) + (t - last)
        cur += d
        last = t
    # gaps (GPU fully idle)
    gaps, end = [], evs[0][1]
    prev_name = evs[0][2]
    for s, e, n, _ in evs[1:]:
        if s > end:
            gaps.append((s - end, prev_name, n))
        if e > end:
            end, prev_name = e, n
    gaps.sort(reverse=True)
    agg = {}
    for s, e, n, _ in evs:
        short = n.split("(")[0][:90]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += e - s
    streams = {}
    for s, e, n, st in evs:
        a = streams.setdefault(str(st), [0, 0.0])
        a[0] += 1
        a[1] += e - s
    out = {
        "span_us": t1 - t0, "busy_us": busy, "idle_us": (t1 - t0) - busy, "kernels": len(evs),
        "sum_kernel_us": sum(e - s for s, e, _, _ in evs),
        "concurrency_us": {str(k): v for k, v in sorted(conc.items())},
        "streams": streams,
        "idle_gap_sum_us": sum(g_[0] for g_ in gaps), "n_gaps": len(gaps),
        "largest_gaps": [{"us": g_[0], "after": g_[1][:80], "before": g_[2][:80]} for g_ in gaps[:25]],
        "by_kernel": [{"kernel": k, "n": a[0], "us": a[1]} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])],
        "events": [[s - t0, e - s, n.split("(")[0][:60], st] for s, e, n, st in evs],
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"))
    print(json.dumps({k: out[k] for k in ("span_us", "busy_us", "idle_us", "kernels", "sum_kernel_us", "concurrency_us",
                                          "streams", "n_gaps", "idle_gap_sum_us")}, indent=1))
    for r in out["by_kernel"][:30]:
        print(f"{r['us']:9.1f} us  n={r['n']:4d}  {r['kernel']}")


if __name__ == "__main__":
    main()
