"""Launch-geometry sweeps for the HBM-bound channel-lane kernels (accx_set_knob): each kernel alone, back-to-back
launches over rotating buffers larger than L2, CUDA-event timing -> GB/s per (shape, knob setting).  Diagnostic
tool (not a test):   python tests/bench_knobs.py [se_squeeze se_apply se_bwd_reduce se_bwd_apply]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from accx import _lib  # noqa: E402
from accx import engine as E  # noqa: E402

DEV = "cuda"
KNOBS = {n: i for i, n in enumerate(
    ["SE_SQUEEZE_BLOCKS", "SE_SQUEEZE_U", "SE_APPLY_BLOCKS", "SE_APPLY_STATS_BLOCKS", "SE_APPLY_U", "SE_BWD_REDUCE_BLOCKS",
     "SE_BWD_REDUCE_U", "SE_BWD_APPLY_BLOCKS", "SE_BWD_APPLY_BN_BLOCKS", "SE_BWD_APPLY_U", "BN_REDUCE_BLOCKS", "EW_BLOCKS",
     "POOL_BLOCKS", "TC_SMEM_KB", "TC_MAX_STAGES", "WGRAD_MIN_STAGES", "WGRAD_SMEM_KB", "SE_BWD_VEC"])}
SHAPES = [(16, 50176, 32), (16, 12544, 64), (16, 3136, 128), (16, 784, 256)]
ROT = 4


def set_knob(name, v):
    _lib.call("accx_set_knob", KNOBS[name], int(v))


def timeit(fn, iters=24):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3          # us


def bufs(B, HW, C, n):
    return [torch.randn(B, HW, C, device=DEV).to(torch.bfloat16) for _ in range(n)]


def run(kernel, B, HW, C):
    st = E.stream()
    x = bufs(B, HW, C, ROT)
    d = bufs(B, HW, C, ROT)
    o = bufs(B, HW, C, ROT)
    sc, sh = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    gate = torch.rand(B * C, device=DEV)
    S = torch.zeros(2 * B * C, device=DEV)
    G = torch.zeros(2 * B * C, device=DEV)
    PQR = torch.randn(3 * B * C, device=DEV)
    stats = torch.zeros(2 * C, device=DEV)
    mean, rstd = torch.randn(C, device=DEV), torch.rand(C, device=DEV) + 0.5
    p = E.ptr
    nbytes = x[0].numel() * 2
    if kernel == "se_squeeze":
        f = lambda i: _lib.call("accx_se_squeeze", E.BF16, B, HW, C, p(x[i % ROT]), p(sc), p(sh), 2, p(S), st)
        traffic = nbytes
    elif kernel == "se_apply":
        f = lambda i: _lib.call("accx_se_apply", E.BF16, B, HW, C, p(x[i % ROT]), p(sc), p(sh), 2, p(gate), p(sc), p(sh), 0, 0,
                                p(o[i % ROT]), 0, st)
        traffic = 2 * nbytes
    elif kernel == "se_apply_res_stats":
        f = lambda i: _lib.call("accx_se_apply", E.BF16, B, HW, C, p(x[i % ROT]), p(sc), p(sh), 2, p(gate), p(sc), p(sh),
                                p(d[i % ROT]), 0, p(o[i % ROT]), p(stats), st)
        traffic = 3 * nbytes
    elif kernel == "se_bwd_reduce":
        f = lambda i: _lib.call("accx_se_bwd_reduce", E.BF16, B, HW, C, p(x[i % ROT]), p(sc), p(sh), 2, p(gate), p(sc), p(sh),
                                p(d[i % ROT]), 0, 0, 0, p(G), st)
        traffic = 2 * nbytes
    elif kernel == "se_bwd_apply":
        f = lambda i: _lib.call("accx_se_bwd_apply", E.BF16, B, HW, C, p(x[i % ROT]), p(sc), p(sh), 2, p(gate), p(sc), p(sh),
                                p(d[i % ROT]), 0, p(PQR), p(o[i % ROT]), 0, p(mean), p(rstd), p(stats), st)
        traffic = 3 * nbytes
    elif kernel == "bn_bwd_reduce":
        f = lambda i: _lib.call("accx_bn_bwd_reduce", E.BF16, B * HW, C, p(x[i % ROT]), p(sc), p(sh), 2, p(mean), p(rstd),
                                p(d[i % ROT]), p(stats), st)
        traffic = 2 * nbytes
    else:
        raise SystemExit(f"unknown kernel {kernel}")
    return f, traffic


SWEEPS = {
    "se_squeeze": ("SE_SQUEEZE_BLOCKS", "SE_SQUEEZE_U"),
    "se_apply": ("SE_APPLY_BLOCKS", "SE_APPLY_U"),
    "se_apply_res_stats": ("SE_APPLY_STATS_BLOCKS", "SE_APPLY_U"),
    "se_bwd_reduce": ("SE_BWD_REDUCE_BLOCKS", "SE_BWD_REDUCE_U"),
    "se_bwd_apply": ("SE_BWD_APPLY_BN_BLOCKS", "SE_BWD_APPLY_U"),
    "bn_bwd_reduce": ("BN_REDUCE_BLOCKS", None),
}


def main():
    _lib.load()
    if os.environ.get("SE_BWD_VEC"):
        set_knob("SE_BWD_VEC", int(os.environ["SE_BWD_VEC"]))
    kernels = sys.argv[1:] or list(SWEEPS)
    for k in kernels:
        kb, ku = SWEEPS[k]
        for (B, HW, C) in SHAPES:
            f, traffic = run(k, B, HW, C)
            row = []
            for blocks in (0, 2, 4, 8, 16):
                for u in ((0, 4, 8, 16) if ku else (0,)):
                    set_knob(kb, blocks)
                    if ku:
                        set_knob(ku, u)
                    t = timeit(f)
                    row.append((t, blocks, u))
            set_knob(kb, 0)
            if ku:
                set_knob(ku, 0)
            base = row[0][0]
            best = min(row)
            print(f"{k:20s} B={B} HW={HW} C={C}: default {base:6.1f} us {traffic / base / 1e3:5.0f} GB/s | best {best[0]:6.1f} us "
                  f"{traffic / best[0] / 1e3:5.0f} GB/s at blocks={best[1]} U={best[2]} | "
                  + " ".join(f"b{b}u{u}:{t:.0f}" for t, b, u in row), flush=True)


if __name__ == "__main__":
    main()
