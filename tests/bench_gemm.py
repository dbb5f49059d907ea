"""Micro-benchmark (not a test): accx pointwise contractions on model shapes, CUDA-event timed.
    python tests/bench_gemm.py [shape_index ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]
from accx import engine as E  # noqa: E402

SHAPES = [  # (P, K, N, act, stats)
    (50176 * 16, 32, 96, 0, 1), (50176 * 16, 96, 32, 2, 1), (50176 * 16, 32, 32, 1, 1),
    (12544 * 16, 64, 192, 0, 1), (12544 * 16, 192, 64, 2, 1),
    (3136 * 16, 128, 4352, 0, 1), (3136 * 16, 4352, 128, 2, 1), (3136 * 16, 128, 384, 0, 1), (3136 * 16, 384, 128, 2, 1),
    (784 * 16, 256, 768, 0, 1), (784 * 16, 768, 256, 2, 1), (196 * 16, 512, 1536, 0, 1), (196 * 16, 1536, 512, 2, 1),
]


def run_graph(i, n=20):
    """the same launch 20 times inside one CUDA graph (no launch-side CPU time between kernels), L2-warm inputs"""
    P, K, N, act, st = SHAPES[i]
    x = torch.randn(1, 1, P, K, device="cuda").to(torch.bfloat16)
    s = torch.rand(K, device="cuda") + 0.5
    t = torch.randn(K, device="cuda") * 0.1
    L = E.Lazy(x, s, t, act) if act else E.Lazy(x)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    stats = torch.zeros(2 * N, device="cuda") if st else None
    y = torch.empty(1, 1, P, N, device="cuda", dtype=torch.bfloat16)
    ops = [E.Op(L, K, E.WV(w, 0, K, 1))]
    for _ in range(3):
        E.conv(ops, N, (1, 1, P), stats=stats, out=y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            E.conv(ops, N, (1, 1, P), stats=stats, out=y)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    byt = P * (K + N) * 2
    print(f"graph P={P:8d} K={K:5d} N={N:5d} act={act}: {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s", flush=True)


def run(i, reps=5):
    P, K, N, act, st = SHAPES[i]
    x = torch.randn(1, 1, P, K, device="cuda").to(torch.bfloat16)
    s = torch.rand(K, device="cuda") + 0.5
    t = torch.randn(K, device="cuda") * 0.1
    L = E.Lazy(x, s, t, act) if act else E.Lazy(x)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    stats = torch.zeros(2 * N, device="cuda") if st else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = E.conv([E.Op(L, K, E.WV(w, 0, K, 1))], N, (1, 1, P), stats=stats)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    byt = P * (K + N) * 2
    print(f"P={P:8d} K={K:5d} N={N:5d} act={act}: {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s  {2 * P * K * N / ms / 1e9:7.1f} TFLOP/s",
          flush=True)


def run_wgrad(i, reps=5):
    """weight gradient of shape i: dW[N, K] = dY[P, N]^T . act(A[P, K])"""
    P, K, N, act, st = SHAPES[i]
    x = torch.randn(1, 1, P, K, device="cuda").to(torch.bfloat16)
    dy = torch.randn(1, 1, P, N, device="cuda").to(torch.bfloat16)
    s = torch.rand(K, device="cuda") + 0.5
    t = torch.randn(K, device="cuda") * 0.1
    L = E.Lazy(x, s, t, act) if act else E.Lazy(x)
    w = torch.zeros(N, K, device="cuda")
    gw = torch.zeros(N, K, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        E.wgrad(E.Op(L, K, E.WV(w, 0, K, 1)), dy, N, (1, 1, P), gw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    byt = P * (K + N) * 2
    print(f"wgrad P={P:8d} K={K:5d} N={N:5d} act={act}: {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s  {2 * P * K * N / ms / 1e9:7.1f} TFLOP/s",
          flush=True)


SHAPES3 = [(16, 224, 224, 32), (16, 112, 112, 64), (16, 56, 56, 128), (16, 28, 28, 256)]
SMALL = [(12544, 256, 256, 2, 1), (12544, 256, 768, 0, 1), (3136, 256, 768, 0, 1), (12544, 256, 128, 2, 1), (3136, 512, 1536, 0, 1),
         (50176, 128, 128, 2, 1), (200704, 64, 64, 2, 1), (802816, 32, 32, 2, 1)]


def run_taps(i, reps=5, wgrad=False):
    """ResPath's dense 3x3 conv (ACC_UNet.py:316-318) at model level i: forward contraction / weight gradient"""
    from accx.modules import ResPath
    B, H, W, C = SHAPES3[i]
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    s = torch.rand(C, device="cuda") + 0.5
    t = torch.randn(C, device="cuda") * 0.1
    L = E.Lazy(x, s, t, 2)
    w = torch.randn(C, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    stats = torch.zeros(2 * C, device="cuda")
    dy = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    gw = torch.zeros_like(w)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if wgrad:
            E.wgrad_conv3x3(L, C, w, dy, C, (B, H, W), gw)
        else:
            E.conv(ResPath._taps(L, w, C), C, (B, H, W), stats=stats)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    P = B * H * W
    print(f"conv3x3{' wgrad' if wgrad else ''} {B}x{H}x{W}x{C}: {ms * 1e3:8.1f} us  {P * C * 4 / ms / 1e6:7.0f} GB/s  "
          f"{2 * P * C * C * 9 / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    if "taps" in args:
        for i in range(len(SHAPES3)):
            run_taps(i)
            run_taps(i, wgrad=True)
        sys.exit(0)
    if "debug" in args:          # pw_fwd_tc with parts of the kernel switched off (knob 19): where does the tile time go?
        from accx import _lib
        SHAPES[:] = [(802816, 32, 32, 2, 1), (802816, 96, 32, 2, 1), (802816, 32, 96, 0, 1), (200704, 64, 64, 2, 1), (12544, 256, 256, 2, 1)]
        for dbg, what in ((0, "full kernel"), (1, "no transform math"), (2, "no statistics pass"), (4, "no TMA store"),
                          (8, "no TMEM drain / staging"), (15, "pipeline only")):
            _lib.call("accx_set_knob", 19, dbg)
            print(f"---- {what}")
            for i in range(len(SHAPES)):
                run_graph(i)
        _lib.call("accx_set_knob", 19, 0)
        sys.exit(0)
    if "trace" in args:          # role timeline of CTA 0 (knob 19 bit 5): who waits for whom, tile by tile
        import ctypes
        from accx import _lib
        SHAPES[:] = [(802816, 32, 32, 2, 1), (802816, 64, 64, 2, 1)]
        names = ["tma_issue", "xf_landed", "xf_arrive", "mma_full", "mma_commit", "mma_acc", "epi_tfull", "epi_release",
                 "epi_store", "epi_stats"]
        for i in range(len(SHAPES)):
            P, K, N, act, st = SHAPES[i]
            x = torch.randn(1, 1, P, K, device="cuda").to(torch.bfloat16)
            L = E.Lazy(x, torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda") * 0.1, act)
            w = torch.randn(N, K, device="cuda") / K ** 0.5
            stats = torch.zeros(2 * N, device="cuda")
            ops = [E.Op(L, K, E.WV(w, 0, K, 1))]
            for _ in range(3):
                E.conv(ops, N, (1, 1, P), stats=stats)
            _lib.call("accx_set_knob", 19, 32)
            E.conv(ops, N, (1, 1, P), stats=stats)
            _lib.call("accx_set_knob", 19, 0)
            buf = (ctypes.c_ulonglong * 640)()
            _lib.call("accx_debug_tc_trace", buf, 640)
            t0 = min(v for v in buf if v)
            print(f"==== K={K} N={N}: ns since the first stamp, tiles 0..13 of CTA 0 (then 30..33)")
            for tl in list(range(14)) + list(range(30, 34)):
                print(f"tile {tl:2d}: " + "  ".join(f"{names[e]} {buf[e * 64 + tl] - t0:6d}" for e in range(10)))
        sys.exit(0)
    if "ring" in args:           # accumulator ring (knob 21) x parts of the kernel off (knob 19)
        from accx import _lib
        SHAPES[:] = [(802816, 32, 32, 2, 1), (802816, 64, 64, 2, 1), (802816, 32, 96, 0, 1)]
        for nacc in (2, 4):
            _lib.call("accx_set_knob", 21, nacc)
            for dbg in (0, 1, 2, 3, 8, 10, 11, 15):
                _lib.call("accx_set_knob", 19, dbg)
                print(f"---- ring {nacc} debug {dbg}")
                for i in range(len(SHAPES)):
                    run_graph(i)
        _lib.call("accx_set_knob", 19, 0)
        _lib.call("accx_set_knob", 21, 0)
        sys.exit(0)
    if "stages" in args:         # sensitivity of the narrow tile rate to the pipeline depth (knob 14 = TC_MAX_STAGES)
        from accx import _lib
        SHAPES[:] = [(802816, 32, 32, 2, 1), (802816, 64, 64, 2, 1), (802816, 128, 32, 2, 1)]
        for dbg in (15, 0):
            _lib.call("accx_set_knob", 19, dbg)
            for S in (1, 2, 3, 4, 8):
                _lib.call("accx_set_knob", 14, S)
                print(f"---- debug {dbg} stages <= {S}")
                for i in range(len(SHAPES)):
                    run_graph(i)
        _lib.call("accx_set_knob", 19, 0)
        _lib.call("accx_set_knob", 14, 0)
        sys.exit(0)
    if "pack" in args:           # share of the in-kernel weight packing (knob 19 bit 4 skips it; results are then wrong)
        from accx import _lib
        SHAPES[:] = SMALL + [(12544, 128, 384, 0, 1), (3136, 256, 256, 2, 1), (50176, 128, 256, 2, 1), (200704, 192, 64, 2, 1)]
        for dbg, what in ((0, "full kernel"), (16, "weights not packed")):
            _lib.call("accx_set_knob", 19, dbg)
            print(f"---- {what}")
            for i in range(len(SHAPES)):
                run_graph(i)
        _lib.call("accx_set_knob", 19, 0)
        sys.exit(0)
    if "fold" in args:           # pixel folding of narrow contiguous contractions (knob 23: 1 = off, default on)
        from accx import _lib
        SHAPES[:] = [(802816, 32, 32, 2, 1), (802816, 32, 32, 0, 1), (802816, 96, 32, 2, 1), (802816, 32, 64, 0, 1),
                     (802816, 64, 32, 2, 1), (802816, 32, 96, 0, 1), (200704, 32, 64, 2, 1), (200704, 64, 32, 2, 1),
                     (200704, 64, 64, 2, 1)]
        SHAPES += [(200704, 32, 96, 0, 1), (200704, 128, 64, 2, 1), (200704, 64, 128, 2, 1), (50176, 64, 64, 2, 1), (50176, 128, 128, 2, 1)]
        for f in (1, 2, 3):
            _lib.call("accx_set_knob", 23, f)
            print(f"---- pixel folding {('off', 'on', 'on, wide tensors too')[f - 1]}")
            for i in range(len(SHAPES)):
                run_graph(i)
        _lib.call("accx_set_knob", 23, 0)
        sys.exit(0)
    if "small" in args:
        SHAPES[:] = SMALL
        args.remove("small")
    wg = "wgrad" in args
    idx = [int(a) for a in args if a != "wgrad"] or range(len(SHAPES))
    for i in idx:
        (run_wgrad if wg else run)(i)
