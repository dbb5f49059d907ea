"""Micro-benchmark (not a test): accx_hanc_unpool_bnred at the model's shapes, channels per thread by knob 24 (4 or 2),
CUDA-event timed with an L2 flush between launches; also checks that the variants agree.
    python tests/bench_unpool.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]
from accx import _lib, engine as E  # noqa: E402

SHAPES = [(16, 224, 224, 96, 2), (16, 224, 224, 192, 2), (16, 112, 112, 384, 2), (16, 56, 56, 4352, 2), (16, 56, 56, 768, 2),
          (16, 28, 28, 1536, 1)]


def run(shape, vec, reps=5):
    B, H, W, C, levels = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    y = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    da0 = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.bfloat16)
    sc, sh = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g) * 0.3
    mean, rstd = torch.randn(C, device="cuda", generator=g) * 0.2, torch.rand(C, device="cuda", generator=g) + 0.5
    L = E.Lazy(y, sc, sh, 2, mean, rstd, None)
    dps = [torch.randn(B, H >> l, W >> l, 2 * C, device="cuda", generator=g) for l in range(1, levels + 1)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _lib.call("accx_set_knob", 24, vec)
    ts, out = [], None
    for _ in range(reps + 2):
        da = da0.clone()
        ar = E.Arena(y.device)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sums = E.hanc_unpool_bnred(L, dps, da, ar)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        out = (da, sums.clone())
    _lib.call("accx_set_knob", 24, 0)
    ms = sorted(ts[2:])[len(ts[2:]) // 2]
    byt = 3 * y.numel() * 2 + sum(d.numel() * 4 for d in dps)
    print(f"unpool_bnred {B}x{H}x{W}x{C} levels={levels} channels/thread={vec}: {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s", flush=True)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":      # ncu target: the library's default geometry on one shape
        run(tuple(int(v) for v in sys.argv[2:7]), 0, reps=2)
        sys.exit(0)
    for shape in SHAPES:
        a, b = run(shape, 4), run(shape, 2)
        assert torch.equal(a[0], b[0]), "gradient differs between the variants"
        assert torch.allclose(a[1], b[1], rtol=1e-3, atol=1e-2), "sums differ"
        if shape[4] == 2:
            c = run(shape, 3)      # two channels per thread held to 128 registers (four blocks per SM)
            assert torch.equal(a[0], c[0]) and torch.allclose(a[1], c[1], rtol=1e-3, atol=1e-2)
