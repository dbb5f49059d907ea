"""Micro-benchmark (not a test): depthwise 3x3 kernels on model shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]
from accx import engine as E  # noqa: E402

SHAPES = [(16, 224, 224, 96), (16, 112, 112, 192), (16, 56, 56, 384), (16, 56, 56, 4352), (16, 28, 28, 768), (16, 14, 14, 1536)]


def timed(fn, reps=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


for (B, H, W, C) in SHAPES:
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    L = E.Lazy(x, torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, 2)
    w = torch.randn(C, 1, 3, 3, device="cuda")
    b = torch.randn(C, device="cuda")
    st = torch.zeros(2 * C, device="cuda")
    gw = torch.zeros_like(w)
    byt = 2 * x.numel() * 2
    f = timed(lambda: E.dw_fwd(L, w, b, st))
    g = timed(lambda: E.dw_wgrad(L, dy, gw))
    print(f"{B}x{H}x{W}x{C}: fwd {f * 1e3:8.1f} us {byt / f / 1e6:7.0f} GB/s | wgrad {g * 1e3:8.1f} us {byt / g / 1e6:7.0f} GB/s", flush=True)
