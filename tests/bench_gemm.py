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


if __name__ == "__main__":
    args = sys.argv[1:]
    wg = "wgrad" in args
    idx = [int(a) for a in args if a != "wgrad"] or range(len(SHAPES))
    for i in idx:
        (run_wgrad if wg else run)(i)
