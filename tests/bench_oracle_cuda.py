"""Informational (not a test, not the product): the oracle restatement of the reference step executed by torch
eager on the GPU (cuDNN / ATen kernels) -- the number a user of the unmodified reference would see on the same
B200 ("informational second baseline").  fp32 and bf16 autocast.
    python tests/bench_oracle_cuda.py [batch]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from oracle import acc_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda"
for mode in ("fp32", "bf16-autocast"):
    torch.manual_seed(2)
    sd = {k: v.to(dev) for k, v in O.init_acc_unet(3, 1, 32).items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    m = (torch.rand(B, 1, 224, 224, generator=g) > 0.5).float().to(dev)
    opt = None
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
            for _ in range(2):
                _, opt = O.train_step(sd, x, m, opt)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                _, opt = O.train_step(sd, x, m, opt)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"torch eager on B200, {mode}, batch {B}: {dt * 1e3:.1f} ms/step = {B / dt:.1f} images/s, "
              f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"torch eager on B200, {mode}, batch {B}: failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
    del sd, opt
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
