"""Timing diagnostic (not a test): step time of the captured train step with one class of accx kernels NOT launched
(ACCX_ABLATE=entry,entry,...; numerics are then meaningless) -- the marginal cost of that class inside the overlapped graph.
    ACCX_ABLATE=accx_pw_wgrad_tc python tests/ablate.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]
import accx  # noqa: E402
from accx.train import TrainStep  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(2)
model = accx.ACC_UNet(3, 1, 32, compute_dtype=torch.bfloat16).to(dev).train()
model.last_activation = None
step = TrainStep(model, lr=1e-3, graph=True)
g = torch.Generator().manual_seed(100)
x = torch.randn(16, 3, 224, 224, generator=g).to(dev)
m = (torch.rand(16, 1, 224, 224, generator=g) > 0.5).float().to(dev)
for _ in range(6):
    step(x, m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step(x, m)
e1.record()
torch.cuda.synchronize()
print(f"{os.environ.get('ACCX_ABLATE', '-'):90s} {e0.elapsed_time(e1) / 10:7.2f} ms/step", flush=True)
