"""Profiling target (not a test): one eager forward + backward of ACC_UNet(3,1,32) on 16x3x224x224 (bf16); the CUDA
profiler range covers the chosen phase only, so that
    ncu --profile-from-start off -k regex:<kernels> -c N python tests/prof_step.py [fwd|bwd]
captures the first N matching launches of that phase (backward starts at the level-1 decoder: the largest shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]
import accx  # noqa: E402
from accx import engine as E  # noqa: E402
from accx.train import dice_bce_loss  # noqa: E402

phase = sys.argv[1] if len(sys.argv) > 1 else "bwd"
dev = torch.device("cuda", 0)
torch.manual_seed(2)
model = accx.ACC_UNet(3, 1, 32, compute_dtype=torch.bfloat16).to(dev).train()
model.last_activation = None
g = torch.Generator().manual_seed(100)
x = torch.randn(16, 3, 224, 224, generator=g).to(dev)
m = (torch.rand(16, 1, 224, 224, generator=g) > 0.5).float().to(dev)
E.SIDE_MODE, E.LANES = 0, 0                 # one stream: kernels run alone
for it in range(2):
    model.zero_grad(set_to_none=True)
    prof = it == 1
    if prof and phase == "fwd":
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    loss = dice_bce_loss(model(x), m)
    if prof and phase == "fwd":
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
    if prof and phase == "bwd":
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    loss.backward()
    if prof and phase == "bwd":
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
torch.cuda.synchronize()
print("done", float(loss))
