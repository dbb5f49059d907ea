"""Diagnostic: spread of the loss trajectory of the flat TrainStep and of the loose torch step over repeated
runs from the same initial weights, with and without side streams / lanes (chaos vs race)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
    sys.path.insert(0, p)
import torch
import accx
from accx import engine as E
from accx.train import TrainStep, dice_bce_loss_torch, dice_bce_loss

DEV = "cuda"
B, HW, F = int(os.environ.get("B", 4)), int(os.environ.get("HW", 64)), int(os.environ.get("F", 8))
torch.manual_seed(2)
m0 = accx.ACC_UNet(3, 1, F).to(DEV).train()
m0.last_activation = None
g = torch.Generator().manual_seed(7)
x = torch.randn(B, 3, HW, HW, generator=g).to(DEV)
m = (torch.rand(B, 1, HW, HW, generator=g) > 0.5).float().to(DEV)


def run(kind, side, lanes, n=3):
    E.SIDE_MODE, E.LANES = side, lanes
    model = copy.deepcopy(m0)
    out = []
    if kind == "flat":
        step = TrainStep(model, lr=1e-3)
        for _ in range(n):
            out.append(float(step(x, m).detach()))
    else:
        lossf = dice_bce_loss if kind == "loose-native" else dice_bce_loss_torch
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        for _ in range(n):
            l = lossf(model(x), m)
            opt.zero_grad(set_to_none=True)
            l.backward()
            opt.step()
            out.append(float(l.detach()))
    torch.cuda.synchronize()
    return out


for kind in ("flat", "loose-torch", "loose-native"):
    for side, lanes in ((1, 1), (0, 0)):
        for t in range(3):
            print(f"{kind:13s} side={side} lanes={lanes} trial {t}: " + " ".join(f"{v:.6f}" for v in run(kind, side, lanes)))
