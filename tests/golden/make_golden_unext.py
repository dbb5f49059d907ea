"""Golden vectors for the UNeXt shifted tokenized-MLP block from the UNMODIFIED reference
(Experiments/nets/UNext.py; run in the build container only).

    python tests/golden/make_golden_unext.py

UNext.py imports `timm.models.layers.{DropPath, to_2tuple, trunc_normal_}` and matplotlib, neither of which is installed
here: three-symbol stubs are registered in sys.modules before the import (trunc_normal_ -> torch.nn.init.trunc_normal_,
the same algorithm timm ships).  Cases: shiftmlp and shiftedBlock on the token shapes of UNeXt at 8x3x256x256
(BASELINE configs[2]): [8, 256, 160] (16 x 16) and [8, 64, 256] (8 x 8), plus [2, 1024, 128] (32 x 32, dblock2's width).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_unext():
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert self.p == 0. or not self.training
            return x

    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.path.insert(0, os.path.join(REF, "Experiments"))          # `from utils import *`
    sys.path.insert(0, os.path.join(REF, "Experiments", "nets"))
    import UNext as U
    return U


def perturb(mod, seed):
    """non-trivial biases / LayerNorm affine (the reference initialises them to 0 / 1); tests repeat this"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif "norm" in n:
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)


def run(name, mod, shape, H, W, seed):
    """inputs / cotangents are NOT stored: x = randn(shape, seed), cot = randn(shape, seed + 1) on the CPU generator"""
    perturb(mod, seed + 2)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
    out = {"H": np.array(H), "W": np.array(W), "shape": np.array(shape), "seed": np.array(seed)}
    for k, v in mod.state_dict().items():
        out["sd/" + k] = v.detach().clone().numpy()
    xx = x.clone().requires_grad_(True)
    y = mod(xx, H, W)
    cot = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 1))
    (y * cot).sum().backward()
    out["out/0"], out["gin/0"] = y.detach().numpy(), xx.grad.numpy()
    for k, p in mod.named_parameters():
        out["gp/" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, tuple(x.shape), "->", tuple(y.shape))


def main():
    U = import_unext()
    g = lambda *s, seed=0: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
    # (the block has no cross-image coupling -- LayerNorm per token, depthwise conv per image -- so the fixtures keep
    #  the token shapes of UNeXt at 256 x 256 with a batch of 2 / 1; the full batch of 8 is checked against the oracle)
    torch.manual_seed(2); run("unext_shiftmlp_160", U.shiftmlp(160, 160), (2, 256, 160), 16, 16, 301)
    torch.manual_seed(2); run("unext_shiftedblock_160", U.shiftedBlock(dim=160, num_heads=1, mlp_ratio=1), (2, 256, 160), 16, 16, 311)
    torch.manual_seed(2); run("unext_shiftedblock_256", U.shiftedBlock(dim=256, num_heads=1, mlp_ratio=1), (2, 64, 256), 8, 8, 321)
    torch.manual_seed(2); run("unext_shiftedblock_128", U.shiftedBlock(dim=128, num_heads=1, mlp_ratio=1), (1, 1024, 128), 32, 32, 331)
    # whole UNeXt: state_dict layout and seed-2 weights (accx.unext.UNext must reproduce both), one small forward
    torch.manual_seed(2)
    m = U.UNext(3, 1, img_size=64)
    sd = m.state_dict()
    names = list(sd.keys())
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(205))
    m.eval()
    with torch.no_grad():
        y = m(x)
    np.savez_compressed(os.path.join(HERE, "unext_model_init.npz"), names=np.array(names),
                        shapes=np.array([str(tuple(sd[k].shape)) for k in names]),
                        sums=np.array([float(sd[k].double().sum()) for k in names]), x=x.numpy(), eval_out=y.numpy())


if __name__ == "__main__":
    main()
