"""Diagnostic (not a test): per-tensor parity metrics of the accx modules vs the golden fixtures.
    python tests/diag_parity.py        (on a GPU box)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200"), os.path.join(ROOT, "tests")]
from helpers import load_case, module_cases, rel_l2  # noqa: E402
from test_modules_gpu import build  # noqa: E402


def frac_out(a, b, rtol, atol_rel):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    lim = atol_rel * float(b.abs().max()) + rtol * b.abs()
    return float(((a - b).abs() > lim).double().mean())


for dtype, rt, at in ((torch.float32, 1e-3, 2e-4), (torch.bfloat16, 2e-2, 2e-2)):
    for name in module_cases():
        case = load_case(name)
        mod = build(name).to("cuda")
        mod.load_state_dict(case["sd"])
        mod.train()
        xs = [x.to("cuda").to(dtype).requires_grad_(True) for x in case["in"]]
        ys = mod(*xs)
        ys = ys if isinstance(ys, tuple) else (ys,)
        torch.autograd.backward(ys, [c.to("cuda").to(dtype) for c in case["cot"]])
        rows = [("out%d" % i, y.float(), case["out"][i]) for i, y in enumerate(ys)]
        rows += [("gin%d" % i, x.grad.float(), case["gin"][i]) for i, x in enumerate(xs)]
        named = dict(mod.named_parameters())
        worst = ("", 0.0, 0.0)
        for k, g in case["gp"].items():
            if k.endswith("weight") and named[k].grad is not None and g.abs().max() > 0:
                r = rel_l2(named[k].grad.float(), g)
                if r > worst[1]:
                    worst = (k, r, frac_out(named[k].grad.float(), g, rt, at))
        msg = " ".join(f"{n}:l2={rel_l2(a, b):.1e},out={frac_out(a, b, rt, at):.1e}" for n, a, b in rows)
        print(f"{str(dtype)[6:]:9s}{name:24s} {msg} worstW {worst[0]}:l2={worst[1]:.1e},out={worst[2]:.1e}")
