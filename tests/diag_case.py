"""Diagnostic (not a test): per-parameter bf16 gradient error of one golden case, accx vs the oracle run in bf16.
    python tests/diag_case.py hancblock_8_8_k3_f34"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200"), os.path.join(ROOT, "tests")]
from helpers import load_case, rel_l2  # noqa: E402
from test_modules_gpu import build, oracle_run  # noqa: E402

name = sys.argv[1]
case = load_case(name)
dtype = torch.bfloat16
mod = build(name).to("cuda")
mod.load_state_dict(case["sd"])
mod.train()
xs = [x.to("cuda").to(dtype).requires_grad_(True) for x in case["in"]]
ys = mod(*xs)
ys = ys if isinstance(ys, tuple) else (ys,)
torch.autograd.backward(ys, [c.to("cuda").to(dtype) for c in case["cot"]])
_, g_in, g_p, _ = oracle_run(name, {"." + k: v for k, v in case["sd"].items()}, case["in"], case["cot"], "cuda", dtype)
g_p = {k[1:]: v for k, v in g_p.items()}
named = dict(mod.named_parameters())
tot = sum(float(g.double().norm() ** 2) for g in case["gp"].values()) ** 0.5
print(f"{'param':28s} {'|ref|/tot':>9s} {'accx':>9s} {'bf16-oracle':>11s}   (rel-l2 vs fp32 reference); inputs {[tuple(x.shape) for x in xs]}")
for k, g in case["gp"].items():
    if named[k].grad is None or float(g.abs().max()) == 0:
        continue
    print(f"{k:28s} {float(g.norm()) / tot:9.3f} {rel_l2(named[k].grad.float(), g):9.2e} {rel_l2(g_p[k].float(), g):11.2e}")
for i, x in enumerate(xs):
    print(f"gin{i:<25d} {'':9s} {rel_l2(x.grad.float(), case['gin'][i]):9.2e} {rel_l2(g_in[i].float(), case['gin'][i]):11.2e}")
for i, y in enumerate(ys):
    print(f"out{i:<25d} {'':9s} {rel_l2(y.float(), case['out'][i]):9.2e}")
