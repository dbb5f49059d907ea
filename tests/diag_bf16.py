"""diagnostic (not a test): per-tensor error of the bf16 runs against the decision-replayed oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "acc-unet-unext_b200"))
import torch
import test_modules_gpu as T
from helpers import load_case, module_cases, rel_l2

def stats(a, b, rtol=2e-2, atol_rel=1e-2):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    scale = float(b.abs().max())
    exc = (a - b).abs() - rtol * b.abs()
    bad = exc > atol_rel * scale
    return f"n={a.numel():7d} out={float(bad.double().mean()):.1e} worst={float(exc.max() / max(scale, 1e-30)):.3f}*max rel-l2={rel_l2(a, b):.2e}"

def report(tag, mod, ys, xs, ref_out, ref_gin, ref_gp):
    for i, y in enumerate(ys):
        print(f"{tag:28s} out{i:<22d} {stats(y.float(), ref_out[i])}")
    for i, x in enumerate(xs):
        print(f"{tag:28s} gin{i:<22d} {stats(x.grad.float(), ref_gin[i])}")
    named = dict(mod.named_parameters())
    for k, g in ref_gp.items():
        if float(g.abs().max()) > 0:
            print(f"{tag:28s} {k:25s} {stats(named[k].grad.float(), g)}")

dtype = torch.bfloat16
for name in module_cases():
    case = load_case(name)
    mod = T.build(name).to("cuda")
    mod.load_state_dict(case["sd"])
    sd_dot = {"." + k: v for k, v in case["sd"].items()}
    ys, xs, dec = T.run_accx(mod, case["in"], case["cot"], dtype)
    _, g_in, g_p, _ = T.oracle_run(name, sd_dot, [T.bf16_round(x) for x in case["in"]], [T.bf16_round(c) for c in case["cot"]],
                                   "cpu", torch.float32, forced=dec)
    report(name, mod, ys, xs, case["out"], g_in, {k[1:]: v for k, v in g_p.items()})
