"""Profiling target (not a test): a few launches of ONE accx contraction, for `ncu -k regex:... --launch-skip 2 -c 1`.
    python tests/prof_one.py pw P K N act        1x1 contraction
    python tests/prof_one.py taps B H W C        ResPath dense 3x3 (nine shifted operands)
    python tests/prof_one.py taps_wgrad B H W C
    python tests/prof_one.py dw | dw_wgrad | dw_dgrad B H W C    depthwise 3x3 forward / weight gradient / input gradient + BN reduction"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]
from accx import engine as E  # noqa: E402

kind, a = sys.argv[1], [int(v) for v in sys.argv[2:]]
torch.manual_seed(0)
if kind == "pw":
    P, K, N, act = a
    x = torch.randn(1, 1, P, K, device="cuda").to(torch.bfloat16)
    L = E.Lazy(x, torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda") * 0.1, act) if act else E.Lazy(x)
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    st = torch.zeros(2 * N, device="cuda")
    for _ in range(4):
        E.conv([E.Op(L, K, E.WV(w, 0, K, 1))], N, (1, 1, P), stats=st)
elif kind.startswith("dw"):
    B, H, W, C = a
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    L = E.Lazy(x, torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, 2)
    w = torch.randn(C, 1, 3, 3, device="cuda") / 3
    bias = torch.randn(C, device="cuda")
    st = torch.zeros(2 * C, device="cuda")
    dy = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    gw = torch.zeros_like(w)
    for _ in range(4):
        if kind == "dw":
            E.dw_fwd(L, w, bias, st)
        elif kind == "dw_wgrad":
            E.dw_wgrad(L, dy, gw)
        else:
            L.mean, L.rstd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
            E.dw_dgrad_bnred(L, dy, w, E.Arena(x.device))
else:
    from accx.modules import ResPath
    B, H, W, C = a
    x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    L = E.Lazy(x, torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1, 2)
    w = torch.randn(C, C, 3, 3, device="cuda") / (9 * C) ** 0.5
    st = torch.zeros(2 * C, device="cuda")
    dy = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
    gw = torch.zeros_like(w)
    for _ in range(4):
        if kind == "taps":
            E.conv(ResPath._taps(L, w, C), C, (B, H, W), stats=st)
        else:
            E.wgrad_conv3x3(L, C, w, dy, C, (B, H, W), gw)
torch.cuda.synchronize()
