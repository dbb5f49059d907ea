"""UNeXt shifted tokenized-MLP block (BASELINE configs[2]; Experiments/nets/UNext.py:38-160).

CPU: the oracle restatement against fixtures generated from the unmodified reference (tests/golden/make_golden_unext.py),
state_dict layout / seed-2 initial weights of the accx drop-ins against the reference's.
GPU: accx.unext.{shiftmlp, shiftedBlock, DWConv} against the same fixtures and against the oracle at the full token shapes
of UNeXt at 8x3x256x256 ([8,256,160], [8,64,256], [8,1024,128]), fp32 (rtol 1e-3) and bf16 (2e-2)."""
import numpy as np
import pytest
import torch

from helpers import close, load_case, rel_l2, deterministic

CASES = ["unext_shiftmlp_160", "unext_shiftedblock_160", "unext_shiftedblock_256", "unext_shiftedblock_128"]


def _inputs(case):
    z = case["raw"]
    shape, seed = tuple(int(v) for v in z["shape"]), int(z["seed"])
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed))
    cot = torch.randn(*case["out"][0].shape, generator=torch.Generator().manual_seed(seed + 1))
    return x, cot, int(z["H"]), int(z["W"])


def _oracle_run(case, name, x, cot, H, W):
    from oracle import acc_oracle as O
    sd = {"." + k: v.clone().requires_grad_(True) for k, v in case["sd"].items()}
    xx = x.clone().requires_grad_(True)
    fn = O.shiftmlp if "shiftmlp" in name else O.shifted_block
    y = fn(O.Ctx(sd, True), "", xx, H, W)
    (y * cot).sum().backward()
    return y.detach(), xx.grad, {k[1:]: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    case = load_case(name)
    x, cot, H, W = _inputs(case)
    y, gin, gp = _oracle_run(case, name, x, cot, H, W)
    close(y, case["out"][0], 1e-5, 1e-6, name + " out")
    close(gin, case["gin"][0], 1e-5, 1e-6, name + " gin")
    for k, v in case["gp"].items():
        close(gp[k], v, 1e-4, 1e-5, f"{name} grad {k}")


def test_token_shift_is_an_offset_read():
    """the identity the CUDA path rests on: chunk g of shift(x) is x read at offset -(g - pad), zero outside the map;
    ragged last chunk (C = 128 -> chunks of 26,26,26,26,24) included"""
    from oracle import acc_oracle as O
    B, H, W, C = 2, 6, 5, 128
    x = torch.randn(B, H * W, C)
    for axis in (2, 3):
        ref = O.token_shift(x, H, W, axis).view(B, H, W, C)
        xm = x.view(B, H, W, C)
        size = -(-C // 5)
        for g in range(5):
            s = g - 2
            c0, c1 = g * size, min(C, (g + 1) * size)
            exp = torch.zeros(B, H, W, c1 - c0)
            for i in range(H if axis == 2 else W):
                j = i - s
                if 0 <= j < (H if axis == 2 else W):
                    if axis == 2:
                        exp[:, i] = xm[:, j, :, c0:c1]
                    else:
                        exp[:, :, i] = xm[:, :, j, c0:c1]
            assert torch.equal(ref[..., c0:c1], exp), (axis, g)


def test_unext_state_dict_layout_and_seeded_init():
    """accx.unext.UNext: same keys, shapes and (seed 2) initial values as the reference's UNext(3, 1, img_size=64)"""
    import accx.unext as U
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "unext_model_init.npz"))
    torch.manual_seed(2)
    m = U.UNext(3, 1, img_size=64)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(n) for n in z["names"]]
    for k, shp, s in zip(sd.keys(), z["shapes"], z["sums"]):
        assert str(tuple(sd[k].shape)) == str(shp), k
        assert abs(float(sd[k].double().sum()) - float(s)) <= 1e-6 * max(1.0, abs(float(s))), k


def test_shiftmlp_rejects_what_the_reference_cannot_run():
    import accx.unext as U
    with pytest.raises(NotImplementedError):
        U.shiftmlp(32, 32, drop=0.5)
    with pytest.raises(NotImplementedError):
        U.shiftedBlock(32, 1, drop_path=0.1)
    m = U.shiftmlp(32, 32)
    with pytest.raises(Exception):                      # CPU tensors: no fallback
        m(torch.randn(1, 16, 32), 4, 4)


# ---------------------------------------------------------------------------------------------------------- GPU
def _build(name, case):
    import accx.unext as U
    C = int(case["raw"]["shape"][2])
    mod = U.shiftmlp(C, C) if "shiftmlp" in name else U.shiftedBlock(dim=C, num_heads=1, mlp_ratio=1)
    mod.load_state_dict(case["sd"])
    return mod.cuda().train()


def _accx_run(mod, x, cot, H, W, dtype):
    mod.zero_grad(set_to_none=True)
    xg = x.cuda().to(dtype).requires_grad_(True)
    y = mod(xg, H, W)
    (y.float() * cot.cuda()).sum().backward()
    torch.cuda.synchronize()
    return y.float().cpu(), xg.grad.float().cpu(), {k: p.grad.float().cpu() for k, p in mod.named_parameters()}


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_accx_matches_reference_fixture_fp32(name):
    case = load_case(name)
    x, cot, H, W = _inputs(case)
    mod = _build(name, case)
    with deterministic():
        y, gin, gp = _accx_run(mod, x, cot, H, W, torch.float32)
    close(y, case["out"][0], 1e-3, 1e-5, name + " out")
    close(gin, case["gin"][0], 1e-3, 1e-4, name + " gin")
    for k, v in case["gp"].items():
        close(gp[k], v, 1e-3, 1e-4, f"{name} grad {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_accx_matches_reference_fixture_bf16(name):
    """bf16 storage: north_star's rtol 2e-2 with atol 1e-2 * max|ref|; no discontinuity on this path (GELU,
    LayerNorm), so the bound is element-wise"""
    case = load_case(name)
    x, cot, H, W = _inputs(case)
    mod = _build(name, case)
    y, gin, gp = _accx_run(mod, x, cot, H, W, torch.bfloat16)
    # the reference given the same bf16-rounded input
    close(y, case["out"][0], 2e-2, 2e-2, name + " out")
    assert rel_l2(y, case["out"][0]) < 2e-2
    close(gin, case["gin"][0], 2e-2, 3e-2, name + " gin")
    assert rel_l2(gin, case["gin"][0]) < 2e-2
    for k, v in case["gp"].items():
        assert rel_l2(gp[k], v) < 2e-2, (k, rel_l2(gp[k], v))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,H,W", [((8, 256, 160), 16, 16), ((8, 64, 256), 8, 8), ((8, 1024, 128), 32, 32),
                                       ((3, 35, 40), 7, 5)])
def test_accx_matches_oracle_full_batch(shape, H, W):
    """the token shapes of UNeXt at 8x3x256x256, batch 8, vs the CPU oracle on the same seeded input; plus a ragged
    map (7 x 5, C = 40: chunks of 8) smaller than the shift reach"""
    import accx.unext as U
    from oracle import acc_oracle as O
    C = shape[2]
    torch.manual_seed(7)
    mod = U.shiftedBlock(dim=C, num_heads=1, mlp_ratio=1)
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.1)
    x = torch.randn(*shape)
    cot = torch.randn(*shape)
    sd = {"." + k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    xx = x.clone().requires_grad_(True)
    yo = O.shifted_block(O.Ctx(sd, True), "", xx, H, W)
    (yo * cot).sum().backward()
    mod = mod.cuda().train()
    with deterministic():
        y, gin, gp = _accx_run(mod, x, cot, H, W, torch.float32)
    close(y, yo, 1e-3, 1e-5, "out")
    close(gin, xx.grad, 1e-3, 1e-4, "gin")
    for k, p in gp.items():
        close(p, sd["." + k].grad, 1e-3, 1e-4, "grad " + k)


@pytest.mark.gpu
def test_dwconv_module_and_whole_unext_forward():
    """DWConv drop-in vs torch's grouped conv; whole UNext eval forward vs the reference's stored output"""
    import os
    import accx.unext as U
    from helpers import GOLDEN
    torch.manual_seed(3)
    dw = U.DWConv(48).cuda()
    x = torch.randn(2, 12 * 9, 48, device="cuda", requires_grad=True)
    y = dw(x, 12, 9)
    cot = torch.randn_like(y)
    (y * cot).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr.transpose(1, 2).reshape(2, 48, 12, 9), dw.dwconv.weight.detach(),
                                    dw.dwconv.bias.detach(), padding=1, groups=48).flatten(2).transpose(1, 2)
    (yr * cot).sum().backward()
    close(y, yr, 1e-3, 1e-5, "dwconv out")
    close(x.grad, xr.grad, 1e-3, 1e-5, "dwconv gin")
    z = np.load(os.path.join(GOLDEN, "unext_model_init.npz"))
    torch.manual_seed(2)
    m = U.UNext(3, 1, img_size=64).cuda().eval()
    with torch.no_grad():
        out = m(torch.from_numpy(z["x"]).cuda())
    close(out.cpu(), torch.from_numpy(z["eval_out"]), 1e-3, 1e-4, "UNext eval forward")
