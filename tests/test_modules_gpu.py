"""GPU parity tests proper: the accx drop-in modules (CUDA, through the C ABI) against
(a) the golden fixtures produced by the reference itself and (b) the CPU oracle on larger
seeded inputs.  fp32 storage: rtol 1e-3; bf16 storage: rtol 2e-2 (BASELINE.json north_star)."""
import pytest
import torch

from helpers import close, load_case, module_cases, rel_l2, whole_model_checks

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(name):
    import accx
    kind = name.split("_")[0]
    if kind == "se":
        return accx.ChannelSELayer(int(name.split("_c")[1]))
    if kind == "hanclayer":
        return accx.HANCLayer(8, 16, int(name[-1]))
    if kind == "hancblock":
        p = name.split("_")
        f = int(p[4][1:]) if len(p) > 4 else 3
        return accx.HANCBlock(int(p[1]), int(p[2]), k=int(p[3][1:]), inv_fctr=f)
    if kind == "respath":
        return accx.ResPath(int(name.split("_c")[1].split("_")[0]), int(name.split("_n")[1]))
    p = name.split("_")
    variant = {"mlfc": "base", "mlfcw": "w", "mlfclite": "lite"}[kind]
    return accx.MLFC(int(p[1]), int(p[2]), int(p[3]), int(p[4]), lenn=2 if name.endswith("len2") else 1, variant=variant)


def tolerances(dtype):
    # (rtol, atol_rel) on outputs / input grads; param grads get a looser atol (reductions over pixels)
    return (1e-3, 2e-4, 1e-3) if dtype == torch.float32 else (2e-2, 2e-2, 4e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", module_cases())
def test_module_matches_reference_golden(name, dtype):
    case = load_case(name)
    mod = build(name).to(DEV)
    mod.load_state_dict(case["sd"])
    mod.train()
    rt, at, atp = tolerances(dtype)
    xs = [x.to(DEV).to(dtype).requires_grad_(True) for x in case["in"]]
    ys = mod(*xs)
    ys = ys if isinstance(ys, tuple) else (ys,)
    for i, y in enumerate(ys):
        assert y.dtype == dtype and y.shape == case["out"][i].shape
        close(y.float(), case["out"][i], rt, at, f"{name} out{i}")
    torch.autograd.backward(ys, [c.to(DEV).to(dtype) for c in case["cot"]])
    for i, x in enumerate(xs):
        close(x.grad.float(), case["gin"][i], rt, 5 * at, f"{name} gin{i}")
    wscale = max(float(v.abs().max()) for k, v in case["gp"].items() if k.endswith("weight"))
    named = dict(mod.named_parameters())
    for k, g in case["gp"].items():
        got = named[k].grad
        assert got is not None, f"{name}: no grad for {k}"
        if float(g.abs().max()) < 1e-4 * wscale:
            close(got, g, 0, 1e-4 if dtype == torch.float32 else 1e-2, f"{name} grad {k}", zero_scale=wscale)
        else:
            close(got.float(), g, rt, atp, f"{name} grad {k}")
    for k, p in named.items():
        if k not in case["gp"]:
            assert p.grad is None, f"{name}: reference leaves {k} without a gradient"
    sd = mod.state_dict()
    for k, v in case["upd"].items():
        close(sd[k].float(), v.float(), rt, at, f"{name} buffer {k}")
    mod.eval()
    with torch.no_grad():
        ys = mod(*[x.detach() for x in xs])
    ys = ys if isinstance(ys, tuple) else (ys,)
    for i, y in enumerate(ys):
        close(y.float(), case["eval"][i], rt, at, f"{name} eval{i}")


def _oracle_vs_accx(mod, oracle_fn, xs_cpu, dtype, tag):
    """same seeded weights + inputs: CPU oracle (fp32) vs accx on the GPU"""
    from oracle import acc_oracle as O
    sd = {"." + k: v.detach().clone() for k, v in mod.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    xo = [x.clone().requires_grad_(True) for x in xs_cpu]
    cx = O.Ctx(sd, True)
    yo = oracle_fn(cx, xo)
    yo = yo if isinstance(yo, tuple) else (yo,)
    cots = [torch.randn(y.shape, generator=torch.Generator().manual_seed(50 + i)) for i, y in enumerate(yo)]
    sum((y * c).sum() for y, c in zip(yo, cots)).backward()
    m = mod.to(DEV).train()
    xg = [x.to(DEV).to(dtype).requires_grad_(True) for x in xs_cpu]
    yg = m(*xg)
    yg = yg if isinstance(yg, tuple) else (yg,)
    rt, at, atp = tolerances(dtype)
    for i, (a, b) in enumerate(zip(yg, yo)):
        close(a.float(), b, rt, at, f"{tag} out{i}")
    torch.autograd.backward(yg, [c.to(DEV).to(dtype) for c in cots])
    for i, (a, b) in enumerate(zip(xg, xo)):
        close(a.grad.float(), b.grad, rt, 5 * at, f"{tag} gin{i}")
    wscale = max(float(sd["." + k].grad.abs().max()) for k, _ in m.named_parameters()
                 if k.endswith("weight") and sd["." + k].grad is not None)
    for k, p in m.named_parameters():
        go = sd["." + k].grad
        if go is None:
            assert p.grad is None
            continue
        if float(go.abs().max()) < 1e-4 * wscale:
            close(p.grad, go, 0, 1e-4 if dtype == torch.float32 else 1e-2, f"{tag} grad {k}", zero_scale=wscale)
        else:
            close(p.grad.float(), go, rt, atp, f"{tag} grad {k}")
    for k, v in cx.updates.items():
        close(m.state_dict()[k[1:]].float(), v.float(), rt, at, f"{tag} buffer {k}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_hancblock_model_shape_vs_oracle(dtype):
    """HANCBlock(32, 32, k=3) -- the cnv12/cnv92 configuration of ACC_UNet -- at 2x32x56x56"""
    import accx
    from oracle import acc_oracle as O
    torch.manual_seed(2)
    mod = accx.HANCBlock(32, 32, k=3, inv_fctr=3)
    x = torch.randn(2, 32, 56, 56, generator=torch.Generator().manual_seed(3))
    _oracle_vs_accx(mod, lambda cx, xs: O.hanc_block(cx, "", xs[0], 3), [x], dtype, "hancblock32")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_hancblock_concat_input_k2_vs_oracle(dtype):
    """HANCBlock(64, 32, k=2): decoder-style block (input = concat, more channels than output)"""
    import accx
    from oracle import acc_oracle as O
    torch.manual_seed(2)
    mod = accx.HANCBlock(64, 32, k=2, inv_fctr=3)
    x = torch.randn(3, 64, 14, 14, generator=torch.Generator().manual_seed(4))
    _oracle_vs_accx(mod, lambda cx, xs: O.hanc_block(cx, "", xs[0], 2), [x], dtype, "hancblock64")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_respath_vs_oracle(dtype):
    import accx
    from oracle import acc_oracle as O
    torch.manual_seed(2)
    mod = accx.ResPath(32, 3)
    x = torch.randn(2, 32, 28, 28, generator=torch.Generator().manual_seed(5))
    _oracle_vs_accx(mod, lambda cx, xs: O.respath(cx, "", xs[0], 3), [x], dtype, "respath32")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_mlfc_model_shape_vs_oracle(dtype):
    """MLFC(32, 64, 128, 256) -- ACC_UNet's own configuration -- on a 32x32 pyramid"""
    import accx
    from oracle import acc_oracle as O
    torch.manual_seed(2)
    mod = accx.MLFC(32, 64, 128, 256)
    xs = [torch.randn(2, c, 32 >> i, 32 >> i, generator=torch.Generator().manual_seed(6 + i))
          for i, c in enumerate((32, 64, 128, 256))]
    _oracle_vs_accx(mod, lambda cx, v: O.mlfc(cx, "", list(v), 1, "base"), xs, dtype, "mlfc32")


def test_nchw_contiguous_and_channels_last_inputs_agree():
    import accx
    torch.manual_seed(0)
    m = accx.HANCBlock(16, 16, k=2).to(DEV).eval()
    x = torch.randn(2, 16, 8, 8, device=DEV)
    with torch.no_grad():
        a = m(x)
        b = m(x.contiguous(memory_format=torch.channels_last))
    assert a.shape == (2, 16, 8, 8)
    assert torch.equal(a, b)


def test_shape_errors_match_reference_constraints():
    import accx
    m = accx.HANCLayer(8, 8, 3).to(DEV)
    with pytest.raises(ValueError):
        m(torch.randn(1, 8, 6, 8, device=DEV))          # H not divisible by 4
    with pytest.raises(accx.AccxError):
        accx.ChannelSELayer(16)(torch.randn(1, 16, 4, 4))   # CPU tensor: no CPU path


@pytest.mark.parametrize("name,variant", [("accunet_f8", "base"), ("accunetw_f8", "w"), ("accunetlite_f8", "lite")])
def test_whole_model_against_reference_golden(name, variant):
    import accx
    case = load_case(name)
    cls = {"base": accx.ACC_UNet, "w": accx.ACC_UNet_W, "lite": accx.ACC_UNet_Lite}[variant]
    torch.manual_seed(2)
    m = cls(3, 1, 8).to(DEV).train()
    x = case["in"][0].to(DEV).requires_grad_(True)
    y = m(x)
    (y * case["cot"][0].to(DEV)).sum().backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    sd = m.state_dict()
    whole_model_checks(name, case, y, x.grad, grads, sd)
    m.eval()
    with torch.no_grad():
        ye = m(case["in"][0].to(DEV))
    close(ye, case["eval"][0], 1e-3, 5e-3, f"{name} eval")


def test_whole_model_bf16_runs_and_tracks_fp32():
    import accx
    torch.manual_seed(2)
    m = accx.ACC_UNet(3, 1, 8).to(DEV).train()
    x = torch.randn(2, 3, 64, 64, device=DEV)
    y32 = m(x)
    m.compute_dtype = torch.bfloat16
    y16 = m(x)
    assert y16.dtype == torch.float32 and torch.isfinite(y16).all()
    assert rel_l2(y16, y32) < 0.15
    y16.mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
