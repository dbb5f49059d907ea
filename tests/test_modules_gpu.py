"""GPU parity tests proper: the accx drop-in modules (CUDA, through the C ABI) against
(a) the golden fixtures produced by the reference itself and (b) the CPU oracle on larger
seeded inputs.  fp32 storage: rtol 1e-3; bf16 storage: rtol 2e-2 (BASELINE.json north_star)."""
import pytest
import torch

from helpers import (accx_decisions, close, close_frac, deterministic, load_case, module_cases, record_decisions, rel_l2,
                     whole_model_checks)

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


# ---- tolerance policy (BASELINE.json north_star) -------------------------------------
# fp32 storage, run in DETERMINISTIC reduction mode (engine.set_deterministic: no run-to-run spread, so a miss is a
#   bug and not a draw of the atomics): every element of every output, input gradient and parameter gradient within
#   |a-b| <= 1e-3*|b| + atol.  Gradients may miss that on <= 1e-3 of the elements (a max-pool arg-max or LeakyReLU
#   sign at a rounding-level near-tie is decided differently by ANY two fp32 evaluation orders -- the reference on
#   another device does it too) and must agree to a relative L2 of 2e-3 as a whole.
# bf16 storage: outputs element-wise |a-b| <= 2e-2*|b| + 1e-2*max|b| against the fp32 reference.  Gradients are
#   compared ELEMENT-WISE with the same bound against the oracle evaluated with the CUDA run's own discrete
#   decisions replayed (Ctx.forced: the sign under every LeakyReLU, the arg-max of every pool window): bf16 rounding
#   of an activation next to zero flips such a decision, which changes one gradient contribution by a factor 100 --
#   a property of ANY bf16 evaluation (the oracle run in bf16 shows the same) that says nothing about the kernels.
#   With the decisions pinned the remaining error is the smooth rounding error north_star's rtol is about.
BF16_RTOL, BF16_ATOL = 2e-2, 1e-2


def check_out(a, b, dtype, what):
    if dtype == torch.float32:
        close(a.float(), b, 1e-3, 2e-5, what)
    else:
        close(a.float(), b, BF16_RTOL, BF16_ATOL, what)


def check_grad(a, b, dtype, what, atol=1e-3):
    if dtype == torch.float32:
        close_frac(a.float(), b, 1e-3, atol, what, 1e-3)
        assert rel_l2(a, b) <= 2e-3 or float(b.abs().max()) == 0, f"{what}: rel-l2 {rel_l2(a, b):.2e}"
    else:
        assert torch.isfinite(a).all()
        close_frac(a.float(), b, BF16_RTOL, BF16_ATOL, what, 2e-3)
        assert rel_l2(a, b) <= 2e-2 or float(b.abs().max()) == 0, f"{what}: rel-l2 {rel_l2(a, b):.2e}"


def oracle_run(name_or_fn, sd, xs, cots, device, dtype, forced=None):
    """oracle forward+backward on `device` in `dtype` -> (outs, input grads, {param: grad}, ctx).
    forced: discrete decisions to replay (helpers.accx_decisions)"""
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
    if forced is not None:
        forced = {k: v.to(device) for k, v in forced.items()}
    if callable(name_or_fn):
        cx = O.Ctx(sdd, True, forced=forced)
        ys = name_or_fn(cx, xd)
        ys = ys if isinstance(ys, tuple) else (ys,)
    else:
        cx, _, xd, ys = run_oracle(name_or_fn, {"sd": sdd, "in": xd}, True, prepared=True, forced=forced)
    sum((y.float() * c.to(device)).sum() for y, c in zip(ys, cots)).backward()
    return ys, [x.grad for x in xd], {k: v.grad for k, v in sdd.items() if v.is_floating_point() and v.grad is not None}, cx


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def compare_all(tag, dtype, mod, ys, xs, ref_out, ref_gin, ref_gp):
    """ref_out: the fp32 reference's outputs; ref_gin / ref_gp: its gradients (fp32 storage) or the gradients of the
    oracle with the CUDA run's decisions replayed (bf16 storage)"""
    for i, y in enumerate(ys):
        assert y.dtype == dtype and y.shape == ref_out[i].shape
        check_out(y, ref_out[i], dtype, f"{tag} out{i}")
    for i, x in enumerate(xs):
        check_grad(x.grad, ref_gin[i], dtype, f"{tag} gin{i}")
    named = dict(mod.named_parameters())
    wscale = max(float(v.abs().max()) for k, v in ref_gp.items() if k.endswith("weight"))
    for k, g in ref_gp.items():
        got = named[k].grad
        assert got is not None, f"{tag}: no grad for {k}"
        assert torch.isfinite(got).all(), k
        if float(g.abs().max()) < 1e-4 * wscale:       # analytically-zero conv-bias grads (and noise-level ones)
            close(got, g, 0, 1e-4 if dtype == torch.float32 else 2e-2, f"{tag} grad {k}", zero_scale=wscale)
        else:
            check_grad(got, g, dtype, f"{tag} grad {k}", atol=2e-3)
    for k, p in named.items():
        if k not in ref_gp:
            assert p.grad is None, f"{tag}: reference leaves {k} without a gradient"


def run_accx(mod, xs_cpu, cots, dtype):
    """accx forward + backward on the GPU; returns (outputs, input leaves, decisions of the run)"""
    m = mod.to(DEV).train()
    xg = [x.to(DEV).to(dtype).requires_grad_(True) for x in xs_cpu]
    with record_decisions() as rec:
        yg = m(*xg)
    yg = yg if isinstance(yg, tuple) else (yg,)
    torch.autograd.backward(yg, [c.to(DEV).to(dtype) for c in cots])
    torch.cuda.synchronize()
    return yg, xg, accx_decisions(m, rec)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", module_cases())
def test_module_matches_reference_golden(name, dtype):
    case = load_case(name)
    mod = build(name).to(DEV)
    mod.load_state_dict(case["sd"])
    sd_dot = {"." + k: v for k, v in case["sd"].items()}
    with deterministic(dtype == torch.float32):
        ys, xs, decisions = run_accx(mod, case["in"], case["cot"], dtype)
        if dtype == torch.float32:
            ref_gin, ref_gp = case["gin"], case["gp"]
        else:       # gradients of the reference arithmetic with this run's discrete decisions (see the policy above)
            _, g_in, g_p, _ = oracle_run(name, sd_dot, [bf16_round(x) for x in case["in"]],
                                         [bf16_round(c) for c in case["cot"]], "cpu", torch.float32, forced=decisions)
            ref_gin, ref_gp = g_in, {k[1:]: v for k, v in g_p.items()}
        compare_all(name, dtype, mod, ys, xs, case["out"], ref_gin, ref_gp)
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
    with torch.no_grad():
        from oracle import acc_oracle as O
        probe = oracle_fn(O.Ctx({k: v.clone() for k, v in sd.items()}, True), xs_cpu)
    probe = probe if isinstance(probe, tuple) else (probe,)
    cots = [torch.randn(y.shape, generator=torch.Generator().manual_seed(50 + i)) for i, y in enumerate(probe)]
    yo, gin_o, gp_o, cx = oracle_run(oracle_fn, sd, xs_cpu, cots, "cpu", torch.float32)
    with deterministic(dtype == torch.float32):
        yg, xg, decisions = run_accx(mod, xs_cpu, cots, dtype)
        if dtype == torch.bfloat16:
            free = rel_l2(xg[0].grad, gin_o[0])
            _, gin_o, gp_o, _ = oracle_run(oracle_fn, sd, [bf16_round(x) for x in xs_cpu], [bf16_round(c) for c in cots],
                                           "cpu", torch.float32, forced=decisions)
            print(f"{tag} bf16 input gradient: rel-l2 {rel_l2(xg[0].grad, gin_o[0]):.2e} with the run's decisions replayed, "
                  f"{free:.2e} against the free-running fp32 oracle")
        compare_all(tag, dtype, mod, yg, xg, [y.detach() for y in yo], gin_o, {k[1:]: v for k, v in gp_o.items()})
        for k, v in cx.updates.items():
            check_out(mod.state_dict()[k[1:]], v.float(), dtype, f"{tag} buffer {k}")


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
