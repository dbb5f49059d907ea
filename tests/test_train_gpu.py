"""GPU: the steps either side of the blocks inside one training step --
MaxPool2d(2), the final 1x1 conv, the fused WeightedDiceBCE kernels, the flat Adam kernel -- each against
the torch operator the reference calls, and the whole TrainStep (flat parameter / gradient buffers, CUDA
graph) against the same step made of loose torch pieces (torch.optim.Adam, the torch loss) and against the
CPU oracle's train step."""
import copy

import pytest
import torch
import torch.nn.functional as F

from helpers import close, rel_l2

pytestmark = pytest.mark.gpu

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def E():
    from accx import engine
    return engine


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 32, 16, 24), (1, 3, 6, 10), (3, 520, 4, 6), (2, 64, 56, 56)])
def test_maxpool2_matches_torch(dtype, shape):
    from accx.modules import maxpool2
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g)
    x = (x * 2).round() / 2                      # many exact ties inside the 2x2 windows: the tie rule matters
    x = x.to(DEV).to(dtype)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    ya = maxpool2(xa)
    yb = F.max_pool2d(xb, 2)
    assert ya.shape == yb.shape
    assert torch.equal(ya, yb)                   # selection only: bit-exact
    cot = torch.randn(yb.shape, generator=g).to(DEV).to(dtype)
    ya.backward(cot)
    yb.backward(cot)
    assert torch.equal(xa.grad, xb.grad)         # same arg-max (first maximum in window order), bit-exact


def test_maxpool2_odd_size_is_an_error():
    from accx.modules import maxpool2
    with pytest.raises(ValueError):
        maxpool2(torch.randn(1, 8, 7, 8, device=DEV))


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("C,N", [(32, 1), (8, 1), (32, 3)])
def test_out_conv_matches_torch(dtype, C, N):
    from accx.modules import out_conv
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(C, N, 1).to(DEV)
    ref = copy.deepcopy(conv)
    x = torch.randn(2, C, 12, 20, device=DEV).to(dtype)
    xa = x.clone().requires_grad_(True)
    xb = x.float().clone().requires_grad_(True)
    ya = out_conv(xa, conv)
    yb = ref(xb)
    assert ya.dtype == torch.float32 and ya.shape == yb.shape
    rt, at = (1e-3, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    close(ya, yb, rt, at, "out_conv")
    cot = torch.randn_like(yb)
    ya.backward(cot)
    yb.backward(cot)
    close(xa.grad.float(), xb.grad, rt, at, "out_conv dx")
    close(conv.weight.grad, ref.weight.grad, rt, at * 10, "out_conv dW")
    close(conv.bias.grad, ref.bias.grad, 1e-3, 1e-4, "out_conv db")


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("Cin,Co,Cs,H,W", [(16, 8, 8, 6, 10), (64, 32, 32, 14, 14), (512, 256, 256, 7, 7), (24, 6, 10, 5, 3)])
def test_up_cat_matches_torch(dtype, Cin, Co, Cs, H, W):
    """cat([ConvTranspose2d(Cin, Co, 2, 2)(x), skip], dim=1) forward + all gradients vs ATen"""
    from accx.modules import up_cat
    torch.manual_seed(4)
    up = torch.nn.ConvTranspose2d(Cin, Co, kernel_size=(2, 2), stride=2).to(DEV)
    ref = copy.deepcopy(up)
    x = torch.randn(2, Cin, H, W, device=DEV).to(dtype)
    sk = torch.randn(2, Cs, 2 * H, 2 * W, device=DEV).to(dtype)
    xa, sa = x.clone().requires_grad_(True), sk.clone().requires_grad_(True)
    xb, sb = x.float().clone().requires_grad_(True), sk.float().clone().requires_grad_(True)
    ya = up_cat(xa, sa, up)
    yb = torch.cat([ref(xb), sb], dim=1)
    assert ya.shape == yb.shape and ya.dtype == dtype
    rt, at = (1e-3, 1e-5) if dtype == torch.float32 else (2e-2, 1e-2)
    close(ya.float(), yb, rt, at, "up_cat")
    assert torch.equal(ya[:, Co:], sk)                       # the skip half is a copy
    cot = torch.randn_like(yb)
    ya.backward(cot.to(dtype))
    yb.backward(cot.to(dtype).float())
    close(xa.grad.float(), xb.grad, rt, at, "up_cat dx")
    assert torch.equal(sa.grad, cot.to(dtype)[:, Co:])
    close(up.weight.grad, ref.weight.grad, rt, at * 10, "up_cat dW")
    close(up.bias.grad, ref.bias.grad, 1e-3 if dtype == torch.float32 else 2e-2, 1e-3, "up_cat db")


# ---------------------------------------------------------------------------------------------------
def _loss_inputs(B, H, W, seed, dtype=torch.float32, soft=False):
    g = torch.Generator().manual_seed(seed)
    lg = (torch.randn(B, 1, H, W, generator=g) * 3).to(DEV).to(dtype)
    if soft:
        m = torch.rand(B, 1, H, W, generator=g).to(DEV)
    else:
        m = (torch.rand(B, 1, H, W, generator=g) > 0.6).float().to(DEV)
    return lg, m


@pytest.mark.parametrize("B,H,W,soft", [(2, 16, 16, False), (16, 224, 224, False), (3, 10, 14, True), (1, 7, 5, False)])
def test_dice_bce_matches_torch_formula(B, H, W, soft):
    from accx.train import dice_bce_loss, dice_bce_loss_torch
    lg, m = _loss_inputs(B, H, W, 11, soft=soft)
    la = lg.clone().requires_grad_(True)
    lb = lg.clone().requires_grad_(True)
    a = dice_bce_loss(la, m)
    b = dice_bce_loss_torch(lb, m)
    assert abs(float(a) - float(b)) < 1e-5 * max(1.0, abs(float(b)))
    (a * 1.7).backward()                                     # a non-trivial upstream gradient
    (b * 1.7).backward()
    close(la.grad, lb.grad, 1e-3, 1e-5, "dice_bce dlogit")


def test_dice_bce_matches_oracle_and_degenerate_masks():
    from accx.train import dice_bce_loss
    from oracle import acc_oracle as O
    lg, m = _loss_inputs(4, 32, 32, 12)
    for mask in (m, torch.zeros_like(m), torch.ones_like(m)):   # no positives / no negatives: clamp(min=1) branches
        la = lg.clone().requires_grad_(True)
        lo = lg.cpu().clone().requires_grad_(True)
        a = dice_bce_loss(la, mask)
        o = O.dice_bce_loss(lo, mask.cpu())
        assert abs(float(a) - float(o)) < 1e-5
        a.backward()
        o.backward()
        close(la.grad.cpu(), lo.grad, 1e-3, 1e-5, "dice_bce dlogit vs oracle")


def test_dice_bce_bf16_logits():
    from accx.train import dice_bce_loss, dice_bce_loss_torch
    lg, m = _loss_inputs(2, 24, 24, 13, dtype=torch.bfloat16)
    la = lg.clone().requires_grad_(True)
    lb = lg.float().clone().requires_grad_(True)
    a = dice_bce_loss(la, m)
    b = dice_bce_loss_torch(lb, m)
    assert abs(float(a) - float(b)) < 1e-5
    a.backward()
    b.backward()
    assert la.grad.dtype == torch.bfloat16
    close(la.grad.float(), lb.grad, 2e-2, 1e-2, "dice_bce dlogit bf16")


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["random", "edges", "graymask"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_seg_metrics_match_reference_known_answers(case, dtype):
    """accx_seg_metrics vs iou_on_batch / _show_dice of the reference (fixtures made by tests/golden/make_golden.py)"""
    from accx.train import seg_metrics
    from oracle import acc_oracle as O
    from helpers import load_case
    z = load_case("metrics_kat")["raw"]
    lg, tr = torch.from_numpy(z[case + "/logit"]), torch.from_numpy(z[case + "/truth"])
    lgd = lg.to(DEV).to(dtype)
    out = seg_metrics(lgd, tr.to(DEV))
    assert out.is_cuda and out.shape == (2,)
    iou, dice = (float(v) for v in out.cpu())
    if dtype == torch.float32:
        want_iou, want_dice = float(z[case + "/iou"]), float(z[case + "/dice"])
    else:       # bf16 logits: the rounding can move a logit across 0, so the truth is the restatement on the rounded values
        want_iou, want_dice = O.seg_metrics(lgd.float().cpu(), tr)
    assert abs(iou - want_iou) < 1e-6, (iou, want_iou)            # integer counts, ratios in double, one fp32 rounding
    assert abs(dice - want_dice) < 2e-6, (dice, want_dice)


def test_seg_metrics_large_batch_matches_restatement():
    from accx.train import seg_metrics
    from oracle import acc_oracle as O
    g = torch.Generator().manual_seed(11)
    lg = torch.randn(16, 1, 224, 224, generator=g)
    tr = (torch.rand(16, 1, 224, 224, generator=g) > 0.5).float()
    out = seg_metrics(lg.to(DEV), tr.to(DEV)).cpu()
    iou, dice = O.seg_metrics(lg, tr)
    assert abs(float(out[0]) - iou) < 1e-6 and abs(float(out[1]) - dice) < 2e-6, (out, iou, dice)


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_adam_flat_matches_torch_adam(wd):
    e = E()
    n = 4096 + 64
    g = torch.Generator().manual_seed(21)
    p0 = torch.randn(n, generator=g).to(DEV)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=wd)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(4, device=DEV)
    for it in range(5):
        grad = (torch.randn(n, generator=g) * (10.0 ** (it - 2))).to(DEV)
        grad[:64] = 0.0                                      # parameters that never get a gradient stay put (Lite)
        ref.grad = grad.clone()
        opt.step()
        e.adam_step(p, grad, m, v, state, 1e-3, weight_decay=wd)
        close(p, ref.detach(), 1e-5, 1e-6, f"adam step {it + 1}")
    assert float(state[0]) == 5.0
    if wd == 0.0:
        assert torch.equal(p[:64], p0[:64])


# ---------------------------------------------------------------------------------------------------
# The whole ACC-UNet is badly conditioned at random init on test-sized inputs (BatchNorm over a handful of
# values at the 2x2 bottleneck), and Adam's first steps move every weight by +-lr according to the SIGN of its
# gradient: the fp32 atomics of the statistics kernels reorder sums from run to run, a few % of the gradient signs
# are noise, and the loss after one update then varies by +-1e-3 (64x64, batch 4) .. +-5e-3 (32x32, batch 2) between
# two runs of the SAME code, with or without side streams (tests/diag_train.py prints that spread for the flat
# step, the loose torch step and both losses).  So: exact machinery checks run on a small well-conditioned stack
# of accx blocks, and whole-model trajectories are compared with tolerances above that spread.
class _TinyNet(torch.nn.Module):
    """HANCBlock -> MaxPool2d(2) -> HANCBlock -> ConvTranspose2d (torch-native parameters) -> final 1x1 conv"""

    def __init__(self):
        super().__init__()
        import accx
        self.b1 = accx.HANCBlock(3, 16, k=3, inv_fctr=3)
        self.b2 = accx.HANCBlock(16, 16, k=2, inv_fctr=3)
        self.up = torch.nn.ConvTranspose2d(16, 16, kernel_size=(2, 2), stride=2)
        self.idle = torch.nn.Conv2d(4, 4, 1)              # never used: its gradient stays None / zero
        self.out = torch.nn.Conv2d(16, 1, kernel_size=(1, 1))

    def forward(self, x):
        from accx.modules import maxpool2, out_conv
        x = x.contiguous(memory_format=torch.channels_last)
        return out_conv(self.up(self.b2(maxpool2(self.b1(x)))), self.out)


def _loose_step(model, opt, x, m):
    """the same optimisation step from loose torch pieces: torch loss chain, per-parameter .grad, torch Adam"""
    from accx.train import dice_bce_loss_torch
    loss = dice_bce_loss_torch(model(x), m)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss.detach()


def test_train_step_flat_matches_loose_torch_step():
    from accx.train import TrainStep
    torch.manual_seed(2)
    ma = _TinyNet().to(DEV).train()
    mb = copy.deepcopy(ma)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 3, 32, 32, generator=g).to(DEV)
    m = (torch.rand(4, 1, 32, 32, generator=g) > 0.5).float().to(DEV)
    step = TrainStep(ma, lr=1e-3, graph=False)
    opt = torch.optim.Adam([p for p in mb.parameters() if p.requires_grad], lr=1e-3)
    before = {n: p.detach().clone() for n, p in ma.named_parameters()}
    pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
    fs = step.flat
    for it in range(3):
        la = step(x, m)
        lb = _loose_step(mb, opt, x, m)
        assert abs(float(la) - float(lb)) < 1e-3 * max(1.0, abs(float(lb))), (it, float(la), float(lb))
        if it == 0:
            # same gradients (one is a view of the flat buffer, the other a loose tensor)
            gmax = max(float(q.grad.abs().max()) for q in pb.values() if q.grad is not None)
            for n in pa:
                if pb[n].grad is None:
                    assert n.startswith("idle.") and pa[n].grad is None
                    continue
                # tensor by tensor in relative L2 (tensors whose gradient is analytically zero or tiny -- conv biases
                # and scales that a following BatchNorm cancels -- only carry the run-to-run noise of the atomics)
                if float(pb[n].grad.abs().max()) >= 1e-2 * gmax:
                    assert rel_l2(pa[n].grad, pb[n].grad) < 5e-3, f"grad {n}: rel-l2 {rel_l2(pa[n].grad, pb[n].grad):.2e}"
                else:
                    assert float((pa[n].grad - pb[n].grad).abs().max()) < 1e-4 * gmax, f"grad {n}"
                assert pa[n].grad.data_ptr() == fs._view(fs.grad, pa[n]).data_ptr(), f"{n}: grad is not a flat view"
                assert pa[n].data_ptr() == fs._view(fs.param, pa[n]).data_ptr(), f"{n}: parameter is not a flat view"
    moved = 0
    for n in pa:
        d_a, d_b = pa[n].detach() - before[n], pb[n].detach() - before[n]
        if pb[n].grad is None:
            assert torch.equal(pa[n].detach(), before[n]), f"{n}: unused parameter changed"
            continue
        # Adam normalises the update by |g|: elements whose gradient is at rounding level get a +-lr step of
        # arbitrary sign in either implementation, so the update is compared where the gradient carries signal
        sig = pb[n].grad.abs() > 1e-2 * pb[n].grad.abs().max().clamp(min=1e-30)
        if float(pb[n].grad.abs().max()) < 1e-2 * gmax:
            continue
        if sig.any():
            moved += int(sig.sum())
            err = float((d_a - d_b)[sig].abs().max())
            assert err < 3e-4, f"{n}: update differs by {err:.2e} after 3 steps of lr 1e-3"
    assert moved > 1000
    assert float(fs.step_state[0]) == 3.0


@pytest.mark.parametrize("variant", ["base", "lite", "w"])
def test_train_step_whole_model_variants(variant):
    """ACC_UNet / _Lite / _W through the flat train step: first loss equals the loose step's, the trajectory stays
    within the run-to-run spread, Lite's idle MLFC convs never move, every parameter and gradient is a flat view"""
    import accx
    from accx.train import TrainStep
    cls = {"base": accx.ACC_UNet, "lite": accx.ACC_UNet_Lite, "w": accx.ACC_UNet_W}[variant]
    torch.manual_seed(2)
    ma = cls(3, 1, 8).to(DEV).train()
    ma.last_activation = None
    mb = copy.deepcopy(ma)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 3, 64, 64, generator=g).to(DEV)
    m = (torch.rand(4, 1, 64, 64, generator=g) > 0.5).float().to(DEV)
    step = TrainStep(ma, lr=1e-3, graph=False)
    opt = torch.optim.Adam([p for p in mb.parameters() if p.requires_grad], lr=1e-3)
    before = {n: p.detach().clone() for n, p in ma.named_parameters()}
    for it in range(3):
        la, lb = float(step(x, m)), float(_loose_step(mb, opt, x, m))
        print(f"{variant} step {it}: flat {la:.6f} loose {lb:.6f}")
        tol = 1e-4 if it == 0 else 1e-2
        assert abs(la - lb) < tol * max(1.0, abs(lb)), (it, la, lb)
    fs = step.flat
    idle = 0
    gmax = max(float(q.grad.abs().max()) for q in mb.parameters() if q.grad is not None)
    for (n, p), (_, q) in zip(ma.named_parameters(), mb.named_parameters()):
        assert p.data_ptr() == fs._view(fs.param, p).data_ptr()
        assert p.grad is None or p.grad.data_ptr() == fs._view(fs.grad, p).data_ptr()
        if q.grad is None:
            idle += 1
            assert torch.equal(p.detach(), before[n]), f"{n}: parameter without a gradient moved"
        elif float(q.grad.abs().max()) > 1e-6 * gmax:
            # (a gradient at rounding level -- |g| << Adam's eps -- gives an update below the fp32 resolution of the
            # parameter: seen on `out.bias`, whose gradient is a sum of +- terms that cancel to ~1e-13)
            assert not torch.equal(p.detach(), before[n]), f"{n} did not move"
    assert (idle > 0) == (variant == "lite")


def test_train_step_graph_equals_eager_and_loss_decreases():
    import accx
    from accx.train import TrainStep
    torch.manual_seed(2)
    ma = accx.ACC_UNet(3, 1, 8, compute_dtype=torch.bfloat16).to(DEV).train()
    ma.last_activation = None
    mb = copy.deepcopy(ma)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4, 3, 64, 64, generator=g).to(DEV)
    m = (torch.rand(4, 1, 64, 64, generator=g) > 0.5).float().to(DEV)
    sa = TrainStep(ma, lr=1e-3, graph=True, graph_warmup=1)
    sb = TrainStep(mb, lr=1e-3, graph=False)
    la, lb = [], []
    for _ in range(10):
        la.append(float(sa(x, m)))
        lb.append(float(sb(x, m)))
    assert sa.graph is not None
    print("graph vs eager losses:", [f"{a:.5f}/{b:.5f}" for a, b in zip(la, lb)])
    # same weights, bf16 storage: two bf16 evaluations differ by rounding-level noise in the statistics that the
    # BatchNorm stack amplifies (see the note above), so even the first loss only agrees to ~1e-3
    assert abs(la[0] - lb[0]) < 1e-2 * max(1.0, abs(lb[0])), (la, lb)
    for a, b in zip(la, lb):                                              # then within the run-to-run spread
        assert abs(a - b) < 3e-2 * max(1.0, abs(b)), (la, lb)            # measured up to 1e-2 after 10 steps
    assert la[-1] < la[0] and lb[-1] < lb[0], (la, lb)        # memorising one batch: the loss must go down
    # one optimisation step per call, in eager, capture and replay alike
    assert float(sa.flat.step_state[0]) == 10.0 and float(sb.flat.step_state[0]) == 10.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_graph_replay_gradients_are_complete(dtype, monkeypatch):
    """Regression for the stale-gradient race: under TrainStep the weight gradients run on a low-priority side stream
    that is joined once before the optimiser.  The gradients left in the flat buffer by a GRAPH REPLAY must equal the
    ones of a single-stream eager step on the same weights (lr = 0 keeps them fixed), conv weights included, and no
    arena-served gradient may pass through autograd (p.grad is still None when backward returns)."""
    import accx
    from accx import train as T
    from accx.modules import maxpool2, out_conv, up_cat
    from accx.train import TrainStep

    class Small(torch.nn.Module):
        """two levels of the model's dataflow (HANC blocks, ResPath, MLFC-free skip, transposed conv + concat, final conv):
        shallow enough to be well conditioned -- the whole ACC_UNet at random init amplifies the run-to-run rounding
        noise of the fp32 atomics to ~10 % on its first layers, which would drown the comparison"""

        def __init__(self):
            super().__init__()
            self.a = accx.HANCBlock(3, 16, k=3)
            self.b = accx.HANCBlock(16, 32, k=2)
            self.r = accx.ResPath(16, 2)
            self.up = torch.nn.ConvTranspose2d(32, 16, kernel_size=(2, 2), stride=2)
            self.c = accx.HANCBlock(32, 16, k=3)
            self.out = torch.nn.Conv2d(16, 1, kernel_size=(1, 1))

        def forward(self, x):
            x = x.to(dtype).contiguous(memory_format=torch.channels_last)
            e1 = self.a(x)
            e2 = self.b(maxpool2(e1))
            return out_conv(self.c(up_cat(e2, self.r(e1), self.up)), self.out)

    torch.manual_seed(2)
    ma = Small().to(DEV).train()
    mb = copy.deepcopy(ma)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(4, 3, 64, 64, generator=g).to(DEV)
    m = (torch.rand(4, 1, 64, 64, generator=g) > 0.5).float().to(DEV)
    seen = []
    orig = T.FlatState.collect

    def spy(self):
        seen.append(sum(p.grad is not None for p in self.params if id(p) in self.taken))
        return orig(self)

    monkeypatch.setattr(T.FlatState, "collect", spy)
    from helpers import deterministic
    eng = E()
    # deterministic reductions make the comparison exact: the replayed graph (weight gradients on the low-priority side
    # stream, parallel lanes, one join before the optimiser) must reproduce the single-stream eager step BIT FOR BIT
    with deterministic():
        sa = TrainStep(ma, lr=0.0, graph=True, graph_warmup=1)
        for _ in range(4):
            sa(x, m)
        torch.cuda.synchronize()
        assert sa.graph is not None
        assert seen and all(n == 0 for n in seen), f"arena-served gradients went through autograd: {seen}"
        mode, lanes = eng.SIDE_MODE, eng.LANES
        eng.SIDE_MODE, eng.LANES = 0, 0                      # reference: everything on one stream
        try:
            sb = TrainStep(mb, lr=0.0, graph=False)
            sb(x, m)
            torch.cuda.synchronize()
        finally:
            eng.SIDE_MODE, eng.LANES = mode, lanes
        n_w = 0
        for (n, p), (_, q) in zip(ma.named_parameters(), mb.named_parameters()):
            ga, gb = sa.flat._view(sa.flat.grad, p), sb.flat._view(sb.flat.grad, q)
            assert torch.equal(ga, gb), (f"{n}: gradient left by the graph replay differs from the single-stream step "
                                         f"(rel L2 {rel_l2(ga, gb):.3e})")
            n_w += int(p.dim() >= 2 and float(gb.abs().max()) > 0)
        assert n_w >= 20                                     # the conv / FC weights the side stream produces were compared


def test_train_step_optimizer_state_roundtrips_through_torch_adam():
    """TrainStep.state_dict() is torch.optim.Adam's format (the reference checkpoints optimizer.state_dict(),
    Experiments/train_model.py:125-145, and restores it on resume, :677-689): it loads into torch.optim.Adam, comes back
    through load_state_dict into a fresh TrainStep, and the resumed run continues bit-identically (deterministic mode);
    set_lr changes the step size of an already captured graph."""
    import accx
    from accx.train import TrainStep
    from helpers import deterministic
    torch.manual_seed(2)
    ma = accx.ACC_UNet(3, 1, 8).to(DEV).train()
    ma.last_activation = None
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 3, 32, 32, generator=g).to(DEV)
    m = (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float().to(DEV)
    with deterministic():
        sa = TrainStep(ma, lr=1e-3, metrics=True)
        for _ in range(2):
            sa(x, m)
        assert sa.last_metrics is not None and sa.last_metrics.shape == (2,) and bool((sa.last_metrics >= 0).all())
        mb = copy.deepcopy(ma)
        sd = sa.state_dict()
        opt = torch.optim.Adam([p for p in mb.parameters() if p.requires_grad], lr=5e-4)
        opt.load_state_dict(sd)                                        # torch accepts the format
        assert opt.param_groups[0]["lr"] == 1e-3
        p0 = next(iter(mb.parameters()))
        assert float(opt.state[p0]["step"]) == 2.0
        assert torch.equal(opt.state[p0]["exp_avg"], sa.flat._view(sa.flat.exp_avg, sa.params[0]))
        sb = TrainStep(mb, lr=123.0)
        sb.load_state_dict(opt.state_dict())                           # and back
        assert sb.lr == 1e-3 and float(sb.flat.step_state[0]) == 2.0 and float(sb.flat.step_state[1]) == pytest.approx(1e-3)
        la, lb = float(sa(x, m)), float(sb(x, m))
        assert la == lb
        for p, q in zip(ma.parameters(), mb.parameters()):
            assert torch.equal(p, q)
        before = [p.detach().clone() for p in ma.parameters()]
        sa.set_lr(0.0)                                                  # device-side scalar: no update at lr = 0
        sa(x, m)
        assert all(torch.equal(p, b) for p, b in zip(ma.parameters(), before))


def test_train_step_matches_cpu_oracle_step():
    """optimisation steps of the accx TrainStep (fp32 storage) vs oracle.train_step on the CPU"""
    import accx
    from accx.train import TrainStep
    from oracle import acc_oracle as O
    torch.manual_seed(2)
    model = accx.ACC_UNet(3, 1, 8).to(DEV).train()
    model.last_activation = None
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 3, 64, 64, generator=g)
    m = (torch.rand(4, 1, 64, 64, generator=g) > 0.5).float()
    step = TrainStep(model, lr=1e-3)
    opt = None
    for it in range(3):
        lo, opt = O.train_step(sd, x, m, opt)
        la = float(step(x.to(DEV), m.to(DEV)))
        print(f"oracle step {it}: accx {la:.6f} oracle {lo:.6f}")
        tol = 1e-4 if it == 0 else 1e-2          # see the note on conditioning above
        assert abs(la - lo) < tol * max(1.0, abs(lo)), (it, la, lo)
