"""GPU: the BENCHMARKED model -- ACC_UNet(3, 1, 32), Dice+BCE on logits -- against vectors the UNMODIFIED reference
produced at full width (tests/golden/make_golden_full.py: 2 x 3 x 224 x 224 train step in fp32, the same in eval mode,
the Lite / W variants, one 512 x 512 forward).

Tolerances.  The whole net at random init is badly conditioned: the reference's OWN fp32 run differs from its fp64 run
by 3.6e-4 (logits), 6e-2 (input gradient), 6e-3 (per-parameter gradient norms) in relative L2 -- 220 training-mode
BatchNorms with channels of nearly zero variance amplify rounding noise by ~1e4 (`ref_err/*` in the fixtures).  A
bound on |accx - reference| therefore only means something relative to |reference32 - reference64|:
  fp32 storage (DETERMINISTIC reductions): error <= 3 x the reference's own error, per quantity; two runs bit-identical.
  bf16 storage: the loss (a well-conditioned aggregate) within 1e-2; logits / gradients no further from the reference
      than 1.5 x the REFERENCE ARITHMETIC run with bf16 storage (the oracle on this GPU in bfloat16) -- the element-wise
      2e-2 bounds live where the conditioning allows them, in tests/test_modules_gpu.py.
"""
import numpy as np
import pytest
import torch

from helpers import close, deterministic, load_case, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
SLACK = 3.0
# fp32 storage with 3 x TF32 tensor-core contractions (the free-running fp32 mode; ~2^-21 per product instead of 2^-24):
# the whole-model deviation is 2-6x the reference's own fp32-vs-fp64 error (measured on B200, printed by the tests); the
# exact fp32-FMA contractions of the deterministic parity mode stay within SLACK.  Module-level parity (rtol 1e-3) holds
# for both (tests/test_modules_gpu.py::test_module_matches_reference_golden[fp32tc]).
SLACK_TF32 = 8.0


class _contractions:
    def __init__(self, tf32):
        self.tf32 = tf32

    def __enter__(self):
        from accx import engine
        self.old, engine.TC_F32_IN_DET = engine.TC_F32_IN_DET, bool(self.tf32)

    def __exit__(self, *exc):
        from accx import engine
        engine.TC_F32_IN_DET = self.old
        return False


def inputs(B, hw, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, hw, hw, generator=g)
    m = (torch.rand(B, 1, hw, hw, generator=g) > 0.5).float()
    return x, m


def build(cls_name, dtype=None):
    import accx
    torch.manual_seed(2)
    model = getattr(accx, cls_name)(3, 1, 32, compute_dtype=dtype).to(DEV)
    model.last_activation = None
    return model


def step(model, x, m):
    from accx.train import dice_bce_loss
    model.zero_grad(set_to_none=True)
    xx = x.to(DEV).requires_grad_(True)
    logits = model(xx)
    loss = dice_bce_loss(logits, m.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    return logits.detach().float().cpu(), float(loss), xx.grad.detach().float().cpu()


def norms_err(model, z, tag=""):
    names = [str(n) for n in z[tag + "gpsum/names"]]
    ref = z[tag + "gpsum/l2"]
    params = dict(model.named_parameters())
    ours = np.array([float(params[n].grad.double().norm()) if params[n].grad is not None else -1.0 for n in names])
    for n, r, o in zip(names, ref, ours):
        assert (r < 0) == (o < 0), f"{n}: reference grad {'None' if r < 0 else r}, ours {'None' if o < 0 else o}"
    # (conv biases in front of a training-mode BatchNorm: exactly 0 here, rounding noise ~1e-5 in the reference)
    keep = ref > 1e-3 * np.median(ref[ref > 0])
    return float(np.linalg.norm(ours[keep] - ref[keep]) / np.linalg.norm(ref[keep]))


def full_grads_err(model, z, tag=""):
    """worst relative L2 over the stored full / probed parameter gradients (noise-level tensors compared by atol)"""
    params = dict(model.named_parameters())
    scale = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith(tag + "gp/") and k.endswith("weight"))
    worst = ("", 0.0)
    for k in z.files:
        if k.startswith(tag + "gp/"):
            ref = torch.from_numpy(z[k])
            got = params[k[len(tag) + 3:]].grad.float().cpu()
        elif k.startswith(tag + "gprobe/"):
            ref = torch.from_numpy(z[k])
            got = params[k[len(tag) + 7:]].grad.float().cpu().reshape(-1)[:ref.numel()]
        else:
            continue
        if float(ref.abs().max()) < 1e-4 * scale:            # analytically zero (e.g. a bias in front of a train-mode BN)
            assert float(got.abs().max()) < 1e-3 * scale, k
            continue
        e = rel_l2(got, ref)
        if e > worst[1]:
            worst = (k, e)
    return worst


@pytest.mark.parametrize("tf32", [False, True], ids=["fma", "tf32x3"])
@pytest.mark.parametrize("cls_name,fixture", [("ACC_UNet", "full_accunet_224"), ("ACC_UNet_Lite", "full_accunetlite_224"),
                                              ("ACC_UNet_W", "full_accunetw_224")])
def test_full_width_train_step_fp32_matches_reference(cls_name, fixture, tf32):
    z = load_case(fixture)["raw"]
    x, m = inputs(2, 224, 2024)
    SLACK = SLACK_TF32 if tf32 else globals()["SLACK"]
    with deterministic(), _contractions(tf32):
        model = build(cls_name).train()
        logits, loss, gin = step(model, x, m)
        ref_logits = torch.from_numpy(z["logits"])
        e_log, r_log = rel_l2(logits, ref_logits), float(z["ref_err/logits_rel_l2"])
        e_max = float((logits - ref_logits).abs().max())
        r_max = float(z["ref_err/logits_max_abs"])
        e_loss, r_loss = abs(loss - float(z["loss"])), float(z["ref_err/loss_abs"])
        probe = torch.from_numpy(z["gin_probe"])
        e_gin = rel_l2(gin, torch.from_numpy(z["gin"])) if "gin" in z.files else rel_l2(gin[:, :, ::8, ::8], probe)
        r_gin = float(z["ref_err/gin_rel_l2"])
        e_nrm, r_nrm = norms_err(model, z), float(z["ref_err/gnorms_rel_l2"])
        wk, e_full = full_grads_err(model, z)
        r_full = float(z["ref_err/full_grads_rel_l2"])
        print(f"{fixture} fp32 (deterministic) vs reference | reference's own fp32-vs-fp64 error:\n"
              f"  logits rel-l2 {e_log:.2e} | {r_log:.2e}   max abs {e_max:.2e} | {r_max:.2e}   loss {e_loss:.2e} | {r_loss:.2e}\n"
              f"  input grad rel-l2 {e_gin:.2e} | {r_gin:.2e}   grad norms {e_nrm:.2e} | {r_nrm:.2e}   "
              f"worst full grad {wk} {e_full:.2e}")
        assert e_log <= SLACK * r_log, (e_log, r_log)
        assert e_max <= SLACK * r_max, (e_max, r_max)
        assert e_loss <= SLACK * r_loss + 1e-6, (e_loss, r_loss)
        assert e_gin <= SLACK * r_gin, (e_gin, r_gin)
        assert e_nrm <= SLACK * r_nrm, (e_nrm, r_nrm)
        # (the variants' fixtures took the maximum over tensors that are ~0 in fp64 as well: not a usable yardstick; the
        #  worst-conditioned tensor is the first layer's weight, whose reference error is 0.146 in the base model)
        assert e_full <= SLACK * (r_full if r_full < 1.0 else 0.146), (wk, e_full, r_full)
        sd = model.state_dict()
        for k in z.files:
            if k.startswith("upd/"):
                close(sd[k[4:]].float(), torch.from_numpy(z[k]), 1e-3, 1e-3, f"{fixture} buffer {k[4:]}")
        # deterministic mode: a second run from the same state is bit-identical
        grads1 = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        model2 = build(cls_name).train()
        logits2, loss2, gin2 = step(model2, x, m)
        assert torch.equal(logits, logits2) and loss == loss2 and torch.equal(gin, gin2)
        for n, p in model2.named_parameters():
            if p.grad is not None:
                assert torch.equal(p.grad, grads1[n]), f"{n}: gradient differs between two deterministic runs"


def _eval_model(z, dtype=None):
    model = build("ACC_UNet", dtype)
    sd = model.state_dict()
    for k in z.files:
        if k.startswith("eval_sd/"):
            sd[k[8:]].copy_(torch.from_numpy(z[k]))
    return model.eval()


@pytest.mark.parametrize("tf32", [False, True], ids=["fma", "tf32x3"])
def test_full_width_eval_mode_forward_backward_fp32_matches_reference(tf32):
    """eval(): BatchNorm on running statistics, differentiated as a fixed affine (the reference's modules do that too)"""
    z = load_case("full_accunet_224")["raw"]
    x2, m2 = inputs(2, 224, 2025)
    SLACK = SLACK_TF32 if tf32 else globals()["SLACK"]
    with deterministic(), _contractions(tf32):
        model = _eval_model(z)
        logits, loss, gin = step(model, x2, m2)
    e_log, r_log = rel_l2(logits, torch.from_numpy(z["eval/logits"])), float(z["ref_err/eval_logits_rel_l2"])
    e_gin, r_gin = rel_l2(gin, torch.from_numpy(z["eval/gin"])), float(z["ref_err/eval_gin_rel_l2"])
    e_nrm = norms_err(model, z, "eval/")
    wk, e_full = full_grads_err(model, z, "eval/")
    print(f"eval-mode fp32 vs reference | reference's own error: logits {e_log:.2e} | {r_log:.2e}, input grad {e_gin:.2e} | "
          f"{r_gin:.2e}, grad norms {e_nrm:.2e}, worst full grad {wk} {e_full:.2e}, loss {loss:.6f} vs {float(z['eval/loss']):.6f}")
    assert e_log <= SLACK * r_log and e_gin <= SLACK * r_gin
    assert abs(loss - float(z["eval/loss"])) < 1e-4
    assert e_nrm <= SLACK * r_gin and e_full <= SLACK * r_gin
    # conv biases are live parameters in eval mode (nothing cancels them): non-zero gradients, all finite
    g = dict(model.named_parameters())["cnv12.conv1.bias"].grad
    assert g is not None and float(g.abs().max()) > 0 and torch.isfinite(g).all()


def _oracle_bf16(z, x, m, training, variant="base"):
    """the reference arithmetic with bf16 storage (the oracle on this GPU in bfloat16) -> logits, loss, gin"""
    from oracle import acc_oracle as O
    torch.manual_seed(2)
    sd = O.init_acc_unet(3, 1, 32, variant)
    if not training:
        for k in z.files:
            if k.startswith("eval_sd/"):
                sd[k[8:]] = torch.from_numpy(z[k])
    sd = {k: (v.to(DEV).to(torch.bfloat16) if v.is_floating_point() else v.to(DEV)) for k, v in sd.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    xx = x.to(DEV).to(torch.bfloat16).requires_grad_(True)
    logits = O.acc_unet(O.Ctx(sd, training), xx, variant, logits=True)
    loss = O.dice_bce_loss(logits, m.to(DEV))
    loss.backward()
    return logits.detach().float().cpu(), float(loss), xx.grad.detach().float().cpu()


@pytest.mark.parametrize("training", [True, False], ids=["train", "eval"])
def test_full_width_bf16_no_worse_than_reference_arithmetic_in_bf16(training):
    """the bench configuration's storage mode against the reference's full-width vectors"""
    z = load_case("full_accunet_224")["raw"]
    x, m = inputs(2, 224, 2024 if training else 2025)
    pre = "" if training else "eval/"
    ref_logits, ref_loss, ref_gin = torch.from_numpy(z[pre + "logits"]), float(z[pre + "loss"]), torch.from_numpy(z[pre + "gin"])
    model = build("ACC_UNet", torch.bfloat16).train() if training else _eval_model(z, torch.bfloat16)
    logits, loss, gin = step(model, x, m)
    o_logits, o_loss, o_gin = _oracle_bf16(z, x, m, training)
    e = {"logits": rel_l2(logits, ref_logits), "gin": rel_l2(gin, ref_gin), "norms": norms_err(model, z, pre)}
    o = {"logits": rel_l2(o_logits, ref_logits), "gin": rel_l2(o_gin, ref_gin)}
    print(f"bf16 {'train' if training else 'eval'} vs reference (fp32): accx logits {e['logits']:.2e} gin {e['gin']:.2e} "
          f"grad norms {e['norms']:.2e} loss {loss:.5f} | bf16 oracle logits {o['logits']:.2e} gin {o['gin']:.2e} loss "
          f"{o_loss:.5f} | reference loss {ref_loss:.5f}")
    assert torch.isfinite(logits).all() and torch.isfinite(gin).all()
    assert abs(loss - ref_loss) <= 1e-2 * abs(ref_loss), (loss, ref_loss)
    assert e["logits"] <= 1.5 * o["logits"] + 1e-2, (e, o)
    assert e["gin"] <= 1.5 * o["gin"] + 5e-2, (e, o)


def test_full_width_512_forward_fp32_matches_reference():
    z = load_case("full_accunet_512_fwd")["raw"]
    x, _ = inputs(1, 512, 2026)
    with deterministic():
        model = build("ACC_UNet").train()
        with torch.no_grad():
            y = model(x.to(DEV)).float().cpu()
    ref = torch.from_numpy(z["logits"])
    e = rel_l2(y, ref)
    print(f"512x512 train-mode forward (1 image) fp32 vs reference: rel-l2 {e:.2e}, max abs {float((y - ref).abs().max()):.2e} "
          f"of {float(ref.abs().max()):.2e}")
    assert e <= 5e-3, e          # one image per BatchNorm batch: the 224^2 case's conditioning x ~4
