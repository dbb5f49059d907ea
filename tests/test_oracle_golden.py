"""CPU: pin the oracle (oracle/acc_oracle.py) against fixtures produced by the reference itself."""
import numpy as np
import pytest
import torch

from helpers import close, load_case, module_cases, whole_model_checks
from oracle import acc_oracle as O

RTOL, ATOL = 1e-4, 2e-5


def run_oracle(name, case, training=True, prepared=False, forced=None, record=None):
    if prepared:        # caller already placed / typed / marked the tensors
        sd, xs = case["sd"], case["in"]
    else:
        sd = {k: v.clone() for k, v in case["sd"].items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(training)
        xs = [x.clone().requires_grad_(training) for x in case["in"]]
    cx = O.Ctx(sd, training, forced=forced, record=record)
    kind = name.split("_")[0]
    if kind == "se":
        ys = (O.se_layer(cx, "", xs[0]),)
    elif kind == "hanclayer":
        ys = (O.hanc_layer(cx, "", xs[0], int(name[-1])),)
    elif kind == "hancblock":
        k = int(name.split("_k")[1][0])
        ys = (O.hanc_block(cx, "", xs[0], k),)
    elif kind == "respath":
        ys = (O.respath(cx, "", xs[0], int(name.split("_n")[1])),)
    elif kind == "convbn":
        ys = (O.conv_bn_se(cx, "", xs[0]),)
    else:
        variant = {"mlfc": "base", "mlfcw": "w", "mlfclite": "lite"}[kind]
        lenn = 2 if name.endswith("len2") else 1
        ys = O.mlfc(cx, "", xs, lenn, variant)
    return cx, sd, xs, ys


def strip(sd):
    # fixtures were saved from a bare module: keys have no leading prefix; oracle uses name + ".x"
    return {"." + k: v for k, v in sd.items()}


@pytest.mark.parametrize("name", module_cases())
def test_module_matches_reference(name):
    case = load_case(name)
    case["sd"] = strip(case["sd"])
    cx, sd, xs, ys = run_oracle(name, case, True)
    loss = sum((y * r).sum() for y, r in zip(ys, case["cot"]))
    loss.backward()
    for i, y in enumerate(ys):
        close(y, case["out"][i], RTOL, ATOL, f"{name} out{i}")
    for i, x in enumerate(xs):
        close(x.grad, case["gin"][i], RTOL, ATOL, f"{name} gin{i}")
    wscale = max(float(v.abs().max()) for k, v in case["gp"].items() if k.endswith("weight"))
    for k, g in case["gp"].items():
        got = sd["." + k].grad
        assert got is not None, k
        if float(g.abs().max()) < 1e-4 * wscale:     # analytically-zero conv bias grads
            close(got, g, 0, 1e-4, f"{name} grad {k}", zero_scale=wscale)
        else:
            close(got, g, 1e-3, 2e-4, f"{name} grad {k}")
    for k, v in case["upd"].items():
        got = cx.updates.get("." + k, case["sd"]["." + k])   # Lite leaves unused BN buffers untouched
        close(got.float(), v.float(), RTOL, ATOL, f"{name} buffer {k}")
    # eval mode uses the updated running stats
    for k, v in cx.updates.items():
        case["sd"][k] = v
    _, _, _, ys = run_oracle(name, case, False)
    for i, y in enumerate(ys):
        close(y, case["eval"][i], RTOL, ATOL, f"{name} eval{i}")


def test_init_matches_reference_constructors():
    z = np.load("tests/golden/init_seed2_f8.npz") if False else load_case("init_seed2_f8")["raw"]
    torch.manual_seed(2)
    sd = O.init_acc_unet(3, 1, 8)
    names = [str(n) for n in z["names"]]
    assert set(names) == set(sd.keys())
    for n, shp, s, a in zip(names, z["shapes"], z["sums"], z["abssums"]):
        t = sd[n]
        assert str(tuple(t.shape)) == str(shp), n
        assert abs(float(t.double().sum()) - s) <= 1e-9 + 1e-9 * abs(a), n
        assert abs(float(t.double().abs().sum()) - a) <= 1e-9 * abs(a) + 1e-9, n
    for k in z.files:
        if k.startswith("t/"):
            assert torch.equal(sd[k[2:]], torch.from_numpy(z[k])), k


@pytest.mark.parametrize("name,variant", [("accunet_f8", "base"), ("accunetw_f8", "w"), ("accunetlite_f8", "lite")])
def test_whole_model_matches_reference(name, variant):
    case = load_case(name)
    z = case["raw"]
    torch.manual_seed(2)
    sd = O.init_acc_unet(3, 1, 8, variant)
    for k in O.trainable(sd):
        sd[k].requires_grad_(True)
    x = case["in"][0].clone().requires_grad_(True)
    cx = O.Ctx(sd, True)
    y = O.acc_unet(cx, x, variant)
    (y * case["cot"][0]).sum().backward()
    whole_model_checks(name, case, y, x.grad, {n: sd[n].grad for n in O.trainable(sd)}, cx.updates)
    for k, v in cx.updates.items():
        sd[k] = v
    with torch.no_grad():
        ye = O.acc_unet(O.Ctx(sd, False), case["in"][0], variant)
    close(ye, case["eval"][0], 1e-3, 2e-3, f"{name} eval")


def test_loss_known_answer():
    z = load_case("loss_dicebce")["raw"]
    lg = torch.from_numpy(z["logit"]).requires_grad_(True)
    loss = O.dice_bce_loss(lg, torch.from_numpy(z["truth"]))
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    close(lg.grad, torch.from_numpy(z["glogit"]), 1e-4, 1e-5, "dloss/dlogit")


@pytest.mark.parametrize("name", ["hancblock_8_16_k3", "respath_c16_n2", "mlfc_8_16_32_64", "se_c16"])
def test_decision_record_and_replay(name):
    """Ctx.record / Ctx.forced (the flip-free gradient comparison of the bf16 GPU tests): replaying an evaluation's
    own decisions reproduces it bit for bit; replaying them on PERTURBED inputs keeps every LeakyReLU sign and pool
    arg-max where the recording had it (the gradient is then linear in the cotangent with fixed routing)."""
    case = load_case(name)
    case["sd"] = strip(case["sd"])
    rec = {}
    cx, sd, xs, ys = run_oracle(name, case, True, record=rec)
    assert rec and all(v.dtype in (torch.bool, torch.int64) for v in rec.values())
    sum((y * r).sum() for y, r in zip(ys, case["cot"])).backward()
    g0 = [x.grad.clone() for x in xs]
    cx2, sd2, xs2, ys2 = run_oracle(name, case, True, forced=rec)
    sum((y * r).sum() for y, r in zip(ys2, case["cot"])).backward()
    for a, b in zip(ys, ys2):
        assert torch.equal(a, b)
    for a, b in zip(g0, xs2):
        assert torch.equal(a, b.grad)
    # perturbed inputs: decisions stay those of the recording
    pert = dict(case)
    pert["in"] = [x + 0.05 * torch.randn(x.shape, generator=torch.Generator().manual_seed(12345)) for x in case["in"]]
    rec3 = {}
    run_oracle(name, pert, True, forced=rec, record=rec3)
    assert set(rec3) == set(rec) and all(torch.equal(rec3[k], rec[k]) for k in rec)
    free = {}
    run_oracle(name, pert, True, record=free)
    assert any(not torch.equal(free[k], rec[k]) for k in rec)      # left alone, the perturbed run does decide differently


@pytest.mark.parametrize("case", ["random", "edges", "graymask"])
def test_seg_metrics_match_reference_known_answers(case):
    z = load_case("metrics_kat")["raw"]
    iou, dice = O.seg_metrics(torch.from_numpy(z[case + "/logit"]), torch.from_numpy(z[case + "/truth"]))
    assert abs(iou - float(z[case + "/iou"])) < 1e-12, (iou, float(z[case + "/iou"]))
    assert abs(dice - float(z[case + "/dice"])) < 2e-6, (dice, float(z[case + "/dice"]))     # the reference sums in fp32


def test_oracle_full_width_forward_and_loss_match_reference():
    """the benchmarked model at its real width: ACC_UNet(3, 1, 32) logits + WeightedDiceBCE on 2x3x224x224
    (fixture from the unmodified reference, tests/golden/make_golden_full.py)"""
    z = load_case("full_accunet_224")["raw"]
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(2, 3, 224, 224, generator=g)
    m = (torch.rand(2, 1, 224, 224, generator=g) > 0.5).float()
    torch.manual_seed(2)
    sd = O.init_acc_unet(3, 1, 32)
    with torch.no_grad():
        logits = O.acc_unet(O.Ctx(sd, True), x, "base", logits=True)
        loss = O.dice_bce_loss(logits, m)
    ref = torch.from_numpy(z["logits"])
    err = float((logits - ref).norm() / ref.norm())
    # same arithmetic on the same CPU, different operator decomposition (einsum / repeat_interleave vs Conv2d / Upsample):
    # the difference is rounding noise amplified by the net's conditioning.  Two independent fp32 evaluations that are
    # each e away from the exact (fp64) result differ by ~sqrt(2) e; the fixture holds the reference's own e (3.6e-4)
    assert err < 3.0 * float(z["ref_err/logits_rel_l2"]), (err, float(z["ref_err/logits_rel_l2"]))
    assert abs(float(loss) - float(z["loss"])) < 1e-5
