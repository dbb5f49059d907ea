"""CPU: pin the oracle (oracle/acc_oracle.py) against fixtures produced by the reference itself."""
import numpy as np
import pytest
import torch

from helpers import close, load_case, module_cases, whole_model_checks
from oracle import acc_oracle as O

RTOL, ATOL = 1e-4, 2e-5


def run_oracle(name, case, training=True, prepared=False):
    if prepared:        # caller already placed / typed / marked the tensors
        sd, xs = case["sd"], case["in"]
    else:
        sd = {k: v.clone() for k, v in case["sd"].items()}
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(training)
        xs = [x.clone().requires_grad_(training) for x in case["in"]]
    cx = O.Ctx(sd, training)
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
