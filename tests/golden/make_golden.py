"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/ACC_UNet/ACC_UNet{,_w,_lite}.py and Experiments/utils.py
read-only, runs each hot-path module forward + backward on seeded inputs in fp32 on
the CPU and stores inputs, initial state, outputs, gradients and updated BN buffers
as small .npz fixtures next to this script.  The GPU box has no /root/reference; the
tests there only read the fixtures.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "ACC_UNet"))
HERE = os.path.dirname(os.path.abspath(__file__))

import ACC_UNet as R  # noqa: E402
import ACC_UNet_w as RW  # noqa: E402
import ACC_UNet_lite as RL  # noqa: E402


def perturb(mod, seed):
    """Make BN affine/running stats (and the W blend scalar) non-trivial."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in mod.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
                m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) + 0.5)
        for n, p in mod.named_parameters():
            if n.endswith("W") and p.numel() == 1:
                p.fill_(0.3)


def run(name, mod, inputs, seed, whole_model=False):
    """whole_model=True: weights are NOT stored (the oracle's initialiser reproduces them from
    seed 2, see init_seed2_f8.npz) and parameter gradients are stored as per-tensor summaries."""
    torch.manual_seed(seed + 1)
    out = {}
    if not whole_model:
        perturb(mod, seed + 2)
        for k, v in mod.state_dict().items():
            out["sd/" + k] = v.detach().clone().numpy()
    xs = [x.clone().requires_grad_(True) for x in inputs]
    mod.train()
    ys = mod(*xs)
    ys = ys if isinstance(ys, (tuple, list)) else (ys,)
    loss = 0
    for i, y in enumerate(ys):
        r = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 10 + i))
        out[f"cot/{i}"] = r.numpy()
        out[f"out/{i}"] = y.detach().numpy()
        loss = loss + (y * r).sum()
    loss.backward()
    for i, x in enumerate(xs):
        out[f"in/{i}"] = inputs[i].numpy()
        out[f"gin/{i}"] = x.grad.numpy()
    if whole_model:
        # The whole net is badly conditioned at random init (the reference's own fp32 and fp64
        # runs disagree by ~10 % on gradients), so also store the fp64 run: tests bound
        # |ours - ref64| by a multiple of |ref32 - ref64|.
        # (rebuilt from seed 2 so the running stats are the initial ones again)
        torch.manual_seed(2)
        m64 = type(mod)(3, 1, 8).double()
        m64.train()
        x64 = inputs[0].double().requires_grad_(True)
        y64 = m64(x64)
        (y64 * torch.from_numpy(out["cot/0"]).double()).sum().backward()
        out["out64/0"] = y64.detach().numpy()
        out["gin64/0"] = x64.grad.numpy()
        out["gpsum/l2_64"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                       for _, p in m64.named_parameters()], dtype=np.float64)
        names = [k for k, p in mod.named_parameters()]
        out["gpsum/names"] = np.array(names)
        out["gpsum/l2"] = np.array([float(p.grad.norm()) if p.grad is not None else -1.0
                                    for _, p in mod.named_parameters()], dtype=np.float64)
        out["gpsum/sum"] = np.array([float(p.grad.double().sum()) if p.grad is not None else 0.0
                                     for _, p in mod.named_parameters()], dtype=np.float64)
        for k, p in mod.named_parameters():
            if p.grad is not None and (p.numel() <= 1024 or k in ("cnv11.conv1.weight", "cnv92.conv3.weight")):
                if k.startswith(("cnv11.", "cnv92.", "mlfc3.sqe", "rspth4.", "out.", "mlfc1.W")):
                    out["gp/" + k] = p.grad.numpy()
        for k, v in mod.state_dict().items():
            if "running_" in k and k.startswith(("cnv11.", "cnv92.", "rspth4.", "mlfc2.sqe")):
                out["upd/" + k] = v.detach().numpy()
    else:
        for k, p in mod.named_parameters():
            if p.grad is not None:
                out["gp/" + k] = p.grad.numpy()
        for k, v in mod.state_dict().items():
            if "running_" in k or "num_batches" in k:
                out["upd/" + k] = v.detach().numpy()
    mod.eval()
    with torch.no_grad():
        ys = mod(*inputs)
    ys = ys if isinstance(ys, (tuple, list)) else (ys,)
    for i, y in enumerate(ys):
        out[f"eval/{i}"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith(("sd/", "gp/", "upd/"))})


def main():
    torch.set_num_threads(8)
    g = lambda *s, seed=0: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
    torch.manual_seed(2); run("se_c16", R.ChannelSELayer(16), [g(2, 16, 8, 8, seed=1)], 100)
    torch.manual_seed(2); run("se_c40", R.ChannelSELayer(40), [g(3, 40, 6, 10, seed=2)], 101)
    for k in (1, 2, 3, 4, 5):
        hw = 16 if k == 5 else 8
        torch.manual_seed(2); run(f"hanclayer_k{k}", R.HANCLayer(8, 16, k), [g(2, 8, hw, hw, seed=3 + k)], 110 + k)
    torch.manual_seed(2); run("hancblock_8_16_k3", R.HANCBlock(8, 16, k=3, inv_fctr=3), [g(2, 8, 16, 16, seed=20)], 120)
    torch.manual_seed(2); run("hancblock_3_8_k3", R.HANCBlock(3, 8, k=3, inv_fctr=3), [g(2, 3, 16, 12, seed=21)], 121)
    torch.manual_seed(2); run("hancblock_16_16_k2", R.HANCBlock(16, 16, k=2, inv_fctr=3), [g(2, 16, 8, 8, seed=22)], 122)
    torch.manual_seed(2); run("hancblock_16_8_k1", R.HANCBlock(16, 8, k=1, inv_fctr=3), [g(3, 16, 4, 4, seed=23)], 123)
    torch.manual_seed(2); run("hancblock_8_8_k3_f34", R.HANCBlock(8, 8, k=3, inv_fctr=34), [g(1, 8, 8, 8, seed=24)], 124)
    torch.manual_seed(2); run("respath_c16_n2", R.ResPath(16, 2), [g(2, 16, 8, 8, seed=30)], 130)
    torch.manual_seed(2); run("respath_c8_n4", R.ResPath(8, 4), [g(1, 8, 12, 8, seed=31)], 131)
    ml_in = lambda cs, hw, seed: [g(2, c, hw >> i, hw >> i, seed=seed + i) for i, c in enumerate(cs)]
    torch.manual_seed(2); run("mlfc_8_16_32_64", R.MLFC(8, 16, 32, 64, lenn=1), ml_in((8, 16, 32, 64), 16, 40), 140)
    torch.manual_seed(2); run("mlfc_16_32_128_160", R.MLFC(16, 32, 128, 160, lenn=1), ml_in((16, 32, 128, 160), 8, 50), 141)
    torch.manual_seed(2); run("mlfc_8_8_16_16_len2", R.MLFC(8, 8, 16, 16, lenn=2), ml_in((8, 8, 16, 16), 8, 60), 142)
    torch.manual_seed(2); run("mlfcw_8_16_32_64", RW.MLFC(8, 16, 32, 64, lenn=1), ml_in((8, 16, 32, 64), 16, 70), 143)
    torch.manual_seed(2); run("mlfclite_8_16_32_64", RL.MLFC(8, 16, 32, 64, lenn=1), ml_in((8, 16, 32, 64), 16, 80), 144)
    # standalone Conv2d_batchnorm (ACC_UNet.py:146-186) and the channel set the fKAN consumer instantiates
    # (Experiments/nets/archs/archs_InceptionNext_MLFC_fKAN.py:428)
    torch.manual_seed(2); run("convbn_24_16", R.Conv2d_batchnorm(24, 16, (1, 1)), [g(2, 24, 8, 8, seed=35)], 135)
    ml1 = lambda cs, hw, seed: [g(1, c, hw >> i, hw >> i, seed=seed + i) for i, c in enumerate(cs)]
    torch.manual_seed(2); run("mlfc_80_128_160_160", R.MLFC(80, 128, 160, 160, lenn=1), ml1((80, 128, 160, 160), 16, 85), 145)
    # whole models, narrow (n_filts=8) so the fixture stays small
    for nm, cls in (("accunet_f8", R.ACC_UNet), ("accunetw_f8", RW.ACC_UNet_W), ("accunetlite_f8", RL.ACC_UNet_Lite)):
        torch.manual_seed(2)
        m = cls(3, 1, 8)
        run(nm, m, [g(2, 3, 32, 32, seed=90)], 150, whole_model=True)
    # loss known-answer vectors (Experiments/utils.py WeightedDiceBCE(0.5, 0.5))
    sys.path.insert(0, os.path.join(REF, "Experiments"))
    import utils as U
    lg = g(3, 1, 16, 16, seed=95) * 3
    tr = (torch.rand(3, 1, 16, 16, generator=torch.Generator().manual_seed(96)) > 0.5).float()
    lgr = lg.clone().requires_grad_(True)
    loss = U.WeightedDiceBCE(dice_weight=0.5, BCE_weight=0.5)(lgr, tr)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "loss_dicebce.npz"), logit=lg.numpy(), truth=tr.numpy(),
                        loss=loss.detach().numpy(), glogit=lgr.grad.numpy())
    # init parity: seed 2 -> first weights of the reference constructors
    torch.manual_seed(2)
    m = R.ACC_UNet(3, 1, 8)
    sd = m.state_dict()
    keys = ["cnv11.conv1.weight", "cnv52.sqe.fc2.bias", "rspth1.convs.3.bias", "mlfc2.cnv_mrg3.0.conv1.weight",
            "up7.bias", "out.weight", "out.bias"]
    allk = list(sd.keys())
    np.savez_compressed(os.path.join(HERE, "init_seed2_f8.npz"), **{"t/" + k: sd[k].numpy() for k in keys},
                        names=np.array(allk),
                        shapes=np.array([str(tuple(sd[k].shape)) for k in allk]),
                        sums=np.array([float(sd[k].double().sum()) for k in allk]),
                        abssums=np.array([float(sd[k].double().abs().sum()) for k in allk]))

    # the training harness's flavour of the model (Experiments/nets/ACC_UNet.py: cnv72 inv_fctr=3, logits out)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_nets_ACC_UNet", os.path.join(REF, "Experiments", "nets", "ACC_UNet.py"))
    RH = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(RH)
    torch.manual_seed(2)
    m = RH.ACC_UNet(3, 1, 8)
    sd = m.state_dict()
    allk = list(sd.keys())
    x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(7))
    m.eval()
    with torch.no_grad():
        y = m(x)
    np.savez_compressed(os.path.join(HERE, "init_seed2_f8_harness.npz"), names=np.array(allk),
                        shapes=np.array([str(tuple(sd[k].shape)) for k in allk]),
                        sums=np.array([float(sd[k].double().sum()) for k in allk]), x=x.numpy(), eval_logits=y.numpy())

    # per-step metrics known-answer vectors (Experiments/utils.py:478-494 iou_on_batch, :148-157 _show_dice)
    import warnings
    warnings.filterwarnings("ignore")
    gm = torch.Generator().manual_seed(321)
    cases = {}

    def add(name, lg, tr):
        iou = U.iou_on_batch(tr.clone(), lg.clone())
        dice = float(U.WeightedDiceBCE(dice_weight=0.5, BCE_weight=0.5)._show_dice(lg.clone(), tr.clone().float()))
        cases[name + "/logit"], cases[name + "/truth"] = lg.numpy(), tr.numpy()
        cases[name + "/iou"], cases[name + "/dice"] = np.array(float(iou)), np.array(dice)

    lg = torch.randn(4, 1, 32, 32, generator=gm) * 2
    tr = (torch.rand(4, 1, 32, 32, generator=gm) > 0.6).float()
    add("random", lg, tr)
    lg2, tr2 = lg.clone(), tr.clone()
    lg2[0, 0, :4] = 0.0            # exact zeros: sigmoid = 0.5 counts as positive
    lg2[1] = -3.0                  # an image without a positive prediction ...
    tr2[1] = 0.0                   # ... and without a positive mask pixel: empty union -> jaccard_score 0
    tr2[2] = 1.0
    add("edges", lg2, tr2)
    lg3 = torch.randn(3, 1, 20, 28, generator=gm)
    tr3 = (torch.rand(3, 1, 20, 28, generator=gm) * 255).round() * (torch.rand(3, 1, 20, 28, generator=gm) > 0.5)
    add("graymask", lg3, tr3.float())
    np.savez_compressed(os.path.join(HERE, "metrics_kat.npz"), **cases)


if __name__ == "__main__":
    main()
