"""Full-width golden vectors from the UNMODIFIED reference (run in the build container only; ~10 min of CPU).

    python tests/golden/make_golden_full.py

The benchmarked configuration -- ACC_UNet(3, 1, 32), Dice+BCE on logits -- at its real width, on a batch the CPU
can hold (2 x 3 x 224 x 224; the reference saves 4.35 GB of activations per image), plus the Lite / W variants and
one 512 x 512 forward.  Weights are NOT stored: torch.manual_seed(2) + the constructor reproduces them on the GPU
box (tests/test_abi_cpu.py::test_default_init_reproduces_reference_weights pins that), inputs come from seeded CPU
generators.  Stored per case: logits, loss, input gradient, per-parameter gradient norms / sums, a few full parameter
gradients, BN running statistics after the step; for the base model also an EVAL-mode forward + backward (BatchNorm
as a fixed affine: no batch-statistics coupling, so it is well conditioned and bf16 can be held to north_star's rtol
element-wise) with its buffers.  The reference's own fp32 rounding error is measured against an fp64 run of the same
modules and stored as scalars (`ref_err/*`): the whole net at random init amplifies rounding noise through 220
training-mode BatchNorms, and a bound on |ours - reference| is only meaningful relative to |reference32 - reference64|.
"""
import os
import sys
import time

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "ACC_UNet"))
sys.path.insert(0, os.path.join(REF, "Experiments"))
HERE = os.path.dirname(os.path.abspath(__file__))

import ACC_UNet as R  # noqa: E402
import ACC_UNet_w as RW  # noqa: E402
import ACC_UNet_lite as RL  # noqa: E402
import utils as U  # noqa: E402

FULL_GRADS = ("cnv11.conv1.weight", "cnv11.conv2.weight", "cnv11.norm1.weight", "cnv12.hnc.cnv.weight", "cnv32.sqe.fc1.weight",
              "cnv92.conv3.weight", "cnv92.norm3.weight", "cnv92.sqe.fc2.weight", "rspth1.convs.0.weight", "rspth2.bn.weight",
              "mlfc1.cnv_blks1.0.conv1.weight", "mlfc2.sqe3.fc1.weight", "mlfc1.W", "mlfc2.W", "mlfc3.W", "up9.weight",
              "out.weight", "out.bias")
# large tensors: the first PROBE elements of the flattened gradient
PROBE = 4096
PROBE_GRADS = ("cnv52.conv3.weight", "cnv72.conv1.weight", "cnv72.hnc.cnv.weight", "rspth4.convs.0.weight",
               "mlfc3.cnv_mrg4.0.conv1.weight", "up6.weight")


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp(min=1e-300))


def inputs(B, hw, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, hw, hw, generator=g)
    m = (torch.rand(B, 1, hw, hw, generator=g) > 0.5).float()
    return x, m


def step(model, x, m, dtype=torch.float32):
    """forward (logits) + WeightedDiceBCE(0.5, 0.5) + backward -> logits, loss, gin"""
    model.zero_grad(set_to_none=True)
    xx = x.detach().clone().to(dtype).requires_grad_(True)
    logits = model(xx)
    loss = U.WeightedDiceBCE(dice_weight=0.5, BCE_weight=0.5)(logits, m.to(dtype))
    loss.backward()
    return logits.detach(), loss.detach(), xx.grad.detach()


def build(cls, dtype=torch.float32):
    torch.manual_seed(2)
    model = cls(3, 1, 32)
    model.last_activation = None            # logits for the logit-based loss (ACC_UNet.py:653-657)
    return model.to(dtype)


def grads_summary(model, out, tag):
    names = [k for k, _ in model.named_parameters()]
    out[tag + "gpsum/names"] = np.array(names)
    out[tag + "gpsum/l2"] = np.array([float(p.grad.double().norm()) if p.grad is not None else -1.0
                                      for _, p in model.named_parameters()])
    out[tag + "gpsum/sum"] = np.array([float(p.grad.double().sum()) if p.grad is not None else 0.0
                                       for _, p in model.named_parameters()])
    for k, p in model.named_parameters():
        if p.grad is not None and k in FULL_GRADS:
            out[tag + "gp/" + k] = p.grad.float().numpy()
        if p.grad is not None and k in PROBE_GRADS:
            out[tag + "gprobe/" + k] = p.grad.float().reshape(-1)[:PROBE].numpy().copy()


def case(name, cls, with_eval, store_gin):
    t0 = time.time()
    x, m = inputs(2, 224, 2024)
    out = {}
    model = build(cls)
    model.train()
    logits, loss, gin = step(model, x, m)
    out["logits"], out["loss"] = logits.numpy(), loss.numpy()
    if store_gin:
        out["gin"] = gin.numpy()
    out["gin_l2"] = np.array(float(gin.double().norm()))
    out["gin_probe"] = gin[:, :, ::8, ::8].contiguous().numpy()
    grads_summary(model, out, "")
    for k, v in model.state_dict().items():
        if "running_" in k and k.startswith(("cnv11.", "cnv52.", "cnv92.", "rspth4.", "mlfc2.sqe", "mlfc3.cnv_mrg1")):
            out["upd/" + k] = v.detach().numpy()
    print(f"{name}: fp32 train step done ({time.time() - t0:.0f} s), loss {float(loss):.6f}", flush=True)
    # the reference's own rounding error: same modules in fp64
    m64 = build(cls, torch.float64)
    m64.train()
    l64, loss64, g64 = step(m64, x, m, torch.float64)
    n32 = out["gpsum/l2"]
    n64 = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for _, p in m64.named_parameters()])
    keep = n64 >= 0
    out["ref_err/logits_rel_l2"] = np.array(rel_l2(logits, l64))
    out["ref_err/logits_max_abs"] = np.array(float((logits.double() - l64).abs().max()))
    out["ref_err/logits_scale"] = np.array(float(l64.abs().max()))
    out["ref_err/loss_abs"] = np.array(abs(float(loss) - float(loss64)))
    out["ref_err/gin_rel_l2"] = np.array(rel_l2(gin, g64))
    out["ref_err/gnorms_rel_l2"] = np.array(float(np.linalg.norm(n32[keep] - n64[keep]) / np.linalg.norm(n64[keep])))
    full = {k: p.grad for k, p in m64.named_parameters() if p.grad is not None and k in FULL_GRADS}
    out["ref_err/full_grads_rel_l2"] = np.array(max(rel_l2(torch.from_numpy(out["gp/" + k]), v) for k, v in full.items()))
    print(f"{name}: fp64 run done ({time.time() - t0:.0f} s); reference fp32-vs-fp64: logits rel-l2 "
          f"{float(out['ref_err/logits_rel_l2']):.2e}, gin {float(out['ref_err/gin_rel_l2']):.2e}, grad norms "
          f"{float(out['ref_err/gnorms_rel_l2']):.2e}, full grads {float(out['ref_err/full_grads_rel_l2']):.2e}", flush=True)
    del m64
    if with_eval:
        # EVAL mode with meaningful running statistics: one training-mode forward with momentum 1 makes them the batch
        # statistics of x (attributes of the reference's own BatchNorm2d modules; its code is untouched)
        model = build(cls)
        bns = [mod for mod in model.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
        for b in bns:
            b.momentum = 1.0
        model.train()
        with torch.no_grad():
            model(x)
        for b in bns:
            b.momentum = 0.1
        model.eval()
        x2, m2 = inputs(2, 224, 2025)
        for k, v in model.state_dict().items():
            if "running_" in k:
                out["eval_sd/" + k] = v.detach().numpy().copy()
        le, losse, ge = step(model, x2, m2)
        out["eval/logits"], out["eval/loss"], out["eval/gin"] = le.numpy(), losse.numpy(), ge.numpy()
        grads_summary(model, out, "eval/")
        m64 = build(cls, torch.float64)
        sd = {k: (v.double() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
        m64.load_state_dict(sd)
        m64.eval()
        l64, _, g64 = step(m64, x2, m2, torch.float64)
        out["ref_err/eval_logits_rel_l2"] = np.array(rel_l2(le, l64))
        out["ref_err/eval_gin_rel_l2"] = np.array(rel_l2(ge, g64))
        print(f"{name}: eval-mode fwd+bwd done ({time.time() - t0:.0f} s); reference fp32-vs-fp64: logits "
              f"{float(out['ref_err/eval_logits_rel_l2']):.2e}, gin {float(out['ref_err/eval_gin_rel_l2']):.2e}", flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "written", os.path.getsize(os.path.join(HERE, name + ".npz")) >> 10, "KiB", flush=True)


def forward_512():
    x, _ = inputs(1, 512, 2026)
    model = build(R.ACC_UNet)
    model.train()
    with torch.no_grad():
        y = model(x)
    np.savez_compressed(os.path.join(HERE, "full_accunet_512_fwd.npz"), logits=y.numpy())
    print("full_accunet_512_fwd written", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["base", "lite", "w", "512"]
    if "base" in which:
        case("full_accunet_224", R.ACC_UNet, with_eval=True, store_gin=True)
    if "lite" in which:
        case("full_accunetlite_224", RL.ACC_UNet_Lite, with_eval=False, store_gin=False)
    if "w" in which:
        case("full_accunetw_224", RW.ACC_UNet_W, with_eval=False, store_gin=False)
    if "512" in which:
        forward_512()
