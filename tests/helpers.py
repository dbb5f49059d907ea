"""Shared helpers for the parity tests (fixtures loading, tolerance policy)."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {"sd": {}, "gp": {}, "upd": {}, "in": [], "out": [], "cot": [], "gin": [], "eval": [], "raw": z}
    for k in z.files:
        head, _, rest = k.partition("/")
        if head in ("sd", "gp", "upd"):
            d[head][rest] = torch.from_numpy(z[k])
    for head in ("in", "out", "cot", "gin", "eval"):
        i = 0
        while f"{head}/{i}" in z.files:
            d[head].append(torch.from_numpy(z[f"{head}/{i}"]))
            i += 1
    return d


def module_cases():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    return [n for n in names if n.startswith(("se_", "hanclayer_", "hancblock_", "respath_", "mlfc", "convbn_"))]


def close(a, b, rtol, atol_rel, what="", zero_scale=None):
    """|a-b| <= atol + rtol*|b| with atol = atol_rel * max|b| (SURVEY 8c tolerance policy).
    zero_scale: for tensors that are analytically zero (conv bias grads upstream of a
    training BN) compare against zero with atol = atol_rel * zero_scale."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    if zero_scale is not None:
        lim = atol_rel * zero_scale
        worst = float(a.abs().max()) if a.numel() else 0.0
        assert worst <= lim + 1e-30, f"{what}: expected ~0, got max {worst:.3e} > {lim:.3e}"
        return
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = (a - b).abs() - rtol * b.abs()
    worst = float(err.max()) if err.numel() else 0.0
    assert worst <= atol_rel * scale + 1e-30, (
        f"{what}: max excess {worst:.3e} > atol {atol_rel * scale:.3e} (scale {scale:.3e}, "
        f"max abs diff {float((a - b).abs().max()):.3e})")


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def whole_model_checks(name, case, y, gin, grads, updates, slack=3.0, floor=2e-3):
    """End-to-end tolerance calibrated on the reference itself: the whole net is badly
    conditioned at random init (reference fp32 vs fp64 gradients differ by ~10 %), so the
    error against the fp64 reference run must stay within `slack` x the fp32 reference's own
    error (+ floor).  Module-level tests carry the tight rtol-1e-3 parity."""
    z = case["raw"]
    out32, out64 = case["out"][0], torch.from_numpy(z["out64/0"])
    gin32, gin64 = case["gin"][0], torch.from_numpy(z["gin64/0"])
    e_out = float((y.detach().double().cpu() - out64).abs().max())
    ref_out = float((out32.double() - out64).abs().max())
    assert e_out <= slack * ref_out + floor, f"{name} out: {e_out:.3e} vs reference's own {ref_out:.3e}"
    e_g, ref_g = rel_l2(gin, gin64), rel_l2(gin32, gin64)
    assert e_g <= slack * ref_g + floor, f"{name} gin rel-l2 {e_g:.3e} vs reference's own {ref_g:.3e}"
    names = [str(n) for n in z["gpsum/names"]]
    l2_32, l2_64 = z["gpsum/l2"], z["gpsum/l2_64"]
    ours = []
    for n, r64 in zip(names, l2_64):
        g = grads.get(n)
        if r64 < 0:
            assert g is None or float(g.abs().max()) == 0.0, n
            ours.append(-1.0)
        else:
            assert g is not None and torch.isfinite(g).all(), n
            ours.append(float(g.norm()))
    import numpy as np
    ours = np.array(ours)
    keep = l2_64 >= 0
    e_n = float(np.linalg.norm(ours[keep] - l2_64[keep]) / np.linalg.norm(l2_64[keep]))
    ref_n = float(np.linalg.norm(l2_32[keep] - l2_64[keep]) / np.linalg.norm(l2_64[keep]))
    assert e_n <= slack * ref_n + 0.02, f"{name} per-tensor grad norms rel-l2 {e_n:.3e} vs reference's own {ref_n:.3e}"
    for k, v in case["upd"].items():
        if k in updates:
            close(updates[k].float(), v, 1e-3, 2e-3, f"{name} buffer {k}")


def close_frac(a, b, rtol, atol_rel, what, max_frac):
    """like close(), but lets a fraction `max_frac` of the elements miss the tolerance: gradients that
    pass through max-pool / LeakyReLU are discontinuous in the activations, so a rounding-level
    difference at a near-tie re-routes an isolated element (the reference on another device does too)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    scale = float(b.abs().max()) if b.numel() else 0.0
    bad = (a - b).abs() > atol_rel * scale + rtol * b.abs()
    frac = float(bad.double().mean()) if bad.numel() else 0.0
    assert frac <= max_frac, (f"{what}: {frac:.2e} of elements outside rtol {rtol} / atol {atol_rel * scale:.2e} "
                              f"(allowed {max_frac:.1e}); rel-l2 {rel_l2(a, b):.2e}")


def flat_cat(tensors):
    return torch.cat([t.detach().double().cpu().reshape(-1) for t in tensors])


class deterministic:
    """with deterministic(on): accx runs with fixed-order reductions (engine.set_deterministic) -- the mode every
    fp32 parity bound in tests/ is stated for: no run-to-run spread, so a miss is a bug and not a draw of the atomics"""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        if self.on:
            from accx import engine
            engine.set_deterministic(True)
        return self

    def __exit__(self, *exc):
        if self.on:
            from accx import engine
            torch.cuda.synchronize()
            engine.set_deterministic(False)
        return False


def accx_decisions(mod, rec, dotted=True):
    """engine.RECORD of ONE accx forward of `mod` -> {oracle site: decision}: the sign under every LeakyReLU
    (bool, NCHW) and the first-maximum index of every max-pool window, exactly as the CUDA run took them.
    dotted: module-level oracle calls use the name "" + ".child" (sites start with a dot); acc_unet() does not."""
    from oracle import acc_oracle as O

    def activated(L):
        y = L.y.float()
        if L.act:
            y = y * L.scale + L.shift
            if L.act == 2:
                y = torch.where(y > 0, y, 0.01 * y)
        return y

    def windows(a, s):                       # NCHW -> [B, C, H/s, W/s, s*s], row-major inside the window
        B, C, H, W = a.shape
        return a.reshape(B, C, H // s, s, W // s, s).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // s, W // s, s * s)

    out = {}
    for name, m in mod.named_modules():
        pre = ("." + name if name else "") if dotted else name
        L = rec.get(("bn", id(m)))
        if L is not None:
            out[pre] = ((L.y.float() * L.scale + L.shift) > 0).permute(0, 3, 1, 2)
        c = rec.get(("se", id(m)))
        if c is not None:
            a = activated(c.L)
            B, H, W, C = a.shape
            z = a * (c.gate.view(B, 1, 1, C) * c.scale) + c.shift
            out[pre + ".bn"] = (z > 0).permute(0, 3, 1, 2)
            out[pre + ".fc1"] = c.hidden.view(B, -1) > 0
        L2 = rec.get(("hanc", id(m)))
        if L2 is not None:
            a = activated(L2).permute(0, 3, 1, 2)
            for j in range(1, m.k):
                out[f"{pre}.max{j}"] = O.first_argmax(windows(a, 2 ** j))
    for i, x in enumerate(rec.get("pool", [])):
        out[f"pool{i + 1}"] = O.first_argmax(windows(x.float().permute(0, 3, 1, 2), 2))
    return out


class record_decisions:
    """with record_decisions() as rec: accx forwards fill rec (see engine.RECORD)"""

    def __enter__(self):
        from accx import engine
        self.engine = engine
        self.old = engine.RECORD
        engine.RECORD = {}
        return engine.RECORD

    def __exit__(self, *exc):
        self.engine.RECORD = self.old
        return False
