"""CPU oracle for the ACC-UNet HANC/MLFC hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import it.  The product path (``acc-unet-unext_b200/accx``) never does
and fails loudly when its CUDA library is missing.

It restates, as pure functions over a flat ``state_dict`` (name -> tensor), the
arithmetic of the reference's ``ACC_UNet/ACC_UNet.py`` blocks.  Tensors are
NCHW-shaped like the reference's; everything runs on the CPU through basic
torch ops so ``torch.autograd`` supplies the gradient oracle as well.

Parity pin: the reference has no tests or golden vectors of its own for this
path, so the oracle is pinned against outputs of the
reference itself, imported in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures.

Reference lines each function follows (relative to /root/reference):
  batchnorm       torch.nn.BatchNorm2d as used at ACC_UNet/ACC_UNet.py:34,74,244,...
  se_layer        ACC_UNet/ACC_UNet.py:37-49
  hanc_pyramid    ACC_UNet/ACC_UNet.py:83-138
  hanc_layer      ACC_UNet/ACC_UNet.py:77-142
  hanc_block      ACC_UNet/ACC_UNet.py:267-286
  conv_bn_se      ACC_UNet/ACC_UNet.py:182-186   (Conv2d_batchnorm)
  respath         ACC_UNet/ACC_UNet.py:323-328
  mlfc            ACC_UNet/ACC_UNet.py:420-527  (W blend: ACC_UNet_w.py:497-522,
                                                  Lite: ACC_UNet_lite.py:424-427)
  acc_unet        ACC_UNet/ACC_UNet.py:601-659
  dice_bce_loss   Experiments/utils.py:21-74,109-171
  seg_metrics     Experiments/utils.py:478-494 (iou_on_batch), :148-157 (_show_dice)
  token_shift / shiftmlp / shifted_block   Experiments/nets/UNext.py:38-113, :117-147, :150-160
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]
LRELU_SLOPE = 0.01
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


class Ctx:
    """Carries the state dict, the train/eval switch and collects BN buffer updates.

    forced / record: the path's only discrete decisions are the sign under every LeakyReLU and the arg-max of
    every max-pool window.  `record` (a dict) collects this evaluation's decisions per site; `forced` replays
    decisions recorded elsewhere (site -> bool mask / window index).  The bf16 parity tests use it to compare
    gradients element-wise: a bf16 evaluation flips a fraction of these decisions at rounding-level near-ties, which
    moves whole gradient contributions by a factor 100 and says nothing about the kernels; with the decisions of
    the CUDA run replayed here, what is left is the smooth rounding error that north_star's rtol bounds."""

    def __init__(self, state: State, training: bool = True, forced: Optional[dict] = None, record: Optional[dict] = None):
        self.sd = state
        self.training = training
        self.updates: State = {}
        self.forced = forced
        self.record = record

    def p(self, name: str) -> torch.Tensor:
        return self.sd[name]


def lrelu(x, cx: Optional[Ctx] = None, site: Optional[str] = None):
    m = x > 0
    if cx is not None and site is not None:
        if cx.forced is not None and site in cx.forced:
            m = cx.forced[site].to(x.device)
        if cx.record is not None:
            cx.record[site] = m
    return torch.where(m, x, x * LRELU_SLOPE)


def batchnorm(cx: Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    """BatchNorm2d: batch statistics (biased var) in training, running stats in eval.
    Running buffers: momentum 0.1, unbiased variance; recorded in cx.updates."""
    g, b = cx.p(name + ".weight"), cx.p(name + ".bias")
    rm, rv = cx.p(name + ".running_mean"), cx.p(name + ".running_var")
    if cx.training:
        n = x.numel() // x.shape[1]
        mu = x.mean(dim=(0, 2, 3))
        var = ((x - mu[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        with torch.no_grad():
            cx.updates[name + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mu
            cx.updates[name + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
            cx.updates[name + ".num_batches_tracked"] = cx.p(name + ".num_batches_tracked") + 1
    else:
        mu, var = rm, rv
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mu[None, :, None, None]) * (inv * g)[None, :, None, None] + b[None, :, None, None]


def pointwise(cx: Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    """1x1 convolution = per-pixel matrix product with W[out,in] plus bias."""
    w = cx.p(name + ".weight")
    y = torch.einsum("bchw,oc->bohw", x, w.reshape(w.shape[0], w.shape[1]))
    return y + cx.p(name + ".bias")[None, :, None, None]


def first_argmax(win: torch.Tensor) -> torch.Tensor:
    """index of the FIRST maximum along the last dim (row-major window order: ATen's max-pool tie rule)"""
    n = win.shape[-1]
    hit = win == win.amax(dim=-1, keepdim=True)
    pos = torch.arange(n, device=win.device).expand_as(win)
    return torch.where(hit, pos, torch.full_like(pos, n)).amin(dim=-1)


def block_reduce(x: torch.Tensor, s: int, how: str, cx: Optional[Ctx] = None, site: Optional[str] = None) -> torch.Tensor:
    """Non-overlapping s x s pooling computed directly on x."""
    B, C, H, W = x.shape
    v = x.reshape(B, C, H // s, s, W // s, s)
    if how == "avg":
        return v.mean(dim=(3, 5))
    if cx is None or site is None or (cx.forced is None and cx.record is None):
        return v.amax(dim=(3, 5))
    win = v.permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // s, W // s, s * s)      # row-major inside the window
    idx = first_argmax(win.detach())
    if cx.forced is not None and site in cx.forced:
        idx = cx.forced[site].to(x.device)
    if cx.record is not None:
        cx.record[site] = idx
    return win.gather(-1, idx.unsqueeze(-1)).squeeze(-1)


def replicate(x: torch.Tensor, s: int) -> torch.Tensor:
    """Nearest-neighbour upsampling by an integer factor."""
    return x.repeat_interleave(s, dim=2).repeat_interleave(s, dim=3)


def se_layer(cx: Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    """Squeeze-excite gate followed by BN and LeakyReLU (non-standard tail)."""
    m = x.mean(dim=(2, 3))
    h = lrelu(m @ cx.p(name + ".fc1.weight").t() + cx.p(name + ".fc1.bias"), cx, name + ".fc1")
    g = torch.sigmoid(h @ cx.p(name + ".fc2.weight").t() + cx.p(name + ".fc2.bias"))
    return lrelu(batchnorm(cx, name + ".bn", x * g[:, :, None, None]), cx, name + ".bn")


def hanc_pyramid(x: torch.Tensor, k: int, cx: Optional[Ctx] = None, name: str = "") -> torch.Tensor:
    """[x, avg2, avg4, .., max2, max4, ..] interleaved so channel index = c*(2k-1)+j."""
    if k == 1:
        return x
    maps = [x]
    for j in range(1, k):
        maps.append(replicate(block_reduce(x, 2 ** j, "avg"), 2 ** j))
    for j in range(1, k):
        maps.append(replicate(block_reduce(x, 2 ** j, "max", cx, f"{name}.max{j}"), 2 ** j))
    B, C, H, W = x.shape
    return torch.stack(maps, dim=2).reshape(B, C * (2 * k - 1), H, W)


def hanc_layer(cx: Ctx, name: str, x: torch.Tensor, k: int) -> torch.Tensor:
    return lrelu(batchnorm(cx, name + ".bn", pointwise(cx, name + ".cnv", hanc_pyramid(x, k, cx, name))), cx, name + ".bn")


def hanc_block(cx: Ctx, name: str, inp: torch.Tensor, k: int) -> torch.Tensor:
    x = lrelu(batchnorm(cx, name + ".norm1", pointwise(cx, name + ".conv1", inp)), cx, name + ".norm1")
    w2 = cx.p(name + ".conv2.weight")
    x = F.conv2d(x, w2, cx.p(name + ".conv2.bias"), padding=1, groups=w2.shape[0])
    x = lrelu(batchnorm(cx, name + ".norm2", x), cx, name + ".norm2")
    x = hanc_layer(cx, name + ".hnc", x, k)
    x = batchnorm(cx, name + ".norm", x + inp)
    x = lrelu(batchnorm(cx, name + ".norm3", pointwise(cx, name + ".conv3", x)), cx, name + ".norm3")
    return se_layer(cx, name + ".sqe", x)


def conv_bn_se(cx: Ctx, name: str, x: torch.Tensor) -> torch.Tensor:
    x = lrelu(batchnorm(cx, name + ".batchnorm", pointwise(cx, name + ".conv1", x)), cx, name + ".batchnorm")
    return se_layer(cx, name + ".sqe", x)


def respath(cx: Ctx, name: str, x: torch.Tensor, n_lvl: int) -> torch.Tensor:
    for i in range(n_lvl):
        y = F.conv2d(x, cx.p(f"{name}.convs.{i}.weight"), cx.p(f"{name}.convs.{i}.bias"), padding=1)
        x = x + se_layer(cx, f"{name}.sqes.{i}", lrelu(batchnorm(cx, f"{name}.bns.{i}", y), cx, f"{name}.bns.{i}"))
    # the module registered under the name 'sqe' is a BatchNorm2d
    return batchnorm(cx, name + ".sqe", lrelu(batchnorm(cx, name + ".bn", x), cx, name + ".bn"))


def _to_level(x: torch.Tensor, src: int, dst: int) -> torch.Tensor:
    """Bring pyramid level `src` to the resolution of level `dst` with chained
    avg-pool(2) / nearest(2), exactly as many times as the reference chains them."""
    for _ in range(dst - src):
        x = block_reduce(x, 2, "avg")
    for _ in range(src - dst):
        x = replicate(x, 2)
    return x


def mlfc(cx: Ctx, name: str, xs: List[torch.Tensor], lenn: int = 1, variant: str = "base"):
    if variant == "lite":
        return tuple(se_layer(cx, f"{name}.sqe{l + 1}", xs[l]) for l in range(4))
    xc = list(xs)
    for i in range(lenn):
        xc = []
        for l in range(4):
            gathered = torch.cat([_to_level(xs[s], s, l) for s in range(4)], dim=1)
            t = conv_bn_se(cx, f"{name}.cnv_blks{l + 1}.{i}", gathered)
            xc.append(lrelu(batchnorm(cx, f"{name}.bns{l + 1}.{i}", t), cx, f"{name}.bns{l + 1}.{i}"))
        for l in range(4):
            B, C, H, W = xs[l].shape
            merged = torch.stack([xc[l], xs[l]], dim=2).reshape(B, 2 * C, H, W)
            t = conv_bn_se(cx, f"{name}.cnv_mrg{l + 1}.{i}", merged)
            if variant == "w":
                wmix = cx.p(name + ".W")
                t = t * wmix + xs[l] * (1 - wmix)
            else:
                t = t + xs[l]
            xc[l] = lrelu(batchnorm(cx, f"{name}.bns_mrg{l + 1}.{i}", t), cx, f"{name}.bns_mrg{l + 1}.{i}")
    return tuple(se_layer(cx, f"{name}.sqe{l + 1}", xc[l]) for l in range(4))


# (name, k) of the 18 HANC blocks in execution order, ACC_UNet.py:554-592
ENC = [("cnv11", 3), ("cnv12", 3), ("cnv21", 3), ("cnv22", 3), ("cnv31", 3), ("cnv32", 3),
       ("cnv41", 2), ("cnv42", 2), ("cnv51", 1), ("cnv52", 1)]
DEC = [("up6", "cnv61", "cnv62", 2), ("up7", "cnv71", "cnv72", 3),
       ("up8", "cnv81", "cnv82", 3), ("up9", "cnv91", "cnv92", 3)]


def acc_unet(cx: Ctx, x: torch.Tensor, variant: str = "base", logits: bool = False) -> torch.Tensor:
    """Whole ACC-UNet forward (variant in base|w|lite).  With logits=True the
    final sigmoid is skipped (reference: last_activation=None, ACC_UNet.py:653-657)."""
    skips = []
    for idx, (nm, k) in enumerate(ENC):
        x = hanc_block(cx, nm, x, k)
        if idx % 2 == 1 and idx < 9:
            skips.append(x)
            x = block_reduce(x, 2, "max", cx, f"pool{idx // 2 + 1}") if (cx.forced is not None or cx.record is not None) \
                else F.max_pool2d(x, 2)
    for l, n_lvl in enumerate((4, 3, 2, 1)):
        skips[l] = respath(cx, f"rspth{l + 1}", skips[l], n_lvl)
    for m in ("mlfc1", "mlfc2", "mlfc3"):
        skips = list(mlfc(cx, m, skips, 1, variant))
    for (up, a, b, k), skip in zip(DEC, reversed(skips)):
        x = F.conv_transpose2d(x, cx.p(up + ".weight"), cx.p(up + ".bias"), stride=2)
        x = hanc_block(cx, a, torch.cat([x, skip], dim=1), k)
        x = hanc_block(cx, b, x, k)
    y = pointwise(cx, "out", x)
    if not logits and cx.p("out.weight").shape[0] == 1:
        y = torch.sigmoid(y)
    return y


def dice_bce_loss(logit: torch.Tensor, truth: torch.Tensor, dice_weight=0.5, bce_weight=0.5) -> torch.Tensor:
    """WeightedDiceBCE(dice_weight, BCE_weight) on logits with class weights [0.5, 0.5]."""
    B = logit.shape[0]
    lg, tr = logit.reshape(B, -1).float(), truth.reshape(B, -1).float()
    # dice part: both weights 0.5 -> every pixel scaled by 0.5
    p = torch.sigmoid(lg) * 0.5
    t = tr * 0.5
    inter = (p * t).sum(-1)
    union = (p * p).sum(-1) + (t * t).sum(-1)
    dice = (1 - (2 * inter + 1e-5) / (union + 1e-5)).mean()
    # BCE-with-logits, normalised separately over positives and negatives
    l = torch.clamp(lg, min=0) - lg * tr + torch.log1p(torch.exp(-lg.abs()))
    pos = (tr > 0.5).float()
    neg = 1 - pos
    bce = (0.5 * pos * l / pos.sum().clamp(min=1.0) + 0.5 * neg * l / neg.sum().clamp(min=1.0)).sum()
    return dice_weight * dice + bce_weight * bce


# ---- UNeXt shifted tokenized-MLP block (Experiments/nets/UNext.py:38-160) -------------------------------------
def token_shift(x: torch.Tensor, H: int, W: int, axis: int, shift_size: int = 5) -> torch.Tensor:
    """pad -> chunk(shift_size) over channels -> roll chunk g by g - pad along H (axis 2) or W (axis 3) -> narrow
    (UNext.py:78-84, :97-103) on tokens [B, N, C]; equals: chunk g read at offset -(g - pad), zero outside the map."""
    B, N, C = x.shape
    pad = shift_size // 2
    xn = x.transpose(1, 2).reshape(B, C, H, W)
    xn = F.pad(xn, (pad, pad, pad, pad), "constant", 0)
    xs = torch.chunk(xn, shift_size, 1)
    xs = [torch.roll(c, s, axis) for c, s in zip(xs, range(-pad, pad + 1))]
    xc = torch.cat(xs, 1)[:, :, pad:pad + H, pad:pad + W]
    return xc.reshape(B, C, H * W).transpose(1, 2)


def shiftmlp(cx: Ctx, name: str, x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """shift_H -> fc1 -> DWConv (3x3 depthwise + bias) -> GELU -> shift_W -> fc2   (UNext.py:72-113, drop = 0)"""
    B, N, C = x.shape
    y = F.linear(token_shift(x, H, W, 2), cx.p(name + ".fc1.weight"), cx.p(name + ".fc1.bias"))
    w = cx.p(name + ".dwconv.dwconv.weight")
    y = F.conv2d(y.transpose(1, 2).reshape(B, -1, H, W), w, cx.p(name + ".dwconv.dwconv.bias"), padding=1, groups=w.shape[0])
    y = F.gelu(y.flatten(2).transpose(1, 2))
    return F.linear(token_shift(y, H, W, 3), cx.p(name + ".fc2.weight"), cx.p(name + ".fc2.bias"))


def shifted_block(cx: Ctx, name: str, x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """x + shiftmlp(LayerNorm(x))   (UNext.py:144-147, drop_path = 0)"""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), cx.p(name + ".norm2.weight"), cx.p(name + ".norm2.bias"), 1e-5)
    return x + shiftmlp(cx, name + ".mlp", y, H, W)


def seg_metrics(logit: torch.Tensor, truth: torch.Tensor) -> Tuple[float, float]:
    """(iou_on_batch, WeightedDiceBCE._show_dice) of the reference's training loop
    (Experiments/utils.py:478-494, :148-157; called every step at Train_one_epoch.py:134-135)."""
    B = logit.shape[0]
    pred = (torch.sigmoid(logit.reshape(B, -1).float()) >= 0.5)
    mask = truth.reshape(B, -1) > 0
    ious, dices = [], []
    s1 = 1.0 / (1.0 + math.exp(-1.0))
    for b in range(B):
        tp = float((pred[b] & mask[b]).sum())
        n_pred, n_mask, n = float(pred[b].sum()), float(mask[b].sum()), float(pred.shape[1])
        union = n_pred + n_mask - tp
        ious.append(tp / union if union > 0 else 0.0)                  # sklearn jaccard_score(zero_division -> 0)
        # _show_dice feeds the 0/1 prediction to WeightedDiceLoss, which applies sigmoid once more (utils.py:121)
        inter = 0.25 * (s1 * tp + 0.5 * (n_mask - tp))
        pp, tt = 0.25 * (s1 * s1 * n_pred + 0.25 * (n - n_pred)), 0.25 * n_mask
        dices.append((2 * inter + 1e-5) / (pp + tt + 1e-5))
    return sum(ious) / B, sum(dices) / B


# ---------------------------------------------------------------------------------
# Parameter construction with the torch default initialisers, in the same RNG order
# the reference's constructors draw them, so torch.manual_seed(s) reproduces its
# weights.  Used when /root/reference is not there (GPU box, cpu_baseline timing).
# ---------------------------------------------------------------------------------

def _conv(sd: State, name: str, cout: int, cin_per_group: int, kh: int, kw: int):
    w = torch.empty(cout, cin_per_group, kh, kw)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1 / math.sqrt(cin_per_group * kh * kw)
    sd[name + ".weight"] = w
    sd[name + ".bias"] = torch.empty(cout).uniform_(-bound, bound)


def _linear(sd: State, name: str, cin: int, cout: int):
    w = torch.empty(cout, cin)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1 / math.sqrt(cin)
    sd[name + ".weight"] = w
    sd[name + ".bias"] = torch.empty(cout).uniform_(-bound, bound)


def _bn(sd: State, name: str, c: int):
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.zeros(c)
    sd[name + ".running_mean"] = torch.zeros(c)
    sd[name + ".running_var"] = torch.ones(c)
    sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def init_se(sd: State, name: str, c: int):
    _linear(sd, name + ".fc1", c, c // 8)
    _linear(sd, name + ".fc2", c // 8, c)
    _bn(sd, name + ".bn", c)


def init_hanc_block(sd: State, name: str, c: int, cout: int, k: int, f: int):
    e = c * f
    _conv(sd, name + ".conv1", e, c, 1, 1)
    _bn(sd, name + ".norm1", e)
    _conv(sd, name + ".conv2", e, 1, 3, 3)
    _bn(sd, name + ".norm2", e)
    _conv(sd, name + ".hnc.cnv", c, (2 * k - 1) * e, 1, 1)
    _bn(sd, name + ".hnc.bn", c)
    _bn(sd, name + ".norm", c)
    _conv(sd, name + ".conv3", cout, c, 1, 1)
    _bn(sd, name + ".norm3", cout)
    init_se(sd, name + ".sqe", cout)


def init_respath(sd: State, name: str, c: int, n_lvl: int):
    _bn(sd, name + ".bn", c)
    _bn(sd, name + ".sqe", c)
    for i in range(n_lvl):
        _conv(sd, f"{name}.convs.{i}", c, c, 3, 3)
        _bn(sd, f"{name}.bns.{i}", c)
        init_se(sd, f"{name}.sqes.{i}", c)


def init_mlfc(sd: State, name: str, filters, lenn: int = 1, variant: str = "base"):
    tot = sum(filters)
    if variant == "w":
        sd[name + ".W"] = torch.zeros(1)
    for i in range(lenn):
        for l, c in enumerate(filters):
            for kind, cin in (("cnv_blks", tot), ("cnv_mrg", 2 * c)):
                pre = f"{name}.{kind}{l + 1}.{i}"
                _conv(sd, pre + ".conv1", c, cin, 1, 1)
                _bn(sd, pre + ".batchnorm", c)
                init_se(sd, pre + ".sqe", c)
            _bn(sd, f"{name}.bns{l + 1}.{i}", c)
            _bn(sd, f"{name}.bns_mrg{l + 1}.{i}", c)
    for l, c in enumerate(filters):
        init_se(sd, f"{name}.sqe{l + 1}", c)


def init_acc_unet(n_channels=3, n_classes=1, n_filts=32, variant="base") -> State:
    sd: State = {}
    f = n_filts
    chans = [(n_channels, f), (f, f), (f, 2 * f), (2 * f, 2 * f), (2 * f, 4 * f), (4 * f, 4 * f),
             (4 * f, 8 * f), (8 * f, 8 * f), (8 * f, 16 * f), (16 * f, 16 * f)]
    for (nm, k), (ci, co) in zip(ENC, chans):
        init_hanc_block(sd, nm, ci, co, k, 3)
    for l, n_lvl in enumerate((4, 3, 2, 1)):
        init_respath(sd, f"rspth{l + 1}", f * 2 ** l, n_lvl)
    for m in ("mlfc1", "mlfc2", "mlfc3"):
        init_mlfc(sd, m, (f, 2 * f, 4 * f, 8 * f), 1, variant)
    dec = [(16 * f, 8 * f), (8 * f, 4 * f), (4 * f, 2 * f), (2 * f, f)]
    for (up, a, b, k), (ci, co) in zip(DEC, dec):
        w = torch.empty(ci, co, 2, 2)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        bound = 1 / math.sqrt(co * 4)
        sd[up + ".weight"] = w
        sd[up + ".bias"] = torch.empty(co).uniform_(-bound, bound)
        init_hanc_block(sd, a, 2 * co, co, k, 3)
        init_hanc_block(sd, b, co, co, k, 34 if b == "cnv72" else 3)
    _conv(sd, "out", n_classes if n_classes == 1 else n_classes + 1, f, 1, 1)
    return sd


def trainable(sd: State) -> List[str]:
    return [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]


def train_step(sd: State, x: torch.Tensor, mask: torch.Tensor, opt: Optional[torch.optim.Optimizer],
               variant: str = "base") -> Tuple[float, Optional[torch.optim.Optimizer]]:
    """One full training step on the CPU: forward, Dice+BCE on logits, backward, Adam(lr 1e-3).
    (Experiments/Train_one_epoch.py:107-129, train_model.py:647,719.)"""
    names = trainable(sd)
    for n in names:
        sd[n].requires_grad_(True)
    if opt is None:
        opt = torch.optim.Adam([sd[n] for n in names], lr=1e-3)
    cx = Ctx(sd, training=True)
    loss = dice_bce_loss(acc_unet(cx, x, variant, logits=True), mask)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    with torch.no_grad():
        for k, v in cx.updates.items():
            sd[k] = v
    return float(loss), opt
