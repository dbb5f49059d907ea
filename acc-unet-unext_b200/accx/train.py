"""Training step around the accx model: Dice+BCE loss on logits, Adam, optional data-parallel
gradient averaging, optional whole-step CUDA-graph capture.

Mirrors the reference harness (Experiments/Train_one_epoch.py:107-129, train_model.py:647,719):
    preds = model(images); loss = WeightedDiceBCE(0.5, 0.5)(preds, masks); zero_grad; backward; Adam(lr=1e-3).step
Data parallel (new functionality -- the reference has no distributed code): one
process per GPU, rank-local BatchNorm statistics, gradients averaged with ONE NCCL all-reduce of a
flat fp32 buffer per step; parameters whose gradient is None on every rank (ACC_UNet_Lite's unused
MLFC convs) are left out of the bucket.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import engine as E


def dice_bce_loss(logit: torch.Tensor, truth: torch.Tensor, dice_weight: float = 0.5, bce_weight: float = 0.5):
    """WeightedDiceBCE(dice_weight, BCE_weight) with class weights [0.5, 0.5]
    (Experiments/utils.py:21-74 BCE-with-logits normalised over positives / negatives,
    :109-138 weighted soft Dice on sigmoid(logit), :140-171 the sum)."""
    B = logit.shape[0]
    lg = logit.reshape(B, -1).float()
    tr = truth.reshape(B, -1).float()
    p = torch.sigmoid(lg) * 0.5
    t = tr * 0.5
    inter = (p * t).sum(-1)
    union = (p * p).sum(-1) + (t * t).sum(-1)
    dice = (1 - (2 * inter + 1e-5) / (union + 1e-5)).mean()
    l = torch.nn.functional.binary_cross_entropy_with_logits(lg, tr, reduction="none")
    pos = (tr > 0.5).float()
    neg = 1.0 - pos
    bce = (0.5 * pos * l / pos.sum().clamp(min=1.0) + 0.5 * neg * l / neg.sum().clamp(min=1.0)).sum()
    return dice_weight * dice + bce_weight * bce


class GradAverager:
    """Average gradients over the ranks of `group` with one all-reduce of a flat buffer."""

    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = list(params)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def __call__(self):
        if self.world == 1:
            return
        # ranks must agree on the bucket layout: ACC_UNet_Lite leaves the same parameters unused everywhere
        grads = [p.grad for p in self.params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])            # one gather kernel for the ~900 tensors
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)                          # scatter back with multi-tensor kernels


class TrainStep:
    """One optimisation step as a callable: step(images, masks) -> loss tensor (device).

    graph=True captures forward + loss + backward + (all-reduce) + Adam in a CUDA graph after
    `graph_warmup` eager steps; inputs are then copied into static buffers each step."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, graph: bool = False, graph_warmup: int = 2):
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=lr, capturable=graph, fused=True)     # one multi-tensor kernel chain
        self.avg = GradAverager(self.params)
        self.use_graph = graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_warmup = graph_warmup
        self.calls = 0
        self.static_x = self.static_m = self.static_loss = None

    def _eager(self, x, m):
        logits = self.model(x)
        loss = dice_bce_loss(logits, m)
        self.opt.zero_grad(set_to_none=True)
        mode, E.SIDE_MODE = E.SIDE_MODE, (2 if E.SIDE_MODE else 0)     # weight gradients overlap the whole backward ...
        try:
            loss.backward()
        finally:
            E.join_side()                                              # ... and are joined once, before the optimiser
            E.SIDE_MODE = mode
        self.avg()
        self.opt.step()
        return loss.detach()

    def __call__(self, x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
        self.calls += 1
        if not self.use_graph:
            return self._eager(x, m)
        if self.graph is None:
            if self.calls <= self.graph_warmup:
                return self._eager(x, m)
            self.static_x, self.static_m = x.clone(), m.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):            # one more eager step on the side stream (allocator warm-up)
                self._eager(self.static_x, self.static_m)
            torch.cuda.current_stream().wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._eager(self.static_x, self.static_m)
        self.static_x.copy_(x, non_blocking=True)
        self.static_m.copy_(m, non_blocking=True)
        self.graph.replay()
        return self.static_loss
