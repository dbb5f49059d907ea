"""Training step around the accx model: Dice+BCE loss on logits, Adam, optional data-parallel
gradient averaging, optional whole-step CUDA-graph capture.

Mirrors the reference harness (Experiments/Train_one_epoch.py:107-129, train_model.py:647,719):
    preds = model(images); loss = WeightedDiceBCE(0.5, 0.5)(preds, masks); zero_grad; backward; Adam(lr=1e-3).step

Memory layout of the optimiser state (`FlatState`): ALL parameters live back to back in one flat fp32 buffer
(every `p.data` is a view into it), and so do the gradients, `exp_avg` and `exp_avg_sq`.  Per step:
  * one memset zeroes the flat gradient buffer; the accx backward kernels accumulate every parameter gradient
    straight into its slice (`engine.GRAD_ARENA`), so `p.grad` IS a view of the flat buffer;
  * data parallel (new functionality -- the reference has no distributed code): one process per
    GPU, rank-local BatchNorm statistics, ONE in-place NCCL all-reduce (average) of the flat gradient buffer --
    no gather / scatter copies; parameters that receive no gradient on any rank (ACC_UNet_Lite's unused MLFC
    convs) keep a zero gradient, which leaves them and their Adam state untouched;
  * one `accx_adam_step` launch updates all parameters (4 reads + 3 writes per element).
"""
from __future__ import annotations

from typing import List, Optional

import os

import torch
import torch.distributed as dist

from . import engine as E


def dice_bce_loss_torch(logit: torch.Tensor, truth: torch.Tensor, dice_weight: float = 0.5, bce_weight: float = 0.5):
    """WeightedDiceBCE(dice_weight, BCE_weight) with class weights [0.5, 0.5] as a chain of torch operators
    (Experiments/utils.py:21-74 BCE-with-logits normalised over positives / negatives, :109-138 weighted soft
    Dice on sigmoid(logit), :140-171 the sum).  Host-logic tests (gloo, CPU) and the parity test of the fused
    kernels use it; the train step does not."""
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


class _DiceBCEFn(torch.autograd.Function):
    """the same loss as two accx kernels: one reduction pass (per-image partial sums + the scalar), one
    element-wise pass for d loss / d logit"""

    @staticmethod
    def forward(ctx, logit, truth, dice_w, bce_w):
        E.require_cuda(logit)
        if logit.dim() == 4 and logit.shape[1] != 1:
            raise E._lib.AccxError("accx dice_bce_loss implements the one-class case (n_classes = 1)")
        B = logit.shape[0]
        lg = logit.detach().reshape(B, -1)
        if lg.dtype not in (torch.float32, torch.bfloat16):
            lg = lg.float()
        lg = lg.contiguous()
        tr = truth.detach().reshape(B, -1).float().contiguous()
        if tr.shape != lg.shape:
            raise ValueError(f"dice_bce_loss: logits {tuple(logit.shape)} vs masks {tuple(truth.shape)}")
        loss, sums = E.dice_bce_fwd(lg, tr, dice_w, bce_w)
        ctx.saved = (lg, tr, sums)
        ctx.w = (dice_w, bce_w)
        ctx.shape, ctx.dtype = logit.shape, logit.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        lg, tr, sums = ctx.saved
        ctx.saved = None
        gs = g.detach().float().reshape(1).contiguous()
        gdt = ctx.dtype if ctx.dtype in (torch.float32, torch.bfloat16) else torch.float32
        d = E.dice_bce_bwd(lg, tr, sums, ctx.w[0], ctx.w[1], gs, gdt)
        d = d.view(ctx.shape)
        return (d if d.dtype == ctx.dtype else d.to(ctx.dtype)), None, None, None


def dice_bce_loss(logit: torch.Tensor, truth: torch.Tensor, dice_weight: float = 0.5, bce_weight: float = 0.5):
    """WeightedDiceBCE(dice_weight, BCE_weight) on one-class logits [B, 1, H, W] (CUDA, accx kernels; see
    dice_bce_loss_torch for the formula and the reference lines)."""
    return _DiceBCEFn.apply(logit, truth, float(dice_weight), float(bce_weight))


def seg_metrics(logit: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """Per-batch training metrics WITHOUT a device-to-host sync: returns a device tensor [iou, dice] where
    iou = iou_on_batch(masks, preds) (Experiments/utils.py:478-494: sigmoid, 0.5 threshold, sklearn's binary Jaccard
    per image, batch mean) and dice = WeightedDiceBCE._show_dice(preds, masks) (utils.py:148-157).  The reference
    computes both on the host every step (Train_one_epoch.py:134-135); read the tensor only when you log."""
    E.require_cuda(logit)
    B = logit.shape[0]
    lg = logit.detach().reshape(B, -1)
    if lg.dtype not in (torch.float32, torch.bfloat16):
        lg = lg.float()
    tr = truth.detach().reshape(B, -1).float().contiguous()
    if tr.shape != lg.shape:
        raise ValueError(f"seg_metrics: logits {tuple(logit.shape)} vs masks {tuple(truth.shape)}")
    return E.seg_metrics(lg.contiguous(), tr)


class FlatState:
    """Parameters, gradients and Adam moments of a model as four flat fp32 buffers (see the module docstring).
    While it is installed as `engine.GRAD_ARENA` (TrainStep does that around forward + backward) the accx backward
    kernels accumulate parameter gradients straight into the flat gradient buffer."""

    ALIGN = 64          # elements: every slice starts 256-byte aligned (vector loads, TMA-able weight views)

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatState: no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatState expects fp32 parameters on one device")
        self.offsets, off = {}, 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.n = off
        self.param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.step_state = torch.zeros(4, dtype=torch.float32, device=dev)      # [0] = step count
        with torch.no_grad():
            for p in self.params:
                v = self._view(self.param, p)
                v.copy_(p.data)
                p.data = v                                                     # the parameter now lives in the flat buffer
        self.taken = set()

    def _view(self, flat, p):
        o = self.offsets[id(p)]
        return flat[o:o + p.numel()].view(p.shape)

    # ---- engine.GRAD_ARENA protocol ------------------------------------------------------------------
    def serves(self, p) -> bool:
        return id(p) in self.offsets and id(p) not in self.taken

    def take(self, p):
        """fresh view of p's (zeroed) gradient slice; None if p is not ours or was already handed out this step
        (a second use of the parameter then accumulates through autograd into the same slice)"""
        if id(p) not in self.offsets or id(p) in self.taken:
            return None
        self.taken.add(id(p))
        return self._view(self.grad, p)

    def owns(self, p, g) -> bool:
        """is `g` the slice of the flat gradient buffer that take(p) handed out?"""
        o = self.offsets.get(id(p))
        return o is not None and id(p) in self.taken and g.data_ptr() == self.grad.data_ptr() + 4 * o

    def begin_step(self):
        """zero the flat gradients (one memset) and drop the p.grad views of the previous step"""
        for p in self.params:
            p.grad = None
        self.grad.zero_()
        self.taken.clear()

    def collect(self):
        """after backward (and after the side stream has been joined): p.grad of every parameter whose gradient was
        accumulated in place is pointed at its slice -- those gradients never pass through autograd (see
        engine.param_grads), so nothing is copied.  Gradients that autograd produced outside the flat buffer
        (parameters of torch-native layers mixed into the model, a second use of a parameter) are added / copied
        into the slice."""
        for p in self.params:
            g = p.grad
            mine = id(p) in self.taken
            if g is None and not mine:
                continue                              # no gradient this step (ACC_UNet_Lite's idle MLFC convs)
            v = self._view(self.grad, p)
            if g is not None and g.data_ptr() != v.data_ptr():
                if mine:
                    v.add_(g)
                else:
                    v.copy_(g)
            p.grad = v

    def adam(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        """lr < 0: the kernel reads the learning rate from step_state[1] (TrainStep.set_lr)"""
        E.adam_step(self.param, self.grad, self.exp_avg, self.exp_avg_sq, self.step_state, lr, betas[0], betas[1], eps,
                    weight_decay)


def plan_ranges(flat: "FlatState", phases):
    """[(trigger, modules)] -> [(trigger, lo, hi, params)]: the slice [lo, hi) of the flat buffers that holds exactly the
    parameters of `modules` (they must be contiguous in parameter order, which is how FlatState lays them out)"""
    out = []
    for trigger, modules in phases:
        ps = [p for m in modules for p in m.parameters() if id(p) in flat.offsets]
        if not ps:
            continue
        ids = {id(p) for p in ps}
        lo = min(flat.offsets[i] for i in ids)
        hi = max(flat.offsets[id(p)] + (p.numel() + flat.ALIGN - 1) // flat.ALIGN * flat.ALIGN for p in ps)
        inside = {id(p) for p in flat.params if lo <= flat.offsets[id(p)] < hi}
        if inside != ids:
            raise ValueError("allreduce_phases: the parameters of a phase are not contiguous in parameter order")
        out.append((trigger, lo, hi, ps))
    return out


class GradAverager:
    """Average gradients over the ranks of `group`: one in-place all-reduce of a flat buffer."""

    def __init__(self, params: List[torch.nn.Parameter], group=None, flat: Optional[FlatState] = None):
        self.params = list(params)
        self.group = group
        self.flat = flat
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending = []          # (lo, hi, work) of reductions started before backward ended

    def _allreduce_avg(self, buf):
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.div_(self.world)

    # ---- parts of the flat buffer reduced early, under the rest of backward (TrainStep drives this) ---------------
    def reduce_range_async(self, lo: int, hi: int):
        """start the all-reduce of flat.grad[lo:hi] on the CURRENT stream's work queue without making it wait"""
        self.pending.append((lo, hi, dist.all_reduce(self.flat.grad[lo:hi], op=dist.ReduceOp.AVG, group=self.group,
                                                     async_op=True)))

    def __call__(self):
        if self.world == 1:
            return
        if self.flat is not None:                  # gradients already live in one buffer: reduce it in place
            done = sorted((lo, hi) for lo, hi, _ in self.pending)
            pos = 0
            for lo, hi in done + [(self.flat.n, self.flat.n)]:       # what no early reduction covered
                if lo > pos:
                    self._allreduce_avg(self.flat.grad[pos:lo])
                pos = max(pos, hi)
            for _, _, work in self.pending:                            # the current stream waits for the early ones
                work.wait()
            self.pending.clear()
            return
        # loose gradients: ranks must agree on the bucket layout (ACC_UNet_Lite leaves the same parameters unused
        # everywhere); one gather kernel, one all-reduce, multi-tensor scatter
        grads = [p.grad for p in self.params if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        self._allreduce_avg(flat)
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)


class TrainStep:
    """One optimisation step as a callable: step(images, masks) -> loss tensor (device).

    graph=True captures forward + loss + backward + (all-reduce) + Adam in a CUDA graph after
    `graph_warmup` eager steps; inputs are then copied into static buffers each step."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, graph: bool = False, graph_warmup: int = 3,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, metrics: bool = False,
                 distributed: bool = True):
        """distributed=False: a rank-local step inside an initialised process group (no start-up broadcast, no gradient
        exchange) -- e.g. an instrumented step that only rank 0 runs; every collective needs ALL ranks"""
        self.model = model
        self.params = [p for p in model.parameters() if p.requires_grad]
        E.require_cuda(self.params[0])
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.flat = FlatState(self.params)
        self.flat.step_state[1] = lr               # the Adam kernel reads the learning rate from the device
        self.avg = GradAverager(self.params, flat=self.flat)
        if not distributed:
            self.avg.world = 1
        # metrics=True: IoU / Dice of every step's predictions are computed on the device inside the step (and the
        # step's CUDA graph) and left in `last_metrics` -- no per-step device-to-host sync (Train_one_epoch.py:134-135)
        self.metrics = metrics
        self.last_metrics: Optional[torch.Tensor] = None
        if self.avg.world > 1:
            # replicas must start identical whatever each rank's seed or checkpoint: rank 0's parameters and module
            # buffers (BatchNorm running statistics) go to everyone; the statistics stay rank-local afterwards
            dist.broadcast(self.flat.param, src=0)
            for b in model.buffers():
                dist.broadcast(b, src=0)
        # backward-overlapped gradient exchange: contiguous ranges of the flat buffer whose gradients are final before
        # backward ends (model.allreduce_phases), reduced on a communication stream while the rest of backward runs.
        # Opt-in (ACCX_ALLREDUCE_OVERLAP=1): measured on 2 and 8 B200s it changes nothing (33.75 vs 33.77 and 34.12 vs
        # 34.15 ms per step) -- the exchange is 0.3 ms and NCCL's blocks compete with the backward kernels for SMs
        self.phases = []
        self._phase = 0
        self._comm = None
        if (self.avg.world > 1 and hasattr(model, "allreduce_phases") and dist.get_backend() == "nccl"
                and os.environ.get("ACCX_ALLREDUCE_OVERLAP", "0") == "1"):
            self.phases = plan_ranges(self.flat, model.allreduce_phases())
            self._comm = torch.cuda.Stream(device=self.params[0].device)
        self.use_graph = graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_warmup = max(1, graph_warmup)
        self.calls = 0
        self.static_x = self.static_m = self.static_loss = None

    def _eager(self, x, m):
        flat = self.flat
        flat.begin_step()
        arena, E.GRAD_ARENA = E.GRAD_ARENA, flat
        mode = E.SIDE_MODE
        try:
            logits = self.model(x)
            loss = dice_bce_loss(logits, m)
            if self.metrics:
                self.last_metrics = seg_metrics(logits, m)
            E.SIDE_MODE = 2 if mode else 0             # weight gradients overlap the whole backward ...
            hook, E.BWD_START_HOOK = E.BWD_START_HOOK, (self._on_backward_start if self.phases else E.BWD_START_HOOK)
            self._phase = 0
            try:
                loss.backward()
            finally:
                E.BWD_START_HOOK = hook
                E.join_side()                          # ... and are joined once, before the optimiser
                E.SIDE_MODE = mode
        finally:
            E.GRAD_ARENA = arena
        flat.collect()
        self.avg()
        idle = None
        if self.weight_decay != 0.0:
            # torch.optim.Adam skips parameters whose grad is None (ACC_UNet_Lite's idle MLFC convs); the flat kernel
            # would still decay them, so their values are put back after the update (their moments stay zero)
            idle = [p for p in self.params if p.grad is None]
            keep = [p.detach().clone() for p in idle]
        flat.adam(-1.0, self.betas, self.eps, self.weight_decay)
        if idle:
            torch._foreach_copy_([p.data for p in idle], keep)
            for p in idle:
                flat._view(flat.exp_avg, p).zero_()
                flat._view(flat.exp_avg_sq, p).zero_()
        return loss.detach()

    def _on_backward_start(self, obj):
        """engine.BWD_START_HOOK: `obj`'s backward is about to be queued, everything upstream of it in backward order is
        queued already -- if that completes the next phase, its slice of the gradient buffer goes out now"""
        if self._phase >= len(self.phases):
            return
        trigger, lo, hi, params = self.phases[self._phase]
        if not ((trigger == "group" and isinstance(obj, list)) or obj is trigger):
            return
        self._phase += 1
        # only gradients that live in the flat buffer already may leave early (a torch-native layer's gradient would
        # be copied in by collect(), after backward): otherwise this range waits for the final reduction
        if any(p.grad is not None and not self.flat.owns(p, p.grad) for p in params):
            return
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self._comm.wait_event(ev)                      # input-gradient chain (BatchNorm / SE / bias gradients) ...
        side = E._SIDE.get(torch.cuda.current_device())
        if side is not None:
            ev2 = torch.cuda.Event()
            ev2.record(side)
            self._comm.wait_event(ev2)                 # ... and the weight gradients queued on the side stream so far
        with torch.cuda.stream(self._comm):
            self.avg.reduce_range_async(lo, hi)

    def set_lr(self, lr: float):
        """new learning rate from the next step on (lr_scheduler.step() of the harness, Train_one_epoch.py:187-188);
        a device-side scalar, so the captured step graph stays valid"""
        self.lr = float(lr)
        self.flat.step_state[1] = self.lr

    # ---- optimiser checkpointing (Experiments/train_model.py:125-145 saves optimizer.state_dict(), :677-689 and
    # :815-818 restore it on resume) in torch.optim.Adam's own format, so reference checkpoints load and vice versa
    def state_dict(self) -> dict:
        step = self.flat.step_state[0].detach().cpu().clone()
        state = {}
        for i, p in enumerate(self.params):
            state[i] = {"step": step.clone(), "exp_avg": self.flat._view(self.flat.exp_avg, p).detach().clone(),
                        "exp_avg_sq": self.flat._view(self.flat.exp_avg_sq, p).detach().clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict):
        """in place (the captured CUDA graph keeps pointing at the same buffers).  Parameters without an entry in
        sd["state"] (never stepped: torch only creates state for parameters that had a gradient) get zero moments."""
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.params):
            raise ValueError(f"optimizer state has {len(ids)} parameters, the model {len(self.params)}")
        g0 = groups[0]
        hyper = (tuple(self.betas), self.eps, self.weight_decay)
        self.betas = tuple(g0.get("betas", self.betas))
        self.eps, self.weight_decay = g0.get("eps", self.eps), g0.get("weight_decay", self.weight_decay)
        if hyper != (tuple(self.betas), self.eps, self.weight_decay):
            self.graph = None                       # baked into the captured launch: capture again on the next call
        self.set_lr(g0.get("lr", self.lr))
        steps = set()
        with torch.no_grad():
            self.flat.exp_avg.zero_()
            self.flat.exp_avg_sq.zero_()
            for p, i in zip(self.params, ids):
                st = sd["state"].get(i)
                if st is None:
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
                self.flat._view(self.flat.exp_avg, p).copy_(st["exp_avg"])
                self.flat._view(self.flat.exp_avg_sq, p).copy_(st["exp_avg_sq"])
                steps.add(float(st["step"]))
            if len(steps) > 1:
                raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the flat Adam kernel keeps one")
            self.flat.step_state[0] = steps.pop() if steps else 0.0

    def __call__(self, x: torch.Tensor, m: torch.Tensor) -> torch.Tensor:
        self.calls += 1
        if not self.use_graph:
            return self._eager(x, m)
        if self.graph is None:
            if self.calls <= self.graph_warmup:
                return self._eager(x, m)
            # the eager warm-up steps above have created every stream / lazily initialised handle the step uses;
            # capturing records the step without running it, the replay below is this call's ONE optimisation step
            self.static_x, self.static_m = x.clone(), m.clone()
            self.graph = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(priority=E.MAIN_PRIORITY)       # kernel nodes inherit the capture stream's priority
            with torch.cuda.graph(self.graph, stream=cap):
                self.static_loss = self._eager(self.static_x, self.static_m)
        self.static_x.copy_(x, non_blocking=True)
        self.static_m.copy_(m, non_blocking=True)
        self.graph.replay()
        return self.static_loss
