"""Host-side launch helpers: torch tensors in, accx C-ABI calls out.

torch is plumbing here (device memory, the current stream, autograd bookkeeping); every
arithmetic pass over an activation is one accx kernel launched on torch's current stream.
All activations are NHWC ([B, H, W, C] contiguous) in fp32 or bf16; parameters are fp32.
"""
from __future__ import annotations

import ctypes
from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import BF16, F32, Operand

import os

TC = os.environ.get("ACCX_TC", "1") != "0"   # tcgen05 tensor-core contraction for bf16 activations
TC_F32 = os.environ.get("ACCX_TC_F32", "1") != "0"   # ... and for fp32 activations (3 x TF32, fp32 output)
TC_F32_IN_DET = False   # tests: keep the 3 x TF32 contraction in deterministic mode too (default there: exact fp32 FMA)
LAUNCHES = 0          # number of accx kernels launched by this process (bench.py reports it)
LAUNCHES_EXTRA = [0]  # kernels launched inside an accx call beyond the first (weight re-pack)
PROFILE = None        # list -> every launch is bracketed by CUDA events on the launching stream and
                      # appended as (kernel, start, end, algorithmic_bytes, flops); see bench.py


RECORD = None         # dict -> every forward notes the operands that carry the path's discrete decisions (LeakyReLU signs,
                      # max-pool arg-maxes), keyed by the torch module that owns them: ("bn", id(BatchNorm2d)) -> Lazy,
                      # ("se", id(ChannelSELayer)) -> SECtx, ("hanc", id(HANCLayer)) -> Lazy input of the pyramid,
                      # "pool" -> [inputs of MaxPool2d(2)].  tests/ replay them in the CPU restatement (flip-free gradient parity).


PROFILE_LEAD = None   # (every, cycles): while profiling, a spin kernel of `cycles` SM clocks is queued before every
                      # `every`-th launch (outside the event pairs).  Eager launching is CPU-bound (~45 us per launch vs
                      # ~25 us per kernel); without the lead the GPU idles between an event and the kernel that follows,
                      # and that idle time would be booked to the kernel.


# Timing diagnostic ONLY (tests/ablate.py): entry points named here are not launched, so that the marginal cost of a
# kernel class inside the overlapped step graph can be read off the step time.  Results are then meaningless.
ABLATE = frozenset(filter(None, os.environ.get("ACCX_ABLATE", "").split(",")))


def _call(name, *args, cost=(0, 0), tag=""):
    global LAUNCHES
    if ABLATE and name in ABLATE:
        return
    LAUNCHES += 1
    if PROFILE is None:
        _lib.call(name, *args)
        return
    if PROFILE_LEAD is not None and len(PROFILE) % PROFILE_LEAD[0] == 0:
        torch.cuda._sleep(PROFILE_LEAD[1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args)
    e1.record()
    PROFILE.append((name, e0, e1, cost[0], cost[1], tag))


# ---- weight gradients on a side stream -------------------------------------------------------------
# In backward, every weight gradient depends only on tensors the main chain has already produced and
# nothing downstream needs it before the optimiser.  Launching them on a second stream lets the GPU overlap
# them with the (often latency-bound, partially filled) kernels of the input-gradient chain; under CUDA
# graph capture the fork/join events become parallel branches of the graph.
#   SIDE_MODE 0: everything on the caller's stream; 1: join at the end of each module's backward (default);
#   2: join only when join_side() is called (TrainStep: once, before the optimiser).  Tensors the side
#   stream reads are kept alive in _KEEP until the join so the caching allocator cannot recycle them.
SIDE_MODE = int(os.environ.get("ACCX_WGRAD_STREAM", "1"))
# CUDA stream priorities (lower number = scheduled first): the chain of input gradients is the critical path, the
# weight gradients only have to be done before the optimiser, so the side stream yields to the main stream / lanes
SIDE_PRIORITY = int(os.environ.get("ACCX_SIDE_PRIO", "0"))
MAIN_PRIORITY = int(os.environ.get("ACCX_MAIN_PRIO", "-3"))
BWD_DEPTH = [0]
BWD_START_HOOK = None  # callable(obj) run when the backward of an accx module (obj) or of a group of chains (a list) starts:
                       # every module whose backward ran before has all its parameter gradients queued (TrainStep overlaps
                       # the gradient all-reduce of finished parts of the model with the rest of backward)
_SIDE = {}
_SIDE_DIRTY = set()
_KEEP = []
_KEEP_BYTES = [0]
KEEP_LIMIT = int(float(os.environ.get("ACCX_SIDE_KEEP_GB", "12")) * (1 << 30))


class side_stream:
    """with side_stream(keep=(tensors...)): launches inside go to the side stream, ordered after everything
    already queued on the current stream."""

    def __init__(self, keep=()):
        self.keep = keep
        self.ctx = None

    def __enter__(self):
        # only inside a module backward (which joins at its end) or a TrainStep (which joins before the
        # optimiser): a direct engine call stays on the caller's stream
        if SIDE_MODE == 0 or (SIDE_MODE == 1 and BWD_DEPTH[0] == 0):
            return self
        dev = torch.cuda.current_device()
        s = _SIDE.get(dev)
        if s is None:
            s = _SIDE[dev] = torch.cuda.Stream(device=dev, priority=SIDE_PRIORITY)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        s.wait_event(ev)
        _SIDE_DIRTY.add(dev)
        # whatever the side stream reads must outlive the join: a temporary the caller drops right after this
        # call would otherwise go back to the allocator and be handed to the next kernel on the caller's stream
        for t in self.keep:
            if t is not None:
                _KEEP.append(t)
                _KEEP_BYTES[0] += t.numel() * t.element_size()
        self.ctx = torch.cuda.stream(s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
            # deferred join (mode 2): everything the side stream reads stays allocated until the join; past
            # KEEP_LIMIT bytes the caller's stream waits for the side stream here and the tensors are released
            # (bounds the peak memory of the 512^2 / large-batch configurations; never reached at 16 x 224^2)
            if _KEEP_BYTES[0] > KEEP_LIMIT:
                join_side()
        return False


def join_side():
    """make the current stream wait for all side-stream work issued so far"""
    dev = torch.cuda.current_device()
    if dev in _SIDE_DIRTY:
        ev = torch.cuda.Event()
        ev.record(_SIDE[dev])
        torch.cuda.current_stream().wait_event(ev)
        _SIDE_DIRTY.discard(dev)
    _KEEP.clear()
    _KEEP_BYTES[0] = 0


def module_backward_end():
    # mode 2 defers the join to the train step, which owns every gradient through its flat buffer (GRAD_ARENA);
    # without an arena the gradients go to autograd right after this call, so the side stream is joined here
    if SIDE_MODE == 1 or (SIDE_MODE == 2 and GRAD_ARENA is None):
        join_side()


def param_grads(params, grads: dict):
    """What a module backward returns to autograd for `params`.  A gradient that was accumulated in place in the
    train step's flat gradient buffer (GRAD_ARENA) is NOT handed to autograd: AccumulateGrad would steal or -- when
    anything else still references the tensor -- CLONE it on the caller's stream while the side stream may not
    have finished (or started) adding to it, and the stale copy would later overwrite the slice.  FlatState.collect()
    points p.grad at the slice after the join instead.  Any other gradient (a parameter the arena does not serve, or
    a second use of a served one) is returned; in deferred-join mode the side stream is joined first."""
    ga = GRAD_ARENA
    out, loose = [], False
    for p in params:
        g = grads.get(id(p))
        if g is not None and ga is not None and ga.owns(p, g):
            g = None
        elif g is not None:
            loose = True
            if g.dtype != p.dtype:
                g = g.to(p.dtype)
        out.append(g)
    if loose and SIDE_MODE == 2:
        join_side()
    return out


# ---- independent chains on parallel streams --------------------------------------------------------
# MLFC's four pyramid levels (and the model's four ResPaths) are independent chains of small, latency-bound
# kernels.  Inside ONE module call they are issued on separate streams between a fork and a join, so the GPU
# (and the captured CUDA graph) runs them concurrently.  Every tensor a lane allocates stays referenced until
# after the join, and a lane starts by waiting for the caller's stream, so the caching allocator never hands
# memory to a stream that could still race with a pending reader.
LANES = int(os.environ.get("ACCX_LANES", "1"))
_LANE_STREAMS = {}


class fork_lanes:
    depth = 0          # forks do not nest: an inner fork (a module used inside a lane) runs in its caller's lane

    def __init__(self, n):
        self.n = n
        self.used = set()
        self.enabled = LANES != 0 and n > 1

    def __enter__(self):
        if fork_lanes.depth > 0:
            self.enabled = False
        fork_lanes.depth += 1
        if self.enabled:
            self.main = torch.cuda.current_stream()
            dev = torch.cuda.current_device()
            pool = _LANE_STREAMS.setdefault(dev, [])
            while len(pool) < self.n:
                pool.append(torch.cuda.Stream(device=dev, priority=MAIN_PRIORITY))
            self.pool = pool
            self.ev = torch.cuda.Event()
            self.ev.record(self.main)
        return self

    def lane(self, i):
        """context manager: launches inside go to lane i (ordered after everything queued before the fork and
        after this lane's earlier work)"""
        if not self.enabled:
            return _NullCtx()
        s = self.pool[i]
        if i not in self.used:
            s.wait_event(self.ev)
            self.used.add(i)
        return torch.cuda.stream(s)

    def __exit__(self, *exc):
        fork_lanes.depth -= 1
        if self.enabled:
            for i in self.used:
                ev = torch.cuda.Event()
                ev.record(self.pool[i])
                self.main.wait_event(ev)
        return False


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


# ---- deterministic reduction mode ------------------------------------------------------------------
_DET = {}


def set_deterministic(on: bool, workspace_mb: int = 512, device=None):
    """Fixed-order reductions in every accx kernel (include/accx.h: accx_set_deterministic): two runs on the same
    inputs are bit-identical, whatever the stream configuration (side stream, lanes, CUDA graph).  Parity / debugging
    mode (slower); the fp32 parity bounds in tests/ are stated for it."""
    _lib.load()
    if not on:
        if _DET:
            torch.cuda.synchronize()
            _lib.call("accx_set_deterministic", 0, 0, 0, 0)
            _DET.clear()
        return
    if _DET:
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ws = torch.empty(workspace_mb << 20, dtype=torch.uint8, device=dev)
    ctr = torch.zeros(1 << 16, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    _lib.call("accx_set_deterministic", ws.data_ptr(), ws.numel(), ctr.data_ptr(), ctr.numel())
    _DET.update(ws=ws, ctr=ctr)


def deterministic() -> bool:
    return bool(_DET)


def nb(*ts) -> int:
    """bytes of the given tensors (algorithmic traffic accounting: each tensor once)"""
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"accx supports float32 / bfloat16 activations, got {t.dtype}")


def tdtype(code: int) -> torch.dtype:
    return torch.float32 if code == F32 else torch.bfloat16


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def f32(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """parameter -> contiguous fp32 device tensor (no copy in the normal case)"""
    if p is None:
        return None
    p = p.detach()
    if p.dtype != torch.float32:
        p = p.float()
    return p if p.is_contiguous() else p.contiguous()


def require_cuda(x: torch.Tensor):
    if not x.is_cuda:
        raise _lib.AccxError("accx modules run on CUDA tensors only (sm_100a kernels; there is no CPU path)")
    _lib.load()


class Arena:
    """Zero-initialised fp32 scratch carved from a few large allocations (one memset per chunk
    instead of one per statistics buffer)."""

    CHUNK = 1 << 16

    def __init__(self, device):
        self.device = device
        self.buf = None
        self.used = 0

    def take(self, n: int) -> torch.Tensor:
        n_al = (n + 63) // 64 * 64
        if self.buf is None or self.used + n_al > self.buf.numel():
            self.buf = torch.zeros(max(self.CHUNK, n_al), dtype=torch.float32, device=self.device)
            self.used = 0
        out = self.buf[self.used:self.used + n]
        self.used += n_al
        return out


class Lazy:
    """raw NHWC tensor + pending per-channel affine/activation (see include/accx.h)."""
    __slots__ = ("y", "scale", "shift", "act", "mean", "rstd", "bn", "train")

    def __init__(self, y, scale=None, shift=None, act=0, mean=None, rstd=None, bn=None, train=True):
        self.y, self.scale, self.shift, self.act = y, scale, shift, act
        self.mean, self.rstd, self.bn = mean, rstd, bn
        self.train = train        # False: the BatchNorm ran on its running statistics (a fixed affine in backward)

    @property
    def dims(self):
        return tuple(self.y.shape[:3])

    @property
    def C(self):
        return self.y.shape[3]


class WV(NamedTuple):
    """strided view into an fp32 weight tensor: element (n, k) = t.flat[off + n*ld + k*ks]"""
    t: torch.Tensor
    off: int
    ld: int
    ks: int


class Op(NamedTuple):
    src: Lazy
    K: int
    wv: WV
    coff: int = 0      # first channel used inside src.y
    dy: int = 0
    dx: int = 0


def _fill(o: Operand, op: Op, wt: Optional[torch.Tensor] = None):
    y = op.src.y
    esz = y.element_size()
    o.data = y.data_ptr() + op.coff * esz
    o.ld = y.shape[-1]
    o.K = op.K
    o.act = op.src.act
    o.scale = (op.src.scale.data_ptr() + op.coff * 4) if op.src.act else 0
    o.shift = (op.src.shift.data_ptr() + op.coff * 4) if op.src.act else 0
    t = op.wv.t if wt is None else wt
    o.w = t.data_ptr() + op.wv.off * 4
    o.w_ld = op.wv.ld
    o.w_ks = op.wv.ks
    o.dy, o.dx = op.dy, op.dx


def conv(ops: Sequence[Op], N: int, dims: Tuple[int, int, int], bias=None, adds: Sequence[Tuple[torch.Tensor, int]] = (),
         stats=None, out_dtype: Optional[int] = None, out: Optional[torch.Tensor] = None, out_coff: int = 0,
         residual: Optional[torch.Tensor] = None):
    """accx_pw_fwd.  Returns the raw [B,H,W,N] output (or writes columns [out_coff, out_coff+N) of `out`).
    residual: a [B,H,W,N] tensor in the output dtype added to the result (in the tcgen05 epilogue; otherwise as one
    more pass); `out` may be the residual itself (accumulation in place)."""
    B, H, W = dims
    arr = (Operand * len(ops))()
    for i, op in enumerate(ops):
        _fill(arr[i], op)
    in_dt = dt(ops[0].src.y)
    for op in ops:
        assert dt(op.src.y) == in_dt
    if out is None:
        odt = in_dt if out_dtype is None else out_dtype
        out = torch.empty((B, H, W, N), dtype=tdtype(odt), device=ops[0].src.y.device)
    odt = dt(out)
    ap = (ctypes.c_void_p * 4)()
    al = (ctypes.c_int * 4)()
    for i, (t, l) in enumerate(adds):
        assert t.dtype == torch.float32 and t.shape[-1] == N
        ap[i], al[i] = t.data_ptr(), l
    P = B * H * W
    seen, rd = set(), 0
    for op in ops:          # shifted taps of one tensor / column slices count once per distinct column range
        key = (op.src.y.data_ptr(), op.coff)
        if key not in seen:
            seen.add(key)
            rd += P * op.K * op.src.y.element_size()
    rd += sum(t.numel() * 4 for t, _ in adds)
    if residual is not None:
        assert residual.dtype == out.dtype and tuple(residual.shape) == (B, H, W, N) and residual.is_contiguous()
        rd += residual.numel() * residual.element_size()
    cost = (rd + P * N * out.element_size(), 2 * P * N * sum(op.K for op in ops))
    yptr = out.data_ptr() + out_coff * out.element_size()
    # fp32 storage: 3 x TF32 on the tensor cores by default; the deterministic parity mode keeps the exact fp32-FMA
    # contraction (whole-model deviation from the reference 2-6x smaller: tests/test_full_width_gpu.py)
    tc_ok = (TC and (in_dt == BF16 or (TC_F32 and in_dt == F32 and odt == F32 and (TC_F32_IN_DET or not deterministic())))
             and all(o.K % 8 == 0 and o.ld % 8 == 0 and o.data % 16 == 0 for o in arr)
             and yptr % 16 == 0 and (out.shape[-1] * out.element_size()) % 16 == 0       # TMA store of the output
             and not (stats is not None and odt == F32 and in_dt == BF16)
             and all(t.data_ptr() % 16 == 0 for t, _ in adds))
    fuse = (residual is not None and tc_ok and N % 8 == 0 and residual.data_ptr() % 16 == 0)
    if residual is not None and not fuse:
        # no fused epilogue on this path: contract into a temporary, then one accx pass adds it
        assert stats is None and out_coff == 0 and out.shape[-1] == N
        tmp = conv(ops, N, dims, bias=bias, adds=adds, out_dtype=odt)
        if residual.data_ptr() == out.data_ptr():
            return add_inplace(out, tmp)
        return _add_into(out, tmp, residual)
    if tc_ok:
        ws_bytes = _lib.load().accx_pw_tc_workspace_bytes(N, arr, len(ops))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device)
        LAUNCHES_EXTRA[0] += 1          # the weight re-pack kernel
        _call("accx_pw_fwd_tc_res", in_dt, odt, B, H, W, N, arr, len(ops), ptr(bias), ap, al, len(adds),
              ptr(residual) if fuse else 0, N, yptr, out.shape[-1], ptr(stats), ptr(ws), ws_bytes, stream(), cost=cost,
              tag=f"P={P} N={N} K={[o.K for o in ops]} shift={any(o.dy or o.dx for o in ops)} adds={len(adds)}"
                  + (" res" if fuse else ""))
    else:
        _call("accx_pw_fwd", in_dt, odt, B, H, W, N, arr, len(ops), ptr(bias), ap, al, len(adds),
              yptr, out.shape[-1], ptr(stats), stream(), cost=cost,
              tag=f"P={P} N={N} K={[o.K for o in ops]} in={in_dt} out={odt}")
    return out


def _add_into(out: torch.Tensor, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """out = a + b (one accx pass; `out` may be `a`)"""
    C = out.shape[-1]
    _call("accx_act_apply", dt(a), a.numel() // C, C, ptr(a), 0, 0, 0, 0, 0, ptr(b), ptr(out), 0, stream(),
          cost=(3 * nb(out), 0))
    return out


def wgrad(op: Op, dy: torch.Tensor, N: int, dims, gw: torch.Tensor, dy_coff: int = 0):
    """accx_pw_wgrad(_tc): accumulate the weight gradient of one operand into gw (same layout as op.wv.t)."""
    B, H, W = dims
    o = Operand()
    _fill(o, op, gw)
    in_dt = dt(op.src.y)
    dy_f32 = 1 if (dy.dtype == torch.float32 and in_dt != F32) else 0
    assert dy_f32 or dt(dy) == in_dt
    P = B * H * W
    cost = (P * (op.K * op.src.y.element_size() + N * dy.element_size()), 2 * P * N * op.K)
    tag = f"P={P} N={N} K={op.K} shift={bool(op.dy or op.dx)}"
    dwp = gw.data_ptr() + op.wv.off * 4
    dyp = dy.data_ptr() + dy_coff * dy.element_size()
    with side_stream(keep=(op.src.y, op.src.scale, op.src.shift, dy)):
        if (TC and in_dt == BF16 and not dy_f32 and op.K % 8 == 0 and N % 8 == 0 and o.ld % 8 == 0
                and dy.shape[-1] % 8 == 0 and o.data % 16 == 0 and dyp % 16 == 0):
            _call("accx_pw_wgrad_tc", B, H, W, N, ctypes.byref(o), dwp, dyp, dy.shape[-1], stream(), cost=cost, tag=tag)
        else:
            _call("accx_pw_wgrad", in_dt, B, H, W, N, ctypes.byref(o), dwp, dyp, dy.shape[-1], dy_f32, stream(),
                  cost=cost, tag=tag)


def wgrad_conv3x3(X: Lazy, C_in: int, w: torch.Tensor, dy: torch.Tensor, N: int, dims, gw: torch.Tensor):
    """weight gradient of a dense 3x3 convolution (weight [N, C_in, 3, 3], padding 1) whose input was the lazy
    tensor X: the nine taps in groups of up to five per launch (accx_pw_wgrad_taps_tc: dY and the activation are
    read once per group instead of once per tap), or tap by tap where the tensor cores cannot be used."""
    B, H, W = dims
    taps = [(ky * 3 + kx, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    in_dt = dt(X.y)
    y = X.y
    nbk = 64 if C_in >= 64 else (C_in + 15) // 16 * 16
    tc_ok = (TC and in_dt == BF16 and dy.dtype == torch.bfloat16 and C_in % 8 == 0 and N % 8 == 0 and y.shape[-1] % 8 == 0
             and dy.shape[-1] % 8 == 0 and y.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0)
    if not tc_ok:
        for t, ddy, ddx in taps:
            wgrad(Op(X, C_in, WV(w, t, C_in * 9, 9), 0, ddy, ddx), dy, N, dims, gw)
        return
    per = max(1, min(5, 512 // nbk))
    o = Operand()
    _fill(o, Op(X, C_in, WV(w, 0, C_in * 9, 9)), gw)
    P = B * H * W
    with side_stream(keep=(y, X.scale, X.shift, dy)):
        for g0 in range(0, 9, per):
            grp = taps[g0:g0 + per]
            n = len(grp)
            tdy = (ctypes.c_int * n)(*[g[1] for g in grp])
            tdx = (ctypes.c_int * n)(*[g[2] for g in grp])
            tof = (ctypes.c_int64 * n)(*[g[0] for g in grp])
            _call("accx_pw_wgrad_taps_tc", B, H, W, N, ctypes.byref(o), n, tdy, tdx, tof, gw.data_ptr(), dy.data_ptr(),
                  dy.shape[-1], stream(), cost=(P * (C_in + N) * 2, 2 * P * N * C_in * n),
                  tag=f"P={P} N={N} K={C_in} taps={n}")


def bn_affine(bn: torch.nn.BatchNorm2d, stats, count: float, arena: Arena, training: bool, conv_bias=None):
    """statistics -> (scale, shift, mean, rstd); updates the running buffers in training."""
    C = bn.num_features
    scale, shift, mean, rstd = arena.take(C), arena.take(C), arena.take(C), arena.take(C)
    mom = 0.1 if bn.momentum is None else bn.momentum
    track = bn.track_running_stats and bn.running_mean is not None
    _call("accx_bn_finalize", C, float(count), ptr(stats), ptr(f32(bn.weight)), ptr(f32(bn.bias)), ptr(f32(conv_bias)),
          float(bn.eps),
          float(mom), 1 if training else 0, ptr(bn.running_mean) if track else 0, ptr(bn.running_var) if track else 0,
          ptr(bn.num_batches_tracked) if track else 0, ptr(scale), ptr(shift), ptr(mean), ptr(rstd), stream())
    return scale, shift, mean, rstd


def bn_lazy(y: torch.Tensor, stats, bn, act: int, arena: Arena, training: bool, conv_bias=None) -> Lazy:
    """conv_bias: bias of the conv that produced y and was NOT added to it (folded into the BN here)"""
    count = y.numel() // y.shape[-1]
    scale, shift, mean, rstd = bn_affine(bn, stats, count, arena, training, conv_bias)
    L = Lazy(y, scale, shift, act, mean, rstd, bn, train=training)
    if RECORD is not None and act == 2:
        RECORD[("bn", id(bn))] = L
    return L


def materialize(L: Lazy, scale2=None, shift2=None, residual=None, stats=None, out=None, stats_only=False):
    y = L.y
    if out is None and not stats_only:
        out = torch.empty_like(y)
    _call("accx_act_apply", dt(y), y.numel() // y.shape[-1], y.shape[-1], ptr(y), ptr(L.scale), ptr(L.shift), L.act,
          ptr(scale2), ptr(shift2), ptr(residual), ptr(out), ptr(stats), stream(), cost=(nb(y, residual, out), 0),
          tag=f"P={y.numel() // y.shape[-1]} C={y.shape[-1]} stats={int(stats is not None)} out={int(out is not None)}")
    return out


def add_fwd(L: Lazy, r: torch.Tensor, stats):
    z = torch.empty_like(r)
    _call("accx_add_fwd", dt(r), r.numel() // r.shape[-1], r.shape[-1], ptr(L.y), ptr(L.scale), ptr(L.shift), L.act,
          ptr(r), ptr(z), ptr(stats), stream(), cost=(nb(L.y, r, z), 0), tag=f"P={r.numel() // r.shape[-1]} C={r.shape[-1]}")
    return z


def bn_bwd(L: Lazy, da: torch.Tensor, grads: dict, arena: Arena, out: Optional[torch.Tensor] = None,
           sums: Optional[torch.Tensor] = None, conv=None) -> torch.Tensor:
    """gradient w.r.t. the raw tensor L.y given the gradient w.r.t. act(BN(L.y)); accumulates
    dgamma/dbeta into grads[bn.weight]/grads[bn.bias].  In place on `da` unless `out` is given.
    `sums`: the (sum g, sum g*xhat) reduction if the producer of `da` already made it (hanc_unpool_bnred).
    `conv`: the convolution that produced L.y with its bias folded into this BatchNorm: its bias gradient is
    analytically zero in training mode (the mean subtraction cancels a per-channel constant) and
    sum_p dy = gamma * rstd * sum g in eval mode, where the BatchNorm is a fixed affine of the running statistics."""
    y = L.y
    C = y.shape[-1]
    P = y.numel() // C
    dtc = dt(y)
    assert da.dtype == y.dtype and da.is_contiguous()
    if sums is None:
        sums = arena.take(2 * C)
        _call("accx_bn_bwd_reduce", dtc, P, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(L.mean), ptr(L.rstd),
              ptr(da), ptr(sums), stream(), cost=(nb(y, da), 0), tag=f"P={P} C={C}")
    dy = da if out is None else out
    gg = grad_buf(grads, L.bn.weight)
    gb = grad_buf(grads, L.bn.bias)
    # eval mode: 1 / count = 0 removes the batch-statistics terms, dy = gamma * rstd * g
    _call("accx_bn_bwd_apply", dtc, P, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(L.mean), ptr(L.rstd),
          ptr(f32(L.bn.weight)), ptr(da), ptr(sums), float(P) if L.train else float("inf"), ptr(dy), ptr(gg), ptr(gb),
          stream(), cost=(nb(y, da, dy), 0), tag=f"P={P} C={C}")
    if conv is not None and conv.bias is not None:
        cb = grad_buf(grads, conv.bias)           # training: stays zero
        if cb is not None and not L.train:
            cb.add_(f32(L.bn.weight) * L.rstd * sums[:C])
    return dy


# Set by accx.train.FlatState for the duration of a training step: parameter gradients are then accumulated
# straight into ONE flat fp32 buffer (zeroed by one memset per step) that the all-reduce and the Adam kernel
# consume as is.  An object with .serves(p) -> bool and .take(p) -> zero-initialised view or None.
GRAD_ARENA = None


class GradPool(dict):
    """id(param) -> fp32 gradient accumulator.  All accumulators of one module backward are allocated up
    front and zeroed with ONE multi-tensor launch (instead of one fill kernel per parameter); a parameter
    only appears in the dict once a kernel has asked for its buffer, so parameters the module never
    touches still report grad=None (ACC_UNet_Lite's unused MLFC convs)."""

    def __init__(self, params=()):
        super().__init__()
        ga = GRAD_ARENA
        self.spare = {id(p): torch.empty(p.shape, dtype=torch.float32, device=p.device)
                      for p in params if p.requires_grad and (ga is None or not ga.serves(p))}
        if self.spare:
            torch._foreach_zero_(list(self.spare.values()))


def grad_buf(grads: dict, p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """zero-initialised fp32 gradient accumulator for parameter p (None if p takes no gradient)."""
    if p is None or not p.requires_grad:
        return None
    g = grads.get(id(p))
    if g is None and GRAD_ARENA is not None:
        g = GRAD_ARENA.take(p)       # a view into the step's flat gradient buffer (already zero), or None
        if g is not None:
            grads[id(p)] = g
    if g is None:
        spare = getattr(grads, "spare", None)
        g = spare.pop(id(p), None) if spare is not None else None
        if g is None:
            g = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
        grads[id(p)] = g
    return g


def dw_fwd(L: Lazy, w, bias, stats, flip=False):
    y = L.y
    B, H, W, C = y.shape
    out = torch.empty_like(y)
    _call("accx_dw3x3_fwd", dt(y), B, H, W, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(w), ptr(bias),
          1 if flip else 0, ptr(out), ptr(stats), stream(), cost=(nb(y, out), 18 * y.numel()),
          tag=f"{B}x{H}x{W}x{C} flip={int(flip)}")
    return out


def dw_dgrad_bnred_ok(L: Lazy, dy: torch.Tensor) -> bool:
    y = L.y
    return (L.mean is not None and L.rstd is not None and (y.shape[-1] * y.element_size()) % 16 == 0
            and y.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0)


def dw_dgrad_bnred(L: Lazy, dy: torch.Tensor, w, arena: Arena):
    """input gradient of the depthwise conv whose input was the lazy tensor L, fused with the BatchNorm-backward
    reduction of L's BatchNorm -> (da, sums) for bn_bwd(L, da, ..., sums=)"""
    y = L.y
    B, H, W, C = y.shape
    da = torch.empty_like(y)
    sums = arena.take(2 * C)
    _call("accx_dw3x3_dgrad_bnred", dt(y), B, H, W, C, ptr(dy), ptr(w), ptr(da), ptr(y), ptr(L.scale), ptr(L.shift), L.act,
          ptr(L.mean), ptr(L.rstd), ptr(sums), stream(), cost=(nb(dy, y, da), 18 * y.numel()), tag=f"{B}x{H}x{W}x{C}")
    return da, sums


def dw_wgrad(L: Lazy, dy: torch.Tensor, gw: torch.Tensor):
    y = L.y
    B, H, W, C = y.shape
    with side_stream(keep=(y, L.scale, L.shift, dy)):
        _call("accx_dw3x3_wgrad", dt(y), B, H, W, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(dy), ptr(gw),
              stream(), cost=(nb(y, dy), 18 * y.numel()), tag=f"{B}x{H}x{W}x{C}")


def hanc_pools(L: Lazy, k: int) -> List[torch.Tensor]:
    """[B,H>>l,W>>l,2C] (avg | max) for l = 1..k-1"""
    B, H, W, C = L.y.shape
    outs = []
    cur, first = L.y, 1
    for l in range(1, k):
        out = torch.empty((B, H >> l, W >> l, 2 * C), dtype=L.y.dtype, device=L.y.device)
        _call("accx_hanc_pool_fwd", dt(L.y), B, H >> (l - 1), W >> (l - 1), C, first, ptr(cur), ptr(L.scale),
              ptr(L.shift), L.act, ptr(out), stream(), cost=(nb(cur, out), 0))
        outs.append(out)
        cur, first = out, 0
    return outs


def hanc_unpool_bwd(L: Lazy, l: int, dpool: torch.Tensor, da: torch.Tensor, accumulate=True):
    B, H, W, C = L.y.shape
    assert dpool.dtype == torch.float32
    _call("accx_hanc_unpool_bwd", dt(L.y), B, H, W, C, l, ptr(L.y), ptr(L.scale), ptr(L.shift), L.act, ptr(dpool),
          ptr(da), 1 if accumulate else 0, stream(), cost=(nb(L.y, dpool, da) + (nb(da) if accumulate else 0), 0))


def hanc_unpool_bnred(L: Lazy, dpools: List[torch.Tensor], da: torch.Tensor, arena: Arena) -> torch.Tensor:
    """all pyramid levels of a k = 2 / 3 HANC backward added into `da` in one pass + the BN-backward reduction of
    L's BatchNorm on the result; returns the sums for bn_bwd(..., sums=)."""
    B, H, W, C = L.y.shape
    sums = arena.take(2 * C)
    levels = len(dpools)
    _call("accx_hanc_unpool_bnred", dt(L.y), B, H, W, C, levels, ptr(L.y), ptr(L.scale), ptr(L.shift), L.act,
          ptr(dpools[0]), ptr(dpools[1]) if levels > 1 else 0, ptr(da), ptr(L.mean), ptr(L.rstd), ptr(sums), stream(),
          cost=(nb(L.y, da, da, *dpools), 0), tag=f"{B}x{H}x{W}x{C} levels={levels}")
    return sums


def hanc_unpool_fusable(L: Lazy, k: int) -> bool:
    return (k in (2, 3) and L.y.dtype == torch.bfloat16 and L.C % 4 == 0 and L.mean is not None and L.rstd is not None
            and L.y.data_ptr() % 8 == 0)


def pool_sum(x: torch.Tensor, l: int, mul: float, out_dtype: Optional[torch.dtype] = None):
    B, H, W, C = x.shape
    out = torch.empty((B, H >> l, W >> l, C), dtype=out_dtype or x.dtype, device=x.device)
    _call("accx_pool_sum", dt(x), dt(out), B, H, W, C, l, float(mul), ptr(x), ptr(out), C, stream(), cost=(nb(x, out), 0),
          tag=f"{B}x{H}x{W}x{C} l={l}")
    return out


def upsample_add(src: torch.Tensor, dst: torch.Tensor, l: int, mul: float, accumulate: bool, src_coff=0, C=None):
    B, H, W, Cd = dst.shape
    C = Cd if C is None else C
    assert C == Cd
    _call("accx_upsample_add", dt(src), dt(dst), B, H, W, C, l, float(mul), src.data_ptr() + src_coff * src.element_size(),
          src.shape[-1], ptr(dst), 1 if accumulate else 0, stream(),
          cost=(nb(dst) * (2 if accumulate else 1) + src.numel() // src.shape[-1] * C * src.element_size(), 0))


def add_inplace(dst: torch.Tensor, other: torch.Tensor):
    """dst += other (same shape/dtype), as one accx pass"""
    C = dst.shape[-1]
    _call("accx_act_apply", dt(dst), dst.numel() // C, C, ptr(dst), 0, 0, 0, 0, 0, ptr(other), ptr(dst), 0, stream(),
          cost=(3 * nb(dst), 0))
    return dst


class SECtx:
    __slots__ = ("L", "S", "gate", "hidden", "scale", "shift", "mean", "rstd", "mod", "residual", "mix", "mix_param", "train")


def se_fwd(L: Lazy, se, arena: Arena, training: bool, residual=None, mix=None, stats=None, mix_param=None):
    """ChannelSELayer on a lazy input -> (materialised output, ctx)."""
    y = L.y
    B, H, W, C = y.shape
    Cr = se.fc1.out_features
    if Cr < 1:
        raise _lib.AccxError(f"ChannelSELayer({C}): num_channels // 8 must be >= 1")
    c = SECtx()
    c.L, c.mod, c.residual, c.mix, c.mix_param = L, se, residual, mix, mix_param
    c.train = training
    if RECORD is not None:
        RECORD[("se", id(se))] = c
    c.S = arena.take(2 * B * C)
    c.gate, c.hidden = arena.take(B * C), arena.take(B * Cr)      # every slice 256-byte aligned
    c.scale, c.shift, c.mean, c.rstd = arena.take(C), arena.take(C), arena.take(C), arena.take(C)
    counter = arena.take(1)
    d = dt(y)
    _call("accx_se_squeeze", d, B, H * W, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(c.S), stream(),
          cost=(nb(y), 0), tag=f"B={B} HW={H * W} C={C}")
    bn = se.bn
    mom = 0.1 if bn.momentum is None else bn.momentum
    _call("accx_se_gate", B, C, Cr, float(H * W), ptr(c.S), ptr(f32(se.fc1.weight)), ptr(f32(se.fc1.bias)),
          ptr(f32(se.fc2.weight)), ptr(f32(se.fc2.bias)), ptr(f32(bn.weight)), ptr(f32(bn.bias)), float(bn.eps),
          float(mom), 1 if training else 0, ptr(bn.running_mean), ptr(bn.running_var), ptr(bn.num_batches_tracked),
          ptr(c.gate), ptr(c.hidden), ptr(c.scale), ptr(c.shift), ptr(c.mean), ptr(c.rstd), ptr(counter), stream())
    out = torch.empty_like(y)
    _call("accx_se_apply", d, B, H * W, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(c.gate), ptr(c.scale),
          ptr(c.shift), ptr(residual), ptr(mix), ptr(out), ptr(stats), stream(), cost=(nb(y, residual, out), 0),
          tag=f"B={B} HW={H * W} C={C} res={int(residual is not None)} stats={int(stats is not None)}")
    return out, c


def se_bwd(c: SECtx, dout: torch.Tensor, grads: dict, arena: Arena, da: Optional[torch.Tensor] = None,
           accumulate=False, bn_sums=False, gmix: Optional[torch.Tensor] = None):
    """gradient w.r.t. the activated SE input given d(out); parameter grads accumulate in `grads`.
    (The residual branch, if any, simply receives dout * (1 - mix) -- handled by the caller.)
    bn_sums=True: -> (da, sums) where sums is the BatchNorm-backward reduction of the SE input's own
    BatchNorm on da (made in the same pass; hand it to bn_bwd(..., sums=)), or None if the input has no BN."""
    L, se = c.L, c.mod
    y = L.y
    B, H, W, C = y.shape
    Cr = se.fc1.out_features
    d = dt(y)
    assert dout.dtype == y.dtype and dout.is_contiguous()
    G = arena.take(2 * B * C)
    PQR = arena.take(3 * B * C)
    # gmix: where the gradient of the blend scalar is accumulated (default: the parameter's own accumulator; MLFC hands
    # every pyramid level its own slot and adds the four in level order, so concurrent lanes never race on one float)
    if gmix is None and c.mix_param is not None:
        gmix = grad_buf(grads, c.mix_param)
    _call("accx_se_bwd_reduce", d, B, H * W, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(c.gate), ptr(c.scale),
          ptr(c.shift), ptr(dout), ptr(c.mix), ptr(c.residual) if gmix is not None else 0, ptr(gmix), ptr(G), stream(),
          cost=(nb(y, dout), 0), tag=f"B={B} HW={H * W} C={C}")
    _call("accx_se_bwd_gate", B, C, Cr, float(H * W), ptr(c.S), ptr(G), ptr(c.gate), ptr(c.hidden),
          ptr(f32(se.fc1.weight)), ptr(f32(se.fc2.weight)), ptr(f32(se.bn.weight)), ptr(c.mean), ptr(c.rstd),
          ptr(grad_buf(grads, se.fc1.weight)), ptr(grad_buf(grads, se.fc1.bias)), ptr(grad_buf(grads, se.fc2.weight)),
          ptr(grad_buf(grads, se.fc2.bias)), ptr(grad_buf(grads, se.bn.weight)), ptr(grad_buf(grads, se.bn.bias)),
          ptr(PQR), 1 if c.train else 0, stream())
    if da is None:
        da = torch.empty_like(y)
        accumulate = False
    sums = arena.take(2 * C) if (bn_sums and L.mean is not None and L.rstd is not None) else None
    _call("accx_se_bwd_apply", d, B, H * W, C, ptr(y), ptr(L.scale), ptr(L.shift), L.act, ptr(c.gate), ptr(c.scale),
          ptr(c.shift), ptr(dout), ptr(c.mix), ptr(PQR), ptr(da), 1 if accumulate else 0, ptr(L.mean) if sums is not None else 0,
          ptr(L.rstd) if sums is not None else 0, ptr(sums), stream(), cost=(nb(y, dout, da), 0),
          tag=f"B={B} HW={H * W} C={C} acc={int(bool(accumulate))} bn={int(sums is not None)}")
    return (da, sums) if bn_sums else da


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    """NCHW-shaped tensor (any strides) -> contiguous [B,H,W,C]; free for channels_last inputs."""
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    if x.is_contiguous():
        B, C, H, W = x.shape
        out = torch.empty((B, H, W, C), dtype=x.dtype, device=x.device)
        _call("accx_nchw_to_nhwc", dt(x), dt(out), B, C, H * W, ptr(x), ptr(out), stream())
        return out
    return v.contiguous()


def to_nchw_view(y: torch.Tensor) -> torch.Tensor:
    """[B,H,W,C] contiguous -> NCHW-shaped channels_last view (no copy)."""
    return y.permute(0, 3, 1, 2)


def input_to_nhwc(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """model entry: NCHW (contiguous) -> [B,H,W,C] in the compute dtype, one accx pass"""
    if not x.is_contiguous():
        x = x.contiguous()
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=dtype, device=x.device)
    _call("accx_nchw_to_nhwc", dt(x), dt(out), B, C, H * W, ptr(x), ptr(out), stream())
    return out


# ---- the steps either side of the blocks inside a training step (include/accx.h, rows f1/f2) -----------------
def maxpool2(x: torch.Tensor) -> torch.Tensor:
    """MaxPool2d(2) on a contiguous [B,H,W,C] tensor"""
    B, H, W, C = x.shape
    if RECORD is not None:
        RECORD.setdefault("pool", []).append(x)
    out = torch.empty((B, H // 2, W // 2, C), dtype=x.dtype, device=x.device)
    _call("accx_maxpool2_fwd", dt(x), B, H, W, C, ptr(x), ptr(out), stream(), cost=(nb(x, out), 0),
          tag=f"{B}x{H}x{W}x{C}")
    return out


def maxpool2_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    B, H, W, C = x.shape
    assert dy.dtype == x.dtype and dy.is_contiguous()
    dx = torch.empty_like(x)
    _call("accx_maxpool2_bwd", dt(x), B, H, W, C, ptr(x), ptr(dy), ptr(dx), stream(), cost=(nb(x, dy, dx), 0),
          tag=f"{B}x{H}x{W}x{C}")
    return dx


def dice_bce_fwd(logit: torch.Tensor, truth: torch.Tensor, dice_w: float, bce_w: float):
    """logit [B, N] (fp32 / bf16), truth [B, N] fp32 -> (loss scalar tensor, sums for the backward)"""
    B, N = logit.shape
    sums = torch.zeros(8 * B + 8, dtype=torch.float32, device=logit.device)
    loss = torch.empty((), dtype=torch.float32, device=logit.device)
    _call("accx_dice_bce_fwd", dt(logit), B, N, ptr(logit), ptr(truth), float(dice_w), float(bce_w), ptr(sums), ptr(loss),
          stream(), cost=(nb(logit, truth), 0))
    return loss, sums


def dice_bce_bwd(logit, truth, sums, dice_w, bce_w, gscale: Optional[torch.Tensor], grad_dtype: torch.dtype):
    B, N = logit.shape
    d = torch.empty((B, N), dtype=grad_dtype, device=logit.device)
    _call("accx_dice_bce_bwd", dt(logit), dt(d), B, N, ptr(logit), ptr(truth), ptr(sums), float(dice_w), float(bce_w),
          ptr(gscale), ptr(d), stream(), cost=(nb(logit, truth, d), 0))
    return d


# ---- UNeXt shifted-MLP block pieces (include/accx.h, row f4) ------------------------------------------------
def layernorm_fwd(x: torch.Tensor, gamma, beta, eps: float, y: torch.Tensor, mean, rstd):
    """x, y: [.., C] contiguous; mean / rstd: fp32 [rows]"""
    C = x.shape[-1]
    R = x.numel() // C
    _call("accx_layernorm_fwd", dt(x), R, C, ptr(x), ptr(gamma), ptr(beta), float(eps), ptr(y), ptr(mean), ptr(rstd), stream(),
          cost=(nb(x, y), 0), tag=f"R={R} C={C}")


def layernorm_bwd(x, gamma, mean, rstd, dy, dx, dgamma, dbeta):
    C = x.shape[-1]
    R = x.numel() // C
    assert dy.dtype == x.dtype and dy.is_contiguous() and dx.dtype == x.dtype
    _call("accx_layernorm_bwd", dt(x), R, C, ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dy), ptr(dx), ptr(dgamma), ptr(dbeta),
          stream(), cost=(nb(x, dy, dx), 0), tag=f"R={R} C={C}")


def gelu(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    _call("accx_gelu_fwd", dt(x), x.numel(), ptr(x), ptr(y), stream(), cost=(nb(x, y), 0))
    return y


def gelu_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    assert dy.dtype == x.dtype and dy.is_contiguous() and x.is_contiguous()
    dx = torch.empty_like(x)
    _call("accx_gelu_bwd", dt(x), x.numel(), ptr(x), ptr(dy), ptr(dx), stream(), cost=(nb(x, dy, dx), 0))
    return dx


def seg_metrics(logit: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    """logit [B, N] (fp32 / bf16), truth [B, N] fp32 -> device tensor [mean IoU, mean hard Dice] (accx_seg_metrics)"""
    B, N = logit.shape
    counts = torch.zeros(4 * B + 1, dtype=torch.int32, device=logit.device)
    out = torch.empty(2, dtype=torch.float32, device=logit.device)
    _call("accx_seg_metrics", dt(logit), B, N, ptr(logit), ptr(truth), ptr(counts), ptr(out), stream(), cost=(nb(logit, truth), 0))
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              grad_scale=1.0):
    """one Adam step over flat fp32 buffers (state[0] = device-side step count, incremented by the call)"""
    n = param.numel()
    global LAUNCHES
    LAUNCHES += 1                      # the step-count tick
    _call("accx_adam_step", n, ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(state), float(lr), float(beta1),
          float(beta2), float(eps), float(weight_decay), float(grad_scale), stream(), cost=(7 * 4 * n, 0))


def upshuffle(temp: torch.Tensor, bias, out: torch.Tensor, Co: int, forward: bool, dbias=None):
    """temp [B,H,W,4*Co] <-> the left Co columns of out [B,2H,2W,ld] (see accx_upshuffle in include/accx.h)"""
    B, H, W, _ = temp.shape
    assert temp.dtype == out.dtype and temp.is_contiguous() and out.is_contiguous()
    _call("accx_upshuffle", dt(temp), 1 if forward else 0, B, H, W, Co, ptr(temp), ptr(bias), ptr(out), out.shape[-1],
          ptr(dbias), stream(), cost=(2 * nb(temp), 0), tag=f"{B}x{H}x{W}x{Co} fwd={int(forward)}")


def copy_cols(src: torch.Tensor, src_coff: int, dst: torch.Tensor, dst_coff: int, C: int):
    """dst[..., dst_coff:dst_coff+C] = src[..., src_coff:src_coff+C] for contiguous [.., ld] matrices of one dtype"""
    assert src.dtype == dst.dtype and src.is_contiguous() and dst.is_contiguous()
    P = src.numel() // src.shape[-1]
    assert P == dst.numel() // dst.shape[-1]
    es = src.element_size()
    _call("accx_copy_cols", dt(src), P, C, src.data_ptr() + src_coff * es, src.shape[-1], dst.data_ptr() + dst_coff * es,
          dst.shape[-1], stream(), cost=(2 * P * C * es, 0), tag=f"P={P} C={C}")
