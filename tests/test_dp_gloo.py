"""Data-parallel host logic on CPU: world_size 2 over gloo (the N>1 path of accx/train.py).

The kernels need a GPU, the gradient exchange does not: GradAverager must average every live
gradient over the ranks with one flat all-reduce, skip parameters whose gradient is None on every
rank (ACC_UNet_Lite's unused MLFC convs, ACC_UNet_lite.py:424-427) and leave ranks bit-identical.
Both layouts are covered: loose p.grad tensors (gather / all-reduce / scatter) and the train step's
FlatState (gradients live in ONE flat buffer that is all-reduced in place).
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, use_flat=False):
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from accx.train import FlatState, GradAverager, dice_bce_loss_torch as dice_bce_loss

    torch.manual_seed(2)                                  # identical replicas
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.LeakyReLU(), torch.nn.Conv2d(4, 1, 1))
    unused = torch.nn.Parameter(torch.zeros(5))           # never receives a gradient on any rank
    params = list(net.parameters()) + [unused]
    g = torch.Generator().manual_seed(100 + rank)         # rank-offset shard of the global batch
    x = torch.randn(2, 3, 8, 8, generator=g)
    m = (torch.rand(2, 1, 8, 8, generator=g) > 0.5).float()
    loss = dice_bce_loss(net(x), m)
    flat = None
    if use_flat:
        flat = FlatState(params)                          # parameters / gradients move into flat buffers
        flat.begin_step()
        loss = dice_bce_loss(net(x), m)
    loss.backward()
    local = [None if p.grad is None else p.grad.clone() for p in params]
    if use_flat:
        flat.collect()                                    # autograd's loose gradients -> slices of flat.grad
        assert all(p.grad is None or p.grad.data_ptr() == flat._view(flat.grad, p).data_ptr() for p in params)
    GradAverager(params, flat=flat)()
    torch.save({"local": local, "avg": [None if p.grad is None else p.grad.clone() for p in params],
                "loss": loss.detach()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("use_flat", [False, True], ids=["loose", "flat"])
def test_grad_averager_world2_gloo(tmp_path, use_flat):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), use_flat), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    n = len(r[0]["local"])
    assert r[0]["avg"][-1] is None and r[1]["avg"][-1] is None          # unused parameter stays grad=None
    for j in range(n - 1):
        want = (r[0]["local"][j] + r[1]["local"][j]) / world
        assert not torch.equal(r[0]["local"][j], r[1]["local"][j])      # shards really differ
        for i in range(world):
            torch.testing.assert_close(r[i]["avg"][j], want, rtol=1e-6, atol=1e-7)
        assert torch.equal(r[0]["avg"][j], r[1]["avg"][j])              # replicas stay bit-identical


def test_reference_arm_other_ranks_do_no_work(monkeypatch, capsys):
    """bench.py --impl reference under torchrun: only rank 0 runs and prints (contract in bench.py docstring)."""
    sys.path.insert(0, ROOT)
    import bench
    import argparse
    bench.run_reference(argparse.Namespace(gpus=2, steps=1, warmup=1), rank=1)
    assert capsys.readouterr().out == ""


def test_flat_state_layout_cpu():
    """FlatState host logic: aligned slices, parameters become views, one-shot gradient hand-out per step"""
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from accx.train import FlatState
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 5, 3), torch.nn.BatchNorm2d(5), torch.nn.Linear(7, 3))
    before = [p.detach().clone() for p in net.parameters()]
    params = list(net.parameters())
    fs = FlatState(params)
    assert fs.n % FlatState.ALIGN == 0 and all(o % FlatState.ALIGN == 0 for o in fs.offsets.values())
    for p, b in zip(params, before):
        assert torch.equal(p.detach(), b)                                       # values preserved
        o = fs.offsets[id(p)]
        assert p.data_ptr() == fs.param.data_ptr() + 4 * o                      # p lives in the flat buffer
    assert list(net.state_dict().keys())[:2] == ["0.weight", "0.bias"]          # names / shapes untouched
    fs.begin_step()
    p0 = params[0]
    assert fs.serves(p0)
    g = fs.take(p0)
    assert g.shape == p0.shape and g.data_ptr() == fs.grad.data_ptr() + 4 * fs.offsets[id(p0)]
    assert not fs.serves(p0) and fs.take(p0) is None                            # handed out once per step
    assert fs.take(torch.nn.Parameter(torch.zeros(2))) is None                  # foreign parameter
    g.add_(1.0)
    params[1].grad = torch.full_like(params[1], 2.0)                            # a loose gradient made by autograd
    fs.collect()
    assert params[1].grad.data_ptr() == fs._view(fs.grad, params[1]).data_ptr()
    assert float(fs.grad.sum()) == p0.numel() * 1.0 + params[1].numel() * 2.0   # padding stays zero
    fs.begin_step()
    assert float(fs.grad.abs().sum()) == 0.0 and all(p.grad is None for p in params) and fs.serves(p0)


def test_grad_arena_hands_flat_views_to_module_backward_cpu():
    """engine.GRAD_ARENA protocol (host logic, no kernels): while a FlatState is the arena, GradPool leaves the
    parameters it serves to the arena, grad_buf returns the arena's view once per step, and a foreign parameter
    (not in the flat buffers) still gets its own zeroed accumulator."""
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from accx import engine as E
    from accx.train import FlatState
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    params = list(net.parameters())
    foreign = torch.nn.Parameter(torch.ones(3))
    fs = FlatState(params)
    fs.begin_step()
    old, E.GRAD_ARENA = E.GRAD_ARENA, fs
    try:
        pool = E.GradPool(params + [foreign])
        assert set(pool.spare) == {id(foreign)}                       # only the foreign parameter gets a loose buffer
        g0 = E.grad_buf(pool, params[0])
        assert g0.data_ptr() == fs._view(fs.grad, params[0]).data_ptr() and float(g0.abs().sum()) == 0.0
        assert E.grad_buf(pool, params[0]) is g0                      # same accumulator within one module backward
        gf = E.grad_buf(pool, foreign)
        assert gf.data_ptr() != fs.grad.data_ptr() and float(gf.abs().sum()) == 0.0
        # a second module backward in the same step asking for the same parameter does NOT get the slice again
        pool2 = E.GradPool([params[0]])
        g1 = E.grad_buf(pool2, params[0])
        assert g1.data_ptr() != g0.data_ptr() and float(g1.abs().sum()) == 0.0
        frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
        assert E.grad_buf(pool, frozen) is None
    finally:
        E.GRAD_ARENA = old
    fs.begin_step()                                                   # next step: the slices are handed out again
    E.GRAD_ARENA = fs
    try:
        assert E.grad_buf(E.GradPool(params), params[0]).data_ptr() == fs._view(fs.grad, params[0]).data_ptr()
    finally:
        E.GRAD_ARENA = old


def test_arena_gradients_never_pass_through_autograd_cpu():
    """Regression for the stale-gradient race (round-1 advisor finding): a gradient accumulated in place in the flat
    buffer must not be returned to autograd -- AccumulateGrad clones a tensor that anything else still references,
    and under TrainStep's deferred side-stream join that clone could be taken before the weight-gradient kernels
    have run.  engine.param_grads returns None for arena-owned gradients; FlatState.collect points p.grad at the
    slice without a copy; a second, loose gradient of a served parameter is ADDED to the slice."""
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from accx import engine as E
    from accx.train import FlatState
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    params = list(net.parameters())
    foreign = torch.nn.Parameter(torch.ones(3))
    fs = FlatState(params)
    fs.begin_step()
    old, E.GRAD_ARENA = E.GRAD_ARENA, fs
    try:
        pool = E.GradPool(params + [foreign])
        g0 = E.grad_buf(pool, params[0])
        gf = E.grad_buf(pool, foreign)
        out = E.param_grads(params + [foreign, None], pool)
        assert out[0] is None                                  # lives in the flat buffer: autograd never sees it
        assert out[1] is None and out[2] is None               # never touched by the module: grad stays None
        assert out[len(params)] is gf and out[-1] is None      # the foreign parameter's loose buffer is returned
        assert fs.owns(params[0], g0) and not fs.owns(foreign, gf) and not fs.owns(params[1], g0)
        g0.add_(1.0)                                           # "the weight-gradient kernel" writes late ...
        pool2 = E.GradPool([params[0]])                        # ... and a second use yields a loose gradient
        g1 = E.grad_buf(pool2, params[0])
        assert E.param_grads([params[0]], pool2)[0] is g1
        g1.add_(0.5)
        params[0].grad = g1
    finally:
        E.GRAD_ARENA = old
    assert params[1].grad is None
    fs.collect()
    v = fs._view(fs.grad, params[0])
    assert params[0].grad.data_ptr() == v.data_ptr()           # p.grad IS the slice, no copy-back of a snapshot
    assert torch.all(v == 1.5)                                 # in-place part + the loose second use
    assert params[1].grad is None                              # untouched parameters keep grad=None


@pytest.mark.timeout(600)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the reference's CPU path on the host cores): stdout is exactly one JSON line
    carrying the keys the driver reads; anything else the process prints goes to stderr."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=580)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("ACC-UNet train images/sec") and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    # the same config object as the accx arm's line for the same flags (the driver pairs the two lines by it)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config("ACC_UNet", bench.PER_GPU_BATCH, bench.HW, 1, 1)
    assert "16x3x224x224 per GPU" in d["config"]["workload"] and "2x3x224x224" in d["cpu_baseline"]["sample"]


# ---- backward-overlapped exchange: ranges of the flat buffer reduced early + the final reduction of the rest ----------
def test_allreduce_phase_plan_covers_contiguous_ranges():
    """ACC_UNet.allreduce_phases -> slices of the flat gradient buffer: decoder (60 % of the parameters) and MLFC are
    contiguous in parameter order, disjoint, and hold exactly their modules' parameters"""
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import accx
    from accx.train import FlatState, plan_ranges
    model = accx.ACC_UNet(3, 1, 8)
    flat = FlatState(list(model.parameters()))
    plan = plan_ranges(flat, model.allreduce_phases())
    assert [t if isinstance(t, str) else "mod" for t, *_ in plan] == ["mod", "group"]
    (_, lo1, hi1, p1), (_, lo2, hi2, p2) = plan
    assert hi1 == flat.n and lo1 == hi2 and 0 < lo2 < hi2                 # head | MLFC | decoder
    names = {id(p): n for n, p in model.named_parameters()}
    assert all(names[id(p)].split(".")[0].startswith(("up", "cnv6", "cnv7", "cnv8", "cnv9", "out")) for p in p1)
    assert all(names[id(p)].startswith("mlfc") for p in p2)
    frac = sum(p.numel() for p in p1) / sum(p.numel() for p in model.parameters())
    assert 0.5 < frac < 0.7
    # a phase whose parameters are not contiguous is refused
    with pytest.raises(ValueError):
        plan_ranges(flat, [("group", [model.cnv11, model.out])])


def _worker_ranges(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from accx.train import FlatState, GradAverager
    torch.manual_seed(2)
    params = [torch.nn.Parameter(torch.randn(n)) for n in (70, 5, 130, 64, 9)]
    flat = FlatState(params)
    flat.grad.copy_(torch.randn(flat.n, generator=torch.Generator().manual_seed(7 + rank)))
    ref = flat.grad.clone()
    dist.all_reduce(ref)
    ref /= world
    avg = GradAverager(params, flat=flat)

    class Done:                      # the early reductions' handle: already complete on this backend
        def wait(self):
            return True

    # two ranges go out "early" (sum + divide here: gloo has no AVG), the final call covers head, gap and tail
    for lo, hi in ((128, 192), (320, 384)):
        v = flat.grad[lo:hi]
        dist.all_reduce(v)
        v /= world
        avg.pending.append((lo, hi, Done()))
    avg()
    assert not avg.pending
    torch.save({"got": flat.grad.clone(), "ref": ref}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_early_ranges_plus_final_reduction_equal_one_all_reduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker_ranges, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2))
    assert torch.allclose(r0["got"], r0["ref"], rtol=0, atol=1e-6) and torch.equal(r0["got"], r1["got"])


def test_rank_local_legs_of_the_bench_never_enter_a_collective():
    """bench.py under torchrun: after the timed region only rank 0 runs the instrumented per-kernel step, the other
    ranks wait in the final barrier -- a TrainStep built there must be rank-local (distributed=False: no start-up
    broadcast, no gradient exchange), and the fp32 / CPU legs run at N = 1 only.  (A plain TrainStep(...) in that leg
    deadlocked every N > 1 run against the barrier until the collective timeout.)"""
    import inspect
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    rank0 = src[src.index("if rank == 0 and not args.no_kernel_table:"):src.index("# ---- secondary result")]
    made = re.findall(r"TrainStep\(([^)]*)\)", rank0)
    assert made and all("distributed=False" in a for a in made), made
    assert "if world == 1 and args.dtype == \"bf16\" and not args.no_fp32:" in src
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from accx.train import TrainStep
    sig = inspect.signature(TrainStep.__init__)
    assert sig.parameters["distributed"].default is True
    body = inspect.getsource(TrainStep.__init__)
    assert body.index("if not distributed:") < body.index("dist.broadcast(self.flat.param"), "world must be 1 before the broadcast"
