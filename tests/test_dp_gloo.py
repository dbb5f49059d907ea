"""Data-parallel host logic on CPU: world_size 2 over gloo (the N>1 path of accx/train.py).

The kernels need a GPU, the gradient exchange does not: GradAverager must average every live
gradient over the ranks with one flat all-reduce, skip parameters whose gradient is None on every
rank (ACC_UNet_Lite's unused MLFC convs, ACC_UNet_lite.py:424-427) and leave ranks bit-identical.
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


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from accx.train import GradAverager, dice_bce_loss

    torch.manual_seed(2)                                  # identical replicas
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.LeakyReLU(), torch.nn.Conv2d(4, 1, 1))
    unused = torch.nn.Parameter(torch.zeros(5))           # never receives a gradient on any rank
    params = list(net.parameters()) + [unused]
    g = torch.Generator().manual_seed(100 + rank)         # rank-offset shard of the global batch
    x = torch.randn(2, 3, 8, 8, generator=g)
    m = (torch.rand(2, 1, 8, 8, generator=g) > 0.5).float()
    loss = dice_bce_loss(net(x), m)
    loss.backward()
    local = [None if p.grad is None else p.grad.clone() for p in params]
    GradAverager(params)()
    torch.save({"local": local, "avg": [None if p.grad is None else p.grad.clone() for p in params],
                "loss": loss.detach()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_grad_averager_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
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
