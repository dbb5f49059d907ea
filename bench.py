#!/usr/bin/env python
"""bench.py -- ACC-UNet training throughput on B200 (BASELINE.json: "ACC-UNet train images/sec @224^2").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl accx|reference] [--dtype bf16|fp32]

Workload (BASELINE.json configs[1]): ACC_UNet(3, 1, 32) full train step -- forward, WeightedDiceBCE
on logits, backward, Adam(lr 1e-3) -- on a 16x3x224x224 GlaS-shaped synthetic batch per GPU.
One step = one pass of that over one batch.  N > 1: one process per GPU (torchrun), batch-sharded
(weak scaling: 16 images per GPU), NCCL gradient all-reduce.

Prints ONE JSON line (rank 0): value = images/s with inputs resident in HBM, e2e = the same through
the public API with pinned-host inputs copied H2D and the loss read back D2H inside the timed region,
roofline = the dominant accx kernel (live CUDA-event timing), cpu_baseline = the CPU oracle port on a
bounded sample.  --impl reference times the reference's CPU implementation (the oracle port:
/root/reference is Python and not present on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "acc-unet-unext_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "ACC-UNet train images/sec @224^2"
UNIT = "images/s"
HW = 224
PER_GPU_BATCH = 16
CPU_SAMPLE_BATCH = 2          # CPU leg: 2 images per step (16 do not fit host RAM: 4.35 GB/img of saved activations)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_images_per_s(steps, warmup, threads):
    """reference's CPU implementation of the step (oracle port), bounded sample"""
    from oracle import acc_oracle as O
    torch.set_num_threads(threads)
    torch.manual_seed(2)
    sd = O.init_acc_unet(3, 1, 32)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(CPU_SAMPLE_BATCH, 3, HW, HW, generator=g)
    m = (torch.rand(CPU_SAMPLE_BATCH, 1, HW, HW, generator=g) > 0.5).float()
    opt = None
    for _ in range(warmup):
        _, opt = O.train_step(sd, x, m, opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, opt = O.train_step(sd, x, m, opt)
    dt = time.perf_counter() - t0
    return CPU_SAMPLE_BATCH * steps / dt, dt / steps


CLS_NAMES = {"base": "ACC_UNet", "w": "ACC_UNet_W", "lite": "ACC_UNet_Lite"}


def workload_config(cls_name, B, hw, world, graph):
    """the `config` object of the result line: the workload both arms are quoted on (BASELINE.json configs[1] by default)"""
    return {"workload": f"{cls_name}(3,1,32) full train step (fwd + Dice/BCE + bwd + Adam), "
                        f"{B}x3x{hw}x{hw} per GPU, GlaS-shaped synthetic",
            "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": bool(graph),
            "l2_policy": "per-step working set (several GB of activations) >> 126 MB L2; no flush needed"}


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    ips, spstep = cpu_oracle_images_per_s(steps, warm, cores)
    sample = (f"{CPU_SAMPLE_BATCH}x3x{HW}x{HW} fp32 train step (fwd+Dice/BCE+bwd+Adam) of ACC_UNet(3,1,32) through the CPU oracle "
              f"port of the reference, {steps} timed steps after {warm} warm-up, normalised to images/s")
    world = max(1, int(getattr(args, "gpus", 1)))
    world_b = getattr(args, "batch", PER_GPU_BATCH)
    if getattr(args, "global_batch", 0):
        world_b = args.global_batch // world
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the same workload / config as the accx arm is quoted on; what was actually timed (a bounded sample of it on the
        # host cores, through the CPU port of the reference's formulas) is stated in cpu_baseline
        "config": workload_config(CLS_NAMES[getattr(args, "variant", "base")], world_b, getattr(args, "hw", HW), world,
                                  getattr(args, "graph", 1)),
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


_RESULT_FD = None


def emit(line: str):
    """the one result line -> the process's original stdout"""
    sys.stdout.flush()
    if _RESULT_FD is None:
        print(line, flush=True)
    else:
        os.write(_RESULT_FD, (line + "\n").encode())


def kernel_table(profile):
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1, nbytes, flops, _tag in profile:
        a = agg.setdefault(name, [0, 0.0, 0, 0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += nbytes
        a[3] += flops
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="accx", choices=["accx", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="images per GPU (weak scaling)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fixed total batch split over the GPUs (strong scaling; BASELINE configs[4]: 64 at --hw 512)")
    ap.add_argument("--variant", default="base", choices=["base", "w", "lite"],
                    help="ACC_UNet / ACC_UNet_W / ACC_UNet_Lite (BASELINE configs[3]; the headline is base)")
    ap.add_argument("--hw", type=int, default=HW, help="image side (BASELINE configs[4] uses 512)")
    ap.add_argument("--graph", type=int, default=1, help="capture the whole step in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary fp32-storage measurement (N = 1 only)")
    ap.add_argument("--no-kernel-table", action="store_true", help="skip the instrumented eager step (roofline = null)")
    ap.add_argument("--kernel-table", default="", help="write the per-kernel timing table (JSON) here")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE line, the JSON result: file descriptor 1 is pointed at stderr for the whole run
    # (NCCL prints its version banner with a C-level printf that no Python-side redirection catches) and the
    # result line is written to the saved descriptor at the end
    sys.stdout.flush()
    global _RESULT_FD
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import accx
    from accx import engine as E
    from accx.train import TrainStep

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries ONE JSON line: whatever NCCL logs (version banner, warnings) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import datetime
        # a rank that dies (e.g. out of memory) must not leave the others in a 10-minute collective timeout
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    accx.load_library()
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch
    if args.global_batch:
        assert args.global_batch % world == 0, "--global-batch must be a multiple of the number of GPUs"
        B = args.global_batch // world
    cd = torch.bfloat16 if args.dtype == "bf16" else torch.float32

    torch.manual_seed(2)                                   # same weights on every rank
    cls = {"base": accx.ACC_UNet, "w": accx.ACC_UNet_W, "lite": accx.ACC_UNet_Lite}[args.variant]
    hw = args.hw
    model = cls(3, 1, 32, compute_dtype=cd).to(dev).train()
    model.last_activation = None                           # logits for the logit-based loss (ACC_UNet.py:653-657)
    step = TrainStep(model, lr=1e-3, graph=bool(args.graph))
    g = torch.Generator().manual_seed(100 + rank)          # rank-offset data
    x_host = torch.randn(B, 3, hw, hw, generator=g).pin_memory()
    m_host = (torch.rand(B, 1, hw, hw, generator=g) > 0.5).float().pin_memory()
    x_dev, m_dev = x_host.to(dev), m_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l0 = E.LAUNCHES
    step(x_dev, m_dev)                                     # first (eager) step: count launches of one step
    launches_per_step = E.LAUNCHES - l0
    for _ in range(W + 2):                                 # eager warm-up + graph capture happen here
        loss = step(x_dev, m_dev)
    barrier()
    assert torch.isfinite(loss).all(), "loss is not finite"

    sampler = ClockSampler(local)
    sampler.start()
    # ---- timed region 1: inputs resident in HBM -------------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = step(x_dev, m_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- timed region 2: end to end (pinned host -> device, loss -> host every step) ------------
    barrier()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        md = m_host.to(dev, non_blocking=True)
        loss = step(xd, md)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller looks at the loss every step
        return loss

    for _ in range(3):                                     # warm-up of THIS path (device buffers of the H2D copies)
        e2e_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    walls = []
    for _ in range(K):
        tw = time.perf_counter()
        loss = e2e_step()
        walls.append((time.perf_counter() - tw) * 1e3)
    if rank == 0:
        print("e2e per-step wall ms: " + " ".join(f"{w:.2f}" for w in walls), file=sys.stderr)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    value = world * B * K / (ms / 1e3)
    e2e_value = world * B * K / (ms_e2e / 1e3)

    peak_mem_gib = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    # the captured step (its private memory pool holds one step's activations) is released before the next legs
    step.graph = None
    step.static_x = step.static_m = step.static_loss = None
    del step
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    # ---- per-kernel table: one instrumented eager step (events around every accx launch) ---------
    roof = None
    if rank == 0 and not args.no_kernel_table:
        hbm, tflops, which = peaks()
        # rank-0-only instrumentation: no collective at all, neither the start-up broadcast nor the gradient exchange
        # (the other ranks are not in it: a broadcast here deadlocks against their final barrier)
        eager = TrainStep(model, lr=1e-3, graph=False, distributed=False)
        # per-kernel timing needs kernels that run alone: no side stream, no parallel lanes in this step
        side_mode, lanes_mode = E.SIDE_MODE, E.LANES
        E.SIDE_MODE, E.LANES = 0, 0
        try:
            eager(x_dev, m_dev)
            torch.cuda.synchronize()
            table_ok = True
        except torch.OutOfMemoryError:
            table_ok = False
            E.SIDE_MODE, E.LANES = side_mode, lanes_mode
            print("kernel table skipped: the instrumented eager step does not fit next to the step's buffers", file=sys.stderr)
    if rank == 0 and not args.no_kernel_table and table_ok:
        E.PROFILE = []
        E.PROFILE_LEAD = (192, 24_000_000)   # keep the (slower) launching CPU ahead of the GPU: see engine._call
        t_e0, t_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_e0.record()
        eager(x_dev, m_dev)
        t_e1.record()
        agg = kernel_table(E.PROFILE)
        calls = sorted(((e0.elapsed_time(e1), n_, tg, nb_, fl_) for n_, e0, e1, nb_, fl_, tg in E.PROFILE), reverse=True)[:40]
        prof_rows = E.PROFILE
        E.PROFILE = None
        E.PROFILE_LEAD = None
        E.SIDE_MODE, E.LANES = side_mode, lanes_mode
        step_ms = t_e0.elapsed_time(t_e1)
        tot = sum(a[1] for a in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1][1])
        name, (n, tms, nbytes, flops) = top
        ai = flops / max(nbytes, 1)
        if flops and ai > tflops * 1e12 / (hbm * 1e9):
            roof = {"bound": "tensor", "achieved": flops / (tms / 1e3) / 1e12, "peak": tflops, "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": nbytes / (tms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof.update({"kernel": name, "launches_per_step": n, "avg_launch_us": tms / n * 1e3,
                     "algorithmic_bytes_per_launch": nbytes / n,
                     "share_of_accx_time": tms / tot, "accx_kernel_ms_per_step": tot, "eager_step_ms": step_ms,
                     "peak_source": which, "traffic": traffic_from_profiles(name),
                     "traffic_source": "profiles/traffic.json: ncu dram__bytes of one eager step at the round-2 head "
                                       "(not this run)"})
        table = {k: {"launches": a[0], "ms": a[1], "alg_gbytes": a[2] / 1e9, "gflop": a[3] / 1e9,
                     "gb_per_s": a[2] / max(a[1], 1e-9) / 1e6, "tflop_per_s": a[3] / max(a[1], 1e-9) / 1e9}
                 for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])}
        if args.kernel_table:
            by_shape = {}
            for n_, e0_, e1_, nb_, fl_, tg in prof_rows:
                a = by_shape.setdefault(f"{n_} {tg}", [0, 0.0, 0, 0])
                a[0] += 1
                a[1] += e0_.elapsed_time(e1_)
                a[2] += nb_
                a[3] += fl_
            shapes = [{"call": k, "n": a[0], "ms": a[1], "gb_per_s": a[2] / max(a[1], 1e-9) / 1e6,
                       "tflop_per_s": a[3] / max(a[1], 1e-9) / 1e9}
                      for k, a in sorted(by_shape.items(), key=lambda kv: -kv[1][1])]
            json.dump({"step_ms_eager": step_ms, "accx_ms": tot, "kernels": table, "by_shape": shapes,
                       "slowest_calls": [{"ms": c[0], "kernel": c[1], "shape": c[2], "gb_per_s": c[3] / max(c[0], 1e-9) / 1e6,
                                          "tflop_per_s": c[4] / max(c[0], 1e-9) / 1e9} for c in calls]},
                      open(args.kernel_table, "w"), indent=1)

    # ---- secondary result: the same step in fp32 storage (the reference's arithmetic; rtol 1e-3 parity mode) ------
    fp32 = None
    if world == 1 and args.dtype == "bf16" and not args.no_fp32:
        try:
            del eager
        except NameError:
            pass
        gc.collect()
        torch.cuda.empty_cache()
        torch.manual_seed(2)
        m32 = cls(3, 1, 32, compute_dtype=torch.float32).to(dev).train()
        m32.last_activation = None
        s32 = TrainStep(m32, lr=1e-3, graph=bool(args.graph))
        for _ in range(W + 3):
            l32 = s32(x_dev, m_dev)
        torch.cuda.synchronize()
        k32 = max(1, min(K, 5))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k32):
            l32 = s32(x_dev, m_dev)
        f1.record()
        torch.cuda.synchronize()
        ms32 = f0.elapsed_time(f1)
        fp32 = {"value": B * k32 / (ms32 / 1e3), "unit": UNIT, "ms_per_step": ms32 / k32, "steps": k32, "dtype": "f32",
                "loss": float(l32), "note": "same workload, fp32 storage + fp32 contractions (inputs resident)"}
        s32.graph = None
        del s32, m32
        gc.collect()
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ips, sp = cpu_oracle_images_per_s(2, 1, cores)
        cpu = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{CPU_SAMPLE_BATCH}x3x{HW}x{HW} fp32 train step of the CPU oracle, 2 timed steps after 1 warm-up "
                         f"({sp:.1f} s/step)"}

    if rank == 0:
        out = {
            "metric": METRIC if hw == HW else f"ACC-UNet train images/sec @{hw}^2", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak",
            "vs_baseline": None,
            "dtype": args.dtype if args.dtype == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(cls.__name__, B, hw, world, args.graph),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": x_host.numel() * 4 + m_host.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * K,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "loss": float(loss_host),
            "fp32": fp32, "peak_mem_gib": peak_mem_gib,
        }
        emit(json.dumps(out))
    if world > 1:
        # Leave without tearing NCCL down: destroying a communicator whose collectives were captured into a
        # still-alive CUDA graph can block at exit.  Everything that matters (the JSON line) is already out.
        try:
            dist.barrier()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def traffic_from_profiles(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches of ONE training step
    captured with ncu (profiles/r02_launches.md lists the same capture; profiles/traffic_from_ncu.py made the file).
    It is a committed measurement of this code, not of this run: ncu cannot run inside the timed bench."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        d = json.load(open(path)).get(kernel)
        return d["dram_bytes_per_launch"] if isinstance(d, dict) else d
    return None


if __name__ == "__main__":
    main()
