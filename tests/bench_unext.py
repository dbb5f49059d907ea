"""Micro-benchmark (not a test): BASELINE.json configs[2] -- UNeXt fwd+bwd on 8x3x256x256 with the shifted tokenized-MLP
blocks on the accx kernels (accx.unext), CUDA-event timed, next to the same blocks as plain torch operators on the same
GPU (the oracle's restatement of Experiments/nets/UNext.py:72-147 run on cuda tensors: pad / chunk / roll / cat / narrow,
F.linear, grouped conv, GELU, LayerNorm).

    python tests/bench_unext.py [--dtype bf16|fp32] [--batch 8] [--hw 256] [--steps 20]
prints one JSON line: whole-model images/s, and per-block (tokens [8,256,160], [8,64,256], [8,1024,128]) fwd+bwd times."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "acc-unet-unext_b200")]


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import accx
    import accx.unext as U
    from accx import engine as E
    from oracle import acc_oracle as O
    accx.load_library()
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    B, hw = args.batch, args.hw
    out = {"workload": f"UNeXt fwd+bwd {B}x3x{hw}x{hw}, shiftedBlocks on accx kernels ({args.dtype} tokens), stem/decoder torch",
           "blocks": []}

    # ---- the four shiftedBlocks alone, accx vs torch operators on the same GPU --------------------------------
    for C, H in ((160, hw // 16), (256, hw // 32), (128, hw // 8)):
        N = H * H
        blk = U.shiftedBlock(dim=C, num_heads=1, mlp_ratio=1).to(dev).train()
        x = torch.randn(B, N, C, device=dev, dtype=dt, requires_grad=True)
        cot = torch.randn(B, N, C, device=dev, dtype=dt)

        def run_accx():
            blk.zero_grad(set_to_none=True)
            x.grad = None
            blk(x, H, H).backward(cot)

        l0 = E.LAUNCHES
        run_accx()
        launches = E.LAUNCHES - l0
        ms_a = timed(run_accx, args.steps)
        sd = {"." + k: v.detach().clone().to(dt).requires_grad_(True) for k, v in blk.state_dict().items()}

        def run_torch():
            for v in sd.values():
                v.grad = None
            x.grad = None
            O.shifted_block(O.Ctx(sd, True), "", x, H, H).backward(cot)

        ms_t = timed(run_torch, args.steps)

        def graphed(fn):
            """the same fwd+bwd captured in a CUDA graph: GPU time without the launching CPU in the way"""
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                for _ in range(3):
                    fn()
            torch.cuda.current_stream().wait_stream(s_)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return timed(g.replay, args.steps)

        ms_ag, ms_tg = graphed(run_accx), graphed(run_torch)
        out["blocks"].append({"tokens": [B, N, C], "accx_ms": ms_a, "torch_ops_ms": ms_t, "accx_graph_ms": ms_ag,
                              "torch_ops_graph_ms": ms_tg, "accx_launches": launches})

    # ---- whole model ------------------------------------------------------------------------------------------
    torch.manual_seed(2)
    model = U.UNext(3, 1, img_size=hw, compute_dtype=dt).to(dev).train()
    xi = torch.randn(B, 3, hw, hw, device=dev)
    mk = (torch.rand(B, 1, hw, hw, device=dev) > 0.5).float()

    def step():
        model.zero_grad(set_to_none=True)
        y = model(xi)
        torch.nn.functional.binary_cross_entropy(y.float(), mk).backward()

    l0 = E.LAUNCHES
    step()
    out["accx_launches_per_step"] = E.LAUNCHES - l0
    ms = timed(step, args.steps)
    out.update({"ms_per_step": ms, "images_per_s": B / ms * 1e3, "dtype": args.dtype})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
