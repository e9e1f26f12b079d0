#!/usr/bin/env python
"""Row f3 timing: lama_b200.refine.refine_predict (evaluation/refinement.py:228-314) on one image with the big-lama
generator, residual blocks on the native forward + input-gradient programs vs the same loop with the blocks under
torch autograd (cuFFT / cuDNN, TF32 on as torch defaults).  Prints one JSON line.

    python tools/refine_bench.py [--size 1024] [--iters 15]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=15)
    args = ap.parse_args()
    from lama_b200 import modules as M, refine as R
    from lama_b200.testing import BIG_LAMA_KWARGS, seeded_parameters_
    dev = torch.device("cuda:0")
    gen = seeded_parameters_(M.FFCResNetGenerator(**BIG_LAMA_KWARGS).eval(), 0).to(dev)
    g = torch.Generator().manual_seed(0)
    S = args.size
    img = torch.rand(1, 3, S, S, generator=g)
    mask = torch.zeros(1, 1, S, S)
    mask[..., S // 4: S // 2, S // 3: 2 * S // 3] = 1
    kw = dict(modulo=8, n_iters=args.iters, lr=0.002, min_side=512, max_scales=3, px_budget=1800000)
    out = {}
    res = {}
    for mode, env in (("native_block_gradients", "1"), ("torch_autograd_blocks", "0"), ("torch_autograd_blocks_fp32", "0")):
        os.environ["LAMA_B200_NATIVE_GRAD"] = env
        tf32 = mode != "torch_autograd_blocks_fp32"          # the third arm: torch without TF32 = the arithmetic yardstick
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        os.environ["LAMA_B200_STRICT"] = "0"
        R.refine_predict(img, mask, gen, **dict(kw, n_iters=2))          # warm-up: programs, cuDNN plans
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res[mode] = R.refine_predict(img, mask, gen, **kw)
        torch.cuda.synchronize()
        out[mode + "_s"] = time.perf_counter() - t0
    def diff(a, b):
        d = (res[a] - res[b]).abs()
        return {"max": float(d.max()), "mean": float(d.mean()), "mean_in_hole": float(d[mask.expand_as(d) > 0].mean())}
    out["native_vs_torch_fp32"] = diff("native_block_gradients", "torch_autograd_blocks_fp32")
    out["torch_tf32_vs_torch_fp32"] = diff("torch_autograd_blocks", "torch_autograd_blocks_fp32")
    out["native_vs_torch_tf32"] = diff("native_block_gradients", "torch_autograd_blocks")
    out["note"] = ("Adam's normalised steps turn a sign flip of a near-zero gradient into a full +-lr move of that feature "
                   "every iteration, so the two arithmetic paths drift apart element-wise (max) while agreeing on average")
    out.update(image=[S, S], n_iters=args.iters, scales="pyramid of refinement.py:176-226 (min_side 512)",
               timer="host wall clock around refine_predict incl. its final .cpu()")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
