#!/usr/bin/env python
"""BASELINE config 4 on hardware: big-lama generator, global batch 64 at 1024x1024 held by rank 0,
batch-sharded over the N GPUs of one box — NCCL scatter of the (64,4,1024,1024) float input (1.07 GB), per-rank
generator step (CUDA-graph replay of the native program), NCCL gather of the (64,3,1024,1024) result (0.81 GB).

    python tools/config4_bench.py --gpus 1                                  (single GPU reference point, no collective)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/config4_bench.py --gpus N [--batch 64] [--size 1024] [--steps 3]

Prints one JSON line (rank 0): images/s with the collectives inside the timed region, compute-only images/s, the
scatter / gather times alone; device-timed (CUDA events), max over ranks.  `--per-gpu-batch b` switches to the
equal-per-GPU-batch reading (global batch = b * N)."""
import argparse
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--per-gpu-batch", type=int, default=0)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["LAMA_B200_STRICT"] = "1"
    from lama_b200 import _lib as L
    from lama_b200 import engine as E
    from lama_b200 import modules as M
    from lama_b200 import parallel as P
    from lama_b200.testing import BIG_LAMA_KWARGS, seeded_parameters_, synthetic_image_mask, generator_input

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev,
                                timeout=datetime.timedelta(seconds=300))
    B = args.per_gpu_batch * world if args.per_gpu_batch else args.batch
    S = args.size
    s0, s1 = P.shard_bounds(B, world)[rank]
    nb = s1 - s0
    gen = seeded_parameters_(M.FFCResNetGenerator(**BIG_LAMA_KWARGS).eval(), 0).to(dev)
    ex = E.get_executor(gen, "generator", (torch.empty(nb, 4, S, S, device="meta"),), math=L.MATH_BF16X3, device=dev)
    graphed = E.GraphedProgram(ex, warmup=1)
    x_full = None
    if rank == 0:
        # the whole batch lives on rank 0's GPU (generated in slices: the CPU generator is slow at this size)
        x_full = torch.empty(B, 4, S, S, device=dev)
        for i in range(0, B, 8):
            img, mask = synthetic_image_mask(min(8, B - i), S, seed=i)
            x_full[i:i + img.shape[0]].copy_(generator_input(img, mask))
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    out_full = [None]

    def scatter():
        if world == 1:
            graphed.static_in["x0"].copy_(x_full)
        else:
            graphed.static_in["x0"].copy_(P.scatter_batch(x_full, (4, S, S), B, device=dev))

    def compute():
        graphed.graph.replay()

    def gather():
        y = ex.outputs["y0"]
        out_full[0] = y if world == 1 else P.gather_batch(y, B)

    def step():
        scatter(); compute(); gather()

    for _ in range(args.warmup):
        step()
    ms_step = timed(step, args.steps)
    ms_compute = timed(compute, args.steps)
    ms_scatter = timed(scatter, args.steps)
    ms_gather = timed(gather, args.steps)
    ok = None
    if rank == 0:
        y = out_full[0]
        ok = bool(torch.isfinite(y).all()) and tuple(y.shape) == (B, 3, S, S)
        print(json.dumps({
            "config": "BASELINE config 4: big-lama generator, rank 0 holds the batch, NCCL scatter + gather timed",
            "n_gpus": world, "global_batch": B, "per_gpu_batch": nb, "size": S, "steps": args.steps,
            "images_per_s_with_collectives": B / (ms_step / 1e3), "ms_per_step": ms_step,
            "images_per_s_compute_only": B / (ms_compute / 1e3), "ms_compute": ms_compute,
            "ms_scatter": ms_scatter, "ms_gather": ms_gather,
            "scatter_bytes": B * 4 * S * S * 4, "gather_bytes": B * 3 * S * S * 4,
            "scaling": "weak (equal per-GPU batch)" if args.per_gpu_batch else "strong (fixed global batch)",
            "launches_per_step": ex.launches_per_run, "result_ok": ok,
            "activation_storage_gb_per_gpu": round(ex.storage_bytes / 1e9, 2),
            "torch_peak_allocated_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2),
            "timer": "CUDA events on the launch stream, barrier + synchronize both sides, max over ranks",
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
