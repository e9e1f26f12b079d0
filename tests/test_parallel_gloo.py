"""world_size-2 gloo tests (CPU) of the batch scatter / shard / gather plumbing and of bench.py's
rank handling for the reference arm."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from lama_b200 import parallel as P
    full = torch.arange(batch * 4 * 2 * 2, dtype=torch.float32).reshape(batch, 4, 2, 2) if rank == 0 else None
    local = P.scatter_batch(full, (4, 2, 2), batch)
    s, e = P.shard_bounds(batch, world)[rank]
    assert local.shape[0] == e - s
    # a per-sample function (no cross-sample coupling), like the generator
    out = P.gather_batch(local[:, :3] * 2 + 1, batch)
    if rank == 0:
        q.put(torch.equal(out, full[:, :3] * 2 + 1))
    dist.barrier()
    dist.destroy_process_group()


def _run(batch, world=2, port=29731):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get() is True


def test_scatter_gather_even():
    _run(batch=8, port=29731)


def test_scatter_gather_ragged_and_tiny():
    _run(batch=5, port=29732)     # 3 + 2
    _run(batch=1, port=29733)     # rank 1 gets an empty shard


def test_shard_bounds_cover_batch():
    from lama_b200.parallel import shard_bounds
    for b in (0, 1, 7, 32, 64):
        for w in (1, 2, 4, 8):
            bounds = shard_bounds(b, w)
            assert bounds[0][0] == 0 and bounds[-1][1] == b
            assert all(x[1] == y[0] for x, y in zip(bounds, bounds[1:]))
            assert max(e - s for s, e in bounds) - min(e - s for s, e in bounds) <= 1


def test_bench_reference_arm_rank_handling():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the JSON line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", LOCAL_RANK="1", RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
