"""Multi-GPU parity (``-m gpu``; skipped on boxes with fewer than two GPUs): BASELINE config 4's data path — rank 0 holds
the batch, NCCL scatter of the inputs, every rank runs the native generator on its shard, NCCL gather of the results —
must equal the single-GPU result bit for bit (the path shards by image: no cross-sample coupling)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LAMA_B200_STRICT="1")
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lama_b200 import modules as M
    from lama_b200 import parallel as P
    from lama_b200.testing import generator_input, seeded_parameters_, small_lama_kwargs, synthetic_image_mask
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = seeded_parameters_(M.FFCResNetGenerator(**small_lama_kwargs(ngf=8, n_blocks=2)).eval(), 3).to(dev)
    batch = 5                                   # ragged: 3 + 2
    x_full = None
    if rank == 0:
        img, mask = synthetic_image_mask(batch, 64, 7)
        x_full = generator_input(img, mask).to(dev)
    with torch.no_grad():
        y = P.sharded_apply(lambda t: g(t.contiguous()), x_full, (4, 64, 64), batch, device=dev)
        if rank == 0:
            want = g(x_full)
            q.put(bool(torch.equal(y, want)))
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_generator_gather_over_nccl_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29741, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get() is True
