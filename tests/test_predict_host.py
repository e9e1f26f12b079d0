"""Host logic of lama_b200.predict that needs no GPU: batching plan, dataset discovery,
config / checkpoint loading, and the "no CPU path" rule."""
import os

import numpy as np
import pytest
import torch

from lama_b200 import modules as M
from lama_b200 import predict as PR
from lama_b200.testing import seeded_parameters_, small_lama_kwargs


def test_plan_groups_by_size_and_splits_batches():
    sizes = [(40, 56), (32, 32), (40, 56), (40, 56), (32, 32), (40, 56), (8, 8)]
    plan = PR.BatchedInpainter.plan(sizes, max_batch=3)
    assert plan == [((40, 56), [0, 2, 3]), ((40, 56), [5]), ((32, 32), [1, 4]), ((8, 8), [6])]
    covered = sorted(i for _, idx in plan for i in idx)
    assert covered == list(range(len(sizes)))
    assert PR.BatchedInpainter.plan([], 4) == []


def test_list_dataset_follows_reference_naming(tmp_path):
    """evaluation/data.py:57-61: masks are **/*mask*.png, the image is the mask name cut at '_mask' + suffix."""
    (tmp_path / "sub").mkdir()
    for n in ["a.png", "a_mask.png", "sub/b.png", "sub/b_mask001.png", "c.jpg"]:
        (tmp_path / n).write_bytes(b"")
    pairs = PR.list_dataset(str(tmp_path) + "/", ".png")
    rel = [(os.path.relpath(i, tmp_path), os.path.relpath(m, tmp_path)) for i, m in pairs]
    assert rel == [("a.png", "a_mask.png"), ("sub/b.png", "sub/b_mask001.png")]


def test_generator_kwargs_and_checkpoint_loading(tmp_path):
    import yaml
    kw = small_lama_kwargs(ngf=8, n_blocks=2)
    cfg = {"generator": dict(kind="ffc_resnet", **kw), "discriminator": {"kind": "pix2pixhd_nlayer"}}
    assert PR.generator_kwargs_from_config(cfg) == kw
    with pytest.raises(ValueError):
        PR.generator_kwargs_from_config({"generator": {"kind": "pix2pixhd_global"}})
    g = seeded_parameters_(M.FFCResNetGenerator(**kw).eval(), seed=3)
    (tmp_path / "models").mkdir()
    with open(tmp_path / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    ckpt = {"state_dict": {**{"generator." + k: v for k, v in g.state_dict().items()},
                           "discriminator.model0.0.weight": torch.zeros(1)}}
    torch.save(ckpt, tmp_path / "models" / "best.ckpt")
    g2 = PR.load_generator(str(tmp_path), device="cpu")
    assert not g2.training
    for (k, a), (_, b) in zip(g.state_dict().items(), g2.state_dict().items()):
        assert torch.equal(a, b), k


def test_no_cpu_path():
    g = M.FFCResNetGenerator(**small_lama_kwargs(ngf=8, n_blocks=1)).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        PR.BatchedInpainter(g)


def test_inpaint_validates_items():
    inp = PR.BatchedInpainter.__new__(PR.BatchedInpainter)      # validation happens before any device work
    with pytest.raises(ValueError):
        inp.inpaint([(np.zeros((8, 8, 3), np.float32), np.zeros((8, 8), np.uint8))])
    with pytest.raises(ValueError):
        inp.inpaint([(np.zeros((8, 8, 3), np.uint8), np.zeros((8, 9), np.uint8))])


def test_shard_pairs_partition_the_directory():
    pairs = [(f"i{k}.png", f"i{k}_mask.png") for k in range(11)]
    for world in (1, 2, 3, 8, 16):
        parts = [PR.shard_pairs(pairs, r, world) for r in range(world)]
        assert sorted(p for part in parts for p in part) == sorted(pairs)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
