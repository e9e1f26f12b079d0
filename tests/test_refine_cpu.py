"""CPU tests of the kornia-free refinement driver (lama_b200/refine.py): the three pyramid operators
the reference takes from kornia are pinned against OpenCV (same definitions: reflect-101 Gaussian blur, flat erosion
whose border never erodes, cv2's own ellipse), and the multi-scale loop runs end to end on a small generator."""
import cv2
import numpy as np
import pytest
import torch

from lama_b200 import modules as M
from lama_b200 import refine as R
from lama_b200.testing import seeded_parameters_, small_lama_kwargs


def test_gaussian_blur_matches_opencv_reflect101():
    rng = np.random.default_rng(0)
    x = rng.random((2, 3, 37, 45)).astype(np.float32)
    got = R.gaussian_blur2d(torch.from_numpy(x)).numpy()
    for b in range(2):
        for c in range(3):
            want = cv2.GaussianBlur(x[b, c], (5, 5), 1.0, borderType=cv2.BORDER_REFLECT_101)
            assert np.abs(got[b, c] - want).max() < 2e-6
    assert np.allclose(R.gaussian_kernel1d(5, 1.0).numpy(), cv2.getGaussianKernel(5, 1.0).ravel(), atol=1e-7)


@pytest.mark.parametrize("k", [3, 5, 15, 21])
def test_ellipse_kernel_is_opencvs(k):
    assert np.array_equal(R.ellipse_kernel(k), cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)).astype(bool))


def test_erosion_matches_opencv_default_border():
    rng = np.random.default_rng(1)
    m = (rng.random((1, 1, 64, 80)) > 0.02).astype(np.float32)
    m[0, 0, :3] = 1.0                                       # hole touching the border: the border must not erode it
    k = R.ellipse_kernel(15)
    got = R.erosion(torch.from_numpy(m), torch.from_numpy(k).float()).numpy()[0, 0]
    want = cv2.erode(m[0, 0], k.astype(np.uint8))
    assert np.array_equal(got, want)
    g = rng.random((1, 1, 20, 20)).astype(np.float32)       # grey values: a true minimum filter
    got = R.erosion(torch.from_numpy(g), torch.from_numpy(R.ellipse_kernel(5)).float()).numpy()[0, 0]
    assert np.array_equal(got, cv2.erode(g[0, 0], R.ellipse_kernel(5).astype(np.uint8)))


def test_pyramid_shapes_and_mask_binarisation():
    img = torch.rand(1, 3, 300, 420)
    mask = torch.zeros(1, 1, 300, 420); mask[..., 100:180, 150:260] = 1
    images, masks = R.image_mask_pyramid(img, mask, min_side=64, max_scales=3, px_budget=10 ** 7)
    assert [tuple(t.shape[2:]) for t in images] == [(75, 105), (150, 210), (300, 420)]
    assert all(set(np.unique(m.numpy())) <= {0.0, 1.0} for m in masks)
    images, masks = R.image_mask_pyramid(img, mask, min_side=64, max_scales=3, px_budget=300 * 420 // 4)
    assert tuple(images[-1].shape[2:]) == (150, 210)        # resized to the pixel budget first


def test_refinement_loop_runs_end_to_end_and_keeps_known_pixels():
    g = seeded_parameters_(M.FFCResNetGenerator(**small_lama_kwargs(ngf=8, n_blocks=2)).eval(), 2, gain=1.0)
    gen = torch.Generator().manual_seed(0)
    img = torch.rand(1, 3, 100, 132, generator=gen)
    mask = torch.zeros(1, 1, 100, 132); mask[..., 30:70, 40:100] = 1
    out = R.refine_predict(img, mask, g, modulo=8, n_iters=3, lr=0.002, min_side=64, max_scales=2, px_budget=10 ** 7)
    assert tuple(out.shape) == (1, 3, 100, 132) and torch.isfinite(out).all()
    keep = (mask == 0).expand_as(img)
    assert torch.equal(out[keep], img[keep])                # inpainted = mask*pred + (1-mask)*image (refinement.py:171)
    assert not torch.equal(out[~keep], img[~keep])
    assert all(not p.requires_grad for p in g.parameters())
