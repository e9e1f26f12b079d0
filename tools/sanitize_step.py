#!/usr/bin/env python
"""One big-lama generator forward (B=1, 512x512, bf16x3) + smoke-sized ops, for compute-sanitizer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

os.environ["LAMA_B200_STRICT"] = "1"
from lama_b200 import modules as M  # noqa: E402
from lama_b200.testing import BIG_LAMA_KWARGS, seeded_parameters_, synthetic_image_mask, generator_input  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

for math in ("bf16x3", "fp32"):
    os.environ["LAMA_B200_MATH"] = math
    gen = seeded_parameters_(M.FFCResNetGenerator(**BIG_LAMA_KWARGS).eval(), 0).to("cuda:0")
    size = 512 if math == "bf16x3" else 128
    img, mask = synthetic_image_mask(1, size, 0)
    with torch.no_grad():
        y = gen(generator_input(img, mask).to("cuda:0"))
    torch.cuda.synchronize()
    print(math, size, float(y.mean()), bool(torch.isfinite(y).all()))
    del gen
os.environ["LAMA_B200_MATH"] = "bf16x3"
ge.smoke()

# uint8 predict path (row f1) on a size whose planes have no compile-time FFT plan (row f2: 200x120 -> 25x15)
import numpy as np  # noqa: E402
from lama_b200.predict import BatchedInpainter  # noqa: E402
from lama_b200.testing import small_lama_kwargs  # noqa: E402

small = seeded_parameters_(M.FFCResNetGenerator(**small_lama_kwargs(ngf=8, n_blocks=2)).eval(), 1).to("cuda:0")
rng = np.random.default_rng(0)
imgs = rng.integers(0, 256, size=(3, 197, 118, 3), dtype=np.uint8)
msks = (rng.random((3, 197, 118)) < 0.3).astype(np.uint8) * 255
out = BatchedInpainter(small, max_batch=2)(imgs, msks)
print("u8 predict", out.shape, out.dtype, int(out.mean()))
