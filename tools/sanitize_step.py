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
