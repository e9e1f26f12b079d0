#!/usr/bin/env python
"""dL/dx of one FFCResnetBlock: native forward+backward program vs torch autograd through the torch composition of the
same module on the same GPU (fp32, TF32 off).  Prints relative 2-norm errors per shape."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lama_b200 import modules as M  # noqa: E402
from lama_b200.testing import seeded_parameters_  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
out = {}
for (b, h, w) in ((1, 64, 64), (1, 128, 128), (1, 17, 25), (1, 96, 128)):
    blk = seeded_parameters_(M.FFCResnetBlock(512, padding_type="reflect", norm_layer=torch.nn.BatchNorm2d,
                                              activation_layer=torch.nn.ReLU, ratio_gin=0.75, ratio_gout=0.75,
                                              enable_lfu=False).eval(), 4, gain=1.0).to(dev)
    for p in blk.parameters():
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(h)
    xl, xg = torch.randn(b, 128, h, w, generator=g).to(dev), torch.randn(b, 384, h, w, generator=g).to(dev)
    gl, gg = torch.randn(b, 128, h, w, generator=g).to(dev), torch.randn(b, 384, h, w, generator=g).to(dev)
    res = {}
    for mode in ("1", "0"):
        os.environ["LAMA_B200_NATIVE_GRAD"] = mode
        os.environ["LAMA_B200_STRICT"] = "0"
        a, c = xl.clone().requires_grad_(True), xg.clone().requires_grad_(True)
        o_l, o_g = blk((a, c))
        ((o_l * gl).sum() + (o_g * gg).sum()).backward()
        res[mode] = (o_l.detach(), o_g.detach(), a.grad.clone(), c.grad.clone())
    rel = lambda x, y: float((x - y).norm() / y.norm())  # noqa: E731
    out[f"{h}x{w}"] = {"fwd_l": rel(res["1"][0], res["0"][0]), "fwd_g": rel(res["1"][1], res["0"][1]),
                       "dx_l": rel(res["1"][2], res["0"][2]), "dx_g": rel(res["1"][3], res["0"][3])}
print(json.dumps(out))
