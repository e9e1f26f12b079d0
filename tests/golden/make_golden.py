"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For each case the reference module (``saicinpainting/training/modules/ffc.py``, loaded by
``oracle/ref_import.py``) is constructed, its parameters are overwritten by the seeded
factory ``lama_b200.testing.seeded_parameters_``, it is run in ``eval()`` under
``no_grad`` on CPU fp32, and input / state_dict / output are stored in one ``.npz``.
The fixtures pin (a) the numpy and torch-CPU restatements in ``oracle/`` and (b), on the
GPU box where the reference tree is absent, the CUDA path itself.

Sizes are small on purpose (the whole directory stays < 2 MB).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import load_reference_ffc  # noqa: E402
from lama_b200.testing import (seeded_parameters_, small_lama_kwargs, synthetic_image_mask,  # noqa: E402
                               generator_input)


def _sd_np(module):
    return {"sd::" + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()
            if not k.endswith("num_batches_tracked")}


def _save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name:32s} {os.path.getsize(path) / 1024:8.1f} KiB")


def _randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@torch.no_grad()
def main():
    ffc = load_reference_ffc()
    torch.set_num_threads(1)

    # ---- FourierUnit: power-of-two, rectangular, odd and non-power-of-two planes
    fu_cases = {
        "fu_c8_16x16": (2, 8, 8, 16, 16),
        "fu_c4to6_8x32": (1, 4, 6, 8, 32),
        "fu_c16_32x32": (1, 16, 16, 32, 32),
        "fu_c4_15x15": (1, 4, 4, 15, 15),      # bin/to_jit.py traces at 120x120 -> 15x15 bottleneck
        "fu_c4_6x9": (2, 4, 4, 6, 9),          # odd width: no Nyquist column
        "fu_c2_20x24": (1, 2, 2, 20, 24),      # 2^a 3^b 5^c sizes (pad_out_to_modulo: 8 images)
    }
    for i, (name, (b, ci, co, h, w)) in enumerate(fu_cases.items()):
        m = seeded_parameters_(ffc.FourierUnit(ci, co).eval(), seed=10 + i, gain=1.0)
        x = _randn((b, ci, h, w), 100 + i)
        _save(name, x=x.numpy(), y=m(x).numpy(), **_sd_np(m))

    # ---- SpectralTransform (stride 1 no LFU = big-lama; stride 2; LFU on)
    st_cases = {
        "st_16to24_8x8": dict(ci=16, co=24, stride=1, lfu=False, hw=(8, 8)),
        "st_16to16_s2_16x16": dict(ci=16, co=16, stride=2, lfu=False, hw=(16, 16)),
        "st_16to16_lfu_8x8": dict(ci=16, co=16, stride=1, lfu=True, hw=(8, 8)),
    }
    for i, (name, c) in enumerate(st_cases.items()):
        m = seeded_parameters_(ffc.SpectralTransform(c["ci"], c["co"], stride=c["stride"],
                                                     enable_lfu=c["lfu"]).eval(), seed=20 + i, gain=1.0)
        x = _randn((2, c["ci"]) + c["hw"], 200 + i)
        _save(name, x=x.numpy(), y=m(x).numpy(), **_sd_np(m))

    # ---- FFC_BN_ACT: resblock flavour (0.75 / 0.75), stem flavour (local only, k7),
    #      downsample flavour (stride 2, local -> local+global)
    def run_ffc(name, seed, ctor_kw, xl_shape, xg_shape):
        m = seeded_parameters_(ffc.FFC_BN_ACT(**ctor_kw).eval(), seed=seed, gain=1.0)
        xl = _randn(xl_shape, seed + 300)
        xg = _randn(xg_shape, seed + 301) if xg_shape else 0
        yl, yg = m((xl, xg) if xg_shape else xl)
        arrays = dict(x_l=xl.numpy(), y_l=yl.numpy())
        if xg_shape:
            arrays["x_g"] = xg.numpy()
        if torch.is_tensor(yg):
            arrays["y_g"] = yg.numpy()
        _save(name, **arrays, **_sd_np(m))

    relu = torch.nn.ReLU
    run_ffc("ffcbnact_32_k3_075", 30,
            dict(in_channels=32, out_channels=32, kernel_size=3, ratio_gin=0.75, ratio_gout=0.75, padding=1,
                 activation_layer=relu, enable_lfu=False), (2, 8, 8, 8), (2, 24, 8, 8))
    run_ffc("ffcbnact_4to8_k7_local", 31,
            dict(in_channels=4, out_channels=8, kernel_size=7, ratio_gin=0, ratio_gout=0, padding=0,
                 activation_layer=relu, enable_lfu=False), (1, 4, 22, 22), None)
    run_ffc("ffcbnact_16to32_s2_to_global", 32,
            dict(in_channels=16, out_channels=32, kernel_size=3, ratio_gin=0, ratio_gout=0.75, stride=2,
                 padding=1, activation_layer=relu, enable_lfu=False), (2, 16, 16, 16), None)

    # ---- FFCResnetBlock (big-lama flavour, 32 channels)
    m = seeded_parameters_(
        ffc.FFCResnetBlock(32, padding_type="reflect", norm_layer=torch.nn.BatchNorm2d, activation_layer=relu,
                           ratio_gin=0.75, ratio_gout=0.75, enable_lfu=False).eval(), seed=40)
    xl, xg = _randn((2, 8, 16, 16), 400), _randn((2, 24, 16, 16), 401)
    yl, yg = m((xl, xg))
    _save("resblock_32_16x16", x_l=xl.numpy(), x_g=xg.numpy(), y_l=yl.numpy(), y_g=yg.numpy(), **_sd_np(m))

    # ---- small generator with the big-lama topology (ngf 8 -> 16+48 bottleneck channels, 2 blocks)
    kw = small_lama_kwargs(ngf=8, n_blocks=2)
    g = seeded_parameters_(ffc.FFCResNetGenerator(**kw).eval(), seed=50, gain=1.0)
    img, mask = synthetic_image_mask(2, 64, seed=5)
    x = generator_input(img, mask)
    _save("generator_ngf8_b2_64x64", image=img.numpy(), mask=mask.numpy(), x=x.numpy(), y=g(x).numpy(), **_sd_np(g))
    # rectangular, not a multiple of 64: 40 x 72 image -> 5 x 9 bottleneck
    img, mask = synthetic_image_mask(1, 40, seed=6, width=72)
    x = generator_input(img, mask)
    # same weights as generator_ngf8_b2_64x64 (state_dict stored there only)
    _save("generator_ngf8_b2_40x72", x=x.numpy(), y=g(x).numpy())


if __name__ == "__main__":
    main()
