"""The drop-in installs under the reference's module name and the reference's own factory builds it."""
import sys

import pytest
import torch


def test_install_registers_and_rebinds():
    import lama_b200.patch as patch
    from lama_b200 import modules as M
    patch.install()
    try:
        assert patch.installed()
        import importlib
        mod = importlib.import_module("saicinpainting.training.modules.ffc") if "saicinpainting" in sys.modules \
            else sys.modules["saicinpainting.training.modules.ffc"]
        assert mod.FFCResNetGenerator is M.FFCResNetGenerator and mod.FFCResnetBlock is M.FFCResnetBlock
    finally:
        patch.uninstall()
    assert not patch.installed()


def test_reference_factory_builds_the_drop_in():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    # run in a fresh interpreter: the ordering (install BEFORE the package import) is the point
    import subprocess, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.append(%r)\n"
        "from oracle import ref_import; ref_import._install_stubs()\n"
        "import lama_b200.patch as p; p.install()\n"
        "from saicinpainting.training.modules import make_generator\n"
        "from lama_b200.testing import BIG_LAMA_KWARGS\n"
        "from lama_b200 import modules as M\n"
        "g = make_generator(None, kind='ffc_resnet', **BIG_LAMA_KWARGS)\n"
        "assert type(g) is M.FFCResNetGenerator, type(g)\n"
        "from saicinpainting.training.modules.pix2pixhd import FFCResnetBlock\n"
        "assert FFCResnetBlock is M.FFCResnetBlock\n"
        "print('ok')\n" % (root, ref_import.REFERENCE_ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_to_jit_style_trace_of_the_drop_in_generator(tmp_path):
    """bin/to_jit.py:14-25,49-72: trace a wrapper around the model with a 120x120 image (15x15 bottleneck planes),
    save, reload, compare.  The drop-in modules must stay traceable (they run the torch operator sequence while
    tracing; a native ctypes call would be invisible to the tracer)."""
    import torch
    from lama_b200 import modules as M
    from lama_b200.testing import seeded_parameters_, small_lama_kwargs

    class JITWrapper(torch.nn.Module):            # to_jit.py:14-25 + trainers/default.py:59-71
        def __init__(self, generator):
            super().__init__()
            self.generator = generator

        def forward(self, image, mask):
            masked = torch.cat([image * (1 - mask), mask], dim=1)
            return mask * self.generator(masked) + (1 - mask) * image

    g = seeded_parameters_(M.FFCResNetGenerator(**small_lama_kwargs(ngf=8, n_blocks=2)).eval(), seed=1)
    w = JITWrapper(g).eval()
    gen = torch.Generator().manual_seed(0)
    image, mask = torch.rand(1, 3, 120, 120, generator=gen), (torch.rand(1, 1, 120, 120, generator=gen) > 0.7).float()
    with torch.no_grad():
        out = w(image, mask)
        traced = torch.jit.trace(w, (image, mask), strict=False)
    path = str(tmp_path / "lama.pt")
    traced.save(path)
    with torch.no_grad():
        jit_out = torch.jit.load(path)(image, mask)
    assert float((out - jit_out).abs().max()) < 1e-6
