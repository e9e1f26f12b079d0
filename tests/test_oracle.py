"""Pin the CPU restatements in oracle/ against the goldens generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ffc_numpy as onp
from oracle import ffc_torch_cpu as otc
from lama_b200.testing import small_lama_kwargs

FU_CASES = ["fu_c8_16x16", "fu_c4to6_8x32", "fu_c16_32x32", "fu_c4_15x15", "fu_c4_6x9", "fu_c2_20x24"]


def _f64(sd):
    return {k: v.astype(np.float64) for k, v in sd.items()}


def _t(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def _close(got, ref, rel):
    """max-abs error relative to max|ref| (the reference's own fp32 noise is ~1e-6 relative)."""
    scale = float(np.max(np.abs(ref))) or 1.0
    err = float(np.max(np.abs(np.asarray(got, dtype=np.float64) - ref.astype(np.float64))))
    assert err <= rel * scale, f"max-abs {err:.3e} > {rel:g} * {scale:.3e}"


@pytest.mark.parametrize("name", FU_CASES)
def test_fourier_unit_numpy(name):
    a, sd = load_golden(name)
    _close(onp.fourier_unit(a["x"].astype(np.float64), _f64(sd)), a["y"], 2e-6)


@pytest.mark.parametrize("name", FU_CASES)
def test_fourier_unit_torch_port(name):
    a, sd = load_golden(name)
    _close(otc.fourier_unit(torch.from_numpy(a["x"]), _t(sd)).numpy(), a["y"], 1e-6)


@pytest.mark.parametrize("h,w", [(8, 8), (6, 9), (15, 15), (4, 10), (64, 64)])
def test_c2r_rule_on_non_hermitian_input(h, w):
    """The explicit inverse (H first, then C2R dropping Im of k_w=0 and k_w=w/2) equals
    numpy's irfftn AND torch's irfftn on a spectrum that is not Hermitian (post-ReLU case)."""
    rng = np.random.default_rng(h * 100 + w)
    wf = w // 2 + 1
    z = rng.standard_normal((2, 3, h, wf)) + 1j * rng.standard_normal((2, 3, h, wf))
    z = np.maximum(z.real, 0) + 1j * np.maximum(z.imag, 0)
    want = torch.fft.irfftn(torch.from_numpy(z), s=(h, w), dim=(-2, -1), norm="ortho").numpy()
    np.testing.assert_allclose(onp.irfft2_explicit(z, h, w), want, atol=1e-12)
    np.testing.assert_allclose(onp.irfft2_ortho(z, h, w), want, atol=1e-12)


def test_fft_known_answers():
    """Known answers independent of any FFT library: impulse -> flat spectrum, Parseval, round trip."""
    x = np.zeros((1, 1, 8, 8)); x[0, 0, 0, 0] = 1.0
    s = onp.rfft2_ortho(x)
    np.testing.assert_allclose(s, np.full((1, 1, 8, 5), 1.0 / 8.0), atol=1e-15)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 12, 10))
    full = np.fft.fftn(x, axes=(-2, -1), norm="ortho")
    np.testing.assert_allclose(np.sum(np.abs(full) ** 2), np.sum(x ** 2), rtol=1e-12)
    np.testing.assert_allclose(onp.irfft2_explicit(onp.rfft2_ortho(x), 12, 10), x, atol=1e-12)


@pytest.mark.parametrize("name,stride,lfu", [("st_16to24_8x8", 1, False), ("st_16to16_s2_16x16", 2, False),
                                              ("st_16to16_lfu_8x8", 1, True), ("st_32to32_lfu_8x8", 1, True),
                                              ("st_32to32_s2_lfu_16x16", 2, True), ("st_32to64_s2_12x20", 2, False)])
def test_spectral_transform(name, stride, lfu):
    a, sd = load_golden(name)
    _close(onp.spectral_transform(a["x"].astype(np.float64), _f64(sd), stride=stride, enable_lfu=lfu), a["y"], 2e-6)
    _close(otc.spectral_transform(torch.from_numpy(a["x"]), _t(sd), stride=stride, enable_lfu=lfu).numpy(),
           a["y"], 1e-6)


def test_conv_primitives_against_torch():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 5, 9, 11)); w = rng.standard_normal((7, 5, 3, 3)); b = rng.standard_normal(7)
    xt, wt, bt = map(torch.from_numpy, (x, w, b))
    want = torch.nn.functional.conv2d(torch.nn.functional.pad(xt, (1, 1, 1, 1), mode="reflect"), wt, bt, stride=2)
    np.testing.assert_allclose(onp.conv2d(x, w, b, stride=2, padding=1, padding_mode="reflect"), want.numpy(),
                               atol=1e-12)
    wt_t = torch.from_numpy(rng.standard_normal((5, 4, 3, 3)))
    want = torch.nn.functional.conv_transpose2d(xt, wt_t, stride=2, padding=1, output_padding=1)
    np.testing.assert_allclose(onp.conv_transpose2d(x, wt_t.numpy()), want.numpy(), atol=1e-12)


@pytest.mark.parametrize("name,kw,has_g", [
    ("ffcbnact_32_k3_075", dict(kernel_size=3, padding=1, ratio_gout=0.75), True),
    ("ffcbnact_4to8_k7_local", dict(kernel_size=7, padding=0, ratio_gout=0), False),
    ("ffcbnact_16to32_s2_to_global", dict(kernel_size=3, stride=2, padding=1, ratio_gout=0.75), False),
])
def test_ffc_bn_act(name, kw, has_g):
    a, sd = load_golden(name)
    xg = a["x_g"].astype(np.float64) if has_g else 0
    yl, yg = onp.ffc_bn_act(a["x_l"].astype(np.float64), xg, _f64(sd), "", **kw)
    _close(yl, a["y_l"], 2e-6)
    if "y_g" in a:
        _close(yg, a["y_g"], 2e-6)
    tkw = {k: v for k, v in kw.items() if k != "kernel_size"}
    tl, tg = otc.ffc_bn_act(torch.from_numpy(a["x_l"]), torch.from_numpy(a["x_g"]) if has_g else 0, _t(sd), "", **tkw)
    _close(tl.numpy(), a["y_l"], 2e-6)
    if "y_g" in a:
        _close(tg.numpy(), a["y_g"], 2e-6)


def test_resnet_block():
    a, sd = load_golden("resblock_32_16x16")
    yl, yg = onp.ffc_resnet_block(a["x_l"].astype(np.float64), a["x_g"].astype(np.float64), _f64(sd), "")
    _close(yl, a["y_l"], 2e-6); _close(yg, a["y_g"], 2e-6)
    tl, tg = otc.ffc_resnet_block(torch.from_numpy(a["x_l"]), torch.from_numpy(a["x_g"]), _t(sd), "")
    _close(tl.numpy(), a["y_l"], 2e-6); _close(tg.numpy(), a["y_g"], 2e-6)


@pytest.mark.parametrize("name", ["generator_ngf8_b2_64x64", "generator_ngf8_b2_40x72"])
def test_generator(name):
    a, _ = load_golden(name)
    _, sd = load_golden("generator_ngf8_b2_64x64")
    kw = small_lama_kwargs(ngf=8, n_blocks=2)
    y = onp.ffc_resnet_generator(a["x"].astype(np.float64), _f64(sd), **kw)
    assert np.max(np.abs(y - a["y"])) < 2e-6          # sigmoid output, absolute
    yt = otc.ffc_resnet_generator(torch.from_numpy(a["x"]), _t(sd), **kw).numpy()
    assert np.max(np.abs(yt - a["y"])) < 2e-6


def test_inpaint_glue():
    a, sd = load_golden("generator_ngf8_b2_64x64")
    kw = small_lama_kwargs(ngf=8, n_blocks=2)
    pred, inp = onp.inpaint_forward(a["image"].astype(np.float64), a["mask"].astype(np.float64), _f64(sd), **kw)
    assert np.max(np.abs(pred - a["y"])) < 2e-6
    m = a["mask"]
    np.testing.assert_allclose(inp, m * pred + (1 - m) * a["image"], atol=1e-12)


def _u8_agreement(got, ref, masks):
    """Outside the hole the bytes are an exact function of the input bytes; inside, a 1e-6 difference in the
    prediction may flip a truncation, so allow one level there (and require that to be rare)."""
    hole = masks > 0
    assert np.array_equal(got[~hole], ref[~hole])
    d = np.abs(got[hole].astype(np.int16) - ref[hole].astype(np.int16))
    assert d.max(initial=0) <= 1 and (d != 0).mean() < 0.01, (int(d.max(initial=0)), float((d != 0).mean()))


def test_predict_u8_glue_pinned_on_reference_dataset_and_generator():
    """oracle/predict_numpy.py against the fixture made by the reference's InpaintingDataset
    (PNG decode, /255, symmetric pad to modulo 8) + reference generator + the restated blend / x255 / uint8."""
    from oracle import predict_numpy as opn
    a, _ = load_golden("predict_ngf8_3x45x52")
    _, sd = load_golden("generator_ngf8_b2_64x64")
    kw = small_lama_kwargs(ngf=8, n_blocks=2)
    x, img, mask = opn.generator_input(a["images"], a["masks"], pad_mod=8)
    assert x.dtype == np.float32 and np.array_equal(x, a["x"])                    # bit-exact front end
    assert np.array_equal(opn.finish(a["predicted"], img, mask, 45, 52), a["out"])  # bit-exact back end
    gen = lambda t: otc.ffc_resnet_generator(torch.from_numpy(t), _t(sd), **kw).numpy()  # noqa: E731
    _u8_agreement(opn.predict_u8(gen, a["images"], a["masks"]), a["out"], a["masks"])
    # outside the hole the float32 round trip u8 -> /255 -> *255 -> astype(uint8) is the identity for all 256 values
    keep = a["masks"] == 0
    assert np.array_equal(a["out"][keep], a["images"][keep])
    v = np.arange(256, dtype=np.uint8)
    assert np.array_equal(np.clip((v.astype("float32") / 255) * 255, 0, 255).astype("uint8"), v)
