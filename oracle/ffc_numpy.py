"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64 by default) of the reference's
FFC inference path, ``/root/reference/saicinpainting/training/modules/ffc.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``lama_b200``) never
does and raises if its CUDA library is missing.

Pinning: every function below is checked against outputs of the *unmodified* reference
module run in the build container (``tests/golden/*.npz`` made by
``tests/golden/make_golden.py``; ``tests/test_oracle.py``).  The reference itself ships
no tests, golden vectors or fixtures for this path, so the goldens
generated from the reference code are the only pin there is.

The arithmetic of the reference lives in PyTorch (pinned torch==1.8.x, README.md:77):
``torch.fft.rfftn/irfftn`` (ffc.py:86,108), ``nn.Conv2d`` (ffc.py:57,129,139,189-196,361),
``nn.BatchNorm2d`` eval mode (ffc.py:60,131,243-244,353), ``nn.ConvTranspose2d``
(ffc.py:350).  Their published definitions are restated here with numpy; the FFT is
``numpy.fft`` (pocketfft) plus an explicit-formula inverse used to pin the C2R rule for
non-Hermitian input.

Everything is functional: a model is a ``dict[str, np.ndarray]`` with the reference's
``state_dict`` key names plus the constructor kwargs.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5  # nn.BatchNorm2d default eps, used by every BN on the path (ffc.py:60,131,243-244,353)


# --------------------------------------------------------------------------- primitives
def _sub(sd: dict, prefix: str) -> dict:
    """Sub-dictionary of ``sd`` below ``prefix`` (keys with the prefix stripped)."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def batchnorm_eval(x, sd, prefix, eps=BN_EPS):
    """Eval-mode BatchNorm2d: ``(x - mean) / sqrt(var + eps) * gamma + beta`` per channel."""
    g = sd[prefix + "weight"].astype(x.dtype)
    b = sd[prefix + "bias"].astype(x.dtype)
    m = sd[prefix + "running_mean"].astype(x.dtype)
    v = sd[prefix + "running_var"].astype(x.dtype)
    scale = g / np.sqrt(v + eps)
    shift = b - m * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def pad2d(x, pad, mode):
    """``pad`` pixels on each side of the last two axes; mode 'reflect' (no edge repeat) or 'zeros'."""
    if pad == 0:
        return x
    widths = ((0, 0), (0, 0), (pad, pad), (pad, pad))
    if mode == "reflect":
        return np.pad(x, widths, mode="reflect")
    if mode == "zeros":
        return np.pad(x, widths, mode="constant")
    raise ValueError(mode)


def conv2d(x, w, bias=None, stride=1, padding=0, padding_mode="zeros", dilation=1):
    """Cross-correlation as ``nn.Conv2d`` defines it (groups=1).

    x: (B, Ci, H, W); w: (Co, Ci, kh, kw).  Output (B, Co, Ho, Wo).
    """
    x = pad2d(x, padding, padding_mode)
    co, ci, kh, kw = w.shape
    win = np.lib.stride_tricks.sliding_window_view(
        x, ((kh - 1) * dilation + 1, (kw - 1) * dilation + 1), axis=(2, 3))
    win = win[:, :, ::stride, ::stride, ::dilation, ::dilation]  # (B, Ci, Ho, Wo, kh, kw)
    out = np.einsum("bchwij,ocij->bohw", win, w.astype(x.dtype), optimize=True)
    if bias is not None:
        out = out + bias.astype(x.dtype)[None, :, None, None]
    return out


def conv_transpose2d(x, w, bias=None, stride=2, padding=1, output_padding=1):
    """``nn.ConvTranspose2d`` (groups=1, dilation=1).  w: (Ci, Co, kh, kw).

    Definition: out[b, o, y*s - p + i, x*s - p + j] += x[b, c, y, x] * w[c, o, i, j].
    Implemented as zero-insertion followed by a correlation with the flipped kernel.
    """
    b, ci, h, wd = x.shape
    _, co, kh, kw = w.shape
    up = np.zeros((b, ci, (h - 1) * stride + 1, (wd - 1) * stride + 1), dtype=x.dtype)
    up[:, :, ::stride, ::stride] = x
    lo_h, lo_w = kh - 1 - padding, kw - 1 - padding
    hi_h, hi_w = lo_h + output_padding, lo_w + output_padding
    up = np.pad(up, ((0, 0), (0, 0), (lo_h, hi_h), (lo_w, hi_w)))
    w_corr = np.flip(w, axis=(2, 3)).transpose(1, 0, 2, 3)  # (Co, Ci, kh, kw)
    return conv2d(up, w_corr, bias=bias)


# --------------------------------------------------------------------------- FFT pieces
def rfft2_ortho(x):
    """ffc.py:86 — ``torch.fft.rfftn(x, dim=(-2,-1), norm='ortho')``."""
    return np.fft.rfftn(x, axes=(-2, -1), norm="ortho")


def irfft2_ortho(z, h, w):
    """ffc.py:108 — ``torch.fft.irfftn(z, s=(h, w), dim=(-2,-1), norm='ortho')``."""
    return np.fft.irfftn(z, s=(h, w), axes=(-2, -1), norm="ortho")


def irfft2_explicit(z, h, w):
    """The inverse written out: complex inverse DFT along H first,
    then the real (C2R) inverse along W, which ignores Im of the k_w = 0 bin and, for even
    w, of the k_w = w/2 bin.  The spectrum after ReLU is NOT Hermitian, so this ordering is
    part of the contract the CUDA kernels must reproduce.
    """
    wf = w // 2 + 1
    assert z.shape[-2] == h and z.shape[-1] == wf
    t = np.fft.ifft(z, axis=-2, norm="ortho")  # (.., h, wf) complex, all columns
    n = np.arange(w)
    out = np.real(t[..., 0:1]) * np.ones(w)
    last = wf - 1 if w % 2 == 0 else wf
    k = np.arange(1, last)
    if k.size:
        ph = np.exp(2j * np.pi * np.outer(k, n) / w)  # (k, n)
        out = out + 2.0 * np.real(np.einsum("...k,kn->...n", t[..., 1:last], ph))
    if w % 2 == 0:
        out = out + np.real(t[..., wf - 1:wf]) * ((-1.0) ** n)
    return out / np.sqrt(w)


# --------------------------------------------------------------------------- FourierUnit
def fourier_unit(x, sd, prefix=""):
    """``FourierUnit.forward`` (ffc.py:76-113) with default options
    (groups=1, no spatial scaling / positional encoding / SE, 2-D, fft_norm='ortho').

    x: (B, c, h, w) real -> (B, c_out, h, w) real.
    """
    b, c, h, w = x.shape
    spec = rfft2_ortho(x)                                            # :86  (B,c,h,wf) complex
    # :87-89 — channels interleaved: 2k = Re(channel k), 2k+1 = Im(channel k)
    s = np.stack((spec.real, spec.imag), axis=2).reshape(b, 2 * c, h, spec.shape[-1])
    wmat = sd[prefix + "conv_layer.weight"].astype(x.dtype)[:, :, 0, 0]   # (2co, 2c)
    z = np.einsum("ok,bkhw->bohw", wmat, s, optimize=True)           # :100 1x1 conv, no bias
    z = relu(batchnorm_eval(z, sd, prefix + "bn."))                  # :101
    co = z.shape[1] // 2
    z = z.reshape(b, co, 2, h, z.shape[-1])                          # :103-105 de-interleave
    zc = z[:, :, 0] + 1j * z[:, :, 1]
    return irfft2_ortho(zc, h, w)                                    # :108


def spectral_transform(x, sd, prefix="", stride=1, enable_lfu=False):
    """``SpectralTransform.forward`` (ffc.py:142-163)."""
    if stride == 2:                                                  # :122-123 AvgPool2d(2,2)
        b, c, h, w = x.shape
        x = x[:, :, : h // 2 * 2, : w // 2 * 2].reshape(b, c, h // 2, 2, w // 2, 2).mean(axis=(3, 5))
    t = conv2d(x, sd[prefix + "conv1.0.weight"])                     # :129 1x1, no bias
    t = relu(batchnorm_eval(t, sd, prefix + "conv1.1."))             # :131-132
    out = fourier_unit(t, sd, prefix + "fu.")                        # :146
    xs = 0.0
    if enable_lfu:                                                   # :148-157
        n, c, h, w = t.shape
        s = h // 2
        q = t[:, : c // 4]
        q = np.concatenate([q[:, :, :s], q[:, :, s:2 * s]], axis=1)
        q = np.concatenate([q[:, :, :, :s], q[:, :, :, s:2 * s]], axis=1)
        q = fourier_unit(q, sd, prefix + "lfu.")
        xs = np.tile(q, (1, 1, 2, 2))
    return conv2d(t + out + xs, sd[prefix + "conv2.weight"])         # :161


# --------------------------------------------------------------------------- FFC family
def ffc(x_l, x_g, sd, prefix, *, kernel_size, stride=1, padding=0, dilation=1,
        ratio_gout, enable_lfu=False, padding_type="reflect"):
    """``FFC.forward`` (ffc.py:205-225), gated=False.  An empty side is the int 0, as in the
    reference; which branches exist is read off the state-dict keys (Identity has none)."""
    out_l, out_g = 0, 0
    kw = dict(stride=stride, padding=padding, padding_mode=padding_type, dilation=dilation)
    if ratio_gout != 1:                                              # :220-221
        out_l = conv2d(x_l, sd[prefix + "convl2l.weight"], **kw)
        if (prefix + "convg2l.weight") in sd:
            out_l = out_l + conv2d(x_g, sd[prefix + "convg2l.weight"], **kw)
    if ratio_gout != 0:                                              # :222-223
        out_g = conv2d(x_l, sd[prefix + "convl2g.weight"], **kw)
        if (prefix + "convg2g.conv2.weight") in sd:
            out_g = out_g + spectral_transform(x_g, sd, prefix + "convg2g.", stride=stride,
                                               enable_lfu=enable_lfu)
    return out_l, out_g


def ffc_bn_act(x_l, x_g, sd, prefix, *, ratio_gout, act=relu, **ffc_kw):
    """``FFC_BN_ACT.forward`` (ffc.py:251-255)."""
    o_l, o_g = ffc(x_l, x_g, sd, prefix + "ffc.", ratio_gout=ratio_gout, **ffc_kw)
    if ratio_gout != 1:
        o_l = act(batchnorm_eval(o_l, sd, prefix + "bn_l."))
    if ratio_gout != 0:
        o_g = act(batchnorm_eval(o_g, sd, prefix + "bn_g."))
    return o_l, o_g


def ffc_resnet_block(x_l, x_g, sd, prefix, *, ratio_gout=0.75, dilation=1, enable_lfu=False,
                     padding_type="reflect"):
    """``FFCResnetBlock.forward`` (ffc.py:277-292), inline=False."""
    kw = dict(kernel_size=3, padding=dilation, dilation=dilation, ratio_gout=ratio_gout,
              enable_lfu=enable_lfu, padding_type=padding_type)
    y_l, y_g = ffc_bn_act(x_l, x_g, sd, prefix + "conv1.", **kw)
    y_l, y_g = ffc_bn_act(y_l, y_g, sd, prefix + "conv2.", **kw)
    return x_l + y_l, x_g + y_g                                      # :288


def ffc_resnet_generator(x, sd, *, input_nc=4, output_nc=3, ngf=64, n_downsampling=3, n_blocks=9,
                         init_conv_kwargs=None, downsample_conv_kwargs=None, resnet_conv_kwargs=None,
                         add_out_act=True, max_features=1024, prefix="model.", return_stages=False, **_):
    """``FFCResNetGenerator.forward`` (ffc.py:306-367), default norm/activation/padding,
    no spatial-transform layers, ``out_ffc=False``.  ``x``: (B, input_nc, H, W)."""
    init_conv_kwargs = init_conv_kwargs or {}
    downsample_conv_kwargs = downsample_conv_kwargs or {}
    resnet_conv_kwargs = resnet_conv_kwargs or {}
    stages = []
    i = 0
    h = pad2d(x, 3, "reflect"); i += 1                                # :315 ReflectionPad2d(3)
    l, g = ffc_bn_act(h, 0, sd, f"{prefix}{i}.", kernel_size=7, padding=0,
                      ratio_gout=init_conv_kwargs.get("ratio_gout", 0),
                      enable_lfu=init_conv_kwargs.get("enable_lfu", True)); i += 1   # :316-317
    stages.append((l, g))
    for d in range(n_downsampling):                                  # :320-332
        kw = dict(downsample_conv_kwargs)
        if d == n_downsampling - 1:
            kw["ratio_gout"] = resnet_conv_kwargs.get("ratio_gin", 0)
        l, g = ffc_bn_act(l, g, sd, f"{prefix}{i}.", kernel_size=3, stride=2, padding=1,
                          ratio_gout=kw.get("ratio_gout", 0), enable_lfu=kw.get("enable_lfu", True)); i += 1
        stages.append((l, g))
    for _ in range(n_blocks):                                        # :338-343
        l, g = ffc_resnet_block(l, g, sd, f"{prefix}{i}.",
                                ratio_gout=resnet_conv_kwargs.get("ratio_gout", 0),
                                enable_lfu=resnet_conv_kwargs.get("enable_lfu", True)); i += 1
        stages.append((l, g))
    h = np.concatenate([l, g], axis=1) if isinstance(g, np.ndarray) else l; i += 1   # :345 ConcatTupleLayer
    for _ in range(n_downsampling):                                  # :348-354
        h = conv_transpose2d(h, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"]); i += 1
        h = relu(batchnorm_eval(h, sd, f"{prefix}{i}.")); i += 2      # BN, ReLU
        stages.append(h)
    h = pad2d(h, 3, "reflect"); i += 1                                # :360
    h = conv2d(h, sd[f"{prefix}{i}.weight"], sd[f"{prefix}{i}.bias"]); i += 1          # :361
    if add_out_act:                                                  # :362-363
        kind = "tanh" if add_out_act is True else add_out_act
        h = np.tanh(h) if kind == "tanh" else sigmoid(h)
    return (h, stages) if return_stages else h


def inpaint_forward(image, mask, sd, **gen_kwargs):
    """``DefaultInpaintingTrainingModule.forward`` glue (trainers/default.py:59,68,70,71):
    returns (predicted_image, inpainted)."""
    masked = image * (1 - mask)
    pred = ffc_resnet_generator(np.concatenate([masked, mask], axis=1), sd, **gen_kwargs)
    return pred, mask * pred + (1 - mask) * image


def state_dict_to_numpy(sd, dtype=np.float64) -> dict:
    """torch ``state_dict`` -> numpy dict (drops ``num_batches_tracked``)."""
    out = {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v).astype(dtype)
    return out
