"""Program builder + CUDA executor behind the drop-in modules.

A *program* is a straight-line list of op records over named channels-last activation buffers
(``Buf``), built once per (module, input shape, device, math mode) from the module's parameters:
BatchNorm folded, weights packed (``lama_b200.packing``), K-segment lists laid out, buffers
wired so that the local|global halves of an FFC feature map share one allocation (split / concat
are free) and residual adds, bias, BN and activations live in GEMM epilogues.

The executor binds every op to one C-ABI call of ``libffc_b200.so`` with pre-built ctypes
descriptors; replaying a program is a loop of foreign calls on the current CUDA stream (no
allocation, no synchronisation) and is therefore CUDA-graph capturable (``GraphedProgram``).

The op records are plain data so that ``tests/spec_interp.py`` can interpret the very same
program with slow torch/numpy restatements on the CPU box — that is test infrastructure; the
product path below only ever executes through the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from . import packing as P

PROGRAM_CACHE_SIZE = 3          # executors (programs + their buffers) kept per module, LRU
MATH_ENV = "LAMA_B200_MATH"     # "bf16x3" (default: tcgen05 arm) | "fp32" (CUDA-core arm)


def default_math() -> int:
    return {"fp32": L.MATH_FP32, "bf16x3": L.MATH_BF16X3}[os.environ.get(MATH_ENV, "bf16x3").lower()]


# ------------------------------------------------------------------------------------------- IR
@dataclass
class Buf:
    """Channels-last activation buffer [B][H+2p][W+2p][C] (fp32) or [2][B][H+2p][W+2p][C] (split bf16)."""
    name: str
    B: int
    H: int
    W: int
    C: int
    pad: int = 0
    fmt: int = L.F32
    reflect_border: int = 0
    cg: int = 0        # > 0: channel-group planar storage [C/cg][B][H][W][cg] (FourierUnit chain, include/ffc_b200.h)
    tile: int = 0      # 128 (with cg == 8, split bf16): tile-blocked [pixel block of 128][C/8][128][8] — tcgen05 operand tiles


@dataclass
class TV:
    """View of a Buf: channel slice [c0, c0+C) and, for transposed-conv outputs, a sub-pixel phase."""
    buf: Buf
    c0: int = 0
    C: Optional[int] = None
    phase: Optional[Tuple[int, int]] = None   # (a, b): pixels (2i+a, 2j+b) of the buffer
    window: int = 0                           # sliding-window view: pixel x exposes pixels x..x+window-1 (C*window channels)
    b0: int = 0                               # batch slice [b0, b0+nb)
    nb: Optional[int] = None
    win: Optional[Tuple[int, int, int, int]] = None   # spatial sub-rectangle (y0, x0, h, w) of the buffer (LFU quadrants)
    bcast: int = 0                            # > 0: a one-image buffer read as a batch of `bcast` images (stride 0)

    @property
    def channels(self) -> int:
        if self.window:
            return self.buf.C * self.window
        return self.buf.C - self.c0 if self.C is None else self.C

    @property
    def batch(self) -> int:
        if self.bcast:
            return self.bcast
        return self.buf.B - self.b0 if self.nb is None else self.nb

    def bslice(self, b0: int, nb: int) -> "TV":
        if self.bcast:
            return TV(self.buf, self.c0, self.C, self.phase, self.window, 0, None, self.win, nb)
        return TV(self.buf, self.c0, self.C, self.phase, self.window, self.b0 + b0, nb, self.win)

    @property
    def hw(self) -> Tuple[int, int]:
        if self.window:
            return (self.buf.H, self.buf.W - self.window)
        if self.win is not None:
            return (self.win[2], self.win[3])
        return (self.buf.H // 2, self.buf.W // 2) if self.phase else (self.buf.H, self.buf.W)


@dataclass
class ToNHWC:
    src: str          # name of an external NCHW float tensor
    out: TV


@dataclass
class ToNCHW:
    inp: TV
    dst: str          # name of an external NCHW float output


@dataclass
class StemOp:
    src: str          # external NCHW input
    cin: int
    w: torch.Tensor   # [(ky*7+kx)*Cin + c][N]
    shift: torch.Tensor
    out: TV


@dataclass
class StemPackOp:
    """NCHW float input -> reflect-padded NHWC8 image for the tensor-core stem (ffcb_stem_pack)."""
    src: str
    cin: int
    out: TV           # Buf (B, H+6, W+8, 8)


@dataclass
class HeadOp:
    inp: TV
    w: torch.Tensor   # [N][49][C]
    bias: torch.Tensor
    n_out: int
    act: int
    dst: str


@dataclass
class HeadGatherOp:
    """y = act(bias + sum_kx q[.., reflect(x+kx-3), n*7+kx]) -> external NCHW output (ffcb_head_gather7)."""
    q: TV
    bias: torch.Tensor
    n_out: int
    act: int
    dst: str


@dataclass
class StemPackU8Op:
    """Decoded RGB bytes + mask bytes -> packed stem image (ffcb_stem_pack_u8): /255, symmetric pad to the
    modulo size, mask > 0, img * (1 - mask), cat(mask), ReflectionPad2d(3)."""
    img: str          # external uint8 (B, H0, W0, 3)
    mask: str         # external uint8 (B, H0, W0)
    h0: int
    w0: int
    out: TV           # Buf (B, H+6, W+8, 8)


@dataclass
class HeadGatherU8Op:
    """ffcb_head_gather7_blend_u8: head gather + activation + blend with the input + crop + x255/clip/truncate."""
    q: TV
    bias: torch.Tensor
    act: int
    img: str
    mask: str
    h0: int
    w0: int
    dst: str          # external uint8 (B, H0, W0, 3)


@dataclass
class ConvOp:
    packed: P.PackedConv
    ins: List[Optional[TV]]
    out: TV
    addend: Optional[TV] = None
    addend_post: bool = False
    tag: str = ""


@dataclass
class RfftOp:
    inp: TV
    spec: TV


@dataclass
class IrfftOp:
    spec: TV
    residual: Optional[TV]
    out: TV


@dataclass
class BorderOp:
    """(Re)build the reflected ring of a padded buffer after a producer that does not write it."""
    view: TV


@dataclass
class ReluBwdOp:
    """out = dy * [y > 0] (ffcb_relu_bwd): ReLU backward with the forward activation."""
    dy: TV
    y: TV
    out: TV


@dataclass
class FoldOp:
    """Adjoint of the 1-pixel reflect padding (ffcb_fold_reflect_border): gradient w.r.t. the padded plane ``gpad``
    (B,H+2,W+2,C) folded onto the interior, plus optional addends written as (view, first output channel)."""
    gpad: TV
    addends: List[Tuple[TV, int]]
    out: TV


@dataclass
class SplitOp:
    """Boundary between the forward and the backward part of a forward+backward program (no kernel)."""


@dataclass
class Program:
    kind: str
    math: int
    bufs: List[Buf] = field(default_factory=list)
    ops: list = field(default_factory=list)
    inputs: Dict[str, Tuple[int, ...]] = field(default_factory=dict)    # name -> NCHW shape
    outputs: Dict[str, Tuple[int, ...]] = field(default_factory=dict)
    dtypes: Dict[str, torch.dtype] = field(default_factory=dict)       # inputs / outputs that are not float32
    meta: Dict[tuple, dict] = field(default_factory=dict)              # buffers a backward program needs (per module)
    consts: Dict[str, torch.Tensor] = field(default_factory=dict)      # buffer name -> initial contents [B,H,W,C] float

    def buf(self, name, B, H, W, C, gemm=False, halo=False, halo_px=1, cg=0) -> Buf:
        """``gemm``: the buffer is an operand of a contraction; ``halo``: that contraction has spatial taps.
        FFCB_MATH_BF16X3 stores gemm operands as split bf16, and gives halo buffers a reflected border
        ring so that the TMA box of tap (dy,dx) is the tile shifted by (dx,dy); everything else (FFT
        inputs, spectra leaving the GEMM, the head's input) stays float32 without padding."""
        tc = self.math == L.MATH_BF16X3 and gemm
        ring = halo_px if (tc and halo) else 0
        assert not (cg and ring) and (cg == 0 or C % cg == 0)
        # channel-group planar contraction operands are tile-blocked: one contiguous 16 KB run per (M tile, K block, plane)
        tile = 128 if (cg == 8 and tc) else 0
        b = Buf(f"{name}#{len(self.bufs)}", B, H, W, C, pad=ring, fmt=L.BF16X2 if tc else L.F32,
                reflect_border=1 if ring else 0, cg=cg, tile=tile)
        self.bufs.append(b)
        return b

    def fft_workspace_bytes(self) -> int:
        need = 0
        for op in self.ops:
            if isinstance(op, RfftOp):
                b, (h, w), c = op.inp.batch, op.inp.hw, op.inp.channels
            elif isinstance(op, IrfftOp):
                b, (h, w), c = op.out.batch, op.out.hw, op.out.channels
            else:
                continue
            need = max(need, 8 * b * h * (w // 2 + 1) * c)
        return need


# ------------------------------------------------------------------------------- support predicates
def _act_code(m: nn.Module) -> Optional[int]:
    if isinstance(m, nn.ReLU):
        return L.ACT_RELU
    if isinstance(m, nn.Identity):
        return L.ACT_NONE
    if isinstance(m, nn.Sigmoid):
        return L.ACT_SIGMOID
    if isinstance(m, nn.Tanh):
        return L.ACT_TANH
    return None


def bn_foldable(bn) -> bool:
    """Eval-mode BatchNorm that can be folded into the preceding weights: it must carry running statistics
    (track_running_stats=False uses batch statistics even in eval mode — the torch composition handles that)."""
    if isinstance(bn, nn.Identity):
        return True
    return (isinstance(bn, nn.BatchNorm2d) and bn.track_running_stats and bn.running_var is not None
            and bn.running_mean is not None)


def _plain_conv(conv, k_ok=(1, 3, 7)) -> bool:
    return (isinstance(conv, nn.Conv2d) and conv.groups == 1 and conv.dilation == (1, 1) and conv.bias is None
            and conv.kernel_size[0] == conv.kernel_size[1] and conv.kernel_size[0] in k_ok
            and conv.stride[0] == conv.stride[1] and conv.stride[0] in (1, 2)
            and conv.padding[0] == conv.padding[1] and isinstance(conv.padding[0], int)
            and (conv.padding_mode == 'reflect' or conv.padding[0] == 0)
            and conv.padding[0] <= 1          # activation buffers carry a 1-pixel reflected ring
            and conv.in_channels % 4 == 0 and conv.out_channels % 4 == 0)


def fft_len_ok(n: int) -> bool:
    """Lengths the shared-memory FFT kernels take (csrc/fft.cu: make_plan / kMaxSmem): powers of two up to 256 have
    compile-time plans; any other length needs 8 * (n + 2 * n * 32) bytes of shared memory <= 227 KB, i.e. n <= 446."""
    if n >= 4 and (n & (n - 1)) == 0:
        return n <= 256
    return 1 <= n and 8 * (n + 64 * n) <= 227 * 1024


def plane_ok(h: int, w: int) -> bool:
    """Plane sizes the native FFT pair accepts (others take the torch composition)."""
    return w >= 2 and fft_len_ok(h) and fft_len_ok(w)


def ffc_bn_act_supported(m) -> bool:
    f = m.ffc
    if m.training or f.gated:
        return False
    convs = [c for c in (f.convl2l, f.convl2g, f.convg2l) if not isinstance(c, nn.Identity)]
    if not convs or not all(_plain_conv(c) for c in convs):
        return False
    c0 = convs[0]
    if any((c.kernel_size, c.stride, c.padding) != (c0.kernel_size, c0.stride, c0.padding) for c in convs):
        return False
    if c0.kernel_size[0] ** 2 + 1 > L.MAX_KSEG:
        return False
    if not isinstance(f.convg2g, nn.Identity):
        if not f.convg2g.native_supported():
            return False
        if isinstance(f.convl2g, nn.Identity):      # global-only input is never produced by the generator
            return False
    for bn in (m.bn_l, m.bn_g):
        if not bn_foldable(bn):
            return False
    return _act_code(m.act_l) is not None and _act_code(m.act_g) is not None


def ffc_bn_act_shapes_ok(m, x_l, x_g) -> bool:
    f = m.ffc
    if not torch.is_tensor(x_l) or x_l.dim() != 4:
        return False
    in_cg = f.global_in_num
    if (in_cg > 0) != torch.is_tensor(x_g):
        return False
    if torch.is_tensor(x_g) and (x_g.shape[0] != x_l.shape[0] or x_g.shape[2:] != x_l.shape[2:]):
        return False
    c = next(c for c in (f.convl2l, f.convl2g, f.convg2l) if not isinstance(c, nn.Identity))
    k, p = c.kernel_size[0], c.padding[0]
    h, w = x_l.shape[2], x_l.shape[3]
    if p > 0 and (h <= p or w <= p):     # reflect padding needs pad < size
        return False
    if h + 2 * p < k or w + 2 * p < k:
        return False
    if not isinstance(f.convg2g, nn.Identity):
        st = f.convg2g
        if not plane_ok(*st_out_hw(st, h, w)) or not st.native_supported((h, w)):   # LFU needs even square planes
            return False
        if st.stride == 2 and (c.stride[0] != 2 or ((h + 2 * p - k) // 2 + 1, (w + 2 * p - k) // 2 + 1) != (h // 2, w // 2)):
            return False
    return True


def _generator_layout(gen):
    """Parse ``gen.model`` into (stem, downs, blocks, ups, head_conv, out_act) or None."""
    from .modules import FFC_BN_ACT, FFCResnetBlock, ConcatTupleLayer
    mods = list(gen.model)
    i = 0
    try:
        if not (isinstance(mods[0], nn.ReflectionPad2d) and tuple(mods[0].padding) == (3, 3, 3, 3)):
            return None
        stem = mods[1]
        if not (isinstance(stem, FFC_BN_ACT) and isinstance(stem.ffc.convl2l, nn.Conv2d)
                and stem.ffc.convl2l.kernel_size == (7, 7) and stem.ffc.convl2l.padding == (0, 0)
                and stem.ffc.convl2l.stride == (1, 1) and isinstance(stem.ffc.convl2g, nn.Identity)
                and stem.ffc.global_in_num == 0 and isinstance(stem.bn_l, nn.BatchNorm2d) and bn_foldable(stem.bn_l)
                and isinstance(stem.act_l, nn.ReLU) and stem.ffc.convl2l.bias is None
                and stem.ffc.convl2l.groups == 1 and stem.ffc.convl2l.in_channels <= 16
                and stem.ffc.convl2l.out_channels % 4 == 0 and not stem.ffc.gated):
            return None
        i = 2
        downs = []
        while isinstance(mods[i], FFC_BN_ACT):
            if not mods[i].native_supported():
                return None
            downs.append(mods[i]); i += 1
        blocks = []
        while isinstance(mods[i], FFCResnetBlock):
            if mods[i].inline or not mods[i].native_supported():
                return None
            blocks.append(mods[i]); i += 1
        if not isinstance(mods[i], ConcatTupleLayer):
            return None
        i += 1
        ups = []
        while isinstance(mods[i], nn.ConvTranspose2d):
            ct, bn, act = mods[i], mods[i + 1], mods[i + 2]
            if not (ct.kernel_size == (3, 3) and ct.stride == (2, 2) and ct.padding == (1, 1)
                    and ct.output_padding == (1, 1) and ct.groups == 1 and ct.dilation == (1, 1)
                    and isinstance(bn, nn.BatchNorm2d) and bn_foldable(bn) and isinstance(act, nn.ReLU)
                    and ct.in_channels % 4 == 0 and ct.out_channels % 4 == 0):
                return None
            ups.append((ct, bn)); i += 3
        out_blk = None
        if isinstance(mods[i], FFCResnetBlock):          # out_ffc=True (ffc.py:356-358): an inline block at full resolution
            if not (mods[i].inline and mods[i].native_supported()):
                return None
            out_blk = mods[i]; i += 1
        if not (isinstance(mods[i], nn.ReflectionPad2d) and tuple(mods[i].padding) == (3, 3, 3, 3)):
            return None
        head = mods[i + 1]
        if not (isinstance(head, nn.Conv2d) and head.kernel_size == (7, 7) and head.padding == (0, 0)
                and head.stride == (1, 1) and head.groups == 1 and head.out_channels <= 4
                and head.in_channels % 4 == 0):
            return None
        i += 2
        out_act = L.ACT_NONE
        if i < len(mods):
            out_act = _act_code(mods[i])
            if out_act is None:
                return None
            i += 1
        if i != len(mods):
            return None
        return stem, downs, blocks, ups, out_blk, head, out_act
    except IndexError:
        return None


def generator_supported(gen, x) -> bool:
    lay = _generator_layout(gen)
    if lay is None or x.dim() != 4:
        return False
    stem, downs, _blocks, _ups, out_blk, _head, _ = lay
    b, c, h, w = x.shape
    if out_blk is not None:        # its FourierUnit transforms full-resolution planes
        f0 = out_blk.conv1.ffc
        if not ffc_bn_act_shapes_ok(out_blk.conv1, torch.empty(1, f0.convl2l.in_channels, h, w, device="meta"),
                                    torch.empty(1, f0.global_in_num, h, w, device="meta")):
            return False
    if c != stem.ffc.convl2l.in_channels or h < 4 or w < 4:
        return False
    f = 2 ** len(downs)
    if h % f or w % f:                      # ConvTranspose doubles sizes: only exact multiples round-trip
        return False
    if blocks_use_fft(gen) and not plane_ok(h // f, w // f):     # e.g. 512-wide bottleneck planes (4096 px images)
        return False
    return h // f >= 2 and w // f >= 2      # reflect pad 1 at the bottleneck


def blocks_use_fft(gen) -> bool:
    lay = _generator_layout(gen)
    return lay is not None and any(not isinstance(b.conv1.ffc.convg2g, nn.Identity) for b in lay[2])


# -------------------------------------------------------------------------------------- builders
def _fold(bn, n, device):
    if isinstance(bn, nn.BatchNorm2d):
        return P.bn_scale_shift(bn)
    return torch.ones(n, dtype=torch.float64, device=device), torch.zeros(n, dtype=torch.float64, device=device)


def fu_batch_chunk(batch: int, h: int, w: int, c: int) -> int:
    """Images per pass of the rfft2 -> GEMM -> irfft2 chain.  Running the chain over slices of the batch keeps
    its intermediates inside the 126 MB L2, but measured on B200 (bs32, 512x512: 624 img/s unsliced vs 600-620
    with 6..16-image slices, profiles/r01_fu_chunk_sweep.txt) the extra launches and partial waves cost more than
    the L2 hits save — the kernels are latency- not bandwidth-bound.  Default: whole batch;
    LAMA_B200_FU_CHUNK=n slices."""
    n = int(os.environ.get("LAMA_B200_FU_CHUNK", "0"))
    return batch if n <= 0 else min(batch, n)


def fu_planar_ok(prog: Program, st, h: int, w: int) -> bool:
    """Channel-group planar storage for the SpectralTransform chain (conv1 -> rfft2 -> spectral conv -> irfft2 ->
    conv2): every (image, 4-channel group) plane set is one dense block for the second-generation plane FFT kernels
    (csrc/fft_plane_cg.cu) and the GEMMs read [K/8][pixel][8] operand tiles.  Needs the tcgen05 arm, 64x64 or 32x32
    planes (the 512x512 / 256x256 bottleneck) and whole 64-channel K blocks on every contraction of the chain.
    LAMA_B200_FU_LAYOUT=nhwc keeps the round-1 channels-last chain (A/B measurements)."""
    if prog.math != L.MATH_BF16X3 or os.environ.get("LAMA_B200_FU_LAYOUT", "planar") != "planar":
        return False
    dev = st.conv2.weight.device
    if dev.type == "cuda" and not planar_selftest(dev):
        return False
    c = st.conv1[0].out_channels
    fu = st.fu
    if st.enable_lfu and not ((h, w) == (64, 64) and c % 32 == 0):      # LFU quadrants: 32x32 planes, c/4 % 8 == 0
        return False
    return ((h, w) in ((64, 64), (32, 32)) and c % 64 == 0 and fu.conv_layer.in_channels == 2 * c + (2 if fu.spectral_pos_encoding else 0)
            and fu.conv_layer.out_channels == 2 * c)


_PLANAR_OK: Dict[int, bool] = {}


def planar_selftest(device: torch.device) -> bool:
    """Once per process and device: run the planar chain's three kernels (plane FFT pair, interleaved-operand GEMM with
    a planar output) on a small random problem and compare with torch on the same device.  The chain depends on
    details no compile-time check covers (tcgen05 no-swizzle descriptor fields, bulk-copy tile layout); if the check
    fails the process keeps the channels-last chain of round 1 (still the native kernels) and says so loudly —
    LAMA_B200_FU_LAYOUT=planar! skips the check and forces the planar chain, =nhwc forces the other."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if os.environ.get("LAMA_B200_FU_LAYOUT") == "planar!":
        return True
    if idx in _PLANAR_OK:
        return _PLANAR_OK[idx]
    ok, why = True, ""
    try:
        with torch.no_grad():
            b, c, h = 2, 64, 64
            wf = h // 2 + 1
            g = torch.Generator().manual_seed(1234)
            prog = Program("planar_selftest", L.MATH_BF16X3)
            X = prog.buf("x", b, h, h, c, cg=4)
            S = prog.buf("s", b, h, wf, 2 * c, gemm=True, cg=8)
            Z = prog.buf("z", b, h, wf, 2 * c, cg=8)
            U = prog.buf("u", b, h, h, c, gemm=True, cg=8)
            wgt = torch.randn(2 * c, 2 * c, 1, 1, generator=g) * 0.1
            pk = P.pack_conv([(wgt, 0, 0, 0)], None, None, act=L.ACT_RELU)
            prog.inputs = {"x0": (b, c, h, h)}
            prog.ops += [ToNHWC("x0", TV(X)), RfftOp(TV(X), TV(S)), ConvOp(pk, [TV(S), None], TV(Z)),
                         IrfftOp(TV(Z), TV(X), TV(U)), ToNCHW(TV(U), "y0")]
            prog.outputs = {"y0": (b, c, h, h)}
            x = torch.randn(b, c, h, h, generator=g).to(device)
            y = CudaExecutor(prog, device).run({"x0": x})["y0"]
            f = torch.fft.rfftn(x.double(), dim=(-2, -1), norm="ortho")
            f = torch.stack((f.real, f.imag), dim=2).reshape(b, 2 * c, h, wf)
            f = torch.relu(torch.einsum("nk,bkyx->bnyx", wgt[:, :, 0, 0].double().to(device), f))
            f = f.reshape(b, c, 2, h, wf)
            want = x.double() + torch.fft.irfftn(torch.complex(f[:, :, 0], f[:, :, 1]), s=(h, h), dim=(-2, -1), norm="ortho")
            err = float((y.double() - want).abs().max()) / float(want.abs().max())
            ok, why = err < 1e-3, f"relative error {err:.2e}"
    except Exception as e:  # noqa: BLE001  (a launch failure is an answer too)
        ok, why = False, f"{type(e).__name__}: {e}"
    _PLANAR_OK[idx] = ok
    if not ok:
        import warnings
        msg = (f"lama_b200: the channel-group planar FourierUnit chain failed its start-up check on cuda:{idx} ({why}); "
               f"using the channels-last chain (native round-1 kernels) instead")
        warnings.warn(msg, RuntimeWarning)
        print(msg, file=__import__("sys").stderr, flush=True)
    return ok


def emit_fourier_unit(prog: Program, fu, t: TV, out: TV, residual: Optional[TV]):
    """FourierUnit (ffc.py:76-113): rfft2 -> [1x1 conv + BN + ReLU] on the interleaved spectrum -> irfft2,
    optionally with the SpectralTransform residual fused into the inverse (out = residual + fu(t))."""
    b = t.batch
    h, w = t.hw
    wf = w // 2 + 1
    cout2 = fu.conv_layer.out_channels
    cin2 = fu.conv_layer.in_channels - (2 if fu.spectral_pos_encoding else 0)       # spectrum channels (ffc.py:57)
    planar = t.buf.cg == 4      # FourierUnit chain in channel-group planar storage (emit_spectral_transform decides)
    S = prog.buf("spectrum", b, h, wf, cin2, gemm=True, cg=8 if planar else 0)
    Z = prog.buf("spectrum_out", b, h, wf, cout2, cg=8 if planar else 0)
    prog.meta[("fu", id(fu))] = dict(S=S, Z=Z)
    scale, shift = P.bn_scale_shift(fu.bn)
    wconv, pos = fu.conv_layer.weight, None
    if fu.spectral_pos_encoding:
        # ffc.py:91-95 prepends two coordinate channels (linspace over H and over W/2+1) to the spectrum.  They do not
        # depend on the data: their contribution W[:, :2] . (v, h) is a per-position addend of the spectral GEMM
        # (BN scale folded), broadcast over the batch.
        wpos = wconv.detach().double()[:, :2, 0, 0] * scale.double()[:, None]                     # [2co, 2]
        wconv = wconv[:, 2:]
        vert = torch.linspace(0, 1, h, dtype=torch.float64, device=wpos.device)
        hor = torch.linspace(0, 1, wf, dtype=torch.float64, device=wpos.device)
        PE = prog.buf("fu.pos_addend", 1, h, wf, cout2)
        prog.consts[PE.name] = (vert[:, None, None] * wpos[:, 0] + hor[None, :, None] * wpos[:, 1])[None].float()
        pos = TV(PE, bcast=b)
    cin2 = wconv.shape[1]
    pk = P.pack_conv([(wconv, 0, 0, 0)], scale, shift, act=L.ACT_RELU, device=fu.conv_layer.weight.device)
    chunk = fu_batch_chunk(b, h, w, cin2 // 2) if prog.math == L.MATH_BF16X3 else b
    if planar and ((chunk * h * wf) % 128 or (chunk * h * w) % 128):
        chunk = b          # tile-blocked buffers can only be sliced on 128-pixel blocks
    for b0 in range(0, b, chunk):
        nb = min(chunk, b - b0)
        prog.ops.append(RfftOp(t.bslice(b0, nb), TV(S).bslice(b0, nb)))
        prog.ops.append(ConvOp(pk, [TV(S).bslice(b0, nb), None], TV(Z).bslice(b0, nb),
                               addend=pos.bslice(b0, nb) if pos is not None else None, tag="fu.conv_layer+bn+relu"))
        prog.ops.append(IrfftOp(TV(Z).bslice(b0, nb), residual.bslice(b0, nb) if residual is not None else None,
                                out.bslice(b0, nb)))


def st_out_hw(st, h: int, w: int) -> Tuple[int, int]:
    """Spatial size SpectralTransform works at: AvgPool2d(2, 2) first when stride == 2 (ffc.py:122-125)."""
    return (h // 2, w // 2) if st.stride == 2 else (h, w)


def lfu_supported(st, h: int, w: int) -> bool:
    """LFU (ffc.py:148-157) natively: the first c/4 channels, cut into 2x2 quadrants stacked as channels — the
    reference splits rows AND columns by h // 2, which only type-checks for even square planes; quadrant views carry
    c/4 channels and every view needs a multiple of 4."""
    c = st.conv1[0].out_channels
    return h == w and h % 2 == 0 and h >= 4 and c % 16 == 0 and st.lfu.native_supported()


def emit_spectral_transform(prog: Program, st, x: TV, u_consumer=None) -> Tuple[TV, P.PackedConv]:
    """SpectralTransform (ffc.py:142-163) up to, but not including, conv2: returns the view holding
    ``x1 + fu(x1) [+ tile(lfu(quadrants(x1)))]`` and lets the caller fuse conv2 into its own contraction.
    stride 2: AvgPool2d(2,2) + the 1x1 conv1 are ONE 2x2 stride-2 contraction (four taps with conv1.weight / 4)."""
    b = x.buf.B
    h, w = st_out_hw(st, *x.hw)
    c = st.conv1[0].out_channels
    dev = st.conv2.weight.device
    planar = fu_planar_ok(prog, st, h, w)
    T = prog.buf("st.t", b, h, w, c, cg=4 if planar else 0)
    U = prog.buf("st.u", b, h, w, c, gemm=True, cg=8 if planar else 0)
    s1, b1 = P.bn_scale_shift(st.conv1[1])
    if st.stride == 2:
        w1 = st.conv1[0].weight.detach().repeat(1, 1, 2, 2) / 4.0
        pk1 = P.pack_conv([(w1, 0, x.c0, 0)], s1, b1, stride=2, act=L.ACT_RELU, device=dev)
        tag = "st.avgpool2x2+conv1+bn+relu"
    else:
        pk1 = P.pack_conv([(st.conv1[0].weight, 0, x.c0, 0)], s1, b1, act=L.ACT_RELU, device=dev)
        tag = "st.conv1+bn+relu"
    prog.ops.append(ConvOp(pk1, [TV(x.buf), None], TV(T), tag=tag))
    prog.meta[("st", id(st))] = dict(T=T, U=U)
    residual = TV(T)
    if st.enable_lfu:
        # xs = lfu(quadrants of the first c/4 channels) tiled 2x2; XS = T + tile(xs) becomes the residual of the main
        # inverse transform, so  U = T + fu(T) + tile(xs)  (ffc.py:161) costs no extra pass over U
        lfu, c4, s = st.lfu, c // 4, h // 2
        sf = s // 2 + 1
        XS = prog.buf("st.xs", b, h, w, c, cg=4 if planar else 0)
        LS = prog.buf("lfu.spectrum", b, s, sf, 2 * c, gemm=True, cg=8 if planar else 0)
        LZ = prog.buf("lfu.spectrum_out", b, s, sf, 2 * c, cg=8 if planar else 0)
        # channel order of the two torch.cat calls (ffc.py:152-155): top-left, bottom-left, top-right, bottom-right
        for q, (qy, qx) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
            prog.ops.append(RfftOp(TV(T, 0, c4, win=(qy * s, qx * s, s, s)), TV(LS, q * 2 * c4, 2 * c4)))
        sc, sh = P.bn_scale_shift(lfu.bn)
        pkl = P.pack_conv([(lfu.conv_layer.weight, 0, 0, 0)], sc, sh, act=L.ACT_RELU, device=dev)
        prog.ops.append(ConvOp(pkl, [TV(LS), None], TV(LZ), tag="lfu.conv_layer+bn+relu"))
        for ty in (0, 1):
            for tx in (0, 1):          # .repeat(1, 1, 2, 2) (ffc.py:157): the same s x s result in all four quadrants
                wq = (ty * s, tx * s, s, s)
                prog.ops.append(IrfftOp(TV(LZ), TV(T, win=wq), TV(XS, win=wq)))
        residual = TV(XS)
    emit_fourier_unit(prog, st.fu, TV(T), TV(U), residual=residual)
    return TV(U)


def emit_ffc_bn_act(prog: Program, m, X: Buf, in_cl: int, in_cg: int, residual: Optional[Buf] = None,
                    Y: Optional[Buf] = None) -> Tuple[Buf, int, int]:
    """FFC + BN + activation (ffc.py:205-225, 251-255) reading [x_l | x_g] from ``X`` and writing
    [y_l | y_g] into ``Y`` (allocated here unless given).  With ``residual`` the block identity
    (ffc.py:288) is added after the activation in the same epilogues.
    Returns (Y, out_cl, out_cg)."""
    f = m.ffc
    dev = next(m.parameters()).device
    conv0 = next(c for c in (f.convl2l, f.convl2g, f.convg2l) if not isinstance(c, nn.Identity))
    k, s, p = conv0.kernel_size[0], conv0.stride[0], conv0.padding[0]
    out_cl = f.convl2l.out_channels if not isinstance(f.convl2l, nn.Identity) else (
        f.convg2l.out_channels if not isinstance(f.convg2l, nn.Identity) else 0)
    out_cg = f.convl2g.out_channels if not isinstance(f.convl2g, nn.Identity) else 0
    ho, wo = (X.H + 2 * p - k) // s + 1, (X.W + 2 * p - k) // s + 1
    if Y is None:
        Y = prog.buf("ffc.out", X.B, ho, wo, out_cl + out_cg, gemm=True, halo=True)
    act_l, act_g = _act_code(m.act_l), _act_code(m.act_g)
    sl, bl = _fold(m.bn_l, out_cl, dev)
    sg, bg = _fold(m.bn_g, out_cg, dev)
    has_spectral = not isinstance(f.convg2g, nn.Identity)
    res_l = TV(residual, 0, out_cl) if residual is not None else None
    res_g = TV(residual, out_cl, out_cg) if residual is not None else None

    if in_cg == 0 and out_cl > 0 and out_cg > 0 and act_l == act_g:
        # local input only (stem-like / downsample-to-global): convl2l and convl2g read the same
        # pixels, so they are ONE contraction with N = out_cl + out_cg.
        wcat = torch.cat([f.convl2l.weight, f.convl2g.weight], dim=0)
        pk = P.pack_conv([(wcat, 0, 0, p)], torch.cat([sl, sg]), torch.cat([bl, bg]), stride=s, act=act_l, device=dev)
        prog.ops.append(ConvOp(pk, [TV(X, 0, in_cl), None], TV(Y), addend=TV(residual) if residual else None,
                               addend_post=True, tag="convl2l|convl2g+bn+act"))
        return Y, out_cl, out_cg

    if out_cl > 0:
        # y_l = act(bn_l(convl2l(x_l) + convg2l(x_g))): x_l|x_g are adjacent channels of X, so the
        # two convolutions are one contraction over C = in_cl + in_cg.
        ws = [f.convl2l.weight] + ([f.convg2l.weight] if in_cg > 0 else [])
        pk = P.pack_conv([(torch.cat(ws, dim=1), 0, 0, p)], sl, bl, stride=s, act=act_l, device=dev)
        prog.ops.append(ConvOp(pk, [TV(X), None], TV(Y, 0, out_cl), addend=res_l, addend_post=True,
                               tag="convl2l+convg2l+bn_l+act"))
    if out_cg > 0:
        parts = [(f.convl2g.weight, 0, 0, p)]     # ffc_bn_act_supported() guarantees convl2g exists
        ins = [TV(X), None]
        pre = None
        if has_spectral:
            U = emit_spectral_transform(prog, f.convg2g, TV(X, in_cl, in_cg))
            if s == 1:
                parts.append((f.convg2g.conv2.weight, 1, 0, 0))
                ins[1] = U
            else:
                # stride-2 FFC (ffc.py:122-125, 221-224): convl2g samples X with stride 2 while conv2 reads the already
                # pooled u with stride 1 — one ffcb_conv has one stride, so conv2 (with bn_g's scale folded, no shift)
                # runs first and joins the 3x3 contraction as its pre-activation addend.  Never a residual layer.
                assert residual is None
                A = prog.buf("st.conv2.out", X.B, ho, wo, out_cg)
                pk2 = P.pack_conv([(f.convg2g.conv2.weight, 0, 0, 0)], sg, None, device=dev)
                prog.ops.append(ConvOp(pk2, [U, None], TV(A), tag="st.conv2 (x bn_g scale)"))
                pre = TV(A)
        # y_g = act(bn_g(convl2g(x_l) + conv2(x1 + fu(x1)))): conv2 rides as one more K-segment.
        pk = P.pack_conv(parts, sg, bg, stride=s, act=act_g, device=dev)
        prog.ops.append(ConvOp(pk, ins, TV(Y, out_cl, out_cg), addend=pre if pre is not None else res_g,
                               addend_post=pre is None, tag="convl2g+st.conv2+bn_g+act"))
    return Y, out_cl, out_cg


def emit_resnet_block(prog: Program, blk, X: Buf, cl: int, cg: int, in_place: bool) -> Buf:
    """FFCResnetBlock (ffc.py:277-292): X <- X + conv2(conv1(X)).  With ``in_place`` the second
    FFC_BN_ACT writes its result over X (each output pixel only reads its own residual pixel)."""
    Y, ycl, ycg = emit_ffc_bn_act(prog, blk.conv1, X, cl, cg)
    out = X if in_place else prog.buf("block.out", X.B, X.H, X.W, X.C, gemm=True, halo=True)
    emit_ffc_bn_act(prog, blk.conv2, Y, ycl, ycg, residual=X, Y=out)
    return out


# Largest plane side the forward+backward block program is used for.  Hardware-validated against autograd: 32x32 and
# 64x64 (planar chain), 12x20 and 17x25 (general FFT kernels) — tests/test_gpu_parity.py.  A late check of round 2
# (tools/grad_check.py, profiles/r02_grad_check.json) found the program's FORWARD off by 8-14 % (2-norm) on 128-wide
# planes (128x128 and 96x128 alike, so an x-direction effect) although the CPU interpretation of the very same program
# matches the oracle to 2e-7 and the whole-generator program is right at 128x128 and 256x256 planes: a kernel-level
# defect specific to the standalone block program's buffers at that width, not yet located.  Until it is, planes wider
# than 64 take the torch-autograd composition (correct, ~1.4x slower in the refinement loop).
BLOCK_GRAD_MAX_PLANE = 64


def block_grad_supported(blk) -> bool:
    """Input gradients exist for the residual-block flavour of the shipped generators: two
    FFC_BN_ACT with local and global halves on both sides, 3x3 reflect convs, stride 1, ReLU, no LFU / gating."""
    if blk.inline or not blk.native_supported():
        return False
    for m in (blk.conv1, blk.conv2):
        f = m.ffc
        if any(isinstance(c, nn.Identity) for c in (f.convl2l, f.convl2g, f.convg2l, f.convg2g)):
            return False
        st = f.convg2g
        if (f.convl2l.kernel_size != (3, 3) or f.convl2l.stride != (1, 1) or f.convl2l.padding != (1, 1)
                or st.enable_lfu or st.stride != 1 or _act_code(m.act_l) != L.ACT_RELU or _act_code(m.act_g) != L.ACT_RELU):
            return False
    return True


def _flip_t(w: torch.Tensor) -> torch.Tensor:
    """[N, C, k, k] forward conv weight -> [C, N, k, k] weight of the gradient convolution (taps flipped)."""
    return w.detach().double().flip(-1, -2).transpose(0, 1).contiguous()


def emit_ffc_bn_act_backward(prog: Program, m, Y: Buf, DO: TV, in_cl: int, in_cg: int,
                             extra: Optional[TV] = None) -> Buf:
    """Gradient of one FFC_BN_ACT (ffc.py:205-225, 251-255; eval-mode BN) w.r.t. its input [x_l | x_g], given the
    gradient ``DO`` w.r.t. its output and the forward activations kept in ``prog`` (Y, t, z).  Every step is one of
    the forward's own operations with transposed weights:
      dP   = dO * [Y > 0]
      d[x_l|x_g]_pad = 3x3 gradient convolutions of dP (flipped taps, zero border) -> folded back (reflect adjoint)
      du   = W2^T (s_g dP_g);  dz = rfft2(du) * [z > 0];  ds = Wf^T (s_f dz);  dt = du + irfft2(ds)
             (the half-spectrum weights 2 and 1/2 of the two adjoint transforms cancel around the per-position GEMM)
      dx_g += W1^T (s_1 (dt * [t > 0]))
    ``extra``: one more gradient added into the result (the block's identity path)."""
    f = m.ffc
    st = f.convg2g
    dev = f.convl2l.weight.device
    b, h, w = Y.B, Y.H, Y.W
    out_cl, out_cg = f.convl2l.out_channels, f.convl2g.out_channels
    sl, _ = _fold(m.bn_l, out_cl, dev)
    sg, _ = _fold(m.bn_g, out_cg, dev)
    s1, _ = P.bn_scale_shift(st.conv1[1])
    sf, _ = P.bn_scale_shift(st.fu.bn)
    saved_st, saved_fu = prog.meta[("st", id(st))], prog.meta[("fu", id(st.fu))]
    T, Z = saved_st["T"], saved_fu["Z"]
    planar = T.cg == 4
    c = st.conv1[0].out_channels
    wf_ = w // 2 + 1
    k4 = lambda t: t[:, :, None, None]      # noqa: E731
    col = lambda v: v.double()[:, None, None, None]   # noqa: E731

    DP = prog.buf("grad.dP", b, h, w, out_cl + out_cg, gemm=True)
    prog.ops.append(ReluBwdOp(DO, TV(Y), TV(DP)))
    GP = prog.buf("grad.gpad", b, h + 2, w + 2, in_cl + in_cg)
    w_to_l = torch.cat([f.convl2l.weight.detach().double() * col(sl), f.convl2g.weight.detach().double() * col(sg)], dim=0)
    pk_a = P.pack_conv([(_flip_t(w_to_l), 0, 0, 2)], None, None, border=L.BORDER_ZERO, device=dev)
    prog.ops.append(ConvOp(pk_a, [TV(DP), None], TV(GP, 0, in_cl), tag="grad: d x_l (3x3^T of dP_l|dP_g)"))
    pk_b = P.pack_conv([(_flip_t(f.convg2l.weight.detach().double() * col(sl)), 0, 0, 2)], None, None,
                       border=L.BORDER_ZERO, device=dev)
    prog.ops.append(ConvOp(pk_b, [TV(DP), None], TV(GP, in_cl, in_cg), tag="grad: d x_g (3x3^T of dP_l)"))

    DU = prog.buf("grad.du", b, h, w, c, cg=4 if planar else 0)
    w2 = st.conv2.weight.detach().double()[:, :, 0, 0] * sg.double()[:, None]            # [out_cg, c]
    pk2 = P.pack_conv([(k4(w2.t().contiguous()), 0, out_cl, 0)], None, None, device=dev)
    prog.ops.append(ConvOp(pk2, [TV(DP), None], TV(DU), tag="grad: du = conv2^T"))
    DZ = prog.buf("grad.dz", b, h, wf_, 2 * c, gemm=True, cg=8 if planar else 0)
    prog.ops.append(RfftOp(TV(DU), TV(DZ)))
    DPZ = prog.buf("grad.dpz", b, h, wf_, 2 * c, gemm=True, cg=8 if planar else 0)
    prog.ops.append(ReluBwdOp(TV(DZ), TV(Z), TV(DPZ)))
    DS = prog.buf("grad.ds", b, h, wf_, 2 * c, cg=8 if planar else 0)
    wfu = st.fu.conv_layer.weight.detach().double()[:, :, 0, 0] * sf.double()[:, None]  # [2c out, 2c in]
    pkf = P.pack_conv([(k4(wfu.t().contiguous()), 0, 0, 0)], None, None, device=dev)
    prog.ops.append(ConvOp(pkf, [TV(DPZ), None], TV(DS), tag="grad: ds = fu.conv_layer^T"))
    DT = prog.buf("grad.dt", b, h, w, c, cg=4 if planar else 0)
    prog.ops.append(IrfftOp(TV(DS), TV(DU), TV(DT)))
    DPT = prog.buf("grad.dpt", b, h, w, c, gemm=True, cg=8 if planar else 0)
    prog.ops.append(ReluBwdOp(TV(DT), TV(T), TV(DPT)))
    DG = prog.buf("grad.dg", b, h, w, in_cg)
    w1 = st.conv1[0].weight.detach().double()[:, :, 0, 0] * s1.double()[:, None]        # [c, in_cg]
    pk1 = P.pack_conv([(k4(w1.t().contiguous()), 0, 0, 0)], None, None, device=dev)
    prog.ops.append(ConvOp(pk1, [TV(DPT), None], TV(DG), tag="grad: d x_g += conv1^T"))
    DX = prog.buf("grad.dx", b, h, w, in_cl + in_cg)
    prog.ops.append(FoldOp(TV(GP), [(TV(DG), in_cl)] + ([(extra, 0)] if extra is not None else []), TV(DX)))
    return DX


def build_block_grad_program(prog: Program, blk, sl: Tuple[int, ...], sg: Tuple[int, ...]):
    """Forward of FFCResnetBlock WITHOUT the identity add (all activations kept) | SplitOp | input-gradient program.
    inputs  x0, x1 (forward), g0, g1 (gradient w.r.t. the block outputs);
    outputs y0, y1 = conv2(conv1(x)) halves, dx0, dx1 = gradients w.r.t. x_l, x_g (identity path included)."""
    b, cl, h, w = sl
    cg = sg[1]
    prog.inputs.update(x0=tuple(sl), x1=tuple(sg), g0=tuple(sl), g1=tuple(sg))
    X = prog.buf("in", b, h, w, cl + cg, gemm=True, halo=True)
    prog.ops.append(ToNHWC("x0", TV(X, 0, cl)))
    prog.ops.append(ToNHWC("x1", TV(X, cl, cg)))
    Y1, _, _ = emit_ffc_bn_act(prog, blk.conv1, X, cl, cg)
    Y2, _, _ = emit_ffc_bn_act(prog, blk.conv2, Y1, cl, cg)
    prog.ops.append(ToNCHW(TV(Y2, 0, cl), "y0")); prog.outputs["y0"] = tuple(sl)
    prog.ops.append(ToNCHW(TV(Y2, cl, cg), "y1")); prog.outputs["y1"] = tuple(sg)
    prog.ops.append(SplitOp())
    DO = prog.buf("grad.dout", b, h, w, cl + cg)
    prog.ops.append(ToNHWC("g0", TV(DO, 0, cl)))
    prog.ops.append(ToNHWC("g1", TV(DO, cl, cg)))
    D1 = emit_ffc_bn_act_backward(prog, blk.conv2, Y2, TV(DO), cl, cg)
    D0 = emit_ffc_bn_act_backward(prog, blk.conv1, Y1, TV(D1), cl, cg, extra=TV(DO))
    prog.ops.append(ToNCHW(TV(D0, 0, cl), "dx0")); prog.outputs["dx0"] = tuple(sl)
    prog.ops.append(ToNCHW(TV(D0, cl, cg), "dx1")); prog.outputs["dx1"] = tuple(sg)


def build_module_program(module, kind: str, shapes: Sequence[Optional[Tuple[int, ...]]], math: int) -> Program:
    """Programs for stand-alone module calls: NCHW float in -> channels-last inside -> NCHW float out."""
    prog = Program(kind=kind, math=math)
    if kind == "fourier_unit":
        b, c, h, w = shapes[0]
        prog.inputs["x0"] = shapes[0]
        X = prog.buf("in", b, h, w, c)
        prog.ops.append(ToNHWC("x0", TV(X)))
        co = module.conv_layer.out_channels // 2
        O = prog.buf("out", b, h, w, co)
        emit_fourier_unit(prog, module, TV(X), TV(O), residual=None)
        prog.ops.append(ToNCHW(TV(O), "y0")); prog.outputs["y0"] = (b, co, h, w)
    elif kind == "spectral_transform":
        b, c, h, w = shapes[0]
        prog.inputs["x0"] = shapes[0]
        X = prog.buf("in", b, h, w, c, gemm=True, halo=module.stride == 2)
        prog.ops.append(ToNHWC("x0", TV(X)))
        U = emit_spectral_transform(prog, module, TV(X))
        co = module.conv2.out_channels
        h, w = st_out_hw(module, h, w)
        O = prog.buf("out", b, h, w, co)
        pk = P.pack_conv([(module.conv2.weight, 0, 0, 0)], None, None, device=module.conv2.weight.device)
        prog.ops.append(ConvOp(pk, [U, None], TV(O), tag="st.conv2"))
        prog.ops.append(ToNCHW(TV(O), "y0")); prog.outputs["y0"] = (b, co, h, w)
    elif kind in ("ffc_bn_act", "resnet_block"):
        sl, sg = shapes
        b, cl, h, w = sl
        cg = sg[1] if sg is not None else 0
        prog.inputs["x0"] = sl
        X = prog.buf("in", b, h, w, cl + cg, gemm=True, halo=True)
        prog.ops.append(ToNHWC("x0", TV(X, 0, cl)))
        if cg:
            prog.inputs["x1"] = sg
            prog.ops.append(ToNHWC("x1", TV(X, cl, cg)))
        if kind == "ffc_bn_act":
            Y, ocl, ocg = emit_ffc_bn_act(prog, module, X, cl, cg)
        else:
            Y = emit_resnet_block(prog, module, X, cl, cg, in_place=False)
            ocl, ocg = cl, cg
        if ocl:
            prog.ops.append(ToNCHW(TV(Y, 0, ocl), "y0")); prog.outputs["y0"] = (b, ocl, Y.H, Y.W)
        if ocg:
            prog.ops.append(ToNCHW(TV(Y, ocl, ocg), "y1")); prog.outputs["y1"] = (b, ocg, Y.H, Y.W)
    elif kind == "resnet_block_grad":
        build_block_grad_program(prog, module, shapes[0], shapes[1])
    elif kind == "generator":
        build_generator_program(prog, module, shapes[0])
    elif kind.startswith("generator_u8"):            # "generator_u8:<pad modulo>", shapes = (img, mask)
        mod = int(kind.split(":")[1]) if ":" in kind else 8
        b, h0, w0, _ = shapes[0]
        h, w = -(-h0 // mod) * mod, -(-w0 // mod) * mod
        build_generator_program(prog, module, (b, 4, h, w), u8_size=(h0, w0))
    else:
        raise ValueError(kind)
    if math == L.MATH_BF16X3 and not tc_compatible(prog):
        if kind.startswith("generator_u8"):
            raise ValueError("the uint8 predict path needs channel counts in multiples of 8 (tensor-core arm)")
        return build_module_program(module, kind, shapes, L.MATH_FP32)
    insert_border_ops(prog)
    return prog


def build_generator_program(prog: Program, gen, shape, u8_size: Optional[Tuple[int, int]] = None):
    """FFCResNetGenerator (ffc.py:306-367) as one program: stem -> stride-2 convs -> residual blocks
    (in place on one 512-channel buffer) -> sub-pixel transposed convs -> head.

    ``u8_size=(H0, W0)``: the predict-path variant.  Inputs are the decoded bytes "img"
    (B,H0,W0,3) and "mask" (B,H0,W0); ``shape`` is the modulo-padded generator input (B,4,H,W); the output "y0" is
    the inpainted RGB bytes (B,H0,W0,3).  Pre/post-processing lives in the pack and gather kernels."""
    stem, downs, blocks, ups, out_blk, head, out_act = _generator_layout(gen)
    b, cin, h, w = shape
    dev = head.weight.device
    if u8_size is None:
        prog.inputs["x0"] = tuple(shape)
    else:
        h0, w0 = u8_size
        prog.inputs["img"], prog.inputs["mask"] = (b, h0, w0, 3), (b, h0, w0)
        prog.dtypes.update(img=torch.uint8, mask=torch.uint8, y0=torch.uint8)
    conv = stem.ffc.convl2l
    n0 = conv.out_channels
    s0, b0 = P.bn_scale_shift(stem.bn_l)
    wst, shst = P.pack_stem(conv.weight, s0, b0, device=dev)
    X = prog.buf("stem", b, h, w, n0, gemm=True, halo=True)
    tc_stem = (prog.math == L.MATH_BF16X3 and cin <= 8 and n0 % 8 == 0
               and os.environ.get("LAMA_B200_STEM", "tc") == "tc")
    if u8_size is not None and not (tc_stem and cin == 4):
        raise ValueError("the uint8 predict path needs the tensor-core stem (bf16x3 arithmetic, 4 input channels)")
    if tc_stem:
        # tensor-core stem: the 7x7 window of the packed image is 7 contiguous 128-byte K blocks per pixel
        Pk = prog.buf("stem.packed", b, h + 6, w + 8, 8, gemm=True)
        if u8_size is None:
            prog.ops.append(StemPackOp("x0", cin, TV(Pk)))
        else:
            prog.ops.append(StemPackU8Op("img", "mask", h0, w0, TV(Pk)))
        pk = P.pack_stem_windowed(conv.weight, s0, b0, device=dev)
        prog.ops.append(ConvOp(pk, [TV(Pk, window=8), None], TV(X), tag="stem 7x7 (windowed)+bn+relu"))
    else:
        prog.ops.append(StemOp("x0", cin, wst, shst, TV(X)))
    cl, cg = n0, 0
    for d in downs:
        X, cl, cg = emit_ffc_bn_act(prog, d, X, cl, cg)
    for blk in blocks:
        X = emit_resnet_block(prog, blk, X, cl, cg, in_place=True)
    # ConcatTupleLayer (ffc.py:295-302) is free: x_l | x_g already share X.
    tc_head = (prog.math == L.MATH_BF16X3 and head.in_channels % 8 == 0 and head.out_channels <= 3
               and min(h, w) > 3 and os.environ.get("LAMA_B200_HEAD", "tc") == "tc")
    for iu, (ct, bn) in enumerate(ups):
        sc, sh = P.bn_scale_shift(bn)
        last = iu == len(ups) - 1
        # tensor-core head: its row contraction reads 3 pixels up and down -> ring of 3; CUDA-core head: float32
        Yb = prog.buf("up", b, X.H * 2, X.W * 2, ct.out_channels, gemm=(not last) or tc_head,
                      halo=(not last) or tc_head, halo_px=3 if last else 1)
        for a, bb, pk in P.pack_conv_transpose_phases(ct.weight, ct.bias, sc, sh, act=L.ACT_RELU, device=dev):
            prog.ops.append(ConvOp(pk, [TV(X), None], TV(Yb, phase=(a, bb)), tag=f"convT phase {a}{bb}+bn+relu"))
        X = Yb
    if out_blk is not None:
        # out_ffc: FFCResnetBlock(inline=True) splits its input as x[:, :-g] | x[:, -g:] (ffc.py:278-280) — exactly the
        # [local | global] channel order of the buffer, so it runs in place like the bottleneck blocks
        ocg = out_blk.conv1.ffc.global_in_num
        X = emit_resnet_block(prog, out_blk, X, X.C - ocg, ocg, in_place=True)
    if u8_size is not None and not (tc_head and ups and head.out_channels == 3):
        raise ValueError("the uint8 predict path needs the tensor-core head with 3 output channels")
    if tc_head and ups:
        pkh = P.pack_head_rows(head.weight, device=dev)
        Q = prog.buf("head.q", b, h, w, pkh.n_out)
        prog.ops.append(ConvOp(pkh, [TV(X), None], TV(Q), tag="head 7x7 rows"))
        bias = head.bias.detach().float().contiguous() if head.bias is not None else torch.zeros(head.out_channels)
        if u8_size is None:
            prog.ops.append(HeadGatherOp(TV(Q), bias.to(dev), head.out_channels, out_act, "y0"))
        else:
            prog.ops.append(HeadGatherU8Op(TV(Q), bias.to(dev), out_act, "img", "mask", h0, w0, "y0"))
    else:
        wh, bh = P.pack_head(head.weight, head.bias, device=dev)
        prog.ops.append(HeadOp(TV(X), wh, bh, head.out_channels, out_act, "y0"))
    prog.outputs["y0"] = (b, head.out_channels, h, w) if u8_size is None else (b, h0, w0, 3)


def tc_compatible(prog: Program) -> bool:
    """The tcgen05 arm needs 16-byte aligned bf16 pixels/slices: channel counts and slice starts in
    multiples of 8.  Programs that do not qualify run the fp32 CUDA-core arm (still native)."""
    for op in prog.ops:
        if isinstance(op, ConvOp):
            for tv in op.ins:
                if tv is not None and (tv.buf.C % 8 or tv.c0 % 8):
                    return False
            if any(sg.c0 % 8 for sg in op.packed.segs):
                return False
            if op.out.buf.fmt == L.BF16X2 and (op.out.buf.C % 8 or op.out.c0 % 8):
                return False
    return True


def insert_border_ops(prog: Program):
    """Producers never write the reflected ring of a padded buffer (the tcgen05 epilogue stores through a
    tensor map of the interior; layout conversions, the stem and the FFT kernels write pixels only).  Insert
    a BorderOp lazily: right before the first contraction that reads a buffer whose interior changed since
    its ring was last rebuilt.  For the in-place residual blocks that is one ring refresh per FFC_BN_ACT."""
    out, dirty = [], {}
    for op in prog.ops:
        if isinstance(op, ConvOp):
            for tv in op.ins:
                if tv is not None and tv.buf.reflect_border and id(tv.buf) in dirty:
                    out.append(BorderOp(TV(tv.buf)))
                    del dirty[id(tv.buf)]
        out.append(op)
        wrote = None
        if isinstance(op, (ToNHWC, StemOp, StemPackOp, StemPackU8Op, IrfftOp, ConvOp)):
            wrote = op.out
        elif isinstance(op, RfftOp):
            wrote = op.spec
        if wrote is not None and wrote.buf.reflect_border and not conv_writes_ring(prog, op):
            dirty[id(wrote.buf)] = wrote.buf
    prog.ops = out


def conv_writes_ring(prog: Program, op) -> bool:
    """The tcgen05 contraction writes the mirrored copies of rows 1 / H-2 and columns 1 / W-2 into a 1-pixel reflected
    ring of its output itself (conv_tc.cu, TcParams::ring) — whole-plane outputs only (a sub-pixel phase or a window
    does not own the ring)."""
    if os.environ.get("LAMA_B200_RING_KERNEL", "0") == "1":        # A/B: always refresh rings with the ring kernel
        return False
    return (prog.math == L.MATH_BF16X3 and isinstance(op, ConvOp) and op.out.phase is None and op.out.win is None
            and not op.out.window and op.out.buf.pad == 1 and op.out.buf.fmt == L.BF16X2 and op.out.buf.H >= 4
            and op.out.buf.W >= 4)


# ------------------------------------------------------------------------------------- executor
def op_views(op) -> Tuple[List[TV], List[TV]]:
    """(views read, views written) by one op of a program — the basis of the buffer liveness analysis."""
    if isinstance(op, (ToNHWC, StemOp, StemPackOp, StemPackU8Op)):
        return [], [op.out]
    if isinstance(op, (ToNCHW, HeadOp)):
        return [op.inp], []
    if isinstance(op, (HeadGatherOp, HeadGatherU8Op)):
        return [op.q], []
    if isinstance(op, ConvOp):
        return [t for t in op.ins if t is not None] + ([op.addend] if op.addend is not None else []), [op.out]
    if isinstance(op, RfftOp):
        return [op.inp], [op.spec]
    if isinstance(op, IrfftOp):
        return [op.spec] + ([op.residual] if op.residual is not None else []), [op.out]
    if isinstance(op, BorderOp):
        return [op.view], [op.view]
    if isinstance(op, ReluBwdOp):
        return [op.dy, op.y], [op.out]
    if isinstance(op, FoldOp):
        return [op.gpad] + [tv for tv, _c0 in op.addends], [op.out]
    if isinstance(op, SplitOp):
        return [], []
    raise TypeError(op)


def storage_key(b: Buf) -> tuple:
    """Buffers with equal keys have byte-identical storage (dtype, shape, ring, layout)."""
    return (b.fmt, b.B, b.H, b.W, b.C, b.pad, b.reflect_border, b.cg, b.tile)


def assign_storage_slots(prog: Program) -> Dict[str, int]:
    """Buffer name -> storage slot.  A program is a straight-line op list on one stream, so a buffer is dead after
    the last op that touches it and its storage can back a later buffer.  Only buffers with IDENTICAL storage
    (``storage_key``) share a slot: whatever a kernel leaves unwritten (zero-initialised pixel / channel padding, the
    reflected ring before its producer ran) then holds what the same kind of buffer held there before, never foreign
    bits.  For big-lama this folds the 18 residual blocks' ~150 buffers onto two blocks' worth: 28 GB -> 10 GB at bs32
    512x512, and bs64 1024x1024 (BASELINE config 4 on one GPU) fits a 180 GB B200 at all.  Constant buffers
    (``prog.consts``) keep their own storage; ``LAMA_B200_POOL=0`` gives every buffer its own."""
    first: Dict[str, int] = {}
    last: Dict[str, int] = {}
    for i, op in enumerate(prog.ops):
        reads, writes = op_views(op)
        for tv in reads + writes:
            first.setdefault(tv.buf.name, i)
            last[tv.buf.name] = i
    pooling = os.environ.get("LAMA_B200_POOL", "1") != "0"
    slots: Dict[str, int] = {}
    free_at: List[Tuple[tuple, int]] = []          # per slot: (storage key, index of the last op that touches it)
    for b in sorted(prog.bufs, key=lambda bb: first.get(bb.name, -1)):
        key = storage_key(b)
        reuse = None
        if pooling and b.name in first and b.name not in prog.consts:
            for si, (k, until) in enumerate(free_at):
                if k == key and until is not None and until < first[b.name]:
                    reuse = si
                    break
        if reuse is None:
            reuse = len(free_at)
            free_at.append((key, None))
        # constants and never-touched buffers hold their slot for good
        free_at[reuse] = (key, last[b.name] if (b.name in last and b.name not in prog.consts) else None)
        slots[b.name] = reuse
    return slots


class CudaExecutor:
    """Binds a Program to device buffers and pre-built C-ABI calls."""

    def __init__(self, prog: Program, device: torch.device):
        self.prog = prog
        self.device = device
        self.lib = L.get_lib()
        self.dev_index = device.index if device.index is not None else torch.cuda.current_device()
        L.check(self.lib.ffcb_check_device(self.dev_index), "ffcb_check_device")
        self.storage: Dict[str, torch.Tensor] = {}
        self.slots = assign_storage_slots(prog)          # buffers whose lifetimes do not overlap share storage
        slot_tensor: Dict[int, torch.Tensor] = {}
        for b in prog.bufs:
            si = self.slots[b.name]
            if si in slot_tensor:
                self.storage[b.name] = slot_tensor[si]
                continue
            shape = (b.B, b.H + 2 * b.pad, b.W + 2 * b.pad, b.C)
            if b.cg:
                shape = (b.C // b.cg, b.B, b.H, b.W, b.cg)
            if b.tile:
                shape = (-(-(b.B * b.H * b.W) // 128), b.C // 8, 128, 8)      # zero-initialised: the tail block stays finite
            if b.fmt == L.F32:
                slot_tensor[si] = torch.empty(shape, dtype=torch.float32, device=device)
            else:
                slot_tensor[si] = torch.zeros((2,) + shape, dtype=torch.bfloat16, device=device)
            self.storage[b.name] = slot_tensor[si]
        self.storage_bytes = sum(t.numel() * t.element_size() for t in slot_tensor.values())
        for name, val in prog.consts.items():
            assert self.storage[name].dtype == torch.float32 and tuple(self.storage[name].shape) == tuple(val.shape)
            self.storage[name].copy_(val.to(device=device, dtype=torch.float32))
        ws_bytes = prog.fft_workspace_bytes()
        self.ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=device)
        self.ws_bytes = ws_bytes
        self.outputs = {k: torch.empty(v, dtype=prog.dtypes.get(k, torch.float32), device=device)
                        for k, v in prog.outputs.items()}
        self._launches = None
        self._keep = []          # ctypes objects / tensors that must outlive the calls
        self.calls = []          # (fn, args) with a trailing stream argument appended at run time
        self.input_slots: Dict[str, List[Tuple[int, int]]] = {}   # input name -> [(call idx, arg idx)]
        self.split = None        # forward+backward programs: index of the first backward call (SplitOp)
        self.generation = 0      # bumped by every forward part: a stale backward must not read newer activations
        for op in prog.ops:
            self._bind(op)

    # -- view construction
    def tensor(self, tv: TV) -> L.Tensor:
        b = tv.buf
        st = self.storage[b.name]
        es = 4 if b.fmt == L.F32 else 2
        if b.tile:
            assert (tv.phase is None and not tv.window and tv.win is None and b.pad == 0 and tv.c0 % 8 == 0
                    and tv.channels % 8 == 0)
            m0 = tv.b0 * b.H * b.W
            if m0 % 128:
                raise ValueError("a batch slice of a tile-blocked buffer must start on a 128-pixel block")
            t = L.Tensor()
            nblk, groups = -(-(b.B * b.H * b.W) // 128), b.C // 8
            t.cg, t.tile, t.sg = 8, 128, groups * 1024
            t.sx, t.sy, t.sb = 8, b.W * 8, b.H * b.W * 8
            t.lo_off = nblk * groups * 1024
            t.ptr = st.data_ptr() + ((tv.c0 // 8) * 1024 + (m0 // 128) * t.sg) * es
            t.B, t.H, t.W, t.C = tv.batch, b.H, b.W, tv.channels
            t.fmt, t.pad, t.reflect_border = b.fmt, 0, 0
            return t
        if b.cg:
            assert tv.phase is None and not tv.window and b.pad == 0 and tv.c0 % b.cg == 0 and tv.channels % b.cg == 0
            t = L.Tensor()
            t.sx, t.sy, t.sb = b.cg, b.W * b.cg, b.H * b.W * b.cg
            t.sg, t.cg = b.B * b.H * b.W * b.cg, b.cg
            t.lo_off = b.C * b.B * b.H * b.W if b.fmt == L.BF16X2 else 0
            off = (tv.c0 // b.cg) * t.sg + tv.b0 * t.sb
            h, w = b.H, b.W
            if tv.win is not None:
                y0, x0, h, w = tv.win
                off += y0 * t.sy + x0 * t.sx
            t.ptr = st.data_ptr() + off * es
            t.B, t.H, t.W, t.C = tv.batch, h, w, tv.channels
            t.fmt, t.pad, t.reflect_border = b.fmt, 0, 0
            return t
        wp, hp = b.W + 2 * b.pad, b.H + 2 * b.pad
        sx, sy, sb = b.C, wp * b.C, hp * wp * b.C
        off = (b.pad * wp + b.pad) * b.C + tv.c0
        h, w = b.H, b.W
        if tv.phase is not None:
            a, bb = tv.phase
            off += a * sy + bb * sx
            sy, sx, h, w = 2 * sy, 2 * sx, b.H // 2, b.W // 2
        if tv.win is not None:
            assert tv.phase is None and not tv.window
            y0, x0, h, w = tv.win
            off += y0 * sy + x0 * sx
        t = L.Tensor()
        t.ptr = 0  # set below (after the batch-slice offset)
        t.sb, t.sy, t.sx = sb, sy, sx
        t.lo_off = b.B * hp * wp * b.C if b.fmt == L.BF16X2 else 0
        off += tv.b0 * sb
        t.ptr = st.data_ptr() + off * es
        t.B, t.H, t.W, t.C = tv.batch, h, w, tv.channels
        if tv.bcast:
            t.sb = 0                 # every image of the batch reads the same plane (position-dependent addends)
        t.fmt, t.pad, t.reflect_border = b.fmt, b.pad, b.reflect_border
        if tv.phase is not None or tv.win is not None:     # not a whole image: no ring semantics
            t.pad, t.reflect_border = 0, 0
        if tv.window:                # pixel x exposes the buf.C * window contiguous elements starting at x * sx
            t.W, t.C, t.window = b.W - tv.window, b.C * tv.window, 1
        return t

    def _ref(self, obj):
        self._keep.append(obj)
        return obj

    def _dev(self, t: torch.Tensor) -> torch.Tensor:
        """Packed parameter on the executor's device, kept alive with the executor."""
        t = t.to(self.device).contiguous()
        self._keep.append(t)
        return t

    def _bind(self, op):
        lib = self.lib
        if isinstance(op, ToNHWC):
            bb, c, h, w = self.prog.inputs[op.src]
            t = self._ref(self.tensor(op.out))
            self.input_slots.setdefault(op.src, []).append((len(self.calls), 0))
            self.calls.append(("ffcb_nchw_to_nhwc", lib.ffcb_nchw_to_nhwc, [None, bb, c, h, w, C.byref(t)]))
        elif isinstance(op, ToNCHW):
            t = self._ref(self.tensor(op.inp))
            self.calls.append(("ffcb_nhwc_to_nchw", lib.ffcb_nhwc_to_nchw,
                               [C.byref(t), self.outputs[op.dst].data_ptr()]))
        elif isinstance(op, StemOp):
            bb, c, h, w = self.prog.inputs[op.src]
            t = self._ref(self.tensor(op.out))
            wd, sd = self._dev(op.w), self._dev(op.shift)
            self.input_slots.setdefault(op.src, []).append((len(self.calls), 0))
            self.calls.append(("ffcb_stem_conv7", lib.ffcb_stem_conv7,
                               [None, bb, c, h, w, wd.data_ptr(), sd.data_ptr(), wd.shape[1], C.byref(t)]))
        elif isinstance(op, StemPackOp):
            bb, c, h, w = self.prog.inputs[op.src]
            t = self._ref(self.tensor(op.out))
            self.input_slots.setdefault(op.src, []).append((len(self.calls), 0))
            self.calls.append(("ffcb_stem_pack", lib.ffcb_stem_pack, [None, bb, c, h, w, C.byref(t)]))
        elif isinstance(op, StemPackU8Op):
            bb = self.prog.inputs[op.img][0]
            t = self._ref(self.tensor(op.out))
            self.input_slots.setdefault(op.img, []).append((len(self.calls), 0))
            self.input_slots.setdefault(op.mask, []).append((len(self.calls), 1))
            self.calls.append(("ffcb_stem_pack_u8", lib.ffcb_stem_pack_u8, [None, None, bb, op.h0, op.w0, C.byref(t)]))
        elif isinstance(op, HeadGatherU8Op):
            t = self._ref(self.tensor(op.q))
            bd = self._dev(op.bias)
            self.input_slots.setdefault(op.img, []).append((len(self.calls), 3))
            self.input_slots.setdefault(op.mask, []).append((len(self.calls), 4))
            self.calls.append(("ffcb_head_gather7_blend_u8", lib.ffcb_head_gather7_blend_u8,
                               [C.byref(t), bd.data_ptr(), op.act, None, None, op.h0, op.w0,
                                self.outputs[op.dst].data_ptr()]))
        elif isinstance(op, HeadOp):
            t = self._ref(self.tensor(op.inp))
            wd, bd = self._dev(op.w), self._dev(op.bias)
            self.calls.append(("ffcb_head_conv7", lib.ffcb_head_conv7,
                               [C.byref(t), wd.data_ptr(), bd.data_ptr(), op.n_out, op.act,
                                self.outputs[op.dst].data_ptr()]))
        elif isinstance(op, HeadGatherOp):
            t = self._ref(self.tensor(op.q))
            bd = self._dev(op.bias)
            self.calls.append(("ffcb_head_gather7", lib.ffcb_head_gather7,
                               [C.byref(t), bd.data_ptr(), op.n_out, op.act, self.outputs[op.dst].data_ptr()]))
        elif isinstance(op, ConvOp):
            d = self._ref(L.ConvDesc())
            pk = op.packed
            d.inp[0] = self.tensor(op.ins[0])
            if op.ins[1] is not None:
                d.inp[1] = self.tensor(op.ins[1])
            d.out = self.tensor(op.out)
            if op.addend is not None:
                d.addend = self.tensor(op.addend)
            if self.prog.math == L.MATH_BF16X3:
                wt = pk.split_weights()
            else:
                wt = pk.w_kn
            d.weight = self._dev(wt).data_ptr()
            if pk.shift is not None:
                d.shift = self._dev(pk.shift).data_ptr()
            d.n_out, d.stride, d.border, d.act = pk.n_out, pk.stride, pk.border, pk.act
            d.nseg, d.math, d.addend_post = len(pk.segs), self.prog.math, int(op.addend_post)
            for i, s in enumerate(pk.segs):
                d.seg[i] = L.KSeg(s.src, s.dy, s.dx, s.c0, s.nch)
            self.calls.append(("ffcb_conv:" + op.tag, lib.ffcb_conv, [C.byref(d)]))
        elif isinstance(op, SplitOp):
            self.split = len(self.calls)
        elif isinstance(op, ReluBwdOp):
            a, y, o = (self._ref(self.tensor(v)) for v in (op.dy, op.y, op.out))
            self.calls.append(("ffcb_relu_bwd", lib.ffcb_relu_bwd, [C.byref(a), C.byref(y), C.byref(o)]))
        elif isinstance(op, FoldOp):
            g, o = self._ref(self.tensor(op.gpad)), self._ref(self.tensor(op.out))
            adds = [(C.byref(self._ref(self.tensor(tv))), c0) for tv, c0 in op.addends] + [(None, 0)] * 2
            self.calls.append(("ffcb_fold_reflect_border", lib.ffcb_fold_reflect_border,
                               [C.byref(g), adds[0][0], adds[0][1], adds[1][0], adds[1][1], C.byref(o)]))
        elif isinstance(op, BorderOp):
            t = self._ref(self.tensor(op.view))
            self.calls.append(("ffcb_fill_reflect_border", lib.ffcb_fill_reflect_border, [C.byref(t)]))
        elif isinstance(op, RfftOp):
            a, s = self._ref(self.tensor(op.inp)), self._ref(self.tensor(op.spec))
            self.calls.append(("ffcb_rfft2", lib.ffcb_rfft2, [C.byref(a), C.byref(s), self.ws.data_ptr(), self.ws_bytes]))
        elif isinstance(op, IrfftOp):
            s, o = self._ref(self.tensor(op.spec)), self._ref(self.tensor(op.out))
            r = C.byref(self._ref(self.tensor(op.residual))) if op.residual is not None else None
            self.calls.append(("ffcb_irfft2", lib.ffcb_irfft2, [C.byref(s), r, C.byref(o), self.ws.data_ptr(), self.ws_bytes]))
        else:
            raise TypeError(op)

    def run(self, inputs: Dict[str, torch.Tensor], stream: Optional[int] = None,
            part: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Issue every call on ``stream`` (default: torch's current stream).  Outputs are the
        executor's own tensors (overwritten by the next run).  ``part``: 0 / 1 run only the forward / backward half
        of a forward+backward program (inputs of the other half may be absent)."""
        if torch.cuda.current_device() != self.dev_index:
            # the module lives on another GPU than the caller's current device (one process driving several GPUs):
            # kernels must be launched with that device current, as torch's own ops do through their device guards
            with torch.cuda.device(self.dev_index):
                return self.run(inputs, stream, part)
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        if part is not None:
            assert self.split is not None, "not a forward+backward program"
            lo, hi = (0, self.split) if part == 0 else (self.split, len(self.calls))
            for name, slots in self.input_slots.items():
                if name in inputs:
                    for ci, ai in slots:
                        self.calls[ci][2][ai] = inputs[name].data_ptr()
            for name, fn, args in self.calls[lo:hi]:
                rc = fn(*args, stream)
                if rc != 0:
                    L.check(rc, name)
            if part == 0:
                self.generation += 1
            return self.outputs
        for name, slots in self.input_slots.items():
            t = inputs[name]
            dt = self.prog.dtypes.get(name, torch.float32)
            assert t.is_cuda and t.dtype == dt and t.is_contiguous() and tuple(t.shape) == tuple(
                self.prog.inputs[name]), f"input {name}: expected contiguous {dt} {self.prog.inputs[name]}"
            for ci, ai in slots:
                self.calls[ci][2][ai] = t.data_ptr()
        first = self._launches is None
        if first:
            self.lib.ffcb_reset_launch_count()
        for name, fn, args in self.calls:
            rc = fn(*args, stream)
            if rc != 0:
                L.check(rc, name)
        if first:
            self._launches = int(self.lib.ffcb_launch_count())
        return self.outputs

    @property
    def launches_per_run(self) -> int:
        """Kernel launches of one replay, counted by the library itself during the first run."""
        assert self._launches is not None, "run the executor once first"
        return self._launches


class GraphedProgram:
    """CUDA-graph replay of an executor with static input tensors (bench / serving path)."""

    def __init__(self, ex: CudaExecutor, warmup: int = 2):
        self.ex = ex
        self.static_in = {k: torch.empty(v, dtype=ex.prog.dtypes.get(k, torch.float32), device=ex.device)
                          for k, v in ex.prog.inputs.items()}
        side = torch.cuda.Stream(device=ex.device)
        side.wait_stream(torch.cuda.current_stream(ex.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                ex.run(self.static_in)
        torch.cuda.current_stream(ex.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            ex.run(self.static_in)

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        for k, t in inputs.items():
            self.static_in[k].copy_(t, non_blocking=True)
        self.graph.replay()
        return self.ex.outputs


# ---------------------------------------------------------------------------- module entry point
# Executor caches live OUTSIDE the module (weak keys): they hold ctypes pointers / byref objects, which must never
# end up in `module.__dict__` — the reference deep-copies the generator for its EMA copy (trainers/base.py:168) and
# `torch.save(module)` / pickling / DataParallel replication walk `__dict__`.
import weakref

_PROGRAMS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()      # module -> {key: (signature, executor)}
_TENSORS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()       # module -> [flat tensor list, calls since walk]
_REWALK_EVERY = 8
_SMALL_MODULE = 64       # modules with fewer tensors are re-walked on every call (the walk is cheap there)


def invalidate(module) -> None:
    """Drop every cached program / packed weight of ``module`` (call after editing weights in a way the automatic
    checks cannot see; load_state_dict, .to(), in-place ops and ``.data`` edits ARE seen)."""
    _PROGRAMS.pop(module, None)
    _TENSORS.pop(module, None)


def _content_checksum(tensors) -> float:
    """One number per weight version: the sum of all tensors' L2 norms (a handful of fused multi-tensor kernels and
    one scalar read).  Catches what (data_ptr, _version) cannot: edits through ``.data`` — the reference's EMA update
    (trainers/base.py:40) and the common ``weight.data.copy_()`` loading idiom leave ``_version`` unchanged."""
    fl = [t for t in tensors if t.is_floating_point() and t.numel()]
    if not fl:
        return 0.0
    groups = {}
    for t in fl:
        groups.setdefault((t.device, t.dtype), []).append(t.detach())
    total = 0.0
    for ts in groups.values():
        total += float(torch.stack(torch._foreach_norm(ts)).double().sum())
    return total


def _weights_signature(module, content: bool = True) -> Tuple:
    """(data_ptr, version) of every parameter / buffer — changes on load_state_dict, .to(), in-place edits — plus,
    with ``content``, a checksum of the values (``.data`` edits).  Walking the module tree costs ~1.5 ms for big-lama
    (989 tensors), so the flat tensor list is cached and re-walked every few calls (every call for small modules);
    a replaced Parameter object whose storage was freed changes data_ptr and is seen at once in practice.
    LAMA_B200_TRUST_WEIGHTS=1 skips the checksum (serving loops that never touch the weights)."""
    st = _TENSORS.get(module)
    if st is None or st[1] >= _REWALK_EVERY or len(st[0]) <= _SMALL_MODULE:
        st = [list(module.parameters()) + list(module.buffers()), 0]
        _TENSORS[module] = st
    st[1] += 1
    sig = tuple((t.data_ptr(), t._version) for t in st[0])
    if content and os.environ.get("LAMA_B200_TRUST_WEIGHTS", "0") != "1":
        sig = sig + (_content_checksum(st[0]),)
    return sig


def get_executor(module, kind: str, tensors, math: Optional[int] = None,
                 device: Optional[torch.device] = None) -> CudaExecutor:
    """``tensors`` only contribute their shapes (meta tensors are fine when ``device`` is given)."""
    math = default_math() if math is None else math
    shapes = tuple(tuple(t.shape) if torch.is_tensor(t) else None for t in tensors)
    dev = device if device is not None else next(t for t in tensors if torch.is_tensor(t)).device
    key = (kind, shapes, str(dev), math)
    cache = _PROGRAMS.setdefault(module, {})
    sig = _weights_signature(module)
    hit = cache.get(key)
    if hit is not None and hit[0] == sig:
        cache[key] = cache.pop(key)          # LRU: most recently used last
        return hit[1]
    with torch.no_grad():
        prog = build_module_program(module, kind, shapes, math)
    ex = CudaExecutor(prog, dev)
    cache.pop(key, None)                      # stale weights
    # every executor owns its activation buffers (~0.3 GB per 512x512 image for big-lama): keep only the few most
    # recently used shapes per module so that a stream of differently sized images cannot exhaust HBM
    while len(cache) >= PROGRAM_CACHE_SIZE:
        cache.pop(next(iter(cache)))
    cache[key] = (sig, ex)
    return ex


class _BlockGradFn(torch.autograd.Function):
    """FFCResnetBlock with native forward AND native input gradients: what the reference's
    refinement loop (evaluation/refinement.py:137-167) needs — it optimises the block inputs, the weights are frozen.
    Forward runs the first half of a ``resnet_block_grad`` program (activations stay in the executor's buffers),
    backward the second half.  One executor per (module, shape): a second forward before the backward of the first
    would overwrite those activations, which raises instead of returning wrong gradients."""

    @staticmethod
    def forward(ctx, module, x_l, x_g):
        ex = get_executor(module, "resnet_block_grad", (x_l, x_g))
        xl, xg = x_l.detach().contiguous(), x_g.detach().contiguous()
        outs = ex.run({"x0": xl, "x1": xg}, part=0)
        ctx.ex, ctx.generation = ex, ex.generation
        return xl + outs["y0"], xg + outs["y1"]

    @staticmethod
    def backward(ctx, g_l, g_g):
        ex = ctx.ex
        if ex.generation != ctx.generation:
            raise RuntimeError("lama_b200: the same FFCResnetBlock ran forward again (same shape) before this backward; "
                               "its saved activations were overwritten")
        outs = ex.run({"g0": g_l.contiguous(), "g1": g_g.contiguous()}, part=1)
        return None, outs["dx0"].clone(), outs["dx1"].clone()


def block_with_input_grad(module, x_l, x_g):
    """(out_l, out_g) of an FFCResnetBlock, differentiable w.r.t. x_l / x_g on the native path."""
    return _BlockGradFn.apply(module, x_l, x_g)


def run_module(module, kind: str, tensors):
    """Execute ``module`` natively on NCHW float CUDA tensors; returns fresh tensors (or the int 0
    for an empty FFC side)."""
    first = next(t for t in tensors if torch.is_tensor(t))
    if first.shape[0] == 0:
        # empty batch (the reference returns empty tensors of the right shape): nothing to launch; the output
        # shapes come from the program of a one-image batch
        shapes = tuple((1,) + tuple(t.shape[1:]) if torch.is_tensor(t) else None for t in tensors)
        with torch.no_grad():
            prog = build_module_program(module, kind, shapes, L.MATH_FP32)
        outs = {k: torch.empty((0,) + tuple(v[1:]), dtype=prog.dtypes.get(k, torch.float32), device=first.device)
                for k, v in prog.outputs.items()}
        y0 = outs.get("y0", 0)
        if kind in ("ffc_bn_act", "resnet_block"):
            return y0, outs.get("y1", 0)
        return (y0,)
    ex = get_executor(module, kind, tensors)
    feed = {}
    i = 0
    for t in tensors:
        if torch.is_tensor(t):
            feed[f"x{i}"] = t.contiguous()
            i += 1
    outs = ex.run(feed)
    y0 = outs["y0"].clone() if "y0" in outs else 0
    if kind in ("ffc_bn_act", "resnet_block"):
        return y0, (outs["y1"].clone() if "y1" in outs else 0)
    return (y0,)
