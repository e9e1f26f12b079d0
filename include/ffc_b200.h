/*
 * ffc_b200.h — C ABI of libffc_b200.so: the B200 (sm_100a) kernels behind the drop-in
 * replacements for advimman/lama's FFC inference path
 * (reference: saicinpainting/training/modules/ffc.py).
 *
 * The reference has no FFI: its seam is the Python nn.Module surface, and every FLOP runs
 * inside torch (cuFFT / cuDNN / ATen).  This library is what a maintainer binds *instead of*
 * those torch calls; each entry point names the reference lines it replaces.  The binding
 * itself (ctypes) is lama_b200/_lib.py and is documented in INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no C++/torch types, no exceptions.
 *  - every function returns 0 (FFCB_OK) or a negative FFCB_E* code; the message of the last
 *    failure on the calling thread is ffcb_last_error().
 *  - all device pointers are owned by the caller (torch's caching allocator in the Python
 *    binding) and must stay alive until the stream work completes.  The library allocates
 *    nothing (FFT twiddles are generated in shared memory by the kernels themselves).
 *  - every launch goes to the caller's stream; no call synchronises or allocates, so all
 *    entry points are CUDA-graph capturable (run each op once eagerly first: kernels that need
 *    more than 48 KB of shared memory set their function attribute on first use).
 *  - activations inside the path are channels-last ("NHWC"): element (b, y, x, c) of a tensor
 *    lives at ptr + b*sb + y*sy + x*sx + c.  Strides let one allocation hold a reflect-padded
 *    plane ([B][H+2][W+2][C], ptr at the interior origin) or a channel slice of a wider tensor
 *    (the local|global halves of an FFC feature map share one 512-channel buffer).
 *    Tensors of the FourierUnit chain may instead be "channel-group planar" (ffcb_tensor.cg / .sg below).
 *  - storage formats: FFCB_F32 (float) and FFCB_BF16X2 ("split" bfloat16: value = hi + lo with
 *    hi = bf16(v), lo = bf16(v - hi); hi plane at ptr, lo plane at ptr + lo_off elements).
 *    The split format is what the tcgen05 path multiplies (3 bf16 products, fp32 accumulate:
 *    hi*hi + lo*hi + hi*lo, relative error ~2^-16, see DESIGN.md "precision").
 */
#ifndef FFC_B200_H_
#define FFC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFCB_VERSION 111 /* 0.1.1: ffcb_tensor gained the channel-group fields cg / tile / sg */

enum {
  FFCB_OK = 0,
  FFCB_EINVAL = -1,  /* bad shape / alignment / unsupported combination */
  FFCB_EARCH = -2,   /* device is not sm_100 */
  FFCB_ECUDA = -3,   /* CUDA runtime / driver error (text in ffcb_last_error) */
  FFCB_ENOMEM = -4   /* caller-provided workspace too small */
};

enum { FFCB_F32 = 0, FFCB_BF16X2 = 1 };
enum { FFCB_ACT_NONE = 0, FFCB_ACT_RELU = 1, FFCB_ACT_SIGMOID = 2, FFCB_ACT_TANH = 3 };
enum { FFCB_BORDER_ZERO = 0, FFCB_BORDER_REFLECT = 1 };
/* arithmetic of the contraction kernels */
enum {
  FFCB_MATH_FP32 = 0,   /* CUDA-core FFMA, fp32 operands (reference-grade path) */
  FFCB_MATH_BF16X3 = 1  /* tcgen05.mma kind::f16 on split-bf16 operands, fp32 accumulators in TMEM */
};

typedef void* ffcb_stream_t; /* cudaStream_t */

/* Channels-last tensor view.  Strides are in elements of the storage type. */
typedef struct {
  void* ptr;       /* element (0,0,0,0); for FFCB_BF16X2 the hi plane */
  int64_t sb, sy, sx;
  int64_t lo_off;  /* FFCB_BF16X2: offset (elements) from the hi to the lo plane */
  int32_t B, H, W, C;
  int32_t fmt;     /* FFCB_F32 | FFCB_BF16X2 */
  int32_t pad;     /* physical border pixels around the interior (0..3).  With reflect_border != 0 the ring holds
                      the reflected image (ffcb_fill_reflect_border), so that the TMA tile of tap (dy,dx),
                      |dy|,|dx| <= pad, is the output tile shifted by (dx,dy) — no index math. */
  int32_t reflect_border;
  int32_t window;  /* != 0: "sliding window" view — consecutive pixels overlap (sx < C): pixel x exposes the C
                      contiguous elements starting at x*sx.  Used to feed the 7x7 stem to ffcb_conv as 7 K-segments
                      of (8 taps x 8 channels) read straight out of a packed NHWC8 image (see ffcb_stem_pack). */
  int32_t cg;      /* 0: plain channels-last.  > 0: "channel-group planar" — channels are stored in groups of cg
                      (4 or 8): element (b,y,x,c) lives at ptr + (c/cg)*sg + b*sb + y*sy + x*sx + (c%cg), i.e. every
                      group of cg channels is its own dense little channels-last image.  This is the layout of the
                      FourierUnit chain (SpectralTransform.conv1 -> rfft2 -> spectral conv -> irfft2 -> conv2,
                      ffc.py:145-161): one (image, group) plane set is ONE contiguous block for the plane FFT kernels
                      and the [K/8][pixel][8] "interleaved" (no-swizzle, K-major) operand tile of tcgen05.mma. */
  int32_t tile;    /* 0, or 128 with cg == 8: "tile-blocked" variant for tcgen05 operands — pixels are flattened
                      (m = (b*H + y)*W + x over the view) and stored in blocks of 128:
                        element (m, c) at ptr + (m/128)*sg + (c/8)*1024 + (m%128)*8 + c%8
                      so the [8 groups][128 pixels][8] operand tile of one 64-channel K block of one 128-pixel M tile is
                      ONE contiguous 16 KB run (a single cp.async.bulk per plane); sg = elements per 128-pixel block
                      (all groups of the allocation), sb / sy / sx are ignored.  Written by the plane FFT kernels,
                      read by ffcb_conv (tcgen05 arm; 1x1 taps; M tiles that coincide with the blocks). */
  int64_t sg;      /* cg > 0: stride between channel groups (elements), or between 128-pixel blocks when tile != 0 */
} ffcb_tensor;

/* One K-segment of an implicit-GEMM convolution: `nch` input channels starting at channel
 * `c0` of input tensor `src`, sampled at input pixel (y*stride + dy, x*stride + dx). */
typedef struct {
  int32_t src, dy, dx, c0, nch;
} ffcb_kseg;

#define FFCB_MAX_KSEG 64

/*
 * v            = sum_seg sum_k in[seg.src][b, y*stride+dy, x*stride+dx, c0+k] * W[n][koff(seg)+k] + shift[n]
 * out[b,y,x,n] = addend_post ? act(v) + addend[b,y,x,n]      (residual: id + act(bn(conv)), ffc.py:288)
 *                            : act(v + addend[b,y,x,n])
 *
 * Replaces, with BatchNorm folded into W/shift by the caller (lama_b200/packing.py):
 *   nn.Conv2d k in {1,3} of FFC.convl2l/convl2g/convg2l            ffc.py:189-196, 221, 223
 *   SpectralTransform.conv1 (1x1 + BN + ReLU) and .conv2 (1x1)     ffc.py:128-133, 139-140, 145, 161
 *   FourierUnit.conv_layer + bn + relu on the interleaved spectrum ffc.py:57-61, 100-101
 *   FFC_BN_ACT.bn_l/bn_g + act                                     ffc.py:243-249, 253-254
 *   the residual add of FFCResnetBlock                             ffc.py:288
 *   nn.ConvTranspose2d(k3,s2,p1,op1)+BN+ReLU as four sub-pixel phases  ffc.py:350-354
 * The out view's H,W are the output grid; input coordinates outside the input's interior are
 * zero (FFCB_BORDER_ZERO) or reflected without edge repeat (FFCB_BORDER_REFLECT, ffc.py:189
 * padding_mode='reflect').
 * Output ring (FFCB_MATH_BF16X3): when `out` is a whole split-bf16 plane with a 1-pixel reflected ring
 * (out.reflect_border != 0, out.pad == 1, H, W >= 4) the kernel also writes the mirrored copies of rows 1 / H-2 and columns 1 / W-2 into the
 * ring, so the result can feed the next 3x3 reflect contraction without ffcb_fill_reflect_border.  The FP32 arm and
 * every other producer leave the ring to ffcb_fill_reflect_border.
 *
 * weight: FFCB_MATH_FP32  -> float  [Ktot][N]           (N contiguous)
 *         FFCB_MATH_BF16X3 -> bf16  [2][N][Ktot] hi|lo   (K contiguous), Ktot = sum nch
 */
typedef struct {
  ffcb_tensor in[2];
  ffcb_tensor out;
  ffcb_tensor addend;   /* addend.ptr == NULL: none */
  const void* weight;
  const float* shift;   /* [N] or NULL */
  int32_t n_out;        /* N */
  int32_t stride;       /* 1 or 2 */
  int32_t border;       /* FFCB_BORDER_* */
  int32_t act;          /* FFCB_ACT_* */
  int32_t nseg;
  int32_t math;         /* FFCB_MATH_* */
  int32_t addend_post;  /* 0: addend joins the pre-activation sum; 1: added after the activation */
  int32_t _reserved;
  ffcb_kseg seg[FFCB_MAX_KSEG];
} ffcb_conv_desc;

int ffcb_version(void);
const char* ffcb_last_error(void);
/* 0 if `device` is an sm_100 part this library can run on */
int ffcb_check_device(int device);
void ffcb_shutdown(void);

/* Generic fused convolution / pointwise contraction (see ffcb_conv_desc). */
int ffcb_conv(const ffcb_conv_desc* desc, ffcb_stream_t stream);

/*
 * Stem: ReflectionPad2d(3) + Conv2d(Cin -> N, k7, no bias) + folded BN + ReLU.
 * ffc.py:315-317 (FFC_BN_ACT with ratio 0/0 -> convl2l only) and :253.
 * x: NCHW float [B][Cin][H][W] contiguous (what the caller of the generator passes);
 * w: float [7*7*Cin][N] (k index = (ky*7+kx)*Cin + c); out: channels-last view.
 */
int ffcb_stem_conv7(const float* x_nchw, int B, int Cin, int H, int W, const float* w, const float* shift,
                    int N, const ffcb_tensor* out, ffcb_stream_t stream);

/*
 * Stem, tensor-core form.  ffcb_stem_pack writes the generator input as a reflect-padded (3 pixels) channels-last
 * image with 8 channels per pixel (Cin real + zeros), rows of W+8 pixels (the tail is zero), in split bf16:
 *     packed[b][yp][xp][c],  yp in [0,H+6), xp in [0,W+8),  = x[b][c][reflect(yp-3)][reflect(xp-3)]      (c < Cin)
 * and, when Cin <= 4 ("two-row" packing), packed[b][yp][xp][4+c] = packed[b][yp+1][xp][c] (zero below the last row).
 * A window view of it (C = 64, sx = 8, window = 1) exposes, at pixel x, the 8 taps x 8 channels of one kernel
 * row (two kernel rows with the two-row packing) as ONE contiguous 128-byte K block, so ReflectionPad2d(3) +
 * Conv2d(k7) (ffc.py:315-317) becomes an ffcb_conv with seven K-segments (dy = 0..6, dx = 0) and zero-padded
 * weights [N][7][8 taps][8 channels] — or four K-segments (dy = 0, 2, 4, 6) with weights
 * [N][4][8 taps][row dy | row dy+1][4 channels] when Cin <= 4 (43 % fewer tensor-core MACs).
 */
int ffcb_stem_pack(const float* x_nchw, int B, int Cin, int H, int W, const ffcb_tensor* packed, ffcb_stream_t stream);

/*
 * Head: ReflectionPad2d(3) + Conv2d(C -> N<=4, k7, bias) + activation, NCHW float output.
 * ffc.py:360-363.  w: float [N][7*7][C]; bias [N]; y: [B][N][H][W].
 */
int ffcb_head_conv7(const ffcb_tensor* in, const float* w, const float* bias, int N, int act, float* y_nchw,
                    ffcb_stream_t stream);

/*
 * Real 2-D FFT pair, norm='ortho', over the (H, W) axes of a channels-last tensor.
 *   ffcb_rfft2 : torch.fft.rfftn(x, dim=(-2,-1), norm='ortho') + the stack/permute/view that
 *                interleaves Re/Im as channels 2k / 2k+1                       ffc.py:86-89
 *                in (B,H,W,C) real -> spec (B,H,W/2+1,2C)
 *   ffcb_irfft2: the inverse view/permute/complex + torch.fft.irfftn(s=(H,W), norm='ortho'),
 *                with the residual of SpectralTransform fused: out = residual + irfft2(spec)
 *                                                                     ffc.py:103-108, 161
 *                spec (B,H,W/2+1,2C) float -> out (B,H,W,C)
 * The inverse transforms along H first (all W/2+1 columns, complex) and then C2R along W,
 * ignoring Im of the k_w=0 and (even W) k_w=W/2 bins — exactly what torch/cuFFT/MKL do for the
 * non-Hermitian post-ReLU spectrum.
 * Power-of-two H, W in [4, 256] take the shared-memory Stockham kernels; any other size takes
 * a direct-DFT kernel (exact same results, O(n^2)).
 * ws: caller workspace of ffcb_fft2_workspace_bytes(B,H,W,C) bytes (row-pass intermediate).
 */
size_t ffcb_fft2_workspace_bytes(int B, int H, int W, int C);
int ffcb_rfft2(const ffcb_tensor* in, const ffcb_tensor* spec, void* ws, size_t ws_bytes, ffcb_stream_t stream);
int ffcb_irfft2(const ffcb_tensor* spec, const ffcb_tensor* residual /* nullable */, const ffcb_tensor* out,
                void* ws, size_t ws_bytes, ffcb_stream_t stream);

/*
 * Head, tensor-core form (ffc.py:360-363).  The 7x7 convolution to N <= 3 outputs is split into
 *   (1) an ffcb_conv over the kernel ROWS only — seven K-segments (dy = -3..3, dx = 0) producing, for every input
 *       column, the 7*N partial sums q[b,y,x',n*7+kx] = sum_ky sum_c in[b,y+ky-3,x',c] * w[n,c,ky,kx], and
 *   (2) this gather: y[b,n,y,x] = act(bias[n] + sum_kx q[b,y,reflect(x+kx-3),n*7+kx])   (NCHW float output).
 * q: float view (B,H,W,>=7N).
 */
int ffcb_head_gather7(const ffcb_tensor* q, const float* bias, int N, int act, float* y_nchw, ffcb_stream_t stream);

/*
 * uint8 image I/O of the predict path — the elementwise work the reference does around the
 * generator, fused into the two kernels that touch the full-resolution image anyway.
 *
 * ffcb_stem_pack_u8 replaces, for a batch of decoded RGB images [B][H0][W0][3] and masks [B][H0][W0]:
 *     load_image: u8 -> float32 / 255                         saicinpainting/evaluation/data.py:11-19
 *     pad_img_to_modulo(mode='symmetric') to (H, W)           evaluation/data.py:32-36 (bottom / right only)
 *     mask = (mask > 0) * 1                                   bin/predict.py:83
 *     masked_img = img * (1 - mask); cat([masked_img, mask])  training/trainers/default.py:59, 68
 *     ReflectionPad2d(3) of the stem                          ffc.py:315
 * and writes the packed stem image of ffcb_stem_pack (view (B, H+6, W+8, 8); H, W = padded size, H-H0 <= H0).
 *
 * ffcb_head_gather7_blend_u8 is ffcb_head_gather7 (N = 3) followed by
 *     inpainted = mask * predicted + (1 - mask) * image       training/trainers/default.py:71
 *     crop to unpad_to_size (H0, W0)                          bin/predict.py:86-91
 *     np.clip(res * 255, 0, 255).astype('uint8')              bin/predict.py:93   (truncation)
 * writing RGB bytes [B][H0][W0][3].  Pixels outside the hole reproduce the reference's u8 -> /255 -> *255 -> u8
 * round trip bit for bit (IEEE float32 division and product).
 */
int ffcb_stem_pack_u8(const uint8_t* image_hwc, const uint8_t* mask_hw, int B, int H0, int W0,
                      const ffcb_tensor* packed, ffcb_stream_t stream);
int ffcb_head_gather7_blend_u8(const ffcb_tensor* q, const float* bias, int act, const uint8_t* image_hwc,
                               const uint8_t* mask_hw, int H0, int W0, uint8_t* out_hwc, ffcb_stream_t stream);

/* Layout/format conversion at the module boundary (the reference's tensors are NCHW float):
 * ffc.py has no counterpart — these replace nothing, they adapt torch's layout to the path's. */
int ffcb_nchw_to_nhwc(const float* x_nchw, int B, int C, int H, int W, const ffcb_tensor* out, ffcb_stream_t stream);
int ffcb_nhwc_to_nchw(const ffcb_tensor* in, float* y_nchw, ffcb_stream_t stream);
/* (re)write the reflected 1-pixel border ring of a pad==1 view from its interior */
int ffcb_fill_reflect_border(const ffcb_tensor* t, ffcb_stream_t stream);

/*
 * Input gradients through FFCResnetBlock (reference: evaluation/refinement.py:137-167 optimises the
 * block inputs by back-propagation).  With eval-mode BN folded, the backward pass of the block re-uses ffcb_conv (3x3
 * taps flipped, zero border, transposed weights) and the ffcb_rfft2 / ffcb_irfft2 pair itself; these two elementwise
 * steps complete it:
 *   ffcb_relu_bwd:            out = dy * [y > 0]   — nn.ReLU backward with the forward activation y (ffc.py:101,
 *                             133, 253-254); any storage format / layout on each of the three views
 *   ffcb_fold_reflect_border: adjoint of the reflect padding of ffc.py:189 (padding_mode='reflect', pad 1):
 *                             out[b,y,x,c] = sum of gpad over the padded positions that reflect onto (y,x)
 *                                           (+ add0[b,y,x,c-add0_c0] + add1[b,y,x,c-add1_c0] where defined);
 *                             gpad is (B,H+2,W+2,C); addends are optional (NULL) channel slices of the output
 */
int ffcb_relu_bwd(const ffcb_tensor* dy, const ffcb_tensor* y, const ffcb_tensor* out, ffcb_stream_t stream);
int ffcb_fold_reflect_border(const ffcb_tensor* gpad, const ffcb_tensor* add0, int add0_c0, const ffcb_tensor* add1,
                             int add1_c0, const ffcb_tensor* out, ffcb_stream_t stream);

/* number of kernel launches issued by this library on the calling thread since the last
 * ffcb_reset_launch_count() — bench.py reports it as "gpu_launches" */
long long ffcb_launch_count(void);
void ffcb_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FFC_B200_H_ */
