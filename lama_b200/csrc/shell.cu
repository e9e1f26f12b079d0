// Generator shell kernels and layout adapters:
//   stem_conv7  : ReflectionPad2d(3) + Conv2d(Cin->N, 7x7) + folded BN + ReLU, NCHW float in -> NHWC out
//                 (reference ffc.py:315-317, 253)
//   head_conv7  : ReflectionPad2d(3) + Conv2d(C->N<=4, 7x7, bias) + activation, NHWC in -> NCHW float out
//                 (reference ffc.py:360-363)
//   nchw<->nhwc : module-boundary layout conversion (the reference's tensors are NCHW float)
//   reflect ring: (re)build the 1-pixel reflected border of a padded view
#include "common.cuh"

namespace ffcb {
namespace {

constexpr int TILE = 16;          // output pixels per CTA edge
constexpr int HALO = 3;           // 7x7
constexpr int PT = TILE + 2 * HALO;  // 22

// ---------------------------------------------------------------------------------------- stem
// One thread = one output pixel, NACC output channels per CTA pass (grid.z walks channel groups).
// smem: patch[Cin][22][22] floats, then w[49*Cin][NACC].
template <int NACC>
__global__ void __launch_bounds__(TILE * TILE) stem_conv7_kernel(const float* __restrict__ x, int B, int Cin, int H,
                                                                 int W, const float* __restrict__ w,
                                                                 const float* __restrict__ shift, int N, View out) {
  extern __shared__ __align__(16) float smem[];
  float* patch = smem;
  float* ws = smem + ((Cin * PT * PT + 3) & ~3);
  const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
  const int tiles_x = (W + TILE - 1) / TILE;
  const int x0 = (blockIdx.x % tiles_x) * TILE, y0 = (blockIdx.x / tiles_x) * TILE;
  const int b = blockIdx.y;
  const int n0 = blockIdx.z * NACC;
  const int K = 49 * Cin;

  for (int i = threadIdx.x; i < Cin * PT * PT; i += blockDim.x) {
    const int c = i / (PT * PT), r = i % (PT * PT);
    const int yy = reflect_idx(y0 + r / PT - HALO, H), xx = reflect_idx(x0 + r % PT - HALO, W);
    // tiles hanging over the image edge: clamp (values unused by in-range pixels)
    const int yc = min(max(yy, 0), H - 1), xc = min(max(xx, 0), W - 1);
    patch[i] = __ldg(x + (((long long)b * Cin + c) * H + yc) * W + xc);
  }
  for (int i = threadIdx.x; i < K * NACC; i += blockDim.x) {
    const int k = i / NACC, j = i % NACC;
    ws[i] = (n0 + j < N) ? __ldg(w + (long long)k * N + n0 + j) : 0.f;
  }
  __syncthreads();

  float acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < 7; ++ky)
    for (int kx = 0; kx < 7; ++kx)
      for (int c = 0; c < Cin; ++c) {
        const float a = patch[(c * PT + ty + ky) * PT + tx + kx];
        const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 7 + kx) * Cin + c) * NACC);
#pragma unroll
        for (int j = 0; j < NACC / 4; ++j) {
          const float4 wv = wr[j];
          acc[4 * j + 0] = fmaf(a, wv.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(a, wv.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a, wv.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a, wv.w, acc[4 * j + 3]);
        }
      }
  const int y = y0 + ty, xo = x0 + tx;
  if (y >= H || xo >= W) return;
  const long long o = pix_off(out, b, y, xo);
#pragma unroll
  for (int j = 0; j < NACC / 4; ++j) {
    const int n = n0 + 4 * j;
    if (n >= N) break;
    float4 v = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    if (shift != nullptr) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(shift + n));
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    store4(out, o + n, v);
  }
}

// ---------------------------------------------------------------------------------------- head
// One thread = one output pixel, all N<=4 outputs.  Channels are staged 16 at a time:
// patch[22*22][PSTR] floats (PSTR = 20 keeps float4 reads conflict-free), w[N][49][16].
constexpr int HC = 16, PSTR = 20;

__global__ void __launch_bounds__(TILE * TILE) head_conv7_kernel(View in, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, int N, int act,
                                                                 float* __restrict__ y_out) {
  extern __shared__ __align__(16) float smem[];
  float* patch = smem;                  // [PT*PT][PSTR]
  float* ws = smem + PT * PT * PSTR;    // [4][49][HC]
  const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
  const int H = in.H, W = in.W, C = in.C;
  const int tiles_x = (W + TILE - 1) / TILE;
  const int x0 = (blockIdx.x % tiles_x) * TILE, y0 = (blockIdx.x / tiles_x) * TILE;
  const int b = blockIdx.y;

  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < C; c0 += HC) {
    for (int i = threadIdx.x; i < PT * PT * (HC / 4); i += blockDim.x) {
      const int pix = i / (HC / 4), q = i % (HC / 4);
      const int yy = reflect_idx(y0 + pix / PT - HALO, H), xx = reflect_idx(x0 + pix % PT - HALO, W);
      const int yc = min(max(yy, 0), H - 1), xc = min(max(xx, 0), W - 1);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + 4 * q < C) v = load4(in, pix_off(in, b, yc, xc) + c0 + 4 * q);
      *reinterpret_cast<float4*>(patch + pix * PSTR + 4 * q) = v;
    }
    for (int i = threadIdx.x; i < 4 * 49 * HC; i += blockDim.x) {
      const int n = i / (49 * HC), r = i % (49 * HC), t = r / HC, c = r % HC;
      ws[i] = (n < N && c0 + c < C) ? __ldg(w + ((long long)n * 49 + t) * C + c0 + c) : 0.f;
    }
    __syncthreads();
    for (int ky = 0; ky < 7; ++ky)
      for (int kx = 0; kx < 7; ++kx) {
        const float* pp = patch + ((ty + ky) * PT + tx + kx) * PSTR;
        const float* wp = ws + (ky * 7 + kx) * HC;
#pragma unroll
        for (int q = 0; q < HC / 4; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(pp + 4 * q);
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + n * 49 * HC + 4 * q);
            acc[n] = fmaf(a.x, wv.x, fmaf(a.y, wv.y, fmaf(a.z, wv.z, fmaf(a.w, wv.w, acc[n]))));
          }
        }
      }
    __syncthreads();
  }
  const int y = y0 + ty, xo = x0 + tx;
  if (y >= H || xo >= W) return;
  for (int n = 0; n < N; ++n) {
    const float v = apply_act(acc[n] + (bias ? __ldg(bias + n) : 0.f), act);
    y_out[(((long long)b * N + n) * H + y) * W + xo] = v;
  }
}

// ---------------------------------------------------------------------------------------- layout
// [C][W] <-> [W][C] transposes per (b, y) row through a 32x33 shared tile.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int H, int W, View out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z / H, y = blockIdx.z % H;
  const int c0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, xx = x0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && xx < W) ? __ldg(x + (((long long)b * C + c) * H + y) * W + xx) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int xx = x0 + i, c = c0 + threadIdx.x;
    if (xx < W && c < C) store1(out, pix_off(out, b, y, xx) + c, tile[threadIdx.x][i]);
  }
}

__global__ void nhwc_to_nchw_kernel(View in, float* __restrict__ yo) {
  __shared__ float tile[32][33];
  const int H = in.H, W = in.W, C = in.C;
  const int b = blockIdx.z / H, y = blockIdx.z % H;
  const int c0 = blockIdx.y * 32, x0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int xx = x0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (xx < W && c < C) ? load1(in, pix_off(in, b, y, xx) + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, xx = x0 + threadIdx.x;
    if (c < C && xx < W) yo[(((long long)b * C + c) * H + y) * W + xx] = tile[threadIdx.x][i];
  }
}

// Reflected ring of a pad==1 view: ring pixel (y, x) with y in {-1, H} or x in {-1, W}
// copies interior pixel (reflect(y), reflect(x)).
__global__ void reflect_ring_kernel(View t) {
  const int ring = 2 * (t.W + 2) + 2 * t.H;  // ring pixels per image
  const int c4 = t.C / 4;
  const long long total = (long long)t.B * ring * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % c4);
    const long long pi = i / c4;
    const int r = (int)(pi % ring), b = (int)(pi / ring);
    int y, x;
    if (r < t.W + 2) { y = -1; x = r - 1; }
    else if (r < 2 * (t.W + 2)) { y = t.H; x = r - (t.W + 2) - 1; }
    else { const int s = r - 2 * (t.W + 2); y = s >> 1; x = (s & 1) ? t.W : -1; }
    const float4 v = load4(t, pix_off(t, b, reflect_idx(y, t.H), reflect_idx(x, t.W)) + 4 * q);
    store4(t, pix_off(t, b, y, x) + 4 * q, v);
  }
}

template <int NACC>
int launch_stem(const float* x, int B, int Cin, int H, int W, const float* w, const float* shift, int N,
                const View& out, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (((size_t)Cin * PT * PT + 3) / 4 * 4 + (size_t)49 * Cin * NACC);
  if (smem > 48 * 1024)
    FFCB_CUDA(cudaFuncSetAttribute(stem_conv7_kernel<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE), B, (N + NACC - 1) / NACC);
  stem_conv7_kernel<NACC><<<grid, TILE * TILE, smem, stream>>>(x, B, Cin, H, W, w, shift, N, out);
  FFCB_LAUNCH_CHECK("stem_conv7_kernel");
  return FFCB_OK;
}

}  // namespace

int stem_conv7(const float* x, int B, int Cin, int H, int W, const float* w, const float* shift, int N,
               const ffcb_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(out, "stem_conv7.out"))) return rc;
  FFCB_REQUIRE(x && w, "stem_conv7: null pointer");
  FFCB_REQUIRE(Cin >= 1 && Cin <= 16, "stem_conv7: Cin=%d outside [1,16]", Cin);
  FFCB_REQUIRE(H >= 4 && W >= 4, "stem_conv7: reflect pad 3 needs H,W >= 4 (got %dx%d)", H, W);
  FFCB_REQUIRE(N % 4 == 0 && N >= 4, "stem_conv7: N=%d must be a positive multiple of 4", N);
  FFCB_REQUIRE(out->B == B && out->H == H && out->W == W && out->C == N, "stem_conv7: out view shape mismatch");
  if (B == 0) return FFCB_OK;
  const View vo = make_view(*out);
  if (N >= 64) return launch_stem<64>(x, B, Cin, H, W, w, shift, N, vo, stream);
  if (N >= 32) return launch_stem<32>(x, B, Cin, H, W, w, shift, N, vo, stream);
  if (N >= 16) return launch_stem<16>(x, B, Cin, H, W, w, shift, N, vo, stream);
  return launch_stem<8>(x, B, Cin, H, W, w, shift, N, vo, stream);
}

int head_conv7(const ffcb_tensor* in, const float* w, const float* bias, int N, int act, float* y,
               cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "head_conv7.in"))) return rc;
  FFCB_REQUIRE(w && y, "head_conv7: null pointer");
  FFCB_REQUIRE(N >= 1 && N <= 4, "head_conv7: N=%d outside [1,4]", N);
  FFCB_REQUIRE(in->H >= 4 && in->W >= 4, "head_conv7: reflect pad 3 needs H,W >= 4");
  if (in->B == 0) return FFCB_OK;
  dim3 grid(((in->W + TILE - 1) / TILE) * ((in->H + TILE - 1) / TILE), in->B);
  constexpr size_t smem = sizeof(float) * (PT * PT * PSTR + 4 * 49 * HC);
  FFCB_CUDA(cudaFuncSetAttribute(head_conv7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  head_conv7_kernel<<<grid, TILE * TILE, smem, stream>>>(make_view(*in), w, bias, N, act, y);
  FFCB_LAUNCH_CHECK("head_conv7_kernel");
  return FFCB_OK;
}

int nchw_to_nhwc(const float* x, int B, int C, int H, int W, const ffcb_tensor* out, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(out, "nchw_to_nhwc.out"))) return rc;
  FFCB_REQUIRE(out->B == B && out->C == C && out->H == H && out->W == W, "nchw_to_nhwc: shape mismatch");
  if ((long long)B * C * H * W == 0) return FFCB_OK;
  FFCB_REQUIRE((long long)B * H <= 65535, "nchw_to_nhwc: B*H=%lld exceeds grid.z", (long long)B * H);
  dim3 grid((W + 31) / 32, (C + 31) / 32, B * H), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, stream>>>(x, C, H, W, make_view(*out));
  FFCB_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return FFCB_OK;
}

int nhwc_to_nchw(const ffcb_tensor* in, float* y, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "nhwc_to_nchw.in"))) return rc;
  if ((long long)in->B * in->C * in->H * in->W == 0) return FFCB_OK;
  FFCB_REQUIRE((long long)in->B * in->H <= 65535, "nhwc_to_nchw: B*H exceeds grid.z");
  dim3 grid((in->W + 31) / 32, (in->C + 31) / 32, in->B * in->H), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, stream>>>(make_view(*in), y);
  FFCB_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return FFCB_OK;
}

int fill_reflect_border(const ffcb_tensor* t, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(t, "fill_reflect_border"))) return rc;
  FFCB_REQUIRE(t->pad == 1, "fill_reflect_border: view has no border ring (pad=%d)", t->pad);
  FFCB_REQUIRE(t->H >= 2 && t->W >= 2, "fill_reflect_border: reflect needs H,W >= 2");
  const long long total = (long long)t->B * (2 * (t->W + 2) + 2 * t->H) * (t->C / 4);
  if (total == 0) return FFCB_OK;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  reflect_ring_kernel<<<blocks, 256, 0, stream>>>(make_view(*t));
  FFCB_LAUNCH_CHECK("reflect_ring_kernel");
  return FFCB_OK;
}

}  // namespace ffcb
