// Batched real 2-D FFT pair over channels-last tensors (norm='ortho').
//
// Replaces torch.fft.rfftn / irfftn and the stack/permute/view shuffles around them
// (reference ffc.py:86-89, 103-108) — no cuFFT.
//
// Mapping: lane == channel.  In NHWC the 32 lanes of a warp read 32 consecutive channels of one
// pixel (one 128-byte line), every lane owns one independent 1-D transform, and the transform's
// points live in shared memory as data[point][lane] (float2), so shared-memory accesses are
// conflict-free and twiddles are warp-uniform broadcasts.  A 2-D transform is a row pass and a
// column pass with the half-spectrum intermediate in a caller workspace (L2-resident at the
// sizes of the path); rows are transformed two at a time ("two-for-one": z = row_a + i*row_b).
//
// Sizes: power-of-two lengths 4..256 run a mixed-radix (8/4) Stockham autosort (ping-pong buffers);
// every other length runs a direct DFT (same kernels, O(n^2)) so odd / non-power-of-two planes
// (bin/predict.py pads images to multiples of 8 only -> e.g. 125x188 bottlenecks) stay native.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "fft_core.cuh"

namespace ffcb {
namespace {

constexpr int kLanes = 32;
using namespace fftc;

// Complex FFT of length N (compile-time power of two, or runtime n when N == 0) for this lane.
// `a` holds the input (already synchronised), `b` is scratch of the same size; returns the
// buffer holding the result.  Ends with a barrier.
template <int N, bool INV>
__device__ __forceinline__ float2* fft_dispatch(float2* a, float2* b, const float2* tw, int n, int lane, int worker,
                                                int nworkers, const RtPlan& rp) {
  if constexpr (N == 0) {
    if (rp.np < 0) {                       // direct DFT
      dft_pass<INV, kLanes>(a, b, tw, n, lane, worker, nworkers);
      __syncthreads();
      return b;
    }
    int ns = 1;                            // runtime mixed-radix Stockham (row f2)
    for (int p = 0; p < rp.np; ++p) {
      const int R = rp.radix[p];
      generic_pass<INV, kLanes>(a, b, tw, n, R, ns, lane, worker, nworkers);
      __syncthreads();
      ns *= R;
      float2* t = a; a = b; b = t;
    }
    return a;
  } else {
    stockham_pass<N, 0, INV, kLanes>(a, b, tw, lane, worker, nworkers);
    __syncthreads();
    if constexpr (Plan<N>::P == 1) return b;
    else {
      stockham_pass<N, 1, INV, kLanes>(b, a, tw, lane, worker, nworkers);
      __syncthreads();
      if constexpr (Plan<N>::P == 2) return a;
      else {
        stockham_pass<N, 2, INV, kLanes>(a, b, tw, lane, worker, nworkers);
        __syncthreads();
        return b;
      }
    }
  }
}

__device__ __forceinline__ void make_twiddles(float2* tw, int n) {
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nthreads = blockDim.x * blockDim.y * blockDim.z;
  for (int t = tid; t < n; t += nthreads) {
    float s, c;
    sincospif(2.0f * (float)t / (float)n, &s, &c);
    tw[t] = make_float2(c, -s);  // exp(-2 pi i t / n)
  }
}

// Shared-memory carve-up: [twiddles n][group g: ping n*32 | pong n*32]
template <int N>
__device__ __forceinline__ void carve(float2* smem, int n, int group, float2*& tw, float2*& data, float2*& tmp) {
  tw = smem;
  data = smem + n + (size_t)group * 2 * n * kLanes;
  tmp = data + n * kLanes;
}

// ---------------------------------------------------------------------------------------------
// Row pass, forward.  grid.x = ceil(B*ceil(H/2) / G), grid.y = ceil(C/32).
// in (B,H,W,C) real  ->  ws[b][y][k][c] complex, k = 0..W/2   (unscaled)
template <int N>
__global__ void __launch_bounds__(1024) rfft_rows_kernel(View in, float2* __restrict__ ws, int n, RtPlan rp) {
  extern __shared__ float2 smem_f2[];
  const int W = (N > 0) ? N : n;
  const int lane = threadIdx.x, worker = threadIdx.y, nworkers = blockDim.y, group = threadIdx.z;
  float2 *tw, *data, *tmp;
  carve<N>(smem_f2, W, group, tw, data, tmp);
  make_twiddles(tw, W);

  const int hp = (in.H + 1) / 2;
  const int pair = blockIdx.x * blockDim.z + group;       // (b, y-pair)
  const bool live = pair < in.B * hp;
  const int b = live ? pair / hp : 0;
  const int y0 = live ? (pair % hp) * 2 : 0;
  const int c = blockIdx.y * kLanes + lane;
  const bool cok = live && c < in.C;
  const bool row1 = (y0 + 1) < in.H;

  for (int x = worker; x < W; x += nworkers) {
    float2 z = make_float2(0.f, 0.f);
    if (cok) {
      z.x = load1(in, pix_off(in, b, y0, x) + c);
      if (row1) z.y = load1(in, pix_off(in, b, y0 + 1, x) + c);
    }
    data[x * kLanes + lane] = z;
  }
  __syncthreads();
  const float2* res = fft_dispatch<N, false>(data, tmp, tw, W, lane, worker, nworkers, rp);

  const int wf = W / 2 + 1;
  if (cok) {
    for (int k = worker; k < wf; k += nworkers) {
      float2 a, bb;
      r2c_pair_post<kLanes>(res, W, k, lane, a, bb);
      const size_t o = (((size_t)b * in.H + y0) * wf + k) * in.C + c;
      ws[o] = a;
      if (row1) ws[o + (size_t)wf * in.C] = bb;
    }
  }
}

// Column pass, forward.  grid.x = ceil(B*Wf / G), grid.y = ceil(C/32).
// ws[b][y][k][c] complex -> spec (B,H,Wf,2C): channel 2c = Re, 2c+1 = Im, scaled by `scale`.
template <int N>
__global__ void __launch_bounds__(1024) fft_cols_fwd_kernel(const float2* __restrict__ ws, View spec, int n,
                                                            int C, float scale, RtPlan rp) {
  extern __shared__ float2 smem_f2[];
  const int H = (N > 0) ? N : n;
  const int lane = threadIdx.x, worker = threadIdx.y, nworkers = blockDim.y, group = threadIdx.z;
  float2 *tw, *data, *tmp;
  carve<N>(smem_f2, H, group, tw, data, tmp);
  make_twiddles(tw, H);

  const int wf = spec.W;
  const int col = blockIdx.x * blockDim.z + group;  // (b, k)
  const bool live = col < spec.B * wf;
  const int b = live ? col / wf : 0;
  const int k = live ? col % wf : 0;
  const int c = blockIdx.y * kLanes + lane;
  const bool cok = live && c < C;

  for (int y = worker; y < H; y += nworkers) {
    float2 z = make_float2(0.f, 0.f);
    if (cok) z = ws[(((size_t)b * H + y) * wf + k) * C + c];
    data[y * kLanes + lane] = z;
  }
  __syncthreads();
  const float2* res = fft_dispatch<N, false>(data, tmp, tw, H, lane, worker, nworkers, rp);

  if (cok) {
    for (int y = worker; y < H; y += nworkers) {
      const float2 z = res[y * kLanes + lane];
      const long long o = pix_off(spec, b, y, k) + 2 * c;
      if (spec.fmt == FFCB_F32) {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(spec.ptr) + o) = make_float2(z.x * scale, z.y * scale);
      } else {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(z.x * scale, h0, l0);
        split_bf16(z.y * scale, h1, l1);
        unsigned* p = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned short*>(spec.ptr) + o);
        p[0] = pack_bf16(h0, h1);
        *reinterpret_cast<unsigned*>(reinterpret_cast<unsigned short*>(spec.ptr) + o + spec.lo_off) =
            pack_bf16(l0, l1);
      }
    }
  }
}

// Column pass, inverse: spec (B,H,Wf,2C) -> ws[b][y][k][c] complex (unscaled inverse along H).
template <int N>
__global__ void __launch_bounds__(N > 0 && N <= 64 ? 256 : 1024, N > 0 && N <= 64 ? 6 : 1) fft_cols_inv_kernel(View spec, float2* __restrict__ ws, int n, int C, RtPlan rp) {
  extern __shared__ float2 smem_f2[];
  const int H = (N > 0) ? N : n;
  const int lane = threadIdx.x, worker = threadIdx.y, nworkers = blockDim.y, group = threadIdx.z;
  float2 *tw, *data, *tmp;
  carve<N>(smem_f2, H, group, tw, data, tmp);
  make_twiddles(tw, H);

  const int wf = spec.W;
  const int col = blockIdx.x * blockDim.z + group;
  const bool live = col < spec.B * wf;
  const int b = live ? col / wf : 0;
  const int k = live ? col % wf : 0;
  const int c = blockIdx.y * kLanes + lane;
  const bool cok = live && c < C;

  for (int y = worker; y < H; y += nworkers) {
    float2 z = make_float2(0.f, 0.f);
    if (cok) {
      const long long o = pix_off(spec, b, y, k) + 2 * c;
      if (spec.fmt == FFCB_F32) {
        z = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(spec.ptr) + o));
      } else {
        z.x = load1(spec, o);
        z.y = load1(spec, o + 1);
      }
    }
    data[y * kLanes + lane] = z;
  }
  __syncthreads();
  const float2* res = fft_dispatch<N, true>(data, tmp, tw, H, lane, worker, nworkers, rp);

  if (cok) {
    for (int y = worker; y < H; y += nworkers)
      ws[(((size_t)b * H + y) * wf + k) * C + c] = res[y * kLanes + lane];
  }
}

// Row pass, inverse (C2R, two rows at a time): ws[b][y][k][c] -> out (B,H,W,C) real,
// out = residual + scale * c2r(ws).  Im of bins 0 and (even W) W/2 is ignored.
template <int N>
__global__ void __launch_bounds__(N > 0 && N <= 64 ? 256 : 1024, N > 0 && N <= 64 ? 6 : 1) irfft_rows_kernel(const float2* __restrict__ ws, View res, View out, int n,
                                                          float scale, RtPlan rp) {
  extern __shared__ float2 smem_f2[];
  const int W = (N > 0) ? N : n;
  const int lane = threadIdx.x, worker = threadIdx.y, nworkers = blockDim.y, group = threadIdx.z;
  float2 *tw, *data, *tmp;
  carve<N>(smem_f2, W, group, tw, data, tmp);
  make_twiddles(tw, W);

  const int hp = (out.H + 1) / 2;
  const int pair = blockIdx.x * blockDim.z + group;
  const bool live = pair < out.B * hp;
  const int b = live ? pair / hp : 0;
  const int y0 = live ? (pair % hp) * 2 : 0;
  const int c = blockIdx.y * kLanes + lane;
  const bool cok = live && c < out.C;
  const bool row1 = (y0 + 1) < out.H;
  const int wf = W / 2 + 1;

  for (int k = worker; k < wf; k += nworkers) {
    float2 x1 = make_float2(0.f, 0.f), x2 = make_float2(0.f, 0.f);
    if (cok) {
      const size_t o = (((size_t)b * out.H + y0) * wf + k) * out.C + c;
      x1 = ws[o];
      if (row1) x2 = ws[o + (size_t)wf * out.C];
    }
    c2r_pair_pre<kLanes>(data, W, k, lane, x1, x2);
  }
  __syncthreads();
  const float2* fin = fft_dispatch<N, true>(data, tmp, tw, W, lane, worker, nworkers, rp);

  if (cok) {
    if (N > 0) {
      // power-of-two lengths: every worker owns exactly 8 pixels.  Put all residual loads in flight first —
      // the compiler cannot hoist them above the stores itself (res / out may alias as far as it knows).
      float ra[8], rb[8];
      const long long q0 = res.ptr ? pix_off(res, b, y0, 0) + c : 0, o0 = pix_off(out, b, y0, 0) + c;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int x = worker + i * nworkers;
        ra[i] = (res.ptr != nullptr && x < W) ? load1(res, q0 + x * res.sx) : 0.f;
        rb[i] = (res.ptr != nullptr && x < W && row1) ? load1(res, q0 + res.sy + x * res.sx) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int x = worker + i * nworkers;
        if (x < W) {
          const float2 z = fin[x * kLanes + lane];
          store1(out, o0 + x * out.sx, fmaf(z.x, scale, ra[i]));
          if (row1) store1(out, o0 + out.sy + x * out.sx, fmaf(z.y, scale, rb[i]));
        }
      }
    } else {
      for (int x = worker; x < W; x += nworkers) {
        const float2 z = fin[x * kLanes + lane];
        float r0 = z.x * scale, r1 = z.y * scale;
        if (res.ptr != nullptr) {
          r0 += load1(res, pix_off(res, b, y0, x) + c);
          if (row1) r1 += load1(res, pix_off(res, b, y0 + 1, x) + c);
        }
        store1(out, pix_off(out, b, y0, x) + c, r0);
        if (row1) store1(out, pix_off(out, b, y0 + 1, x) + c, r1);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct LaunchPlan {
  int N;       // template length (0 = direct DFT)
  int n;       // runtime length
  dim3 block;  // (32, workers, groups)
  size_t smem;
  RtPlan rp;   // N == 0: runtime radix plan (np < 0: direct DFT)
};

// Lengths without a compile-time plan: runtime mixed-radix Stockham (FFCB_FFT_MIXED_RADIX=0 selects the O(n^2)
// direct DFT they ran in the first revision — same results to round-off, kept as the cross-check).
bool mixed_radix_enabled() {
  const char* e = getenv("FFCB_FFT_MIXED_RADIX");
  return e ? atoi(e) != 0 : true;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

LaunchPlan make_plan(int n) {
  LaunchPlan p;
  p.n = n;
  p.rp.np = -1;
  for (int i = 0; i < kMaxRtPasses; ++i) p.rp.radix[i] = 1;
  if (is_pow2(n) && n >= 4 && n <= 256) {
    p.N = n;
    const int workers = fftc::workers_for(n);
    const int groups = workers >= 8 ? 1 : 8 / workers;
    p.block = dim3(kLanes, workers, groups);
    p.smem = sizeof(float2) * ((size_t)n + (size_t)groups * 2 * n * kLanes);
  } else {
    p.N = 0;
    int workers = n >= 8 ? 8 : (n >= 4 ? 4 : 1);
    if (mixed_radix_enabled()) {
      p.rp = make_rt_plan(n);
      // one output per worker-iteration: ~8 outputs per worker and pass, up to a full 1024-thread CTA
      if (n > 64) workers = (n + 7) / 8 < 32 ? (n + 7) / 8 : 32;
    }
    const int groups = n <= 32 ? (8 / workers > 0 ? 8 / workers : 1) : 1;
    p.block = dim3(kLanes, workers, groups);
    p.smem = sizeof(float2) * ((size_t)n + (size_t)groups * 2 * n * kLanes);
  }
  return p;
}

constexpr size_t kMaxSmem = 227 * 1024;

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) FFCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return FFCB_OK;
}

#define FFCB_DISPATCH_N(PLAN, ...)                      \
  switch ((PLAN).N) {                                   \
    case 0: { constexpr int NN = 0; __VA_ARGS__; } break;      \
    case 4: { constexpr int NN = 4; __VA_ARGS__; } break;      \
    case 8: { constexpr int NN = 8; __VA_ARGS__; } break;      \
    case 16: { constexpr int NN = 16; __VA_ARGS__; } break;    \
    case 32: { constexpr int NN = 32; __VA_ARGS__; } break;    \
    case 64: { constexpr int NN = 64; __VA_ARGS__; } break;    \
    case 128: { constexpr int NN = 128; __VA_ARGS__; } break;  \
    case 256: { constexpr int NN = 256; __VA_ARGS__; } break;  \
    default: set_error("fft: internal plan error"); return FFCB_EINVAL; \
  }

int check_fft_shapes(const ffcb_tensor* real, const ffcb_tensor* spec, const char* who) {
  FFCB_REQUIRE(real->H >= 1 && real->W >= 2, "%s: plane %dx%d too small", who, real->H, real->W);
  FFCB_REQUIRE(spec->B == real->B && spec->H == real->H && spec->W == real->W / 2 + 1 && spec->C == 2 * real->C,
               "%s: spectrum view must be (B,H,W/2+1,2C) = (%d,%d,%d,%d), got (%d,%d,%d,%d)", who, real->B, real->H,
               real->W / 2 + 1, 2 * real->C, spec->B, spec->H, spec->W, spec->C);
  LaunchPlan pw = make_plan(real->W), ph = make_plan(real->H);
  FFCB_REQUIRE(pw.smem <= kMaxSmem && ph.smem <= kMaxSmem, "%s: plane %dx%d exceeds the shared-memory FFT limits",
               who, real->H, real->W);
  return FFCB_OK;
}

}  // namespace

// fused whole-plane path (fft_plane.cu)
bool plane64_eligible(const ffcb_tensor* real);
int rfft2_plane64(const ffcb_tensor* in, const ffcb_tensor* spec, cudaStream_t stream);
int irfft2_plane64(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out, cudaStream_t stream);
int inv_plane_variant();
// channel-group planar plane kernels (fft_plane_cg.cu)
bool plane64_cg_fwd_eligible(const ffcb_tensor* in, const ffcb_tensor* spec);
bool plane64_cg_inv_eligible(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out);
int rfft2_plane64_cg(const ffcb_tensor* in, const ffcb_tensor* spec, cudaStream_t stream);
int irfft2_plane64_cg(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out, cudaStream_t stream);

size_t fft2_workspace_bytes(int B, int H, int W, int C) {
  return sizeof(float2) * (size_t)B * H * (W / 2 + 1) * C;
}

int rfft2(const ffcb_tensor* in, const ffcb_tensor* spec, void* ws, size_t ws_bytes, cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(in, "rfft2.in", true)) || (rc = check_tensor(spec, "rfft2.spec", true))) return rc;
  if ((rc = check_fft_shapes(in, spec, "rfft2"))) return rc;
  if (ws_bytes < fft2_workspace_bytes(in->B, in->H, in->W, in->C)) {
    set_error("rfft2: workspace %zu < %zu bytes", ws_bytes, fft2_workspace_bytes(in->B, in->H, in->W, in->C));
    return FFCB_ENOMEM;
  }
  if (in->B == 0 || in->C == 0) return FFCB_OK;
  if (in->cg != 0 || spec->cg != 0) {
    FFCB_REQUIRE(plane64_cg_fwd_eligible(in, spec),
                 "rfft2: channel-group planar views need a 64x64 / 32x32 float32 cg=4 input and a split-bf16 cg=8 spectrum");
    return rfft2_plane64_cg(in, spec, stream);
  }
  if (plane64_eligible(in) && !getenv("FFCB_FFT_TWO_PASS")) return rfft2_plane64(in, spec, stream);
  const View vin = make_view(*in), vspec = make_view(*spec);
  float2* w2 = reinterpret_cast<float2*>(ws);
  const int cblocks = (in->C + kLanes - 1) / kLanes;
  const float scale = (float)(1.0 / sqrt((double)in->H * (double)in->W));
  {
    LaunchPlan p = make_plan(in->W);
    const int pairs = in->B * ((in->H + 1) / 2);
    dim3 grid((pairs + p.block.z - 1) / p.block.z, cblocks);
    FFCB_DISPATCH_N(p, {
      if ((rc = set_smem(rfft_rows_kernel<NN>, p.smem))) return rc;
      rfft_rows_kernel<NN><<<grid, p.block, p.smem, stream>>>(vin, w2, p.n, p.rp);
    });
    FFCB_LAUNCH_CHECK("rfft_rows_kernel");
  }
  {
    LaunchPlan p = make_plan(in->H);
    const int cols = in->B * spec->W;
    dim3 grid((cols + p.block.z - 1) / p.block.z, cblocks);
    FFCB_DISPATCH_N(p, {
      if ((rc = set_smem(fft_cols_fwd_kernel<NN>, p.smem))) return rc;
      fft_cols_fwd_kernel<NN><<<grid, p.block, p.smem, stream>>>(w2, vspec, p.n, in->C, scale, p.rp);
    });
    FFCB_LAUNCH_CHECK("fft_cols_fwd_kernel");
  }
  return FFCB_OK;
}

int irfft2(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out, void* ws, size_t ws_bytes,
           cudaStream_t stream) {
  int rc;
  if ((rc = check_tensor(spec, "irfft2.spec", true)) || (rc = check_tensor(out, "irfft2.out", true))) return rc;
  if ((rc = check_fft_shapes(out, spec, "irfft2"))) return rc;
  View vres = null_view();
  if (residual != nullptr && residual->ptr != nullptr) {
    if ((rc = check_tensor(residual, "irfft2.residual", true))) return rc;
    FFCB_REQUIRE(residual->B == out->B && residual->H == out->H && residual->W == out->W && residual->C == out->C,
                 "irfft2: residual shape differs from output");
    vres = make_view(*residual);
  }
  if (ws_bytes < fft2_workspace_bytes(out->B, out->H, out->W, out->C)) {
    set_error("irfft2: workspace %zu < %zu bytes", ws_bytes, fft2_workspace_bytes(out->B, out->H, out->W, out->C));
    return FFCB_ENOMEM;
  }
  if (out->B == 0 || out->C == 0) return FFCB_OK;
  if (spec->cg != 0 || out->cg != 0 || (residual && residual->ptr && residual->cg != 0)) {
    FFCB_REQUIRE(plane64_cg_inv_eligible(spec, residual, out),
                 "irfft2: channel-group planar views need a 64x64 / 32x32 plane, a float32 cg=8 spectrum, a float32 cg=4 "
                 "residual and a split-bf16 cg=8 or float32 cg=4 output");
    return irfft2_plane64_cg(spec, residual, out, stream);
  }
  // FFCB_FFT_INV_PLANE: 0 = two-pass kernels, 1 / 2 = first-revision plane kernels (slower than two-pass),
  // 3 = second revision; irfft2_plane64 returns 1 when the chosen variant does not apply to these views
  if (plane64_eligible(out) && inv_plane_variant() != 0 && !getenv("FFCB_FFT_TWO_PASS")) {
    rc = irfft2_plane64(spec, residual, out, stream);
    if (rc <= 0) return rc;
  }
  const View vspec = make_view(*spec), vout = make_view(*out);
  float2* w2 = reinterpret_cast<float2*>(ws);
  const int cblocks = (out->C + kLanes - 1) / kLanes;
  const float scale = (float)(1.0 / sqrt((double)out->H * (double)out->W));
  {
    LaunchPlan p = make_plan(out->H);
    const int cols = out->B * spec->W;
    dim3 grid((cols + p.block.z - 1) / p.block.z, cblocks);
    FFCB_DISPATCH_N(p, {
      if ((rc = set_smem(fft_cols_inv_kernel<NN>, p.smem))) return rc;
      fft_cols_inv_kernel<NN><<<grid, p.block, p.smem, stream>>>(vspec, w2, p.n, out->C, p.rp);
    });
    FFCB_LAUNCH_CHECK("fft_cols_inv_kernel");
  }
  {
    LaunchPlan p = make_plan(out->W);
    const int pairs = out->B * ((out->H + 1) / 2);
    dim3 grid((pairs + p.block.z - 1) / p.block.z, cblocks);
    FFCB_DISPATCH_N(p, {
      if ((rc = set_smem(irfft_rows_kernel<NN>, p.smem))) return rc;
      irfft_rows_kernel<NN><<<grid, p.block, p.smem, stream>>>(w2, vres, vout, p.n, scale, p.rp);
    });
    FFCB_LAUNCH_CHECK("irfft_rows_kernel");
  }
  return FFCB_OK;
}

}  // namespace ffcb
