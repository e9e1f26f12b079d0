// Fused 2-D real FFT pair for 64x64 planes (the bottleneck of 512x512 images): one CTA transforms the
// whole plane of 8 channels, so the half-spectrum intermediate lives in shared memory instead of a
// global workspace (halves the HBM traffic of ffcb_rfft2 / ffcb_irfft2 and removes two launches).
//
// Every 64-point transform is held in the registers of ONE thread (fft64_regs: unrolled 8x8 Cooley-Tukey
// with compile-time twiddles) — no shuffles, no shared-memory butterflies, one barrier per plane:
//   forward : 256 threads = 8 channels x 32 row pairs (two-for-one real rows)  -> S[y][kx][c] -> barrier ->
//             8 channels x (31 complex columns + 1 packed DC/Nyquist pair) -> spectrum (Re/Im interleaved, scaled)
//   inverse : columns first (all 33, complex), barrier, then C2R rows (+ residual).
// smem: S[64][P] float2 with row pitch P = 33*8 + 4 (the +4 spreads the four row-pair groups of a warp
// over both halves of the banks).  Lanes: 8 consecutive channels (32 B of a pixel) x 4 rows/columns.
#include <stdlib.h>

#include "common.cuh"
#include "fft_core.cuh"

namespace ffcb {
namespace {

using namespace fftc;
constexpr int PN = 64, PWF = 33;
// Channels per CTA (PCH): 8 = one CTA per SM (137 KB of shared memory), lanes cover one full 32-byte sector of a
// pixel; 4 = 69 KB, two CTAs per SM so that one CTA's load / store phases overlap the other's transforms
// (FFCB_FFT_PLANE_CH=4, an experiment of round 1 — half-sector accesses, the sibling CTA picks up the other half
// from L2).  The row pitch keeps the row-phase stores of a half-warp on 32 distinct banks: 4*P mod 32 = 128 / PCH.
template <int PCH> struct PlaneCfg {
  static constexpr int pitch = PWF * PCH + (PCH == 8 ? 4 : 2);
  static constexpr int row_threads = 32 * PCH;                       // PCH channels x 32 row pairs
  static constexpr int col_threads = (PWF * PCH + 31) / 32 * 32;     // PCH channels x 33 columns, whole warps
  static constexpr size_t smem = sizeof(float2) * PN * pitch;
  static constexpr int ctas_per_sm = PCH == 8 ? 1 : 2;
};

// Forward: rows by the first 32*PCH threads, the 33 x PCH column tasks by the first 33*PCH (for PCH = 8 measured
// faster than the packed 256-thread variant: 152 registers instead of 255, and one more warp to hide latency).
template <int PCH, int OCC>
__global__ void __launch_bounds__(PlaneCfg<PCH>::col_threads, OCC)
rfft2_plane64_kernel(View in, View spec, float scale) {
  extern __shared__ float2 S[];
  constexpr int PPITCH = PlaneCfg<PCH>::pitch;
  const int tid = threadIdx.x, c = tid % PCH, g = tid / PCH;
  const int ch = blockIdx.x * PCH + c, b = blockIdx.y;
  if (tid < PlaneCfg<PCH>::row_threads) {   // g = row pair
    const long long r0 = pix_off(in, b, 2 * g, 0) + ch, r1 = r0 + in.sy;
    plane64_rows_fwd(
        [&](int n) { return make_float2(load1(in, r0 + n * in.sx), load1(in, r1 + n * in.sx)); },
        [&](int k, float2 a, float2 bb) {
          S[(2 * g) * PPITCH + k * PCH + c] = a;
          S[(2 * g + 1) * PPITCH + k * PCH + c] = bb;
        });
  }
  __syncthreads();
  if (tid < PWF * PCH) {   // g = kx
    const long long o0 = pix_off(spec, b, 0, g) + 2 * ch;
    plane64_col<false>(
        [&](int y) { return S[y * PPITCH + g * PCH + c]; },
        [&](int ky, float2 z) {
          const long long o = o0 + ky * spec.sy;
          z.x *= scale; z.y *= scale;
          if (spec.fmt == FFCB_F32) {
            *reinterpret_cast<float2*>(reinterpret_cast<float*>(spec.ptr) + o) = z;
          } else {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(z.x, h0, l0);
            split_bf16(z.y, h1, l1);
            unsigned short* p = reinterpret_cast<unsigned short*>(spec.ptr);
            *reinterpret_cast<unsigned*>(p + o) = pack_bf16(h0, h1);
            *reinterpret_cast<unsigned*>(p + o + spec.lo_off) = pack_bf16(l0, l1);
          }
        });
  }
}

template <int PCH>
__global__ void __launch_bounds__(PlaneCfg<PCH>::row_threads, PlaneCfg<PCH>::ctas_per_sm)
irfft2_plane64_kernel(View spec, View res, View out, float scale) {
  extern __shared__ float2 S[];
  constexpr int PPITCH = PlaneCfg<PCH>::pitch;
  const int tid = threadIdx.x, c = tid % PCH, g = tid / PCH;
  const int ch = blockIdx.x * PCH + c, b = blockIdx.y;
  auto get = [&](long long o) {
    if (spec.fmt == FFCB_F32) return __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(spec.ptr) + o));
    return make_float2(load1(spec, o), load1(spec, o + 1));
  };
  const long long o0 = pix_off(spec, b, 0, 0) + 2 * ch;
  const bool packed = g == 0;
  plane64_col_inv_any(
      packed, [&](int ky) { return get(o0 + ky * spec.sy + g * spec.sx); },
      [&](int ky) { return get(o0 + ky * spec.sy + 32 * spec.sx); },
      [&](int y, float2 z) { S[y * PPITCH + g * PCH + c] = z; },
      [&](int y, float2 z) { S[y * PPITCH + 32 * PCH + c] = z; });
  __syncthreads();
  {   // g = row pair: C2R along W
    const long long r0 = pix_off(out, b, 2 * g, 0) + ch, r1 = r0 + out.sy;
    const bool has_res = res.ptr != nullptr;
    const long long q0 = has_res ? pix_off(res, b, 2 * g, 0) + ch : 0, q1 = q0 + res.sy;
    plane64_rows_inv(
        [&](int k, float2& x1, float2& x2) {
          x1 = S[(2 * g) * PPITCH + k * PCH + c];
          x2 = S[(2 * g + 1) * PPITCH + k * PCH + c];
        },
        [&](int n0, const float2* zb) {
          float ra[16], rb[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ra[j] = has_res ? load1(res, q0 + (n0 + j) * res.sx) : 0.f;
            rb[j] = has_res ? load1(res, q1 + (n0 + j) * res.sx) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            store1(out, r0 + (n0 + j) * out.sx, fmaf(zb[j].x, scale, ra[j]));
            store1(out, r1 + (n0 + j) * out.sx, fmaf(zb[j].y, scale, rb[j]));
          }
        });
  }
}

// Inverse, 9-warp variant (FFCB_FFT_INV_PLANE=2): 264 independent column tasks (no packing), then 256 row tasks.
template <int PCH>
__global__ void __launch_bounds__(PlaneCfg<PCH>::col_threads, PlaneCfg<PCH>::ctas_per_sm)
irfft2_plane64_9w_kernel(View spec, View res, View out, float scale) {
  extern __shared__ float2 S[];
  constexpr int PPITCH = PlaneCfg<PCH>::pitch;
  const int tid = threadIdx.x, c = tid % PCH, g = tid / PCH;
  const int ch = blockIdx.x * PCH + c, b = blockIdx.y;
  if (tid < PWF * PCH) {   // g = kx
    const long long o0 = pix_off(spec, b, 0, g) + 2 * ch;
    plane64_col<true>(
        [&](int ky) {
          const long long o = o0 + ky * spec.sy;
          if (spec.fmt == FFCB_F32) return __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(spec.ptr) + o));
          return make_float2(load1(spec, o), load1(spec, o + 1));
        },
        [&](int y, float2 z) { S[y * PPITCH + g * PCH + c] = z; });
  }
  __syncthreads();
  if (tid < PlaneCfg<PCH>::row_threads) {   // g = row pair
    const long long r0 = pix_off(out, b, 2 * g, 0) + ch, r1 = r0 + out.sy;
    const bool has_res = res.ptr != nullptr;
    const long long q0 = has_res ? pix_off(res, b, 2 * g, 0) + ch : 0, q1 = q0 + res.sy;
    plane64_rows_inv(
        [&](int k, float2& x1, float2& x2) {
          x1 = S[(2 * g) * PPITCH + k * PCH + c];
          x2 = S[(2 * g + 1) * PPITCH + k * PCH + c];
        },
        [&](int n0, const float2* zb) {
          float ra[16], rb[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ra[j] = has_res ? load1(res, q0 + (n0 + j) * res.sx) : 0.f;
            rb[j] = has_res ? load1(res, q1 + (n0 + j) * res.sx) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            store1(out, r0 + (n0 + j) * out.sx, fmaf(zb[j].x, scale, ra[j]));
            store1(out, r1 + (n0 + j) * out.sx, fmaf(zb[j].y, scale, rb[j]));
          }
        });
  }
}

int plane_channels() {
  const char* e = getenv("FFCB_FFT_PLANE_CH");
  return (e && atoi(e) == 4) ? 4 : 8;
}

template <int PCH, int OCC>
int launch_fwd(const ffcb_tensor* in, const ffcb_tensor* spec, cudaStream_t stream) {
  using Cfg = PlaneCfg<PCH>;
  FFCB_CUDA(cudaFuncSetAttribute(rfft2_plane64_kernel<PCH, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
  dim3 grid(in->C / PCH, in->B);
  rfft2_plane64_kernel<PCH, OCC><<<grid, Cfg::col_threads, Cfg::smem, stream>>>(make_view(*in), make_view(*spec), 1.0f / 64.0f);
  FFCB_LAUNCH_CHECK("rfft2_plane64_kernel");
  return FFCB_OK;
}

template <int PCH>
int launch_inv(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out, cudaStream_t stream) {
  using Cfg = PlaneCfg<PCH>;
  dim3 grid(out->C / PCH, out->B);
  const View vres = (residual && residual->ptr) ? make_view(*residual) : null_view();
  const char* variant = getenv("FFCB_FFT_INV_PLANE");
  if (variant && variant[0] == '2') {
    FFCB_CUDA(cudaFuncSetAttribute(irfft2_plane64_9w_kernel<PCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
    irfft2_plane64_9w_kernel<PCH><<<grid, Cfg::col_threads, Cfg::smem, stream>>>(make_view(*spec), vres, make_view(*out), 1.0f / 64.0f);
    FFCB_LAUNCH_CHECK("irfft2_plane64_9w_kernel");
    return FFCB_OK;
  }
  FFCB_CUDA(cudaFuncSetAttribute(irfft2_plane64_kernel<PCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
  irfft2_plane64_kernel<PCH><<<grid, Cfg::row_threads, Cfg::smem, stream>>>(make_view(*spec), vres, make_view(*out), 1.0f / 64.0f);
  FFCB_LAUNCH_CHECK("irfft2_plane64_kernel");
  return FFCB_OK;
}

}  // namespace

bool plane64_eligible(const ffcb_tensor* real) {
  return real->H == PN && real->W == PN && real->C % 8 == 0 && real->B <= 65535;
}

int rfft2_plane64(const ffcb_tensor* in, const ffcb_tensor* spec, cudaStream_t stream) {
  if (plane_channels() == 4) {
    const char* occ = getenv("FFCB_FFT_PLANE_OCC");      // 3: cap registers at 136 so that three CTAs share an SM
    return (occ && atoi(occ) == 3) ? launch_fwd<4, 3>(in, spec, stream) : launch_fwd<4, 2>(in, spec, stream);
  }
  return launch_fwd<8, 1>(in, spec, stream);
}

int irfft2_plane64(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out, cudaStream_t stream) {
  return plane_channels() == 4 ? launch_inv<4>(spec, residual, out, stream) : launch_inv<8>(spec, residual, out, stream);
}

}  // namespace ffcb
