// 64x64 (and 32x32) plane FFT pair on channel-group planar tensors (include/ffc_b200.h: ffcb_tensor.cg) — the FourierUnit of the
// 512x512 bottleneck, second generation.
//
// Round 1's plane kernels (fft_plane.cu) work on channels-last tensors: one 137 KB CTA per SM transforms 8 channels,
// its load -> rows -> columns -> store phases run back to back and nothing overlaps them (ncu: 22% of DRAM bandwidth,
// 13% warps active).  Here every tensor of the chain is stored in groups of 4 complex / 4 real channels
// ([group][image][y][x][channels]), so one (image, group) plane set is ONE dense 64 KB block:
//   * a CTA is 128 threads and 64 KB of shared memory, three CTAs share an SM and hide each other's load / store
//     phases (12 warps per SM instead of 9 in one block);
//   * every global access is a whole 128-byte line: 16-byte cp.async / LDG.128 with lanes along x (real planes),
//     8 channels x 8 consecutive kx for the spectra, 8-byte half-pixels of the interleaved GEMM operand `u`;
//   * both passes run IN PLACE in the 64 KB block (fft_plane_cg.cuh), conflict-free through XOR swizzles.
// Formats:  forward  in  F32  cg=4  ->  spec BF16X2 cg=8 (operand of the spectral 1x1 GEMM, interleaved K groups)
//           inverse  spec F32 cg=8  (+ residual F32 cg=4)  ->  out BF16X2 cg=8 (operand of conv2)  or  F32 cg=4
#include <stdint.h>

#include "common.cuh"
#include "fft_plane_cg.cuh"

namespace ffcb {
namespace {

using namespace fftc;

struct CgArgs {
  // element strides of each tensor for this launch (group stride, image stride, row stride, pixel stride)
  const float* in;  long long in_sg, in_sb;  unsigned in_sy, in_sx;       // real input / residual (F32, cg 4)
  void* spec;       long long sp_sg, sp_sb;  unsigned sp_sy, sp_sx; long long sp_lo;   // spectrum (cg 8)
  void* out;        long long out_sg, out_sb; unsigned out_sy, out_sx; long long out_lo;
  float scale;
  int hints;      // L2 residency hints on (common.cuh: l2_policy)
  int sp_tiled;   // forward: the spectrum is tile-blocked (ffcb_tensor.tile = 128): sp_sg = elements per 128-position block
  int out_tiled;  // inverse: the split-bf16 output is tile-blocked: out_sg = elements per 128-pixel block
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// grid: (C/4/sets, B).  A plane set = 4-channel group of the real input = 8-channel (4 complex) group of the spectrum.
// DENSE: every tensor is the engine's dense [group][image][y][x][cg] allocation, so in-plane offsets are compile-time
// constants (one base pointer per thread + immediates instead of 64-bit address arithmetic per access — the kernels
// are issue-bound: 5.2 K instructions per thread before, a quarter of them integer).
template <int N, bool DENSE>
__global__ void __launch_bounds__(kCgThreads, CgCfg<N>::ctas_per_sm) rfft2_plane_cg_kernel(CgArgs a) {
  using Cfg = CgCfg<N>;
  constexpr unsigned WF = N / 2 + 1;
  const unsigned in_sy = DENSE ? 4u * N : a.in_sy, in_sx = DENSE ? 4u : a.in_sx;
  const unsigned sp_sy = DENSE ? 8u * WF : a.sp_sy, sp_sx = DENSE ? 8u : a.sp_sx;
  extern __shared__ __align__(16) float smem_all[];
  const int set = threadIdx.x / Cfg::set_threads, tid = threadIdx.x % Cfg::set_threads;
  const int group = blockIdx.x * Cfg::sets + set;
  float* smem = smem_all + set * Cfg::set_floats;
  float2* S = reinterpret_cast<float2*>(smem);
  {   // plane set -> shared memory (real layout), 16 bytes per pixel, lanes along x
    const float* src = a.in + (long long)group * a.in_sg + (long long)blockIdx.y * a.in_sb;
    const uint32_t base = smem_addr(smem);
    const uint64_t pol_in = l2_policy(a.hints ? 1 : 0);       // the real planes are streamed through
#pragma unroll 8
    for (int i = 0; i < Cfg::px_iters; ++i) {
      int y, x;
      cg_pixel_slot<N>(tid, i, y, x);
      cp_async16(base + 4u * (unsigned)cg_real_idx<N>(y, x, 0), src + ((unsigned)y * in_sy + (unsigned)x * in_sx),
                 pol_in);
    }
    cp_async_wait_all();
  }
  // Two passes (rows, columns) as ONE rolled loop around a single copy of the register transform: the unrolled
  // transform is most of the kernel's code, and one copy keeps the kernel close to the instruction-cache size.
  // Column results go back to shared memory in place (the Nyquist column of the packed task into a 2 KB side buffer) and
  // leave in a rolled loop over spectrum positions: 16-byte stores, the whole plane set is one contiguous stream.
  float2* NQ = reinterpret_cast<float2*>(smem_all + Cfg::sets * Cfg::set_floats) + set * (N * 4);
  float2 v[N];
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();                     // pass 0: the cp.async data of all threads; pass 1: the row results
    if (pass == 0) {
      cg_fwd_rows_load<N>(tid, v, [&](int i) { return *reinterpret_cast<const float2*>(smem + i); });
      __syncwarp();                      // a row's two tasks are neighbouring lanes: reads before the in-place writes
    } else {
      cg_fwd_cols_load<N>(tid, v, [&](int i2) { return S[i2]; });
    }
    RegFft<N>::template run<false>(v);
    if (pass == 0) {
      cg_fwd_rows_post<N>(tid, v, [&](int i2, float4 q) { *reinterpret_cast<float4*>(S + i2) = q; });
    } else {
      // a column is read and written by its own task only: in place, no barrier between the loads above and these stores
      auto put = [&](int ky, int kx, int c, float2 z) {
        if (kx == N / 2) NQ[ky * 4 + c] = z;
        else S[cg_cplx_idx<N>(ky, kx, c)] = z;
      };
      if (tid < 32) cg_fwd_cols_post<N, true>(tid, v, put);     // the packed DC / Nyquist task lives in the first warp
      else cg_fwd_cols_post<N, false>(tid, v, put);
    }
  }
  __syncthreads();
  {
    unsigned short* hi = reinterpret_cast<unsigned short*>(a.spec) + (long long)group * a.sp_sg +
                         (long long)blockIdx.y * a.sp_sb;
    const float scale = a.scale;
    const uint64_t pol_sp = l2_policy(a.hints ? 2 : 0);     // the spectrum is the next kernel's GEMM operand: keep it in L2
    int ky, kx;
    cg_spec_pos0<N>(tid, ky, kx);
    // tile-blocked spectrum (the GEMM operand layout): position m = b * N*WF + p of the flattened batch lives at
    // (m / 128) * sp_sg + group * 1024 + (m % 128) * 8 — consecutive positions stay consecutive inside a block
    unsigned short* tbase = reinterpret_cast<unsigned short*>(a.spec) + (long long)group * 1024;
    long long m = (long long)blockIdx.y * Cfg::positions + tid;
#pragma unroll 1
    for (int i = 0; i < Cfg::pos_iters; ++i) {
      if (ky < N) {
        const float4* src = kx == N / 2 ? reinterpret_cast<const float4*>(NQ + ky * 4)
                                        : reinterpret_cast<const float4*>(S + cg_cplx_idx<N>(ky, kx, 0));
        const float4 q0 = src[0], q1 = src[1];               // (re, im) of channels 0,1 | 2,3
        uint4 h, l;
        const float2 z0 = cscale(make_float2(q0.x, q0.y), scale), z1 = cscale(make_float2(q0.z, q0.w), scale);
        const float2 z2 = cscale(make_float2(q1.x, q1.y), scale), z3 = cscale(make_float2(q1.z, q1.w), scale);
        split_pair(z0.x, z0.y, h.x, l.x);
        split_pair(z1.x, z1.y, h.y, l.y);
        split_pair(z2.x, z2.y, h.z, l.z);
        split_pair(z3.x, z3.y, h.w, l.w);
        unsigned short* dst = a.sp_tiled ? tbase + ((m >> 7) * a.sp_sg + (m & 127) * 8)
                                         : hi + ((unsigned)ky * sp_sy + (unsigned)kx * sp_sx);
        st_hint_u4(dst, h, pol_sp);
        st_hint_u4(dst + a.sp_lo, l, pol_sp);
      }
      cg_spec_pos_next<N>(ky, kx);
      m += Cfg::set_threads;
    }
  }
}

// grid: (C/4/sets, B) over the REAL output's 4-channel groups; spectrum group = the same index (4 complex = 8 floats).
template <int N, bool HAS_RES, bool OUT_SPLIT, bool DENSE>
__global__ void __launch_bounds__(kCgThreads, CgCfg<N>::ctas_per_sm) irfft2_plane_cg_kernel(CgArgs a) {
  using Cfg = CgCfg<N>;
  constexpr unsigned WF = N / 2 + 1;
  const unsigned in_sy = DENSE ? 4u * N : a.in_sy, in_sx = DENSE ? 4u : a.in_sx;
  const unsigned sp_sy = DENSE ? 8u * WF : a.sp_sy, sp_sx = DENSE ? 8u : a.sp_sx;
  const unsigned out_sy = DENSE ? (OUT_SPLIT ? 8u : 4u) * N : a.out_sy, out_sx = DENSE ? (OUT_SPLIT ? 8u : 4u) : a.out_sx;
  extern __shared__ __align__(16) float smem_all[];
  const int set = threadIdx.x / Cfg::set_threads, tid = threadIdx.x % Cfg::set_threads;
  const int group = blockIdx.x * Cfg::sets + set;
  float* smem = smem_all + set * Cfg::set_floats;
  float2* S = reinterpret_cast<float2*>(smem);
  {
    const float* sp = reinterpret_cast<const float*>(a.spec) + (long long)group * a.sp_sg +
                      (long long)blockIdx.y * a.sp_sb + 2 * (tid & 3);
    const uint64_t pol_sp = l2_policy(a.hints ? 1 : 0);       // last use of the post-GEMM spectrum
    auto ldz = [&](int ky, int kx) { return ld_hint_f2(sp + ((unsigned)ky * sp_sy + (unsigned)kx * sp_sx), pol_sp); };
    // columns, then rows: one rolled loop around a single copy of the inverse register transform (see the forward kernel)
    float2 v[N];
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 0) {
        if (tid < 32) cg_inv_cols_load<N, true>(tid, v, ldz);
        else cg_inv_cols_load<N, false>(tid, v, ldz);
      } else {
        __syncthreads();
        cg_inv_rows_load<N>(tid, v, [&](int i2) { return *reinterpret_cast<const float4*>(S + i2); });
        __syncwarp();
      }
      RegFft<N>::template run<true>(v);
      if (pass == 0) cg_inv_cols_post<N>(tid, v, [&](int i2, float2 z) { S[i2] = z; });
      else cg_inv_rows_post<N>(tid, v, [&](int i, float2 z) { *reinterpret_cast<float2*>(smem + i) = z; });
    }
  }
  __syncthreads();
  {   // epilogue: whole pixels (4 channels), lanes along x: + residual, scale, convert, store
    const float* res = HAS_RES ? a.in + (long long)group * a.in_sg + (long long)blockIdx.y * a.in_sb : nullptr;
    const float scale = a.scale;
    const uint64_t pol_res = l2_policy(a.hints ? 1 : 0), pol_out = l2_policy(a.hints ? 2 : 0);   // u feeds the next contraction
#pragma unroll 1
    for (int i0 = 0; i0 < Cfg::px_iters; i0 += 8) {
      float4 q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int y, x;
        cg_pixel_slot<N>(tid, i0 + j, y, x);
        q[j] = HAS_RES ? ld_hint_f4(res + ((unsigned)y * in_sy + (unsigned)x * in_sx), pol_res)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int y, x;
        cg_pixel_slot<N>(tid, i0 + j, y, x);
        const float4 v = *reinterpret_cast<const float4*>(smem + cg_real_idx<N>(y, x, 0));
        const float r0 = fmaf(v.x, scale, q[j].x), r1 = fmaf(v.y, scale, q[j].y);
        const float r2 = fmaf(v.z, scale, q[j].z), r3 = fmaf(v.w, scale, q[j].w);
        if constexpr (OUT_SPLIT) {
          // out is cg = 8: this plane set's four channels are one half (8 bytes per plane) of the 16-byte pixel granule
          unsigned short* hi;
          if (a.out_tiled) {     // pixel m of the flattened batch: (m / 128) * out_sg + (group / 2) * 1024 + (m % 128) * 8
            const long long m = ((long long)blockIdx.y * N + y) * N + x;
            hi = reinterpret_cast<unsigned short*>(a.out) + ((m >> 7) * a.out_sg + (long long)(group >> 1) * 1024 +
                                                               (m & 127) * 8 + 4 * (group & 1));
          } else {
            hi = reinterpret_cast<unsigned short*>(a.out) + (long long)(group >> 1) * a.out_sg +
                 (long long)blockIdx.y * a.out_sb + 4 * (group & 1) + ((unsigned)y * out_sy + (unsigned)x * out_sx);
          }
          unsigned h0, l0, h1, l1;
          split_pair(r0, r1, h0, l0);
          split_pair(r2, r3, h1, l1);
          st_hint_v2(hi, make_uint2(h0, h1), pol_out);
          st_hint_v2(hi + a.out_lo, make_uint2(l0, l1), pol_out);
        } else {
          float* op = reinterpret_cast<float*>(a.out) + (long long)group * a.out_sg +
                      (long long)blockIdx.y * a.out_sb + ((unsigned)y * out_sy + (unsigned)x * out_sx);
          st_hint_f4(op, make_float4(r0, r1, r2, r3), pol_out);
        }
      }
    }
  }
}

bool offsets_fit(const ffcb_tensor* t) { return t->sy > 0 && t->sx > 0 && 64 * t->sy + 64 * t->sx < (1LL << 31); }

// plane sizes with a register transform: 64x64 (512x512 images) and 32x32 (256x256); two 32x32 plane sets share a CTA
bool plane_ok(int h, int w, int c) { return (h == 64 && w == 64) || (h == 32 && w == 32 && c % 8 == 0); }

bool real_cg4(const ffcb_tensor* t) {
  return t->cg == 4 && t->fmt == FFCB_F32 && plane_ok(t->H, t->W, t->C) && t->sx % 4 == 0 && t->sy % 4 == 0 &&
         t->sb % 4 == 0 && t->sg % 4 == 0 && ((uintptr_t)t->ptr % 16) == 0 && offsets_fit(t) && t->B <= 65535;
}

}  // namespace

// Do these views take the channel-group planar plane kernels?  (Anything with cg != 0 must: there is no other path.)
bool plane64_cg_fwd_eligible(const ffcb_tensor* in, const ffcb_tensor* spec) {
  if (!(real_cg4(in) && spec->cg == 8 && spec->fmt == FFCB_BF16X2 && spec->lo_off % 8 == 0 &&
        ((uintptr_t)spec->ptr % 16) == 0))
    return false;
  if (spec->tile) return spec->tile == 128 && spec->sg % 8 == 0;
  return spec->sx % 8 == 0 && spec->sy % 8 == 0 && spec->sb % 8 == 0 && spec->sg % 8 == 0 && offsets_fit(spec);
}

bool plane64_cg_inv_eligible(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out) {
  if (!(spec->cg == 8 && spec->fmt == FFCB_F32 && spec->sx % 2 == 0 && spec->sy % 2 == 0 && spec->sb % 2 == 0 &&
        spec->sg % 2 == 0 && ((uintptr_t)spec->ptr % 8) == 0 && offsets_fit(spec) && plane_ok(out->H, out->W, out->C) &&
        out->B <= 65535))
    return false;
  if (residual && residual->ptr && !real_cg4(residual)) return false;
  if (out->fmt == FFCB_BF16X2) {
    if (!(out->cg == 8 && out->lo_off % 4 == 0 && ((uintptr_t)out->ptr % 8) == 0 && out->C % 8 == 0)) return false;
    if (out->tile) return out->tile == 128 && out->sg % 4 == 0;
    return out->sx % 4 == 0 && out->sy % 4 == 0 && out->sb % 4 == 0 && out->sg % 4 == 0 && offsets_fit(out);
  }
  return real_cg4(out);
}

template <int N, bool DENSE>
static int launch_fwd_cg(const CgArgs& a, int groups, int batch, cudaStream_t stream) {
  using Cfg = CgCfg<N>;
  FFCB_CUDA(cudaFuncSetAttribute(rfft2_plane_cg_kernel<N, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::fwd_smem_bytes));
  dim3 grid(groups / Cfg::sets, batch);
  rfft2_plane_cg_kernel<N, DENSE><<<grid, kCgThreads, Cfg::fwd_smem_bytes, stream>>>(a);
  FFCB_LAUNCH_CHECK("rfft2_plane_cg_kernel");
  return FFCB_OK;
}

static bool dense_real(const ffcb_tensor* t, int cg) { return t->sx == cg && t->sy == (int64_t)t->W * cg; }

int rfft2_plane64_cg(const ffcb_tensor* in, const ffcb_tensor* spec, cudaStream_t stream) {
  CgArgs a{};
  a.in = reinterpret_cast<const float*>(in->ptr);
  a.in_sg = in->sg; a.in_sb = in->sb; a.in_sy = (unsigned)in->sy; a.in_sx = (unsigned)in->sx;
  a.spec = spec->ptr; a.sp_sg = spec->sg; a.sp_sb = spec->sb; a.sp_sy = (unsigned)spec->sy; a.sp_sx = (unsigned)spec->sx;
  a.sp_lo = spec->lo_off;
  a.sp_tiled = spec->tile != 0;
  a.scale = 1.0f / (float)in->H;
  a.hints = l2_hints_enabled() ? 1 : 0;
  const bool dense = dense_real(in, 4) && (spec->tile != 0 || dense_real(spec, 8));
  if (in->H == 64)
    return dense ? launch_fwd_cg<64, true>(a, in->C / 4, in->B, stream) : launch_fwd_cg<64, false>(a, in->C / 4, in->B, stream);
  return dense ? launch_fwd_cg<32, true>(a, in->C / 4, in->B, stream) : launch_fwd_cg<32, false>(a, in->C / 4, in->B, stream);
}

template <int N, bool HAS_RES, bool OUT_SPLIT, bool DENSE>
static int launch_inv_cg(const CgArgs& a, int groups, int batch, cudaStream_t stream) {
  using Cfg = CgCfg<N>;
  FFCB_CUDA(cudaFuncSetAttribute(irfft2_plane_cg_kernel<N, HAS_RES, OUT_SPLIT, DENSE>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::smem_bytes));
  dim3 grid(groups / Cfg::sets, batch);
  irfft2_plane_cg_kernel<N, HAS_RES, OUT_SPLIT, DENSE><<<grid, kCgThreads, Cfg::smem_bytes, stream>>>(a);
  FFCB_LAUNCH_CHECK("irfft2_plane_cg_kernel");
  return FFCB_OK;
}

template <int N, bool DENSE>
static int dispatch_inv_cg(const CgArgs& a, int groups, int batch, bool has_res, bool split, cudaStream_t stream) {
  if (has_res)
    return split ? launch_inv_cg<N, true, true, DENSE>(a, groups, batch, stream)
                 : launch_inv_cg<N, true, false, DENSE>(a, groups, batch, stream);
  return split ? launch_inv_cg<N, false, true, DENSE>(a, groups, batch, stream)
               : launch_inv_cg<N, false, false, DENSE>(a, groups, batch, stream);
}

int irfft2_plane64_cg(const ffcb_tensor* spec, const ffcb_tensor* residual, const ffcb_tensor* out,
                      cudaStream_t stream) {
  const bool has_res = residual && residual->ptr;
  CgArgs a{};
  if (has_res) {
    a.in = reinterpret_cast<const float*>(residual->ptr);
    a.in_sg = residual->sg; a.in_sb = residual->sb; a.in_sy = (unsigned)residual->sy; a.in_sx = (unsigned)residual->sx;
  }
  a.spec = spec->ptr; a.sp_sg = spec->sg; a.sp_sb = spec->sb; a.sp_sy = (unsigned)spec->sy; a.sp_sx = (unsigned)spec->sx;
  a.out = out->ptr; a.out_sg = out->sg; a.out_sb = out->sb; a.out_sy = (unsigned)out->sy; a.out_sx = (unsigned)out->sx;
  a.out_lo = out->lo_off;
  a.out_tiled = out->tile != 0;
  a.scale = 1.0f / (float)out->H;
  a.hints = l2_hints_enabled() ? 1 : 0;
  const bool split = out->fmt == FFCB_BF16X2;
  const bool dense = dense_real(spec, 8) && (out->tile != 0 || dense_real(out, split ? 8 : 4)) &&
                     (!has_res || dense_real(residual, 4));
  if (out->H == 64)
    return dense ? dispatch_inv_cg<64, true>(a, out->C / 4, out->B, has_res, split, stream)
                 : dispatch_inv_cg<64, false>(a, out->C / 4, out->B, has_res, split, stream);
  return dense ? dispatch_inv_cg<32, true>(a, out->C / 4, out->B, has_res, split, stream)
               : dispatch_inv_cg<32, false>(a, out->C / 4, out->B, has_res, split, stream);
}

}  // namespace ffcb
