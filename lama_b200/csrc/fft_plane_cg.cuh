// Per-thread phases of the channel-group planar 64x64 plane FFT kernels (fft_plane_cg.cu), written against plain
// pointers so that tests/host_emul/plane_cg_emul.cpp can run every "thread" of a CTA on the host and check the
// shared-memory index algebra (swizzles, in-place row storage, packed DC / Nyquist slot) against a double DFT.
//
// One CTA = 128 threads = one plane set: 4 channels x 64 x 64 pixels of one image, 64 KB of shared memory, used in
// place by both passes:
//   real layout    R[y][x][c]   float  index  y*256 + ((x ^ (y & 7)) << 2) + c            (pixel = 16 B)
//   complex layout Sx[y][k][c]  float2 index  y*128 + ((k ^ (y & 3)) << 2) + c, k = 0..31 (slot = 4 ch x 8 B)
// Row y of either layout is the same 1 KB, so the row transforms run in place.  The XOR terms make every access of
// the kernels conflict-free: row tasks are (row y, channel pair cp) with the two channels packed two-for-one into
// one complex 64-point transform; column tasks are (column kx, channel c).  Slot k = 0 of the complex layout is the
// packed pair (Re X[y][0], Re X[y][32]): the DC and Nyquist columns of a real row transform are real (forward), and
// the C2R rule only uses their real parts (inverse) — 32 complex slots per row, not 33.
//
// Thread numbering:  row tasks  cp = tid & 1, y = tid >> 1  (a row's two tasks are neighbouring lanes of one warp: the
// in-place hand-over needs only a __syncwarp);  column tasks  c = tid & 3, kx = tid >> 2.
#pragma once
#include "fft_core.cuh"

namespace ffcb {
namespace fftc {

// The same construction serves N x N planes for N = 64 (one plane set per 128-thread CTA, 64 KB) and N = 32 (the
// 256x256 bottleneck: a plane set is 16 KB and 64 tasks per pass, so a CTA carries two of them).  Row pitch = 4N floats
// = 2N float2; the XOR swizzles only touch the low three (two) bits of the pixel (slot) index, the bank analysis is
// the same for both sizes.
constexpr int kCgThreads = 128;
template <int N> struct CgCfg {
  static constexpr int set_threads = 2 * N;                   // tasks per pass of one plane set
  static constexpr int sets = kCgThreads / set_threads;       // plane sets per CTA (1 or 2)
  static constexpr int set_floats = N * N * 4;
  static constexpr int smem_bytes = sets * set_floats * 4;
  static constexpr int px_iters = N * N / set_threads;        // pixels per thread in the load / store loops
  static constexpr int positions = N * (N / 2 + 1);           // spectrum positions of a plane
  static constexpr int pos_iters = (positions + set_threads - 1) / set_threads;
  static constexpr int fwd_smem_bytes = smem_bytes + sets * N * 4 * 8;   // + the Nyquist column of the packed task
  static constexpr int ctas_per_sm = N == 64 ? 3 : 4;      // register-limited: 168 / 128 per thread
};
constexpr int kCgSmemBytes = CgCfg<64>::smem_bytes;      // 64 KB

template <int N> FFCB_HD int cg_real_idx(int y, int x, int c) { return y * (4 * N) + ((x ^ (y & 7)) << 2) + c; }   // float index
template <int N> FFCB_HD int cg_cplx_idx(int y, int k, int c) { return y * (2 * N) + ((k ^ (y & 3)) << 2) + c; }   // float2 index

// Every pass is split into a LOAD part (shared / global memory -> the N-point register array, incl. the pre-processing
// of the inverse passes) and a POST part (post-processing + stores), with the register transform between them.  The
// kernels run the two passes of a plane set as ONE rolled loop around a single copy of the transform
// (`for pass in 0..1 { load(pass); fftN; post(pass); }`): the unrolled transform is most of the code, and with one
// copy per direction a kernel is ~25 KB of SASS — it stays in the 32 KB instruction cache instead of streaming
// ~90 KB per CTA (round-2 ncu: `stall_no_instruction` was the top stall of the first version of these kernels).

// ---- forward, row pass: R (real, two channels of row y) -> Sx (half spectra of both channels, packed slot 0)
// `ld2`: callable(float index) -> float2 (two consecutive floats); `st4`: callable(float2 index, float4) writing two
// consecutive float2
template <int N, class Load>
FFCB_HD void cg_fwd_rows_load(int tid, float2* v, Load&& ld2) {
  const int cp = tid & 1, y = tid >> 1;
#pragma unroll
  for (int x = 0; x < N; ++x) v[x] = ld2(cg_real_idx<N>(y, x, 2 * cp));
}

template <int N, class Store>
FFCB_HD void cg_fwd_rows_post(int tid, const float2* v, Store&& st4) {
  using F = RegFft<N>;
  const int cp = tid & 1, y = tid >> 1;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float2 zk = v[F::at(k)], zm = v[F::at((N - k) & (N - 1))];
    const float2 cm = make_float2(zm.x, -zm.y), df = csub(zk, cm);
    float2 a = cscale(cadd(zk, cm), 0.5f);                                   // spectrum of channel 2cp:   (Z[k] + conj Z[-k]) / 2
    float2 b = cscale(make_float2(df.y, -df.x), 0.5f);                       // spectrum of channel 2cp+1: (Z[k] - conj Z[-k]) / 2i
    if (k == 0) {
      const float2 zn = v[F::at(N / 2)];                                     // Nyquist bin: (Re A[N/2], Re B[N/2])
      a = make_float2(zk.x, zn.x);
      b = make_float2(zk.y, zn.y);
    }
    st4(cg_cplx_idx<N>(y, k, 2 * cp), make_float4(a.x, a.y, b.x, b.y));
  }
}

template <int N, class Load, class Sync, class Store>
FFCB_HD void cg_fwd_rows(int tid, Load&& ld2, Sync&& sync, Store&& st4) {
  float2 v[N];
  cg_fwd_rows_load<N>(tid, v, ld2);
  sync();
  RegFft<N>::template run<false>(v);
  cg_fwd_rows_post<N>(tid, v, st4);
}

// ---- forward, column pass: Sx column (kx, c) -> spectrum values (ky, kx) [and (ky, N/2) for the packed task kx = 0]
// `emit(ky, kx, c, value)`: the caller scales and stores
template <int N, class Load>
FFCB_HD void cg_fwd_cols_load(int tid, float2* v, Load&& ld) {
  const int c = tid & 3, kx = tid >> 2;
#pragma unroll
  for (int y = 0; y < N; ++y) v[y] = ld(cg_cplx_idx<N>(y, kx, c));
}

// MAYBE_PACKED = false: the caller guarantees kx != 0 for every thread of the warp (warps 1.. of a plane set), so the
// Hermitian-split arithmetic and the predicated Nyquist stores of the packed task are not even compiled in.
template <int N, bool MAYBE_PACKED, class Emit>
FFCB_HD void cg_fwd_cols_post(int tid, const float2* v, Emit&& emit) {
  using F = RegFft<N>;
  const int c = tid & 3, kx = tid >> 2;
  const bool packed = MAYBE_PACKED && kx == 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float2 wk = v[F::at(k)];
    if constexpr (MAYBE_PACKED) {
      const float2 wm = v[F::at((N - k) & (N - 1))];
      // packed: v = X0 + i X32 with both columns real -> Hermitian split
      const float2 cm = make_float2(wm.x, -wm.y), df = csub(wk, cm);
      const float2 x0 = cscale(cadd(wk, cm), 0.5f);
      const float2 x32 = cscale(make_float2(df.y, -df.x), 0.5f);
      emit(k, kx, c, packed ? x0 : wk);
      if (packed) emit(k, N / 2, c, x32);
    } else {
      emit(k, kx, c, wk);
    }
  }
}

template <int N, bool MAYBE_PACKED, class Load, class Emit>
FFCB_HD void cg_fwd_cols(int tid, Load&& ld, Emit&& emit) {
  float2 v[N];
  cg_fwd_cols_load<N>(tid, v, ld);
  RegFft<N>::template run<false>(v);
  cg_fwd_cols_post<N, MAYBE_PACKED>(tid, v, emit);
}

// ---- inverse, column pass: spectrum column (kx, c), complex inverse along ky -> Sx[y][kx][c].
// The packed task (kx = 0) transforms Herm(Z[.,0]) + i Herm(Z[.,N/2]) whose inverse is Re(ifft Z0) + i Re(ifft Z_N/2):
// exactly the packed slot the row pass wants (C2R: imaginary parts of bins 0 and N/2 are ignored after the H inverse).
// `ld(ky, kx)`: spectrum value of this task's channel; `st(float2 index, value)`
template <int N, bool MAYBE_PACKED, class Load>
FFCB_HD void cg_inv_cols_load(int tid, float2* v, Load&& ld) {
  const int kx = tid >> 2;
  const bool packed = MAYBE_PACKED && kx == 0;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = ld(k, kx);
  if constexpr (MAYBE_PACKED) {
#pragma unroll
    for (int k = 0; k <= N / 2; ++k) {
      const int m = (N - k) & (N - 1);
      const float2 a = v[k], b = v[m];
      float2 cc = make_float2(0.f, 0.f), d = make_float2(0.f, 0.f);
      if (packed) {
        cc = ld(k, N / 2);
        d = (m != k) ? ld(m, N / 2) : cc;
      }
      const float2 h0 = cscale(cadd(a, make_float2(b.x, -b.y)), 0.5f);          // Hermitian parts of the two columns
      const float2 h32 = cscale(cadd(cc, make_float2(d.x, -d.y)), 0.5f);
      const float2 ih = make_float2(-h32.y, h32.x);                            // i * h32
      const float2 lo = cadd(h0, ih), hi = csub(h0, ih);                       // h0 + i h32,  h0 - i h32
      v[k] = packed ? lo : a;
      if (m != k) v[m] = packed ? make_float2(hi.x, -hi.y) : b;                // conj(h0) + i conj(h32) = conj(h0 - i h32)
    }
  }
}

template <int N, class Store>
FFCB_HD void cg_inv_cols_post(int tid, const float2* v, Store&& st) {
  using F = RegFft<N>;
  const int c = tid & 3, kx = tid >> 2;
#pragma unroll
  for (int y = 0; y < N; ++y) st(cg_cplx_idx<N>(y, kx, c), v[F::at(y)]);
}

template <int N, bool MAYBE_PACKED, class Load, class Store>
FFCB_HD void cg_inv_cols(int tid, Load&& ld, Store&& st) {
  float2 v[N];
  cg_inv_cols_load<N, MAYBE_PACKED>(tid, v, ld);
  RegFft<N>::template run<true>(v);
  cg_inv_cols_post<N>(tid, v, st);
}

// ---- inverse, row pass: Sx row y (both channels of pair cp) -> R[y][x][2cp..2cp+1]   (C2R along W, unnormalised)
// `ld4(float2 index)` -> float4 = two consecutive float2 (X1[k], X2[k]); `st2(float index, float2)`
template <int N, class Load>
FFCB_HD void cg_inv_rows_load(int tid, float2* v, Load&& ld4) {
  const int cp = tid & 1, y = tid >> 1;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float4 q = ld4(cg_cplx_idx<N>(y, k, 2 * cp));
    if (k == 0) {
      v[0] = make_float2(q.x, q.z);              // (Re X1[0], Re X2[0])
      v[N / 2] = make_float2(q.y, q.w);          // (Re X1[N/2], Re X2[N/2])
    } else {
      const float2 x1 = make_float2(q.x, q.y), ix2 = make_float2(-q.w, q.z), dd = csub(x1, ix2);
      v[k] = cadd(x1, ix2);                                // X1 + i X2
      v[N - k] = make_float2(dd.x, -dd.y);                 // conj(X1) + i conj(X2) = conj(X1 - i X2)
    }
  }
}

template <int N, class Store>
FFCB_HD void cg_inv_rows_post(int tid, const float2* v, Store&& st2) {
  using F = RegFft<N>;
  const int cp = tid & 1, y = tid >> 1;
#pragma unroll
  for (int x = 0; x < N; ++x) st2(cg_real_idx<N>(y, x, 2 * cp), v[F::at(x)]);
}

template <int N, class Load, class Sync, class Store>
FFCB_HD void cg_inv_rows(int tid, Load&& ld4, Sync&& sync, Store&& st2) {
  float2 v[N];
  cg_inv_rows_load<N>(tid, v, ld4);
  sync();
  RegFft<N>::template run<true>(v);
  cg_inv_rows_post<N>(tid, v, st2);
}

// spectrum position (ky, kx), kx in [0, N/2], handled by task `tid` in the forward kernel's store loop: positions are
// walked in memory order, T = 2N tasks apart (ky >= N: past the end)
template <int N>
FFCB_HD void cg_spec_pos0(int tid, int& ky, int& kx) {
  ky = tid / (N / 2 + 1);
  kx = tid % (N / 2 + 1);
}
template <int N>
FFCB_HD void cg_spec_pos_next(int& ky, int& kx) {
  constexpr int WF = N / 2 + 1, T = 2 * N;
  kx += T % WF;
  ky += T / WF;
  if (kx >= WF) { kx -= WF; ++ky; }
}

// pixel handled by task `tid` (0 .. 2N-1 within its plane set) in iteration i of the load / store loops: lanes run along x
template <int N>
FFCB_HD void cg_pixel_slot(int tid, int i, int& y, int& x) {
  const int p = i * CgCfg<N>::set_threads + tid;
  y = p / N;
  x = p % N;
}

}  // namespace fftc
}  // namespace ffcb
