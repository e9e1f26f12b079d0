// Host emulation of the channel-group planar 64x64 plane FFT kernels (lama_b200/csrc/fft_plane_cg.cu): runs every
// "thread" of a CTA through the phase functions of fft_plane_cg.cuh against a real 64 KB buffer, with the same
// global layouts the kernels address ([group][image][y][x][channels]), and checks
//   forward : spectrum (Re/Im interleaved, ortho) against a double-precision 2-D DFT
//   inverse : irfftn semantics incl. the C2R rule for a NON-Hermitian spectrum (Im of kx = 0 / 32 ignored after the
//             H inverse) + residual, against a double-precision evaluation of SURVEY.md Appendix A
// In-place hazards are modelled the way the hardware resolves them: within a phase all reads of a row see the
// pre-phase contents (the device code separates a row's reads and writes with __syncwarp), so each phase reads a
// snapshot and writes the live buffer.  Exit code 0 = pass.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../lama_b200/csrc/fft_plane_cg.cuh"

using namespace ffcb::fftc;
typedef std::complex<double> cd;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

template <int N>
static bool run_size() {
  const int WF = N / 2 + 1, T = CgCfg<N>::set_threads, PX = CgCfg<N>::px_iters;
  // ---------------------------------------------------------------- forward
  std::vector<float> plane(N * N * 4);            // [y][x][4]  (one (image, group) block of a cg=4 tensor)
  for (auto& v : plane) v = (float)frand();
  std::vector<float> smem(N * N * 4), snap;
  // load phase: pixel (y,x) -> real layout
  for (int tid = 0; tid < T; ++tid)
    for (int i = 0; i < PX; ++i) {
      int y, x;
      cg_pixel_slot<N>(tid, i, y, x);
      memcpy(&smem[cg_real_idx<N>(y, x, 0)], &plane[(y * N + x) * 4], 16);
    }
  snap = smem;
  float2* S = reinterpret_cast<float2*>(smem.data());
  for (int tid = 0; tid < T; ++tid)
    cg_fwd_rows<N>(
        tid, [&](int i) { return make_float2(snap[i], snap[i + 1]); }, [] {},
        [&](int i2, float4 v) { S[i2] = make_float2(v.x, v.y); S[i2 + 1] = make_float2(v.z, v.w); });
  snap = smem;
  const float2* Sn = reinterpret_cast<const float2*>(snap.data());
  std::vector<float> spec(N * WF * 8, 1e30f);     // [ky][kx][8]: (re, im) x 4 channels
  // column results go back in place (Nyquist column of the packed task into a side buffer), then a position loop
  std::vector<float2> NQ(N * 4, make_float2(1e30f, 1e30f));
  auto put = [&](int ky, int kx, int c, float2 z) {
    if (kx == N / 2) NQ[ky * 4 + c] = z;
    else S[cg_cplx_idx<N>(ky, kx, c)] = z;
  };
  for (int tid = 0; tid < T; ++tid) {       // the kernels compile the packed arithmetic into the first warp only
    if (tid < 32) cg_fwd_cols<N, true>(tid, [&](int i2) { return Sn[i2]; }, put);
    else cg_fwd_cols<N, false>(tid, [&](int i2) { return Sn[i2]; }, put);
  }
  int covered = 0;
  for (int tid = 0; tid < T; ++tid) {
    int ky, kx;
    cg_spec_pos0<N>(tid, ky, kx);
    for (int i = 0; i < CgCfg<N>::pos_iters; ++i) {
      if (ky < N) {
        const float2* src = kx == N / 2 ? &NQ[ky * 4] : &S[cg_cplx_idx<N>(ky, kx, 0)];
        for (int c = 0; c < 4; ++c) {
          spec[(ky * WF + kx) * 8 + 2 * c] = src[c].x / (float)N;
          spec[(ky * WF + kx) * 8 + 2 * c + 1] = src[c].y / (float)N;
        }
        ++covered;
      }
      cg_spec_pos_next<N>(ky, kx);
    }
  }
  if (covered != N * WF) { printf("position loop covered %d of %d\n", covered, N * WF); return false; }
  double err_f = 0;
  for (int c = 0; c < 4; ++c)
    for (int ky = 0; ky < N; ++ky)
      for (int kx = 0; kx < WF; ++kx) {
        cd acc = 0;
        for (int y = 0; y < N; ++y)
          for (int x = 0; x < N; ++x)
            acc += (double)plane[(y * N + x) * 4 + c] * std::polar(1.0, -2 * M_PI * (ky * y + kx * x) / (double)N);
        acc /= (double)N;
        err_f = fmax(err_f, std::abs(acc - cd(spec[(ky * WF + kx) * 8 + 2 * c], spec[(ky * WF + kx) * 8 + 2 * c + 1])));
      }
  printf("N=%d forward max abs err %.3e\n", N, err_f);

  // ---------------------------------------------------------------- inverse (generic complex spectrum)
  std::vector<float> z(N * WF * 8), res(N * N * 4);
  for (auto& v : z) v = (float)frand();
  for (auto& v : res) v = (float)frand();
  std::fill(smem.begin(), smem.end(), 1e30f);
  for (int tid = 0; tid < T; ++tid) {
    const int c = tid & 3;
    auto ldz = [&](int ky, int kx) { return make_float2(z[(ky * WF + kx) * 8 + 2 * c], z[(ky * WF + kx) * 8 + 2 * c + 1]); };
    auto sts = [&](int i2, float2 v) { S[i2] = v; };
    if (tid < 32) cg_inv_cols<N, true>(tid, ldz, sts);
    else cg_inv_cols<N, false>(tid, ldz, sts);
  }
  snap = smem;
  for (int tid = 0; tid < T; ++tid)
    cg_inv_rows<N>(
        tid,
        [&](int i2) {
          const float2* q = reinterpret_cast<const float2*>(snap.data());
          return make_float4(q[i2].x, q[i2].y, q[i2 + 1].x, q[i2 + 1].y);
        },
        [] {}, [&](int i, float2 v) { smem[i] = v.x; smem[i + 1] = v.y; });
  std::vector<float> out(N * N * 4, 1e30f);
  for (int tid = 0; tid < T; ++tid)
    for (int i = 0; i < PX; ++i) {
      int y, x;
      cg_pixel_slot<N>(tid, i, y, x);
      for (int c = 0; c < 4; ++c)
        out[(y * N + x) * 4 + c] = smem[cg_real_idx<N>(y, x, c)] / (float)N + res[(y * N + x) * 4 + c];
    }
  double err_i = 0;
  for (int c = 0; c < 4; ++c) {
    // T[r][k] = sum_q Y[q][k] e^{+2 pi i q r / 64}
    std::vector<cd> T(N * WF);
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < WF; ++k) {
        cd acc = 0;
        for (int q = 0; q < N; ++q)
          acc += cd(z[(q * WF + k) * 8 + 2 * c], z[(q * WF + k) * 8 + 2 * c + 1]) * std::polar(1.0, 2 * M_PI * q * r / (double)N);
        T[r * WF + k] = acc;
      }
    for (int r = 0; r < N; ++r)
      for (int n = 0; n < N; ++n) {
        double acc = T[r * WF].real() + ((n & 1) ? -1.0 : 1.0) * T[r * WF + N / 2].real();
        for (int k = 1; k < N / 2; ++k) acc += 2.0 * (T[r * WF + k] * std::polar(1.0, 2 * M_PI * k * n / (double)N)).real();
        acc = acc / (double)N + res[(r * N + n) * 4 + c];
        err_i = fmax(err_i, fabs(acc - out[(r * N + n) * 4 + c]));
      }
  }
  printf("N=%d inverse max abs err %.3e\n", N, err_i);
  return err_f < 2e-5 && err_i < 2e-5;
}

int main() {
  srand(7);
  const bool ok = run_size<64>() & run_size<32>();
  printf(ok ? "OK\n" : "FAILED\n");
  return ok ? 0 : 1;
}
