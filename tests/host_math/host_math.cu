// host_math.cu -- runs the pass kernel's per-pixel code (geocalib_b200/csrc/gclm_device.cuh: make_pass_const,
// row_accumulate -> pixel_accumulate / model_scalars / the divisional special cases) on the HOST, so that the
// CUDA source's math is checked against the oracle without a GPU (tests/test_host_math.py).  Test
// infrastructure only: nothing in geocalib_b200/ uses this file.  The hardware approximations
// (rsqrt.approx / rcp.approx) and the packed fp32x2 intrinsics fall back to plain C on the host, everything
// else is the code the GPU runs.
#include <math.h>

#include "../../geocalib_b200/csrc/gclm_device.cuh"

using namespace gclm;

namespace {

struct Fields {
  int H, W;
  const float *up, *lat, *uconf, *lconf;  // (2,H,W), (H,W), (H,W), (H,W)
};

template <int MODEL, bool GEN_D, bool PACK>
void run(const PassConst& pc, const Fields& f, float up_scale, float lat_scale, double* out) {
  using V = typename std::conditional<PACK, f2, float>::type;
  constexpr int NACC = Dims<MODEL>::NACC;
  const FieldMask fm = {true, true, true, true};
  const RobustConst rc = make_robust_const(up_scale, lat_scale);
  const size_t N = (size_t)f.H * f.W;
  for (int k = 0; k < NACC; k++) out[k] = 0.0;
  for (int y = 0; y < f.H; y++) {
    for (int x0 = 0; x0 < f.W; x0 += 4) {
      const size_t o = (size_t)y * f.W + x0;
      float u[4];
      for (int i = 0; i < 4; i++) u[i] = pixel_u(pc, x0 + i);
      const float4 ox = make_float4(f.up[o], f.up[o + 1], f.up[o + 2], f.up[o + 3]);
      const float4 oy = make_float4(f.up[N + o], f.up[N + o + 1], f.up[N + o + 2], f.up[N + o + 3]);
      // what sin_plane_kernel leaves in scratch
      const float4 ol = make_float4(sinf(f.lat[o]), sinf(f.lat[o + 1]), sinf(f.lat[o + 2]), sinf(f.lat[o + 3]));
      const float4 cu = make_float4(f.uconf[o], f.uconf[o + 1], f.uconf[o + 2], f.uconf[o + 3]);
      const float4 cl = make_float4(f.lconf[o], f.lconf[o + 1], f.lconf[o + 2], f.lconf[o + 3]);
      V acc[NACC];
      for (int k = 0; k < NACC; k++) acc[k] = splat<V>(0.f);
      row_accumulate<MODEL, GEN_D, true, PACK>(pc, fm, make_row_cols<PACK>(pc, u), pixel_v(pc, y), ox, oy, ol, cu, cl, rc, acc);
      for (int k = 0; k < NACC; k++) out[k] += (double)hsum(acc[k]);  // fp64 sum: only the pixel math is tested
    }
  }
  // what the kernel's per-image epilogue does to the reduced sums before the solver sees them
  finish_sums<MODEL, GEN_D, double>(out, pc.k1, pc.k2, up_scale);
}

template <int MODEL>
void dispatch(bool gen_d, bool pack, const PassConst& pc, const Fields& f, float us, float ls, double* out) {
  if (gen_d) { if (pack) run<MODEL, true, true>(pc, f, us, ls, out); else run<MODEL, true, false>(pc, f, us, ls, out); }
  else { if (pack) run<MODEL, false, true>(pc, f, us, ls, out); else run<MODEL, false, false>(pc, f, us, ls, out); }
}

}  // namespace

// out: NACC doubles in the kernel's accumulator layout [H upper triangle | G | cost_up | cost_lat] (sums over
// the pixels, before the 1/N and scale^2 of the epilogue).  force_gen_d runs the general-focal-direction code
// even where the log-focal specialisation applies (both must agree).  W % 4 == 0.  Returns NACC or < 0.
extern "C" int gclm_host_math_system(int model, int H, int W, const float* up, const float* lat, const float* uconf,
                                     const float* lconf, const float* cam, const float* grav, int spherical,
                                     int log_focal, int rpf, int pack, int force_gen_d, float up_scale,
                                     float lat_scale, double* out) {
  if (W % 4 != 0 || H <= 0 || W <= 0 || model < 0 || model > 3) return -1;
  PassConst pc;
  make_pass_const(cam, grav, spherical != 0, log_focal != 0, rpf != 0, pc);
  const Fields f = {H, W, up, lat, uconf, lconf};
  const bool gen_d = force_gen_d || !log_focal || rpf;  // as gclm_solve launches the passes
  switch (model) {
    case PINHOLE: dispatch<PINHOLE>(gen_d, pack, pc, f, up_scale, lat_scale, out); return Dims<PINHOLE>::NACC;
    case SIMPLE_RADIAL: dispatch<SIMPLE_RADIAL>(gen_d, pack, pc, f, up_scale, lat_scale, out); return Dims<SIMPLE_RADIAL>::NACC;
    case RADIAL: dispatch<RADIAL>(gen_d, pack, pc, f, up_scale, lat_scale, out); return Dims<RADIAL>::NACC;
    default: dispatch<SIMPLE_DIVISIONAL>(gen_d, pack, pc, f, up_scale, lat_scale, out); return Dims<SIMPLE_DIVISIONAL>::NACC;
  }
}

// The adjoint pass of the differentiable LM (gclm_device.cuh: adjoint_pixel, the body of lm_adjoint_kernel) on the
// host: out = (5, H, W) planes [d/dox, d/doy, d/dlat, d/dconf_up, d/dconf_lat] of lambda . G.
namespace {
template <int MODEL>
void run_adjoint(const PassConst& pc, const Fields& f, const float* lam, float up_scale, float lat_scale, float* out) {
  const float BIG = 1e18f;
  const RobustConst rc_probe = make_robust_const(BIG, BIG);
  const float unscale = Forms<MODEL>::HUBER ? BIG : 1.f;
  const float iu2 = 1.f / (up_scale * up_scale), il2 = 1.f / (lat_scale * lat_scale);
  const size_t N = (size_t)f.H * f.W;
  for (int y = 0; y < f.H; y++)
    for (int x = 0; x < f.W; x++) {
      const size_t n = (size_t)y * f.W + x;
      float d[5];
      adjoint_pixel<MODEL>(pc, lam, rc_probe, unscale, iu2, il2, pixel_u(pc, x), pixel_v(pc, y), f.up[n], f.up[N + n], f.lat[n],
                           f.uconf[n], f.lconf[n], d);
      for (int k = 0; k < 5; k++) out[(size_t)k * N + n] = d[k];
    }
}
}  // namespace

extern "C" int gclm_host_math_adjoint(int model, int H, int W, const float* up, const float* lat, const float* uconf,
                                      const float* lconf, const float* cam, const float* grav, const float* lam,
                                      float up_scale, float lat_scale, float* out) {
  if (H <= 0 || W <= 0 || model < 0 || model > 3) return -1;
  PassConst pc;
  make_pass_const(cam, grav, true, true, false, pc);  // spherical manifold + log focal: the training configuration
  const Fields f = {H, W, up, lat, uconf, lconf};
  switch (model) {
    case PINHOLE: run_adjoint<PINHOLE>(pc, f, lam, up_scale, lat_scale, out); break;
    case SIMPLE_RADIAL: run_adjoint<SIMPLE_RADIAL>(pc, f, lam, up_scale, lat_scale, out); break;
    case RADIAL: run_adjoint<RADIAL>(pc, f, lam, up_scale, lat_scale, out); break;
    default: run_adjoint<SIMPLE_DIVISIONAL>(pc, f, lam, up_scale, lat_scale, out); break;
  }
  return 0;
}

// ---- the solver-side pieces of the pass kernel's epilogue on the host: fixed-size damped-system solve, SPD inverse
// (covariance), parameter updates.  n = 3, 4, 5. ----
namespace {
template <int N>
int solve_n(const float* A_in, const float* b_in, float* x) {
  float A[N][N], b[N];
  for (int i = 0; i < N; i++) { b[i] = b_in[i]; for (int j = 0; j < N; j++) A[i][j] = A_in[i * N + j]; }
  if (!cholesky_solve_fixed<N>(A, b)) return 1;
  for (int i = 0; i < N; i++) x[i] = b[i];
  return 0;
}
template <int N>
int inverse_n(const float* A_in, float* out) {
  float A[N][N], Ai[N][N];
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) A[i][j] = A_in[i * N + j];
  if (!spd_inverse_fixed<N>(A, Ai)) return 1;
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) out[i * N + j] = Ai[i][j];
  return 0;
}
}  // namespace

extern "C" int gclm_host_math_solve(int n, const float* A, const float* b, float* x) {
  switch (n) { case 3: return solve_n<3>(A, b, x); case 4: return solve_n<4>(A, b, x); case 5: return solve_n<5>(A, b, x); }
  return -1;
}
extern "C" int gclm_host_math_spd_inverse(int n, const float* A, float* Ainv) {
  switch (n) { case 3: return inverse_n<3>(A, Ainv); case 4: return inverse_n<4>(A, Ainv); case 5: return inverse_n<5>(A, Ainv); }
  return -1;
}
// cam (8) and grav (3) are updated in place with the step delta = [g0 g1 | f | k..] (lm_optimizer.py:518-549)
extern "C" void gclm_host_math_update(int model, float* cam, float* grav, const float* delta, int spherical, int log_focal) {
  gravity_update(grav, delta[0], delta[1], spherical != 0);
  focal_update(cam, delta[2], log_focal != 0);
  if (model != PINHOLE) dist_update(model, cam, delta + 3);
}
