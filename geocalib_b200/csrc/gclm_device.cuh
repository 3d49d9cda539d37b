// gclm_device.cuh -- per-pixel perspective-field math and the per-image LM solver (device code).
//
// Written from the closed form in SURVEY.md Appendix A (verified there against the reference's
// setup_system in fp64) -- NOT a translation of the reference's tensor program.  Reference
// functions each block replaces are cited as geocalib/<file>:<lines>.
//
// Algebra used by the kernel (and independently checked by oracle/gclm_oracle.c, which follows
// the reference's explicit 2x2 / 3x2 Jacobian chain instead):
//   * In 2-D, I - n n^T = n_perp n_perp^T, so every column of the up-field Jacobian
//     J_vecnorm(p) z  (misc.py:262-281) is  n_perp * (n_perp . z) / |p|.  The Gauss-Newton
//     system therefore only needs the scalars  s_j = (n_perp . z_j)/|p|  and the scalar
//     residual e = n_perp . up_obs:   J^T W J += w s s^T,   J^T W r += w s e.
//   * The Huber weight still uses the full 2-vector residual (lm_optimizer.py:291-293).
//   * sin(asin(clamp(l))) == clamp(l)  (perspective_fields.py:208-209 + lm_optimizer.py:262).
//   * Scaling symmetry of the radial models: f -> f e^d with k_i -> k_i e^{2 i d} leaves the image
//     unchanged, hence the log-focal column is   J_f = J_X - sum_i 2 i k_i J_{k_i}   with the
//     "pinhole-like" column X:  up row  c * gamma,  latitude row  -s_u * guv  (verified symbolically for
//     simple_radial, radial and simple_divisional against the closed forms of SURVEY Appendix A).  The
//     streaming loop accumulates the columns [g0 g1 X k..] and finish_sums() maps the reduced system to
//     [g0 g1 f k..] once per image -- the Bc / A / Q terms never enter the hot loop.
//   * Huber with the scale folded in: with c2 = |r|^2, i' = rsqrt(max(c2, a^2)) (= 1/(a sqrt(max(chi,1)))),
//     w' = conf i' (= w / a):  cost conf = c2 w' (2/a - i'),  and J^T W J, J^T W r are accumulated with w'
//     and multiplied by a once per image (lm_optimizer.py:61-87, 291-313).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

// The math below also compiles for the host (tests/host_math: the kernel's per-pixel code is checked
// against the oracle without a GPU); hardware approximations and packed intrinsics then fall back to
// plain C.
#define GCLM_HD __host__ __device__

namespace gclm {

enum : int { PINHOLE = 0, SIMPLE_RADIAL = 1, RADIAL = 2, SIMPLE_DIVISIONAL = 3 };

__host__ __device__ constexpr int num_dist(int model) {
  return model == PINHOLE ? 0 : (model == RADIAL ? 2 : 1);
}
template <int MODEL>
struct Dims {
  static constexpr int K = num_dist(MODEL);
  static constexpr int NF = 3 + K;                // full columns: g0 g1 f k..
  static constexpr int NH = NF * (NF + 1) / 2;    // upper triangle of J^T W J
  static constexpr int NACC = NH + NF + 2;        // + J^T W r + two cost sums
};
constexpr int MAX_NF = 5;
constexpr int MAX_NACC = 15 + 5 + 2;

// ------------------------------------------------------------------------------------------
// fast hardware approximations used in the streaming pass only
// ------------------------------------------------------------------------------------------
GCLM_HD __forceinline__ float fast_rsqrt(float x) {
#ifdef __CUDA_ARCH__
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.f / sqrtf(x);
#endif
}
GCLM_HD __forceinline__ float fast_rcp(float x) {
#ifdef __CUDA_ARCH__
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.f / x;
#endif
}

// ------------------------------------------------------------------------------------------
// f2: two pixels per register pair.  Blackwell (sm_100) has packed fp32x2 arithmetic
// (PTX fma.rn.f32x2 / mul / add -> SASS FFMA2 / FMUL2 / FADD2, operands may be a register pair,
// a broadcast scalar or a uniform register, each with a negate modifier): one issue slot per two
// pixels.  Measured on B200 (tools/ubench_fp32x2.cu): FMUL2/FADD2 issue at the scalar rate
// (2x flops), FFMA2 at ~0.55x the scalar rate (1.1x flops) -- the pass kernel is issue-bound, so
// halving the issue slots of ~80 % of its instructions is what matters.
// The same source below compiles for V = float (one pixel) and V = f2 (two pixels).
// ------------------------------------------------------------------------------------------
// keep the scalar math functions visible next to the packed overloads declared below
using ::fmaf;
using ::fmaxf;
using ::fminf;

struct f2 {
  float2 v;
  GCLM_HD __forceinline__ f2() {}
  GCLM_HD __forceinline__ explicit f2(float s) : v(make_float2(s, s)) {}
  GCLM_HD __forceinline__ f2(float a, float b) : v(make_float2(a, b)) {}
  GCLM_HD __forceinline__ explicit f2(float2 a) : v(a) {}
};
GCLM_HD __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
GCLM_HD __forceinline__ float2 bc2(f2 a) { return a.v; }
template <class T> struct is_f2 { static constexpr bool value = false; };
template <> struct is_f2<f2> { static constexpr bool value = true; };
#define GCLM_ANY_F2_2(A, B) typename = typename std::enable_if<is_f2<A>::value || is_f2<B>::value>::type
#define GCLM_ANY_F2_3(A, B, C) \
  typename = typename std::enable_if<is_f2<A>::value || is_f2<B>::value || is_f2<C>::value>::type

template <class A, class B, GCLM_ANY_F2_2(A, B)>
GCLM_HD __forceinline__ f2 operator*(A a, B b) {
#ifdef __CUDA_ARCH__
  return f2(__fmul2_rn(bc2(a), bc2(b)));
#else
  const float2 x = bc2(a), y = bc2(b);
  return f2(x.x * y.x, x.y * y.y);
#endif
}
template <class A, class B, GCLM_ANY_F2_2(A, B)>
GCLM_HD __forceinline__ f2 operator+(A a, B b) {
#ifdef __CUDA_ARCH__
  return f2(__fadd2_rn(bc2(a), bc2(b)));
#else
  const float2 x = bc2(a), y = bc2(b);
  return f2(x.x + y.x, x.y + y.y);
#endif
}
GCLM_HD __forceinline__ f2 operator-(f2 a) { return f2(-a.v.x, -a.v.y); }
template <class A, class B, GCLM_ANY_F2_2(A, B)>
GCLM_HD __forceinline__ f2 operator-(A a, B b) {
  const float2 nb = bc2(b);
#ifdef __CUDA_ARCH__
  return f2(__fadd2_rn(bc2(a), make_float2(-nb.x, -nb.y)));
#else
  const float2 x = bc2(a);
  return f2(x.x - nb.x, x.y - nb.y);
#endif
}
template <class A, class B, class C, GCLM_ANY_F2_3(A, B, C)>
GCLM_HD __forceinline__ f2 fmaf(A a, B b, C c) {
#ifdef __CUDA_ARCH__
  return f2(__ffma2_rn(bc2(a), bc2(b), bc2(c)));
#else
  const float2 x = bc2(a), y = bc2(b), z = bc2(c);
  return f2(::fmaf(x.x, y.x, z.x), ::fmaf(x.y, y.y, z.y));
#endif
}
GCLM_HD __forceinline__ f2 fmaxf(f2 a, float s) { return f2(::fmaxf(a.v.x, s), ::fmaxf(a.v.y, s)); }
GCLM_HD __forceinline__ f2 fminf(f2 a, float s) { return f2(::fminf(a.v.x, s), ::fminf(a.v.y, s)); }
GCLM_HD __forceinline__ f2 fast_rsqrt(f2 a) { return f2(fast_rsqrt(a.v.x), fast_rsqrt(a.v.y)); }
GCLM_HD __forceinline__ f2 fast_rcp(f2 a) { return f2(fast_rcp(a.v.x), fast_rcp(a.v.y)); }
GCLM_HD __forceinline__ float hsum(float a) { return a; }
GCLM_HD __forceinline__ float hsum(f2 a) { return a.v.x + a.v.y; }
template <class V> GCLM_HD __forceinline__ V splat(float s);
template <> GCLM_HD __forceinline__ float splat<float>(float s) { return s; }
template <> GCLM_HD __forceinline__ f2 splat<f2>(float s) { return f2(s); }

// ------------------------------------------------------------------------------------------
// per-image constants of one pass (hoisted out of the pixel loop)
// ------------------------------------------------------------------------------------------
struct PassConst {
  float inv_fx, inv_fy, cx, cy;
  float a, b, c;                     // gravity (unit)
  float T00, T01, T10, T11, T20, T21; // tangent basis (3x2): spherical or (roll,pitch)
  float dsx, dsy;                    // focal direction d = (-u dsx, -v dsy): 1,1 (log) or 1/fx,1/fy
  float k1, k2;
  int div_any_cold;                  // simple_divisional: some pixel of this image needs the literal masked formulas
};

// Per-launch constants of the robust loss with the scale folded in (see the header comment).
struct RobustConst {
  float up_s2, up_2os, lat_s2, lat_2os, lat_ratio;  // a^2, 2/a for both residuals; a_lat / a_up
  float up_inv2, lat_inv2;                          // 1/a^2 (unfolded form)
};
GCLM_HD inline RobustConst make_robust_const(float up_scale, float lat_scale) {
  RobustConst rc;
  rc.up_s2 = up_scale * up_scale; rc.up_2os = 2.f / up_scale;
  rc.lat_s2 = lat_scale * lat_scale; rc.lat_2os = 2.f / lat_scale;
  rc.lat_ratio = lat_scale / up_scale;
  rc.up_inv2 = 1.f / rc.up_s2; rc.lat_inv2 = 1.f / rc.lat_s2;
  return rc;
}

// ---- gravity helpers: geocalib/misc.py:181-259, geocalib/gravity.py:62-101 ----
GCLM_HD inline void householder(const float g[3], float hv[3], float& beta) {
  float sigma = g[0] * g[0] + g[1] * g[1];
  const float xpiv = g[2];
  const float nrm = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  if (sigma < 1e-7f) sigma += 1e-7f;  // misc.py:202-204
  const float vpiv = xpiv < 0.f ? xpiv - nrm : -sigma / (xpiv + nrm);
  beta = 2.f * vpiv * vpiv / (sigma + vpiv * vpiv);
  hv[0] = g[0] / vpiv; hv[1] = g[1] / vpiv; hv[2] = 1.f;
}

GCLM_HD inline void gravity_rp(const float g[3], float& roll, float& pitch) {
  // gravity.py:15,62-67: note the 1e-4 eps in the roll read-back
  const float r = asinf(-g[0] / (sqrtf(1.f - g[2] * g[2]) + 1e-4f));
  const float sgn = g[0] > 0.f ? 1.f : (g[0] < 0.f ? -1.f : 0.f);
  roll = g[1] < 0.f ? r : -r - 3.14159265358979323846f * sgn;
  pitch = asinf(g[2]);
}

GCLM_HD inline void normalize3(float g[3]) {  // F.normalize, eps 1e-12 (gravity.py:26)
  const float n = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);
  g[0] /= n; g[1] /= n; g[2] /= n;
}

// cam = [w h fx fy cx cy k1 k2], g unit.  rpf: use the (roll,pitch) basis and the non-log focal
// direction (estimate_uncertainty -> setup_system(as_rpf=True), lm_optimizer.py:411-416,481-483).
GCLM_HD inline void make_pass_const(const float* cam, const float* g, bool spherical,
                                       bool log_focal, bool rpf, PassConst& pc) {
  pc.inv_fx = 1.f / cam[2]; pc.inv_fy = 1.f / cam[3];
  pc.cx = cam[4]; pc.cy = cam[5]; pc.k1 = cam[6]; pc.k2 = cam[7];
  pc.a = g[0]; pc.b = g[1]; pc.c = g[2];
  if (spherical && !rpf) {  // SphericalManifold.J_plus: (I - beta v v^T)[:, :2]
    float hv[3], beta;
    householder(g, hv, beta);
    pc.T00 = 1.f - beta * hv[0] * hv[0]; pc.T01 = -beta * hv[0] * hv[1];
    pc.T10 = -beta * hv[1] * hv[0];      pc.T11 = 1.f - beta * hv[1] * hv[1];
    pc.T20 = -beta * hv[2] * hv[0];      pc.T21 = -beta * hv[2] * hv[1];
  } else {  // Gravity.J_rp (gravity.py:69-101)
    float roll, pitch;
    gravity_rp(g, roll, pitch);
    float sr, cr, sp, cp;
    sincosf(roll, &sr, &cr);
    sincosf(pitch, &sp, &cp);
    pc.T00 = -cr * cp; pc.T10 = sr * cp; pc.T20 = 0.f;
    pc.T01 = sr * sp;  pc.T11 = cr * sp; pc.T21 = cp;
  }
  if (log_focal && !rpf) { pc.dsx = 1.f; pc.dsy = 1.f; }
  else { pc.dsx = pc.inv_fx; pc.dsy = pc.inv_fy; }
  // simple_divisional: the closed form applies to every pixel unless k1 == 0 (all masked branches fire) or
  // the root is clamped somewhere: 1 - 4 k1 r^2 < 1e-6 first happens in a corner (r^2 is evaluated exactly as
  // pixel_accumulate does, and fp32 u*u, fma are monotone in |u|, |v|, so the corner bounds every pixel)
  {
    const float ux = fmaxf(fabsf((0.f - pc.cx) * pc.inv_fx), fabsf(((cam[0] - 1.f) - pc.cx) * pc.inv_fx));
    const float vy = fmaxf(fabsf((0.f - pc.cy) * pc.inv_fy), fabsf(((cam[1] - 1.f) - pc.cy) * pc.inv_fy));
    const float r2max = fmaf(ux, ux, vy * vy);
    pc.div_any_cold = (pc.k1 == 0.f || fmaf(-4.f * pc.k1, r2max, 1.f) < 1e-6f) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// camera-model scalars at one pixel
//   distort scale s, offset o = phi*uv, J_o = A I + Bc uv uv^T, ds/dk = sk, do/dk = psi*uv,
//   undistort scale su, d(su uv)/duv = su I + Q uv uv^T, d(su uv)/dk = om*uv
// ------------------------------------------------------------------------------------------
template <int K, class V = float>
struct ModelScalars {
  V s, phi, A, Bc, su, Q;
  V sk[K > 0 ? K : 1], psi[K > 0 ? K : 1], om[K > 0 ? K : 1];
};

// Literal restatement of SimpleDivisional with its masked denominators (camera.py:829-942).
// Used only when the cancellation-free closed form does not apply: k1 == 0 (trivial init, the
// masked branches make ds/dk = 0 and J_o = 2/r^2 I - 8/r^4 uv uv^T) or
// 1 - 4 k1 r^2 < 1e-6 (clamped square root).  Cold path.
GCLM_HD __noinline__ void divisional_literal(float k1, float r2, ModelScalars<1, float>& m) {
  auto mf = [](float d) { return d == 0.f ? 1e6f : d; };
  const float x = 1.f - 4.f * k1 * r2;
  const float rad = 1.f - sqrtf(fmaxf(x, 0.f));
  const float den = 2.f * k1 * r2;
  m.s = den == 0.f ? 1.f : rad / mf(den);
  const float t0 = fmaxf(x, 1e-6f), t1 = sqrtf(t0), t32 = t0 * t1;
  {
    const float d1 = t1 * 2.f * r2, d2 = k1 * r2 * r2;
    m.phi = (4.f * d2 - (1.f - t1) * d1) / mf(d1 * d2);
  }
  {
    const float d1 = 2.f * k1 * t1, d2 = 2.f * r2 * k1 * k1;
    m.sk[0] = (2.f * d2 - (1.f - t1) * d1) / mf(d1 * d2);
  }
  m.psi[0] = 16.f / mf(4.f * t32) - 2.f / mf(r2 * t1 * k1) + (1.f - t1) / mf((r2 * k1) * (r2 * k1));
  m.A = 4.f / mf(2.f * r2 * t1) - (1.f - t1) / mf(k1 * r2 * r2);
  m.Bc = -16.f / mf(4.f * t1 * r2 * r2) + (32.f * k1) / mf(4.f * r2 * t32) - 4.f / mf(r2 * r2 * t1) +
         4.f * (1.f - t1) / mf(k1 * r2 * r2 * r2);
  const float dn = 1.f + k1 * r2;
  m.su = 1.f / mf(dn);
  m.Q = -2.f * k1 / (mf(dn) * mf(dn));
  m.om[0] = -r2 / mf(dn * dn);
}

// r2 == 0 ? if_zero : otherwise, per pixel
GCLM_HD __forceinline__ float select_zero(float r2, float if_zero, float otherwise) {
  return r2 == 0.f ? if_zero : otherwise;
}
GCLM_HD __forceinline__ f2 select_zero(f2 r2, float if_zero, f2 otherwise) {
  return f2(r2.v.x == 0.f ? if_zero : otherwise.v.x, r2.v.y == 0.f ? if_zero : otherwise.v.y);
}

// LIT (simple_divisional, scalar only): evaluate the reference's literal masked formulas instead of
// the closed form -- taken for pixels where the closed form does not apply (see divisional_is_cold).
// K0 (simple_divisional): k1 == 0 for the whole image (trivial init, or a zero distortion prior): the
// literal formulas then reduce to the constants below, evaluated packed.
template <int MODEL, class V, bool LIT = false, bool K0 = false>
GCLM_HD __forceinline__ void model_scalars(const PassConst& pc, V r2, ModelScalars<num_dist(MODEL), V>& m) {
  if constexpr (MODEL == PINHOLE) {  // camera.py:532-562 (never read: the K == 0 paths use constants)
    m.s = splat<V>(1.f); m.phi = m.A = m.Bc = m.Q = splat<V>(0.f); m.su = splat<V>(1.f);
  } else if constexpr (MODEL == SIMPLE_RADIAL) {  // camera.py:611-660
    m.s = fmaf(pc.k1, r2, splat<V>(1.f));
    m.phi = splat<V>(2.f * pc.k1); m.A = m.phi; m.Bc = splat<V>(0.f);
    m.sk[0] = r2; m.psi[0] = splat<V>(2.f);
    m.su = fmaf(-pc.k1, r2, splat<V>(1.f));
    m.Q = splat<V>(-2.f * pc.k1);
    m.om[0] = -r2;
  } else if constexpr (MODEL == RADIAL) {  // camera.py:712-786
    const V r4 = r2 * r2;
    m.s = fmaf(pc.k2, r4, fmaf(pc.k1, r2, splat<V>(1.f)));
    m.phi = fmaf(4.f * pc.k2, r2, splat<V>(2.f * pc.k1)); m.A = m.phi; m.Bc = splat<V>(8.f * pc.k2);
    m.sk[0] = r2; m.sk[1] = r4;
    m.psi[0] = splat<V>(2.f); m.psi[1] = 4.f * r2;
    const float b1 = -pc.k1, b2 = fmaf(3.f * pc.k1, pc.k1, -pc.k2);
    m.su = fmaf(b2, r4, fmaf(b1, r2, splat<V>(1.f)));
    m.Q = fmaf(4.f * b2, r2, splat<V>(2.f * b1));
    m.om[0] = fmaf(6.f * pc.k1, r4, -r2);
    m.om[1] = -r4;
  } else {  // SIMPLE_DIVISIONAL, camera.py:829-942
    // With t = sqrt(1 - 4 k1 r^2), w = 1 + t and 1 - t = 4 k1 r^2 / w every masked quotient of
    // the reference collapses to a cancellation-free form that is finite at r^2 = 0:
    //   s = 2/w, phi = A = 8 k1/(t w^2), Bc = 32 k1^2 (1+3t)/(t^3 w^3),
    //   ds/dk = 4 r^2/(t w^2), psi = 4 (1 + 2t - t^2)/(t^3 w^2).
    // The literal masked formulas (cold path) take over when k1 == 0 or the root is clamped.
    if constexpr (K0) {
      // divisional_literal at k1 == 0: every k1-proportional denominator is masked to 1e6 and its
      // numerator is 0; what survives is A = 4/(2 r^2), Bc = -16/(4 r^4) - 4/r^4, psi = 16/4 - 2/1e6
      // (and the masked values at the centre pixel, r^2 == 0)
      const V ir = fast_rcp(r2);
      m.s = splat<V>(1.f); m.phi = splat<V>(0.f);
      m.A = select_zero(r2, 4.f / 1e6f, 2.f * ir);
      m.Bc = select_zero(r2, -16.f / 1e6f - 4.f / 1e6f, -8.f * ir * ir);
      m.sk[0] = splat<V>(0.f); m.psi[0] = splat<V>(16.f / 4.f - 2.f / 1e6f);
      m.su = splat<V>(1.f); m.Q = splat<V>(0.f); m.om[0] = -r2;
      return;
    }
    if constexpr (LIT) {
      static_assert(!LIT || !is_f2<V>::value, "literal path is scalar");
      if constexpr (!is_f2<V>::value) divisional_literal(pc.k1, r2, m);
      return;
    }
    const V x = fmaf(-4.f * pc.k1, r2, splat<V>(1.f));
    const V xs = fmaxf(x, 1e-6f);
    const V it = fast_rsqrt(xs), t = xs * it, w = 1.f + t, iw = fast_rcp(w);
    m.s = 2.f * iw;
    const V g4 = (it * m.s) * m.s, it2 = it * it;  // 4 / (t w^2)
    m.phi = (2.f * pc.k1) * g4; m.A = m.phi;
    m.Bc = (8.f * pc.k1 * pc.k1) * fmaf(3.f, t, splat<V>(1.f)) * g4 * it2 * iw;  // general focal direction only
    m.sk[0] = r2 * g4;
    m.psi[0] = fmaf(t, 2.f - t, splat<V>(1.f)) * g4 * it2;
    const V su = fast_rcp(fmaf(pc.k1, r2, splat<V>(1.f)));
    const V su2 = su * su;
    m.su = su;
    m.Q = (-2.f * pc.k1) * su2;  // general focal direction only
    m.om[0] = -(r2 * su2);
  }
}

// ------------------------------------------------------------------------------------------
// one pixel: residuals, robust weights, Jacobian rows, accumulation
//   acc layout: [H upper triangle row-major (NH)] [G (NF)] [cost_up] [cost_lat]
//   GEN_D: general focal direction (non-log focal or the final as_rpf pass): columns [g0 g1 f k..].
//   !GEN_D (the LM passes with log focal): columns [g0 g1 X k..] (see the header comment; for pinhole
//   X == f), weights w' = w / a_up; finish_sums() turns the reduced sums into the system the solver sees.
//   flags: runtime presence of the inputs (compile-time true when FULL)
// ------------------------------------------------------------------------------------------
struct FieldMask { bool up, lat, up_conf, lat_conf; };

// A/B knobs, measured on B200 (profiles/r2_ab_forms.txt; ncu source view: every packed fp32x2 instruction --
// FFMA2, FMUL2 or FADD2 -- costs the same two FMA-pipe cycles, so only the NUMBER of packed instructions and
// the register pressure matter, not which of the three they are):
//   * pinhole sits at the 128-register cap of its 2 x 8-warp CTA shape and at 0.89-0.97 of the HBM roofline:
//     neither the column-invariant forms (register moves: 0.97 -> 0.92) nor the scale-folded Huber form
//     (-1 %) pay, so it keeps the round-1 formulas;
//   * the distortion models take the X basis and the folded Huber form (fewer packed instructions);
//     trading FMAs for hoisted products + adds (COLS) leaves the instruction count unchanged and costs
//     registers: off by default (0.661 vs 0.652 of the roofline on simple_radial at equal clocks).
#ifndef GCLM_PIN_GEOM_NEW
#define GCLM_PIN_GEOM_NEW 0
#endif
#ifndef GCLM_PIN_HUBER_NEW
#define GCLM_PIN_HUBER_NEW 0
#endif
#ifndef GCLM_DIST_COLS
#define GCLM_DIST_COLS 0
#endif
template <int MODEL> struct Forms {
  static constexpr bool COLS = MODEL != PINHOLE ? GCLM_DIST_COLS != 0 : GCLM_PIN_GEOM_NEW != 0;  // column-invariant forms (LICM food)
  static constexpr bool HUBER = MODEL != PINHOLE || GCLM_PIN_HUBER_NEW;    // scale-folded Huber
};

// accumulator update a*b + c: packed by default; -DGCLM_SCALAR_ACC issues it as two scalar FFMA (A/B knob)
GCLM_HD __forceinline__ float fma_acc(float a, float b, float c) { return ::fmaf(a, b, c); }
GCLM_HD __forceinline__ f2 fma_acc(f2 a, f2 b, f2 c) {
#if defined(GCLM_SCALAR_ACC)
  return f2(::fmaf(a.v.x, b.v.x, c.v.x), ::fmaf(a.v.y, b.v.y, c.v.y));
#else
  return fmaf(a, b, c);
#endif
}

// Column-invariant terms of a pixel (pair): a warp walks DOWN a strip, so anything that depends on x only
// can be computed once per strip (-DGCLM_HOIST; costs registers) instead of once per row:
//   p0x = a - c u,  zx_j = T0j - T2j u  (the up-field gravity rows are alpha zx_j + beta zy_j with the
//   row-uniform zy_j = T1j - T2j v).
template <class V>
struct Col { V u, p0x, zx0, zx1; };
template <class V>
GCLM_HD __forceinline__ Col<V> make_col(const PassConst& pc, V u) {
  Col<V> c;
  c.u = u;
  c.p0x = fmaf(-pc.c, u, splat<V>(pc.a));
  c.zx0 = fmaf(-pc.T20, u, splat<V>(pc.T00));
  c.zx1 = fmaf(-pc.T21, u, splat<V>(pc.T01));
  return c;
}

template <int MODEL, bool GEN_D, bool FULL, class V, bool LIT = false, bool K0 = false, bool UP = true, bool LAT = true>
GCLM_HD __forceinline__ void pixel_accumulate(const PassConst& pc, const FieldMask fm, const Col<V>& col,
                                                 float v, V ox, V oy, V sin_olat, V cu,
                                                 V cl, const RobustConst& rc,
                                                 V (&acc)[Dims<MODEL>::NACC]) {
  const V u = col.u;
  constexpr int K = Dims<MODEL>::K, NF = Dims<MODEL>::NF, NH = Dims<MODEL>::NH;
  // the masked / literal divisional formulas do not obey the scaling symmetry: their focal row is computed
  // directly and (LIT, k1 != 0) moved into the X basis; K0 means k1 == 0, where X == f
  constexpr bool DIRECT_F = GEN_D || K0 || LIT;
  // UP / LAT: compile-time halves of the pixel (a warp may own only the up-field or only the latitude rows)
  const bool has_up = UP && (FULL || fm.up), has_lat = LAT && (FULL || fm.lat);
  V r2;
  if constexpr (Forms<MODEL>::COLS) r2 = (u * u) + splat<V>(v * v);  // u*u is column-invariant: hoisted out of the row loop
  else r2 = fmaf(u, u, splat<V>(v * v));
  ModelScalars<K, V> m;
  model_scalars<MODEL, V, LIT, K0>(pc, r2, m);

  V su_[NF], sl_[NF];  // Jacobian row scalars (up: along n_perp; latitude)
  V e = splat<V>(0.f), w_up = e, r_lat = e, w_lat = e;
  V dx = e, uvd = e;
  float dy = 0.f;
  if constexpr (GEN_D) { dx = -(u * pc.dsx); dy = -v * pc.dsy; uvd = fmaf(u, dx, splat<V>(v * dy)); }

  if (has_up) {  // ---- up field: perspective_fields.py:47-182 ----
    const V p0x = Forms<MODEL>::COLS ? col.p0x : fmaf(-pc.c, u, splat<V>(pc.a));
    const float p0y = fmaf(-pc.c, v, pc.b);
    const float zy0 = fmaf(-pc.T20, v, pc.T10), zy1 = fmaf(-pc.T21, v, pc.T11);  // row-uniform
    V px = p0x, py = splat<V>(p0y), iota = splat<V>(0.f);
    if constexpr (K > 0) {
      if constexpr (Forms<MODEL>::COLS) iota = (u * p0x) + splat<V>(v * p0y);  // column-invariant + row-uniform
      else iota = fmaf(u, p0x, splat<V>(v * p0y));
      const V pi = m.phi * iota;
      px = fmaf(m.s, p0x, pi * u);
      py = fmaf(m.s, splat<V>(p0y), pi * v);
    }
    const V rho2 = fmaf(px, px, py * py);
    const V irho = fast_rsqrt(fmaxf(rho2, 1e-24f));  // F.normalize eps (perspective_fields.py:79)
    const V upx = px * irho, upy = py * irho;
    const V Ap = -(upy * irho), Bp = upx * irho;  // (n_perp . z)/rho = Ap z_x + Bp z_y
    const V Nu = fmaf(Ap, u, Bp * v);
    // residual (lm_optimizer.py:266) and its component along n_perp
    const V rx = ox - upx, ry = oy - upy;
    e = fmaf(upx, oy, -(upy * ox));
    // Huber (lm_optimizer.py:79-87; chi + 1e-8 == chi in fp32 for chi > 1), scale folded in
    const V c2 = fmaf(rx, rx, ry * ry);
    if constexpr (Forms<MODEL>::HUBER) {
      const V isx = fast_rsqrt(fmaxf(c2, rc.up_s2));
      w_up = isx * cu;
      acc[NH + NF] = fmaf(c2 * w_up, rc.up_2os - isx, acc[NH + NF]);
    } else {  // cost*conf = 2 chi w - min(chi,1) conf, w = conf/sqrt(max(chi,1)); unscaled weights (finish_sums: x1)
      const V chi = c2 * rc.up_inv2;
      const V isx = fast_rsqrt(fmaxf(chi, 1.f));
      w_up = isx * cu;
      acc[NH + NF] = fmaf(chi + chi, w_up, fmaf(-fminf(chi, 1.f), cu, acc[NH + NF]));
    }
    // Jacobian columns.  With M z = s z + phi uv (uv.z) and z_j = (T0j - u T2j, T1j - v T2j):
    //   s_j = T0j alpha + T1j beta - T2j gamma,  alpha = s Ap + phi Nu u, beta = s Bp + phi Nu v,
    //   gamma = alpha u + beta v = Nu (s + phi r^2)        (Nu = Ap u + Bp v)
    if constexpr (K == 0 && !Forms<MODEL>::COLS) {
      su_[0] = fmaf(Ap, pc.T00, fmaf(Bp, pc.T10, -pc.T20 * Nu));
      su_[1] = fmaf(Ap, pc.T01, fmaf(Bp, pc.T11, -pc.T21 * Nu));
      su_[2] = GEN_D ? -pc.c * fmaf(Ap, dx, Bp * dy) : pc.c * Nu;
    } else if constexpr (K == 0) {
      su_[0] = fmaf(Ap, col.zx0, Bp * zy0);
      su_[1] = fmaf(Ap, col.zx1, Bp * zy1);
      if constexpr (!GEN_D) su_[2] = pc.c * Nu;
      else su_[2] = -pc.c * fmaf(Ap, dx, Bp * dy);
    } else {
      const V pn = m.phi * Nu;
      const V alpha = fmaf(pn, u, m.s * Ap), beta = fmaf(pn, v, m.s * Bp);
      const V pr2 = m.phi * r2;
      const V gamma = Nu * (m.s + pr2);
      if constexpr (Forms<MODEL>::COLS) {
        su_[0] = fmaf(alpha, col.zx0, beta * zy0);
        su_[1] = fmaf(alpha, col.zx1, beta * zy1);
      } else {
        su_[0] = fmaf(alpha, pc.T00, fmaf(beta, pc.T10, -pc.T20 * gamma));
        su_[1] = fmaf(alpha, pc.T01, fmaf(beta, pc.T11, -pc.T21 * gamma));
      }
      const V Np0 = fmaf(Ap, p0x, Bp * p0y);  // (n_perp . p0)/rho
      const V iN = iota * Nu;
#pragma unroll
      for (int k = 0; k < K; k++) su_[3 + k] = fmaf(m.sk[k], Np0, m.psi[k] * iN);
      if constexpr (!DIRECT_F) {
        su_[2] = pc.c * gamma;  // column X
      } else if constexpr (!GEN_D) {
        // log-focal column (d = -uv): c gamma - iota Nu (Bc r^2 + A + phi) - phi r^2 Np0
        const V t = iN * (fmaf(m.Bc, r2, m.A) + m.phi);
        su_[2] = fmaf(-pr2, Np0, fmaf(pc.c, gamma, -t));
        if constexpr (LIT) su_[2] = fmaf(2.f * pc.k1, su_[3], su_[2]);  // X = f + 2 k1 k (one-parameter model)
      } else {
        const V Nd = fmaf(Ap, dx, Bp * dy);
        const V p0d = fmaf(p0x, dx, splat<V>(p0y * dy));
        const V t = fmaf(iota * m.Bc, uvd, m.phi * (p0d - pc.c * uvd));
        su_[2] = fmaf(fmaf(iota, m.A, -pc.c * m.s), Nd, fmaf(m.phi * uvd, Np0, Nu * t));
      }
    }
  }

  if (has_lat) {  // ---- latitude field: perspective_fields.py:185-275 ----
    const V S = Forms<MODEL>::COLS ? (pc.a * u) + splat<V>(pc.b * v)  // column-invariant + row-uniform
                                   : fmaf(pc.a, u, splat<V>(pc.b * v));
    V m2, ln;
    if constexpr (K == 0) { m2 = r2 + 1.f; ln = S + pc.c; }
    else { m2 = fmaf(m.su * m.su, r2, splat<V>(1.f)); ln = fmaf(m.su, S, splat<V>(pc.c)); }
    const V im = fast_rsqrt(m2);
    const V l = ln * im;
    const V lc = fminf(fmaxf(l, -1.f + 1e-6f), 1.f - 1e-6f);
    r_lat = sin_olat - lc;  // lm_optimizer.py:270-271
    const V c2 = r_lat * r_lat;
    if constexpr (Forms<MODEL>::HUBER) {
      const V isx = fast_rsqrt(fmaxf(c2, rc.lat_s2));
      w_lat = isx * cl;
      acc[NH + NF + 1] = fmaf(c2 * w_lat, rc.lat_2os - isx, acc[NH + NF + 1]);
      if (rc.lat_ratio != 1.f) w_lat = w_lat * rc.lat_ratio;  // both residuals share the a_up of finish_sums
    } else {
      const V chi = c2 * rc.lat_inv2;
      const V isx = fast_rsqrt(fmaxf(chi, 1.f));
      w_lat = isx * cl;
      acc[NH + NF + 1] = fmaf(chi + chi, w_lat, fmaf(-fminf(chi, 1.f), cl, acc[NH + NF + 1]));
    }
    if constexpr (K == 0 && !Forms<MODEL>::COLS) {
      sl_[0] = im * fmaf(pc.T00, u, splat<V>(fmaf(pc.T10, v, pc.T20)));
      sl_[1] = im * fmaf(pc.T01, u, splat<V>(fmaf(pc.T11, v, pc.T21)));
    } else if constexpr (K == 0) {
      sl_[0] = im * ((pc.T00 * u) + splat<V>(fmaf(pc.T10, v, pc.T20)));
      sl_[1] = im * ((pc.T01 * u) + splat<V>(fmaf(pc.T11, v, pc.T21)));
    } else {
      if constexpr (Forms<MODEL>::COLS) {
        sl_[0] = im * fmaf(m.su, (pc.T00 * u) + splat<V>(pc.T10 * v), splat<V>(pc.T20));
        sl_[1] = im * fmaf(m.su, (pc.T01 * u) + splat<V>(pc.T11 * v), splat<V>(pc.T21));
      } else {
        sl_[0] = im * fmaf(m.su, fmaf(pc.T00, u, splat<V>(pc.T10 * v)), splat<V>(pc.T20));
        sl_[1] = im * fmaf(m.su, fmaf(pc.T01, u, splat<V>(pc.T11 * v)), splat<V>(pc.T21));
      }
    }
    const V li = l * im;
    if constexpr (!GEN_D) {
      if constexpr (K == 0) {
        sl_[2] = im * fmaf(-l, im, splat<V>(pc.c));  // = -gamma.uv with r^2 = m^2 - 1
      } else {
        const V guv = im * fmaf(-(li * m.su), r2, S);
#pragma unroll
        for (int k = 0; k < K; k++) sl_[3 + k] = m.om[k] * guv;
        if constexpr (!DIRECT_F) {
          sl_[2] = -(m.su * guv);  // column X
        } else {
          sl_[2] = -(fmaf(m.Q, r2, m.su) * guv);
          if constexpr (LIT) sl_[2] = fmaf(2.f * pc.k1, sl_[3], sl_[2]);
        }
      }
    } else {
      const V U = m.su * u, Vv = m.su * v;
      const V gx = fmaf(-U, li, splat<V>(pc.a)) * im, gy = fmaf(-Vv, li, splat<V>(pc.b)) * im;
      const V gd = fmaf(gx, dx, gy * dy);
      const V guv = fmaf(gx, u, gy * v);
      sl_[2] = fmaf(m.su, gd, m.Q * uvd * guv);
      if constexpr (K > 0) {
#pragma unroll
        for (int k = 0; k < K; k++) sl_[3 + k] = m.om[k] * guv;
      }
    }
  }

  // ---- J^T W J, J^T W r (lm_optimizer.py:346-359) ----
  V wu[NF], wl[NF];
#pragma unroll
  for (int i = 0; i < NF; i++) {
    wu[i] = has_up ? w_up * su_[i] : splat<V>(0.f);
    wl[i] = has_lat ? w_lat * sl_[i] : splat<V>(0.f);
  }
  int idx = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) {
#pragma unroll
    for (int j = i; j < NF; j++) {
      V h = acc[idx];
      if (has_up) h = fma_acc(wu[i], su_[j], h);
      if (has_lat) h = fma_acc(wl[i], sl_[j], h);
      acc[idx++] = h;
    }
  }
#pragma unroll
  for (int i = 0; i < NF; i++) {
    V g = acc[NH + i];
    if (has_up) g = fma_acc(wu[i], e, g);
    if (has_lat) g = fma_acc(wl[i], r_lat, g);
    acc[NH + i] = g;
  }
}

// Reduced sums of one image -> the system the solver sees: undo the folded robust scale (x a_up) and, for
// the LM passes of the distortion models (!GEN_D), map the columns [g0 g1 X k..] to [g0 g1 f k..]:
//   J_f = J_X + sum_i kappa_i J_{k_i},  kappa_i = -2 i k_i.   T = float (kernel) or double (host test).
template <int MODEL, bool GEN_D, class T>
GCLM_HD inline void finish_sums(T* s, float k1, float k2, float up_scale) {
  constexpr int K = Dims<MODEL>::K, NF = Dims<MODEL>::NF, NH = Dims<MODEL>::NH;
  if constexpr (K > 0 && !GEN_D) {
    const T kap[2] = {(T)(-2.f * k1), (T)(-4.f * k2)};
    T Hf[NF][NF];
    {
      int idx = 0;
      for (int i = 0; i < NF; i++)
        for (int j = i; j < NF; j++) { Hf[i][j] = s[idx]; Hf[j][i] = s[idx]; idx++; }
    }
    // column/row 2: f = X + sum kappa_i k_i
    T hff = Hf[2][2];
    for (int i = 0; i < K; i++) {
      hff += (T)2 * kap[i] * Hf[2][3 + i];
      for (int j = 0; j < K; j++) hff += kap[i] * kap[j] * Hf[3 + i][3 + j];
    }
    T col[NF];
    for (int r = 0; r < NF; r++) {
      T c = Hf[r][2];
      for (int i = 0; i < K; i++) c += kap[i] * Hf[r][3 + i];
      col[r] = c;
    }
    for (int r = 0; r < NF; r++) { Hf[r][2] = col[r]; Hf[2][r] = col[r]; }
    Hf[2][2] = hff;
    T gf = s[NH + 2];
    for (int i = 0; i < K; i++) gf += kap[i] * s[NH + 3 + i];
    s[NH + 2] = gf;
    int idx = 0;
    for (int i = 0; i < NF; i++)
      for (int j = i; j < NF; j++) s[idx++] = Hf[i][j];
  }
  if constexpr (Forms<MODEL>::HUBER)
    for (int i = 0; i < NH + NF; i++) s[i] *= (T)up_scale;
}

// simple_divisional: does pixel (u, v) need the literal masked formulas?  k1 == 0 (trivial init:
// every masked branch fires) or 1 - 4 k1 r^2 < 1e-6 (clamped square root).
GCLM_HD __forceinline__ bool divisional_is_cold(const PassConst& pc, float u, float v) {
  return pc.k1 == 0.f || fmaf(-4.f * pc.k1, fmaf(u, u, v * v), 1.f) < 1e-6f;
}
// Cold path: one pixel through the scalar code with the literal formulas; returns its contribution
// in d[NACC] (local memory is fine here) so that the hot loop's packed accumulators never have
// their address taken.
template <int MODEL, bool GEN_D, bool FULL, bool LIT>
GCLM_HD __noinline__ void pixel_scalar(const PassConst& pc, const FieldMask fm, float u, float v, float ox,
                                           float oy, float sin_olat, float cu, float cl, const RobustConst& rc,
                                           float* d) {
  float a1[Dims<MODEL>::NACC];
#pragma unroll
  for (int i = 0; i < Dims<MODEL>::NACC; i++) a1[i] = 0.f;
  pixel_accumulate<MODEL, GEN_D, FULL, float, LIT>(pc, fm, make_col<float>(pc, u), v, ox, oy, sin_olat, cu, cl, rc, a1);
#pragma unroll
  for (int i = 0; i < Dims<MODEL>::NACC; i++) d[i] = a1[i];
}

// Normalised image coordinates of pixel column x / row y (integer pixel centres, camera.py:335-338).
GCLM_HD __forceinline__ float pixel_u(const PassConst& pc, int x) { return ((float)x - pc.cx) * pc.inv_fx; }
GCLM_HD __forceinline__ float pixel_v(const PassConst& pc, int y) { return ((float)y - pc.cy) * pc.inv_fy; }

// ---- differentiable LM: one pixel of the adjoint pass (gclm_engine.cu: lm_adjoint_kernel) ----
// d(lambda . G)/dx for the five input values of one pixel, G = sum_px [w_up s e + w_lat l r_lat] the LM gradient
// J^T W r with s, l the Jacobian row scalars (full columns [g0 g1 f k..]), e = n_perp . o, r_lat = sin(lat_obs) -
// clamp(l), w = conf / sqrt(max(chi, 1)) (Huber, lm_optimizer.py:61-87):
//   d/do        = (lambda.s) [ w n_perp + e dw/do ],       dw/do = -w (o - up)/|o - up|^2 for chi > 1, else 0
//   d/dlat_obs  = (lambda.l) cos(lat_obs) w for chi <= 1, 0 for chi > 1 (there w r_lat = conf a sign(r_lat))
//   d/dconf     = (lambda.s) e / sqrt(max(chi,1)),   (lambda.l) r_lat / sqrt(max(chi,1))
// The rows come from the SAME pixel_accumulate the pass kernel uses: they enter its G linearly (G_i = w s_i e), so it
// is probed with unit confidences, a robust scale that makes every probe an inlier (rc_probe = make_robust_const(BIG,
// BIG), unscale = BIG for the scale-folded Huber builds, else 1) and observations chosen such that e' = 1 (o' = up +
// n_perp) resp. r_lat' = 1.  out = [d/dox, d/doy, d/dlat, d/dconf_up, d/dconf_lat].  Host-compilable
// (tests/host_math): checked per pixel against fp64 finite differences of the oracle system without a GPU.
template <int MODEL>
GCLM_HD inline void adjoint_pixel(const PassConst& pc, const float* lm, const RobustConst& rc_probe, float unscale,
                                  float inv_up_s2, float inv_lat_s2, float u, float v, float ox, float oy, float la,
                                  float cu, float cl, float out[5]) {
  constexpr int K = Dims<MODEL>::K, NF = Dims<MODEL>::NF, NH = Dims<MODEL>::NH, NACC = Dims<MODEL>::NACC;
  const FieldMask fm_up = {true, false, true, false}, fm_lat = {false, true, false, true};
  const Col<float> col = make_col<float>(pc, u);
  const float sl = sinf(la), cla = cosf(la);
  const bool lit = MODEL == SIMPLE_DIVISIONAL && divisional_is_cold(pc, u, v);
  // predicted up vector / latitude and the residuals of the robust terms (general focal direction code: columns
  // [g0 g1 f k..] directly; the log-focal direction d = -uv is selected by dsx = dsy = 1 in make_pass_const)
  ModelScalars<K, float> m;
  const float r2 = u * u + v * v;
  if (lit) model_scalars<MODEL, float, MODEL == SIMPLE_DIVISIONAL>(pc, r2, m); else model_scalars<MODEL, float>(pc, r2, m);
  const float p0x = pc.a - pc.c * u, p0y = pc.b - pc.c * v;
  float px = p0x, py = p0y;
  if (K > 0) { const float io = u * p0x + v * p0y; px = m.s * p0x + m.phi * io * u; py = m.s * p0y + m.phi * io * v; }
  const float irho = fast_rsqrt(fmaxf(px * px + py * py, 1e-24f));
  const float upx = px * irho, upy = py * irho;
  const float rx = ox - upx, ry = oy - upy;
  const float e = upx * oy - upy * ox;
  const float c2u = rx * rx + ry * ry;
  const float chi_u = c2u * inv_up_s2;
  const float isx_u = chi_u > 1.f ? fast_rsqrt(chi_u) : 1.f;
  const float su = K > 0 ? m.su : 1.f;
  const float U = su * u, V = su * v;
  const float l = (U * pc.a + V * pc.b + pc.c) * fast_rsqrt(U * U + V * V + 1.f);
  const float lc = fminf(fmaxf(l, -1.f + 1e-6f), 1.f - 1e-6f);
  const float r_lat = sl - lc;
  const float chi_l = r_lat * r_lat * inv_lat_s2;
  const float isx_l = chi_l > 1.f ? fast_rsqrt(chi_l) : 1.f;
  float ap[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) ap[i] = 0.f;
  const float qx = upx - upy, qy = upy + upx;  // up + n_perp: e(q) = upx qy - upy qx = 1
  if (lit) pixel_accumulate<MODEL, true, false, float, MODEL == SIMPLE_DIVISIONAL>(pc, fm_up, col, v, qx, qy, sl, 1.f, 1.f, rc_probe, ap);
  else pixel_accumulate<MODEL, true, false, float>(pc, fm_up, col, v, qx, qy, sl, 1.f, 1.f, rc_probe, ap);
  float ls = 0.f;
#pragma unroll
  for (int i = 0; i < NF; i++) ls = fmaf(lm[i], ap[NH + i], ls);
  ls *= unscale;
  float aq[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) aq[i] = 0.f;
  if (lit) pixel_accumulate<MODEL, true, false, float, MODEL == SIMPLE_DIVISIONAL>(pc, fm_lat, col, v, ox, oy, lc + 1.f, 1.f, 1.f, rc_probe, aq);
  else pixel_accumulate<MODEL, true, false, float>(pc, fm_lat, col, v, ox, oy, lc + 1.f, 1.f, 1.f, rc_probe, aq);
  float ll = 0.f;
#pragma unroll
  for (int i = 0; i < NF; i++) ll = fmaf(lm[i], aq[NH + i], ll);
  ll *= unscale;
  const float w_u = cu * isx_u, w_l = cl * isx_l;
  float dwx = 0.f, dwy = 0.f;
  if (chi_u > 1.f) { const float k = -(w_u * fast_rcp(c2u)); dwx = k * rx; dwy = k * ry; }
  out[0] = ls * (w_u * (-upy) + e * dwx);
  out[1] = ls * (w_u * upx + e * dwy);
  out[2] = chi_l > 1.f ? 0.f : ll * cla * w_l;
  out[3] = ls * e * isx_u;
  out[4] = ll * r_lat * isx_l;
}

// Four consecutive pixels of one image row (one float4 of every plane; `ol` = sin(latitude)) into the
// accumulators: the body of lm_pass_kernel's streaming loop, also run on the host by tests/host_math.
// The column terms of the four consecutive pixels a lane owns (two pairs when packed).
template <bool PACK>
struct RowCols {
  using V = typename std::conditional<PACK, f2, float>::type;
  Col<V> c[PACK ? 2 : 4];
  float u[4];
};
template <bool PACK>
GCLM_HD __forceinline__ RowCols<PACK> make_row_cols(const PassConst& pc, const float (&u)[4]) {
  RowCols<PACK> r;
#pragma unroll
  for (int i = 0; i < 4; i++) r.u[i] = u[i];
  if constexpr (PACK) {
    r.c[0] = make_col<f2>(pc, f2(u[0], u[1]));
    r.c[1] = make_col<f2>(pc, f2(u[2], u[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) r.c[i] = make_col<float>(pc, u[i]);
  }
  return r;
}

template <int MODEL, bool GEN_D, bool FULL, bool PACK>
GCLM_HD __forceinline__ void row_accumulate(const PassConst& pc, const FieldMask fm, const RowCols<PACK>& cols, float v,
                                            const float4 ox, const float4 oy, const float4 ol, const float4 cu,
                                            const float4 cl, const RobustConst& rc,
                                            typename std::conditional<PACK, f2, float>::type (&acc)[Dims<MODEL>::NACC],
                                            const PassConst* pc_mem = nullptr, const RobustConst* rc_mem = nullptr) {
  // pc_mem / rc_mem: the constants where they live in (shared) memory -- handed to the noinline cold path so
  // that the register copies `pc` / `rc` never have their address taken (a per-thread local-memory copy
  // otherwise: 110 B x threads of DRAM writes per launch, seen in ncu)
  constexpr int NACC = Dims<MODEL>::NACC;
  using V = typename std::conditional<PACK, f2, float>::type;
  const float sl0 = ol.x, sl1 = ol.y, sl2 = ol.z, sl3 = ol.w;
  const float (&u)[4] = cols.u;
  if constexpr (MODEL == SIMPLE_DIVISIONAL) {
    if (pc.div_any_cold) {  // uniform per image: k1 == 0 or a clamped square root somewhere in the image
      if constexpr (PACK) {
        if (pc.k1 == 0.f) {  // whole image (first pass from the trivial init, zero prior): packed k1 == 0 form
          pixel_accumulate<MODEL, GEN_D, FULL, f2, false, true>(pc, fm, cols.c[0], v, f2(ox.x, ox.y), f2(oy.x, oy.y),
                                                                f2(sl0, sl1), f2(cu.x, cu.y), f2(cl.x, cl.y), rc, acc);
          pixel_accumulate<MODEL, GEN_D, FULL, f2, false, true>(pc, fm, cols.c[1], v, f2(ox.z, ox.w), f2(oy.z, oy.w),
                                                                f2(sl2, sl3), f2(cu.z, cu.w), f2(cl.z, cl.w), rc, acc);
          return;
        }
      }
      // rare pixels (clamped square root; with scalar math also the k1 == 0 passes) go through the
      // literal masked formulas, one at a time
      const bool c0 = divisional_is_cold(pc, u[0], v), c1 = divisional_is_cold(pc, u[1], v);
      const bool c2 = divisional_is_cold(pc, u[2], v), c3 = divisional_is_cold(pc, u[3], v);
      if (c0 | c1 | c2 | c3) {
        float d[NACC];
        const float us[4] = {u[0], u[1], u[2], u[3]};
        const float oxs[4] = {ox.x, ox.y, ox.z, ox.w}, oys[4] = {oy.x, oy.y, oy.z, oy.w};
        const float sls[4] = {sl0, sl1, sl2, sl3}, cus[4] = {cu.x, cu.y, cu.z, cu.w}, cls[4] = {cl.x, cl.y, cl.z, cl.w};
        const bool cold[4] = {c0, c1, c2, c3};
        constexpr float share = is_f2<V>::value ? 0.5f : 1.f;  // packed: both halves of an accumulator sum to d
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
          // the k1 == 0 literal path yields the f column directly; X == f there (kappa = 0)
          const PassConst& pcm = pc_mem ? *pc_mem : pc;
          const RobustConst& rcm = rc_mem ? *rc_mem : rc;
          if (cold[i]) pixel_scalar<MODEL, GEN_D, FULL, true>(pcm, fm, us[i], v, oxs[i], oys[i], sls[i], cus[i], cls[i], rcm, d);
          else pixel_scalar<MODEL, GEN_D, FULL, false>(pcm, fm, us[i], v, oxs[i], oys[i], sls[i], cus[i], cls[i], rcm, d);
#pragma unroll
          for (int k = 0; k < NACC; k++) acc[k] = acc[k] + splat<V>(share * d[k]);
        }
        return;
      }
    }
  }
  if constexpr (PACK) {
    pixel_accumulate<MODEL, GEN_D, FULL, f2>(pc, fm, cols.c[0], v, f2(ox.x, ox.y), f2(oy.x, oy.y),
                                             f2(sl0, sl1), f2(cu.x, cu.y), f2(cl.x, cl.y), rc, acc);
    pixel_accumulate<MODEL, GEN_D, FULL, f2>(pc, fm, cols.c[1], v, f2(ox.z, ox.w), f2(oy.z, oy.w),
                                             f2(sl2, sl3), f2(cu.z, cu.w), f2(cl.z, cl.w), rc, acc);
  } else {
    pixel_accumulate<MODEL, GEN_D, FULL, float>(pc, fm, cols.c[0], v, ox.x, oy.x, sl0, cu.x, cl.x, rc, acc);
    pixel_accumulate<MODEL, GEN_D, FULL, float>(pc, fm, cols.c[1], v, ox.y, oy.y, sl1, cu.y, cl.y, rc, acc);
    pixel_accumulate<MODEL, GEN_D, FULL, float>(pc, fm, cols.c[2], v, ox.z, oy.z, sl2, cu.z, cl.z, rc, acc);
    pixel_accumulate<MODEL, GEN_D, FULL, float>(pc, fm, cols.c[3], v, ox.w, oy.w, sl3, cu.w, cl.w, rc, acc);
  }
}

// ------------------------------------------------------------------------------------------
// small dense linear algebra (n <= 5), single thread
// ------------------------------------------------------------------------------------------
// In-place lower Cholesky; returns false when not positive definite (torch.linalg.cholesky raises,
// lm_optimizer.py:129-133).
GCLM_HD inline bool cholesky(float* A, int n) {
  for (int j = 0; j < n; j++) {
    float d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.f)) return false;
    d = sqrtf(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      float s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return true;
}
GCLM_HD inline void cholesky_solve(const float* L, int n, float* b) {
  for (int i = 0; i < n; i++) {
    float s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    float s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}
// The same factorisation and substitution with a compile-time size: fully unrolled, everything in registers
// (the runtime-n versions above index local-memory arrays).  Same operations in the same order -> same bits.
template <int N>
GCLM_HD __forceinline__ bool cholesky_solve_fixed(float (&A)[N][N], float (&b)[N]) {
  // One reciprocal square root per column instead of a square root and N divisions (IEEE sqrt and division are
  // ~100-cycle dependent instruction sequences; this runs on ONE thread per image and is on the critical path of
  // small batches).  id[j] = 1 / L[j][j].
  float id[N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    float d = A[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
    if (!(d > 0.f)) return false;
#ifdef __CUDA_ARCH__
    id[j] = __frsqrt_rn(d);
#else
    id[j] = 1.f / sqrtf(d);
#endif
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      float t = A[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= A[i][k] * A[j][k];
      A[i][j] = t * id[j];
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    float t = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) t -= A[i][k] * b[k];
    b[i] = t * id[i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    float t = b[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) t -= A[k][i] * b[k];
    b[i] = t * id[i];
  }
  return true;
}

// Inverse of a symmetric positive definite matrix through its Cholesky factor, compile-time size, registers
// only.  Returns false when A is not positive definite (the caller then takes the pivoting inverse below).
template <int N>
GCLM_HD __forceinline__ bool spd_inverse_fixed(const float (&A)[N][N], float (&Ainv)[N][N]) {
  float L[N][N];
#pragma unroll
  for (int j = 0; j < N; j++) {
    float d = A[j][j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k];
    if (!(d > 0.f)) return false;
    d = sqrtf(d);
    L[j][j] = d;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      float t = A[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t / d;
    }
  }
#pragma unroll
  for (int c = 0; c < N; c++) {  // solve L L^T x = e_c
    float y[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      float t = i == c ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < i; k++) t -= L[i][k] * y[k];
      y[i] = t / L[i][i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
      float t = y[i];
#pragma unroll
      for (int k = i + 1; k < N; k++) t -= L[k][i] * y[k];
      y[i] = t / L[i][i];
    }
#pragma unroll
    for (int i = 0; i < N; i++) Ainv[i][c] = y[i];
  }
  return true;
}

// Gauss-Jordan inverse with partial pivoting (torch.inverse, lm_optimizer.py:484).
GCLM_HD inline bool invert(const float* A, int n, float* Ainv) {
  float a[MAX_NF][2 * MAX_NF];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { a[i][j] = A[i * n + j]; a[i][n + j] = i == j ? 1.f : 0.f; }
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabsf(a[r][c]) > fabsf(a[piv][c])) piv = r;
    if (a[piv][c] == 0.f || !isfinite(a[piv][c])) return false;
    if (piv != c) for (int j = 0; j < 2 * n; j++) { float t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
    const float inv = 1.f / a[c][c];
    for (int j = 0; j < 2 * n; j++) a[c][j] *= inv;
    for (int r = 0; r < n; r++) if (r != c) {
      const float f = a[r][c];
      if (f != 0.f) for (int j = 0; j < 2 * n; j++) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ainv[i * n + j] = a[i][n + j];
  return true;
}

// ------------------------------------------------------------------------------------------
// parameter updates: lm_optimizer.py:518-549
// ------------------------------------------------------------------------------------------
GCLM_HD inline void gravity_update(float g[3], float d0, float d1, bool spherical) {
  if (spherical) {  // SphericalManifold.plus, misc.py:233-259
    const float nx = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float nd = sqrtf(d0 * d0 + d1 * d1);
    const float nd_ = nd < 1e-7f ? nd + 1e-7f : nd;
    const float sinc = nd < 1e-7f ? 1.f : sinf(nd_) / nd_;
    const float ex[3] = {sinc * d0, sinc * d1, cosf(nd)};
    float hv[3], beta;
    householder(g, hv, beta);
    const float dot = hv[0] * ex[0] + hv[1] * ex[1] + hv[2] * ex[2];
    for (int i = 0; i < 3; i++) g[i] = nx * (ex[i] - hv[i] * (beta * dot));
  } else {  // EuclideanManifold on (roll,pitch) then from_rp (gravity.py:29-40,112-119)
    float roll, pitch;
    gravity_rp(g, roll, pitch);
    roll += d0; pitch += d1;
    float sr, cr, sp, cp;
    sincosf(roll, &sr, &cr);
    sincosf(pitch, &sp, &cp);
    g[0] = -sr * cp; g[1] = -cr * cp; g[2] = sp;
  }
  normalize3(g);
}

GCLM_HD inline void focal_update(float cam[8], float delta, bool as_log) {  // camera.py:136-152
  const float h = cam[1];
  float fy = as_log ? expf(logf(cam[3]) + delta) : cam[3] + delta;
  const float min_f = h / 2.f / 3.7320508075688776f;   // vfov 150 deg: tan(75 deg)
  const float max_f = h / 2.f / 0.04366094290851206f;  // vfov 5 deg: tan(2.5 deg)
  fy = fminf(fmaxf(fy, min_f), max_f);
  const float fx = fy * cam[2] / cam[3];
  cam[2] = fx; cam[3] = fy;
}

GCLM_HD inline void dist_update(int model, float cam[8], const float* d) {
  // camera.py:599-604,700-705,817-822.  `dist` is _data[6:] (2 wide) for one-parameter models, so
  // delta k1 is broadcast into the unused k2 slot as well -- kept for bit-level
  // equality of camera._data.
  const float lim = model == SIMPLE_DIVISIONAL ? 3.0f : 0.7f;
  const float d1 = model == RADIAL ? d[1] : d[0];
  cam[6] = fminf(fmaxf(cam[6] + d[0], -lim), lim);
  cam[7] = fminf(fmaxf(cam[7] + d1, -lim), lim);
}

}  // namespace gclm
