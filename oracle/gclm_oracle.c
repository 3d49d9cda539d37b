/*
 * gclm_oracle.c -- CPU restatement of GeoCalib's Levenberg-Marquardt calibration path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in geocalib_b200/ (the product) may import, link or call
 * this file.  It is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the *checker* of the CUDA engine, never as the thing shipped.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against golden
 * vectors produced by running the reference itself (geocalib/lm_optimizer.py @ 3ee7ec34) in the
 * build container with oracle/make_golden.py (committed next to the fixtures in tests/golden/).
 *
 * The file is compiled twice (REAL=float, SUFFIX=_f32 and REAL=double, SUFFIX=_f64).  The fp32
 * build mirrors the reference's arithmetic type (lm_optimizer.py:50,52 forces .float()); the fp64
 * build is the formula ground truth.
 *
 * The per-pixel math deliberately follows the reference's *operator structure* (explicit 2x2 /
 * 3x2 Jacobian chain, model member functions) rather than the algebraically reduced form the
 * CUDA kernel uses, so that the kernel's shortcuts are verified by an independent derivation.
 *
 * Reference map (all paths relative to /root/reference/geocalib/):
 *   lm_optimizer.py:20-58    get_trivial_estimation   -> (host side, see oracle.py)
 *   lm_optimizer.py:61-87    scaled_loss, huber_loss  -> huber()
 *   lm_optimizer.py:90-106   early_stop, update_lambda-> lm loop
 *   lm_optimizer.py:109-137  optimizer_step           -> damped_cholesky_solve()
 *   lm_optimizer.py:248-315  residuals + costs        -> pixel_eval()
 *   lm_optimizer.py:317-461  setup_system             -> accumulate_image()
 *   lm_optimizer.py:463-516  estimate_uncertainty     -> uncertainty()
 *   lm_optimizer.py:518-549  update_estimate          -> apply_update()
 *   lm_optimizer.py:551-644  optimize                 -> FN(gclm_oracle_lm)
 *   perspective_fields.py:47-275                       -> pixel_eval()
 *   camera.py:136-152, 308-344, 372-394, 522-942      -> model_eval(), update_focal()
 *   gravity.py:17-40, 62-119                          -> gravity_*()
 *   misc.py:181-318                                   -> householder(), sph_plus(), J_vecnorm
 *   utils.py:272-279                                  -> fov2focal (in update_focal)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <tgmath.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_max_threads(void) { return 1; }
#endif

#ifndef REAL
#define REAL float
#define SUFFIX _f32
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)
#define R(x) ((REAL)(x))

#define MAXK 2            /* distortion parameters (radial has 2) */
#define MAXF (3 + MAXK)   /* full column set: g0 g1 f k1 k2       */

enum { PINHOLE = 0, SIMPLE_RADIAL = 1, RADIAL = 2, SIMPLE_DIVISIONAL = 3 };

static int num_dist(int model) { return model == PINHOLE ? 0 : (model == RADIAL ? 2 : 1); }

/* masked_fill(denom == 0, 1e6): camera.py:836,851,857,867,877,880,900-937 */
static inline REAL mf(REAL d) { return d == R(0) ? R(1e6) : d; }

typedef struct {
  /* distort scale + "up projection offset" (camera.py:611-660, 712-786, 829-942) */
  REAL s;            /* distort(uv, return_scale)                    */
  REAL off[2];       /* J_distort(uv, "scale2pts") = offset          */
  REAL Jo[2][2];     /* J_up_projection_offset(uv, "uv")             */
  REAL s_k[MAXK];    /* J_distort(uv, "scale2dist")                  */
  REAL Jo_k[2][MAXK];/* J_up_projection_offset(uv, "dist")           */
  /* undistort */
  REAL und[2];       /* undistort(uv)                                */
  REAL Ju[2][2];     /* J_undistort(uv, "pts")                       */
  REAL Ju_k[2][MAXK];/* J_undistort(uv, "dist")                      */
} model_out;

static void model_eval(int model, REAL u, REAL v, REAL k1, REAL k2, model_out *o) {
  const REAL r2 = u * u + v * v;
  const REAL uv[2] = {u, v};
  memset(o, 0, sizeof(*o));
  if (model == PINHOLE) { /* camera.py:532-562 */
    o->s = R(1);
    o->und[0] = u; o->und[1] = v;
    o->Ju[0][0] = o->Ju[1][1] = R(1);
    return;
  }
  if (model == SIMPLE_RADIAL) { /* camera.py:611-660 */
    o->s = R(1) + k1 * r2;
    for (int i = 0; i < 2; i++) {
      o->off[i] = R(2) * k1 * uv[i];
      o->Jo[i][i] = R(2) * k1;
      o->Jo_k[i][0] = R(2) * uv[i];
    }
    o->s_k[0] = r2;
    const REAL b1 = -k1, radial = R(1) + b1 * r2;
    for (int i = 0; i < 2; i++) {
      o->und[i] = uv[i] * radial;
      o->Ju_k[i][0] = -r2 * uv[i];
      for (int j = 0; j < 2; j++) o->Ju[i][j] = R(2) * b1 * uv[i] * uv[j] + (i == j ? radial : R(0));
    }
    return;
  }
  if (model == RADIAL) { /* camera.py:712-786 */
    const REAL r4 = r2 * r2;
    o->s = R(1) + k1 * r2 + k2 * r4;
    const REAL phi = R(2) * k1 + R(4) * k2 * r2;
    for (int i = 0; i < 2; i++) {
      o->off[i] = phi * uv[i];
      for (int j = 0; j < 2; j++) o->Jo[i][j] = R(8) * k2 * uv[i] * uv[j] + (i == j ? phi : R(0));
      o->Jo_k[i][0] = R(2) * uv[i];
      o->Jo_k[i][1] = R(4) * r2 * uv[i];
    }
    o->s_k[0] = r2; o->s_k[1] = r4;
    const REAL b1 = -k1, b2 = R(3) * k1 * k1 - k2;
    const REAL radial = R(1) + b1 * r2 + b2 * r4;
    for (int i = 0; i < 2; i++) {
      o->und[i] = uv[i] * radial;
      o->Ju_k[i][0] = (R(6) * r4 * k1 - r2) * uv[i];
      o->Ju_k[i][1] = -r4 * uv[i];
      for (int j = 0; j < 2; j++)
        o->Ju[i][j] = (R(4) * r2 * b2 + R(2) * b1) * uv[i] * uv[j] + (i == j ? radial : R(0));
    }
    return;
  }
  /* SIMPLE_DIVISIONAL: camera.py:829-942 */
  {
    /* distort(): clamp(min=0) */
    REAL rad = R(1) - sqrt(fmax(R(1) - R(4) * k1 * r2, R(0)));
    REAL den = R(2) * k1 * r2;
    o->s = (den == R(0)) ? R(1) : rad / mf(den);
    /* J_distort: t0 = sqrt(clamp(1-4 k1 r2, min=1e-6)) */
    const REAL t0 = sqrt(fmax(R(1) - R(4) * k1 * r2, R(1e-6)));
    {
      const REAL d1 = t0 * R(2) * r2, d2 = k1 * r2 * r2, dd = d1 * d2;
      const REAL phi = (R(4) * d2 - (R(1) - t0) * d1) / mf(dd);
      o->off[0] = u * phi; o->off[1] = v * phi;
    }
    {
      const REAL d1 = R(2) * k1 * t0, d2 = R(2) * r2 * k1 * k1, dd = d1 * d2;
      o->s_k[0] = (R(2) * d2 - (R(1) - t0) * d1) / mf(dd);
    }
    /* J_up_projection_offset: t0c = clamp(1-4 k1 r2, 1e-6); t1 = sqrt(t0c) */
    const REAL t0c = fmax(R(1) - R(4) * k1 * r2, R(1e-6));
    const REAL t1 = sqrt(t0c);
    const REAL t0_32 = pow(t0c, R(1.5));
    {
      REAL J = R(16) / mf(R(4) * t0_32);
      J = J - R(2) / mf(r2 * t1 * k1);
      J = J + (R(1) - t1) / mf((r2 * k1) * (r2 * k1));
      o->Jo_k[0][0] = J * u; o->Jo_k[1][0] = J * v;
    }
    {
      REAL di = R(4) / mf(R(2) * r2 * t1);                 /* diag   */
      REAL pp = -R(16) / mf(R(4) * t1 * r2 * r2);          /* * ppT  */
      pp = pp + (R(32) * k1) / mf(R(4) * r2 * t0_32);
      pp = pp - R(4) / mf(r2 * r2 * t1);
      pp = pp + (R(4) * (R(1) - t1) / mf(k1 * r2 * r2 * r2));
      di = di - (R(1) - t1) / mf(k1 * r2 * r2);
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) o->Jo[i][j] = pp * uv[i] * uv[j] + (i == j ? di : R(0));
    }
    /* undistort */
    {
      const REAL dn = R(1) + k1 * r2;
      const REAL radial = R(1) / mf(dn);
      o->und[0] = u * radial; o->und[1] = v * radial;
      const REAL dn2 = dn * dn;
      const REAL jk = -r2 / mf(dn2);
      o->Ju_k[0][0] = jk * u; o->Ju_k[1][0] = jk * v;
      const REAL tt = mf(dn);
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
          o->Ju[i][j] = (i == j ? R(1) / tt : R(0)) - R(2) * k1 * uv[i] * uv[j] / (tt * tt);
    }
  }
}

/* ---------------- gravity / manifold helpers ---------------- */

/* misc.py:181-231 householder_vector + J_plus: T = (I - beta v v^T)[:, :2] */
static void householder(const REAL g[3], REAL hv[3], REAL *beta) {
  REAL sigma = g[0] * g[0] + g[1] * g[1];
  const REAL xpiv = g[2];
  const REAL norm = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  if (sigma < R(1e-7)) sigma = sigma + R(1e-7);
  const REAL vpiv = xpiv < R(0) ? xpiv - norm : -sigma / (xpiv + norm);
  *beta = R(2) * vpiv * vpiv / (sigma + vpiv * vpiv);
  hv[0] = g[0] / vpiv; hv[1] = g[1] / vpiv; hv[2] = R(1);
}

static void tangent_spherical(const REAL g[3], REAL T[3][2]) {
  REAL hv[3], beta;
  householder(g, hv, &beta);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 2; j++) T[i][j] = (i == j ? R(1) : R(0)) - beta * hv[i] * hv[j];
}

/* gravity.py:62-101 roll/pitch (with the 1e-4 eps) and J_rp */
static void gravity_rp(const REAL g[3], REAL *roll, REAL *pitch) {
  const REAL eps = R(1e-4);
  REAL r = asin(-g[0] / (sqrt(R(1) - g[2] * g[2]) + eps));
  const REAL sgn = g[0] > R(0) ? R(1) : (g[0] < R(0) ? R(-1) : R(0));
  const REAL offset = -R(3.14159265358979323846) * sgn;
  *roll = g[1] < R(0) ? r : -r + offset;
  *pitch = asin(g[2]);
}

static void tangent_rp(const REAL g[3], REAL T[3][2]) {
  REAL roll, pitch;
  gravity_rp(g, &roll, &pitch);
  const REAL cp = cos(pitch), sp = sin(pitch), cr = cos(roll), sr = sin(roll);
  T[0][0] = -cr * cp; T[1][0] = sr * cp; T[2][0] = R(0);
  T[0][1] = sr * sp;  T[1][1] = cr * sp; T[2][1] = cp;
}

static void normalize3(REAL g[3]) { /* F.normalize eps=1e-12 (gravity.py:26) */
  REAL n = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  n = fmax(n, R(1e-12));
  g[0] /= n; g[1] /= n; g[2] /= n;
}

/* misc.py:233-259 SphericalManifold.plus, then Gravity() re-normalisation */
static void sph_plus(REAL g[3], const REAL d[2]) {
  const REAL eps = R(1e-7);
  const REAL nx = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  const REAL nd = sqrt(d[0] * d[0] + d[1] * d[1]);
  const REAL nd_ = nd < eps ? nd + eps : nd;
  const REAL sinc = nd < eps ? R(1) : sin(nd_) / nd_;
  const REAL e[3] = {sinc * d[0], sinc * d[1], cos(nd)};
  REAL hv[3], beta;
  householder(g, hv, &beta);
  const REAL dot = hv[0] * e[0] + hv[1] * e[1] + hv[2] * e[2];
  for (int i = 0; i < 3; i++) g[i] = nx * (e[i] - hv[i] * (beta * dot));
  normalize3(g);
}

static void rp_plus(REAL g[3], const REAL d[2]) { /* gravity.py:112-119 euclidean branch */
  REAL roll, pitch;
  gravity_rp(g, &roll, &pitch);
  roll += d[0]; pitch += d[1];
  const REAL sr = sin(roll), cr = cos(roll), sp = sin(pitch), cp = cos(pitch);
  g[0] = -sr * cp; g[1] = -cr * cp; g[2] = sp;
  normalize3(g);
}

/* ---------------- robust loss ---------------- */

/* lm_optimizer.py:61-87: returns cost, writes weight (first derivative) */
static inline REAL huber(REAL sq, REAL a, REAL *w) {
  const REAL a2 = a * a;
  const REAL x = sq / a2;
  const REAL sx = sqrt(x + R(1e-8));
  const REAL isx = fmax(R(1.1920928955078125e-07), R(1) / sx);
  *w = x <= R(1) ? R(1) : isx;
  return (x <= R(1) ? x : R(2) * sx - R(1)) * a2;
}

/* ---------------- per-image accumulation ---------------- */

typedef struct {
  int model, H, W;
  const REAL *up;       /* (2,H,W) or NULL */
  const REAL *lat;      /* (1,H,W) or NULL */
  const REAL *up_conf;  /* (H,W) or NULL   */
  const REAL *lat_conf; /* (H,W) or NULL   */
  REAL up_scale, lat_scale;
} image_in;

/*
 * One pixel: residuals, costs/weights, and (if want_jac) the full-column Jacobians
 * J_up (2 x NF) and J_lat (1 x NF) with NF = 3 + K columns [g0 g1 f k..].
 */
static void pixel_eval(const image_in *im, int x, int y, const REAL cam[8], const REAL g[3],
                       const REAL T[3][2], const REAL fscale[2], int want_jac,
                       REAL r_up[2], REAL *r_lat, REAL J_up[2][MAXF], REAL J_lat[MAXF]) {
  const int K = num_dist(im->model);
  const REAL fx = cam[2], fy = cam[3], cx = cam[4], cy = cam[5];
  const REAL k1 = cam[6], k2 = cam[7];
  /* camera.py:308-311 normalize: no +0.5 on pixel centres (camera.py:335-338) */
  const REAL u = ((REAL)x - cx) / fx, v = ((REAL)y - cy) / fy;
  const REAL uv[2] = {u, v};
  const size_t n = (size_t)y * im->W + x, N = (size_t)im->H * im->W;
  model_out mo;
  model_eval(im->model, u, v, k1, k2, &mo);

  /* ---- up field: perspective_fields.py:47-81 ---- */
  const REAL p0[2] = {g[0] - g[2] * u, g[1] - g[2] * v};
  REAL M[2][2] = {{R(1), R(0)}, {R(0), R(1)}};
  if (K > 0)
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) M[i][j] = (i == j ? mo.s : R(0)) + mo.off[i] * uv[j];
  REAL p[2];
  for (int i = 0; i < 2; i++) p[i] = M[i][0] * p0[0] + M[i][1] * p0[1];
  const REAL rho = sqrt(p[0] * p[0] + p[1] * p[1]);
  const REAL rho_c = fmax(rho, R(1e-12));
  if (im->up) {
    r_up[0] = im->up[n] - p[0] / rho_c;
    r_up[1] = im->up[N + n] - p[1] / rho_c;
  } else r_up[0] = r_up[1] = R(0);

  /* ---- latitude field: perspective_fields.py:185-211 ---- */
  const REAL w3[3] = {mo.und[0], mo.und[1], R(1)};
  const REAL m = sqrt(w3[0] * w3[0] + w3[1] * w3[1] + R(1));
  const REAL mc = fmax(m, R(1e-12));
  const REAL ray[3] = {w3[0] / mc, w3[1] / mc, w3[2] / mc};
  const REAL ell = ray[0] * g[0] + ray[1] * g[1] + ray[2] * g[2];
  const REAL eps = R(1e-6);
  const REAL lat_pred = asin(fmin(fmax(ell, -R(1) + eps), R(1) - eps));
  if (im->lat) *r_lat = sin(im->lat[n]) - sin(lat_pred); /* lm_optimizer.py:262,270 */
  else *r_lat = R(0);

  if (!want_jac) return;

  /* ---- J_up_field: perspective_fields.py:84-182 ---- */
  REAL Jn[2][2]; /* misc.py:262-281 J_vecnorm of the un-normalised projection */
  {
    REAL nr = rho;
    if (nr == R(0)) nr = nr + R(1e-6); /* reference applies this batch-wide; here per pixel */
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++)
        Jn[i][j] = (i == j ? R(1) : R(0)) / nr - p[i] * p[j] / (nr * nr * nr);
  }
  /* gravity: J_proj2abc = [[1,0,-u],[0,1,-v]] (misc.py:308-314), premultiplied by M */
  {
    const REAL Jabc[2][3] = {{R(1), R(0), -u}, {R(0), R(1), -v}};
    REAL MJ[2][3];
    for (int i = 0; i < 2; i++)
      for (int c = 0; c < 3; c++) MJ[i][c] = M[i][0] * Jabc[0][c] + M[i][1] * Jabc[1][c];
    for (int j = 0; j < 2; j++) {
      REAL z[2];
      for (int i = 0; i < 2; i++) z[i] = MJ[i][0] * T[0][j] + MJ[i][1] * T[1][j] + MJ[i][2] * T[2][j];
      for (int i = 0; i < 2; i++) J_up[i][j] = Jn[i][0] * z[0] + Jn[i][1] * z[1];
    }
  }
  /* focal */
  const REAL d[2] = {-u * fscale[0], -v * fscale[1]}; /* J_normalize(.. "f") [* f] .sum(-1) */
  {
    REAL Jp[2][2]; /* J_proj2uv */
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) Jp[i][j] = (i == j ? -g[2] : R(0));
    if (K > 0) {
      const REAL inner = u * p0[0] + v * p0[1];
      REAL Jp2[2][2];
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
          const REAL J_proj2up = M[i][0] * Jp[0][j] + M[i][1] * Jp[1][j];
          const REAL J_proj2duv = p0[i] * mo.off[j];      /* einsum i,j->ji (offset, p0) */
          const REAL J_off1 = inner * mo.Jo[i][j];
          const REAL J_off2 = mo.off[i] * p0[j];
          Jp2[i][j] = (J_proj2duv + J_off1 + J_off2) + J_proj2up;
        }
      memcpy(Jp, Jp2, sizeof(Jp));
    }
    REAL z[2];
    for (int i = 0; i < 2; i++) z[i] = Jp[i][0] * d[0] + Jp[i][1] * d[1];
    for (int i = 0; i < 2; i++) J_up[i][2] = Jn[i][0] * z[0] + Jn[i][1] * z[1];
  }
  /* distortion */
  for (int k = 0; k < K; k++) {
    REAL z[2];
    for (int i = 0; i < 2; i++) {
      const REAL first = p0[i] * mo.s_k[k];
      /* (uv (x) p0) @ J_uvTdist */
      const REAL sec = uv[i] * (p0[0] * mo.Jo_k[0][k] + p0[1] * mo.Jo_k[1][k]);
      z[i] = first + sec;
    }
    for (int i = 0; i < 2; i++) J_up[i][3 + k] = Jn[i][0] * z[0] + Jn[i][1] * z[1];
  }

  /* ---- J_latitude_field: perspective_fields.py:214-275 ---- */
  REAL Jv[3][2]; /* J_vecnorm(uv1)[..., :2] */
  {
    REAL nr = m;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 2; j++)
        Jv[i][j] = (i == j ? R(1) : R(0)) / nr - w3[i] * w3[j] / (nr * nr * nr);
  }
  for (int j = 0; j < 2; j++) J_lat[j] = ray[0] * T[0][j] + ray[1] * T[1][j] + ray[2] * T[2][j];
  {
    REAL z[2];
    for (int i = 0; i < 2; i++) z[i] = mo.Ju[i][0] * d[0] + mo.Ju[i][1] * d[1];
    REAL acc = R(0);
    for (int i = 0; i < 3; i++) acc += (Jv[i][0] * z[0] + Jv[i][1] * z[1]) * g[i];
    J_lat[2] = acc;
  }
  for (int k = 0; k < K; k++) {
    REAL acc = R(0);
    for (int i = 0; i < 3; i++) acc += (Jv[i][0] * mo.Ju_k[0][k] + Jv[i][1] * mo.Ju_k[1][k]) * g[i];
    J_lat[3 + k] = acc;
  }
}

/*
 * Accumulate over one image (rows [y0,y1)) : G (NF), H (NF x NF), cost sums (up, lat).
 * lm_optimizer.py:276-315 (costs), :317-359 (einsum J^T W r, J^T W J).
 */
static void accumulate_rows(const image_in *im, int y0, int y1, const REAL cam[8], const REAL g[3],
                            const REAL T[3][2], const REAL fscale[2], int want_jac,
                            double G[MAXF], double Hs[MAXF][MAXF], double cost[2]) {
  const int NF = 3 + num_dist(im->model);
  /* row-level partials in REAL (mimics fp32 accumulation), folded into double per row */
  for (int y = y0; y < y1; y++) {
    REAL g_row[MAXF] = {0}, h_row[MAXF][MAXF] = {{0}}, c_row[2] = {0, 0};
    for (int x = 0; x < im->W; x++) {
      REAL r_up[2], r_lat, J_up[2][MAXF], J_lat[MAXF];
      pixel_eval(im, x, y, cam, g, T, fscale, want_jac, r_up, &r_lat, J_up, J_lat);
      const size_t n = (size_t)y * im->W + x;
      REAL w_up = R(0), w_lat = R(0);
      if (im->up) {
        REAL c = huber(r_up[0] * r_up[0] + r_up[1] * r_up[1], im->up_scale, &w_up);
        if (im->up_conf) { w_up *= im->up_conf[n]; c *= im->up_conf[n]; }
        c_row[0] += c;
      }
      if (im->lat) {
        REAL c = huber(r_lat * r_lat, im->lat_scale, &w_lat);
        if (im->lat_conf) { w_lat *= im->lat_conf[n]; c *= im->lat_conf[n]; }
        c_row[1] += c;
      }
      if (!want_jac) continue;
      for (int a = 0; a < NF; a++) {
        if (im->up) g_row[a] += w_up * (J_up[0][a] * r_up[0] + J_up[1][a] * r_up[1]);
        if (im->lat) g_row[a] += w_lat * (J_lat[a] * r_lat);
        for (int b = 0; b < NF; b++) {
          if (im->up) h_row[a][b] += w_up * (J_up[0][a] * J_up[0][b] + J_up[1][a] * J_up[1][b]);
          if (im->lat) h_row[a][b] += w_lat * (J_lat[a] * J_lat[b]);
        }
      }
    }
    cost[0] += c_row[0]; cost[1] += c_row[1];
    for (int a = 0; a < NF; a++) {
      G[a] += g_row[a];
      for (int b = 0; b < NF; b++) Hs[a][b] += h_row[a][b];
    }
  }
}

typedef struct {
  REAL G[MAXF];
  REAL H[MAXF][MAXF];
  REAL cost_up, cost_lat; /* means over pixels (lm_optimizer.py:584) */
} image_sys;

static void eval_image(const image_in *im, const REAL cam[8], const REAL g[3], int rpf_basis,
                       int spherical, int log_focal, int want_jac, image_sys *out) {
  REAL T[3][2];
  /* setup_system: spherical = use_spherical and not as_rpf; log_focal likewise (:411-416) */
  if (spherical && !rpf_basis) tangent_spherical(g, T); else tangent_rp(g, T);
  REAL fscale[2];
  if (log_focal && !rpf_basis) { fscale[0] = fscale[1] = R(1); }
  else { fscale[0] = R(1) / cam[2]; fscale[1] = R(1) / cam[3]; }
  double G[MAXF] = {0}, Hs[MAXF][MAXF] = {{0}}, cost[2] = {0, 0};
  const int chunk = 8;
  const int nchunks = (im->H + chunk - 1) / chunk;
  double *pg = (double *)calloc((size_t)nchunks * (MAXF + MAXF * MAXF + 2), sizeof(double));
#pragma omp parallel for schedule(static) if (nchunks > 1)
  for (int c = 0; c < nchunks; c++) {
    double *q = pg + (size_t)c * (MAXF + MAXF * MAXF + 2);
    double lG[MAXF] = {0}, lH[MAXF][MAXF] = {{0}}, lc[2] = {0, 0};
    const int y0 = c * chunk, y1 = (y0 + chunk < im->H) ? y0 + chunk : im->H;
    accumulate_rows(im, y0, y1, cam, g, T, fscale, want_jac, lG, lH, lc);
    memcpy(q, lG, sizeof(lG));
    memcpy(q + MAXF, lH, sizeof(lH));
    q[MAXF + MAXF * MAXF] = lc[0]; q[MAXF + MAXF * MAXF + 1] = lc[1];
  }
  for (int c = 0; c < nchunks; c++) { /* fixed order: thread-count independent */
    const double *q = pg + (size_t)c * (MAXF + MAXF * MAXF + 2);
    for (int a = 0; a < MAXF; a++) {
      G[a] += q[a];
      for (int b = 0; b < MAXF; b++) Hs[a][b] += q[MAXF + a * MAXF + b];
    }
    cost[0] += q[MAXF + MAXF * MAXF]; cost[1] += q[MAXF + MAXF * MAXF + 1];
  }
  free(pg);
  const double N = (double)im->H * im->W;
  for (int a = 0; a < MAXF; a++) {
    out->G[a] = (REAL)G[a];
    for (int b = 0; b < MAXF; b++) out->H[a][b] = (REAL)Hs[a][b];
  }
  out->cost_up = (REAL)(cost[0] / N);
  out->cost_lat = (REAL)(cost[1] / N);
}

/* Batch version: one flat OpenMP loop over (image, 8-row chunk) tasks so that any batch size uses
 * all host threads; per-image partials are then added in chunk order (thread-count independent). */
static void eval_batch(const image_in *ims, int B, const char *skip, const REAL *cam, const REAL *grav,
                       int rpf_basis, int spherical, int log_focal, int want_jac, image_sys *out) {
  const int chunk = 8, H = ims[0].H;
  const int nchunks = (H + chunk - 1) / chunk;
  const size_t stride = MAXF + MAXF * MAXF + 2;
  double *pg = (double *)calloc((size_t)B * nchunks * stride, sizeof(double));
  REAL (*Ts)[3][2] = malloc(sizeof(REAL[3][2]) * B);
  REAL (*fs)[2] = malloc(sizeof(REAL[2]) * B);
  for (int b = 0; b < B; b++) {
    const REAL *g = grav + 3 * b, *cm = cam + 8 * b;
    if (spherical && !rpf_basis) tangent_spherical(g, Ts[b]); else tangent_rp(g, Ts[b]);
    if (log_focal && !rpf_basis) { fs[b][0] = fs[b][1] = R(1); }
    else { fs[b][0] = R(1) / cm[2]; fs[b][1] = R(1) / cm[3]; }
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (long t = 0; t < (long)B * nchunks; t++) {
    const int b = (int)(t / nchunks), c = (int)(t % nchunks);
    if (skip && skip[b]) continue;
    double *q = pg + (size_t)t * stride;
    double lG[MAXF] = {0}, lH[MAXF][MAXF] = {{0}}, lc[2] = {0, 0};
    const int y0 = c * chunk, y1 = (y0 + chunk < H) ? y0 + chunk : H;
    accumulate_rows(&ims[b], y0, y1, cam + 8 * b, grav + 3 * b, Ts[b], fs[b], want_jac, lG, lH, lc);
    memcpy(q, lG, sizeof(lG));
    memcpy(q + MAXF, lH, sizeof(lH));
    q[MAXF + MAXF * MAXF] = lc[0]; q[MAXF + MAXF * MAXF + 1] = lc[1];
  }
  const double N = (double)H * ims[0].W;
  for (int b = 0; b < B; b++) {
    if (skip && skip[b]) continue;
    double G[MAXF] = {0}, Hs[MAXF][MAXF] = {{0}}, cost[2] = {0, 0};
    for (int c = 0; c < nchunks; c++) {
      const double *q = pg + ((size_t)b * nchunks + c) * stride;
      for (int a = 0; a < MAXF; a++) {
        G[a] += q[a];
        for (int k = 0; k < MAXF; k++) Hs[a][k] += q[MAXF + a * MAXF + k];
      }
      cost[0] += q[MAXF + MAXF * MAXF]; cost[1] += q[MAXF + MAXF * MAXF + 1];
    }
    for (int a = 0; a < MAXF; a++) {
      out[b].G[a] = (REAL)G[a];
      for (int k = 0; k < MAXF; k++) out[b].H[a][k] = (REAL)Hs[a][k];
    }
    out[b].cost_up = (REAL)(cost[0] / N);
    out[b].cost_lat = (REAL)(cost[1] / N);
  }
  free(pg); free(Ts); free(fs);
}

/* ---------------- linear algebra ---------------- */

/* In-place Cholesky (lower) of an n x n SPD matrix stored with leading dim ld. 0 = ok. */
static int cholesky(REAL *A, int n, int ld) {
  for (int j = 0; j < n; j++) {
    REAL d = A[j * ld + j];
    for (int k = 0; k < j; k++) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > R(0))) return 1; /* also catches NaN */
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; i++) {
      REAL s = A[i * ld + j];
      for (int k = 0; k < j; k++) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s / d;
    }
  }
  return 0;
}

static void cholesky_solve(const REAL *L, int n, int ld, REAL *b) {
  for (int i = 0; i < n; i++) {
    REAL s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * ld + k] * b[k];
    b[i] = s / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    REAL s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[k * ld + i] * b[k];
    b[i] = s / L[i * ld + i];
  }
}

/* LU inverse with partial pivoting (torch.inverse, lm_optimizer.py:484). Returns 0 on success. */
static int invert(const REAL *A, int n, REAL *Ainv) {
  REAL a[MAXF][2 * MAXF];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { a[i][j] = A[i * n + j]; a[i][n + j] = (i == j) ? R(1) : R(0); }
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    if (a[piv][c] == R(0)) return 1;
    if (piv != c) for (int j = 0; j < 2 * n; j++) { REAL t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
    const REAL inv = R(1) / a[c][c];
    for (int j = 0; j < 2 * n; j++) a[c][j] *= inv;
    for (int r = 0; r < n; r++) if (r != c) {
      const REAL f = a[r][c];
      if (f != R(0)) for (int j = 0; j < 2 * n; j++) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Ainv[i * n + j] = a[i][n + j];
  return 0;
}

/* ---------------- parameter updates ---------------- */

/* camera.py:136-152 update_focal (+ utils.py:272-274 fov2focal) */
static void update_focal(REAL cam[8], REAL delta, int as_log) {
  const REAL h = cam[1];
  REAL f0 = as_log ? exp(log(cam[2]) + delta) : cam[2] + delta;
  REAL f1 = as_log ? exp(log(cam[3]) + delta) : cam[3] + delta;
  const REAL pi = R(3.14159265358979323846);
  const REAL min_f = h / R(2) / tan((R(150) / R(180) * pi) / R(2));
  const REAL max_f = h / R(2) / tan((R(5) / R(180) * pi) / R(2));
  f0 = fmin(fmax(f0, min_f), max_f);
  f1 = fmin(fmax(f1, min_f), max_f);
  (void)f0;
  const REAL fx = f1 * cam[2] / cam[3];
  cam[2] = fx; cam[3] = f1;
}

/* camera.py:599-604, 700-705, 817-822: note dist is _data[6:] (2 wide) for 1-parameter models,
 * so delta k1 is broadcast into the unused k2 slot as well. */
static void update_dist(int model, REAL cam[8], const REAL *delta) {
  const REAL lim = model == SIMPLE_DIVISIONAL ? R(3.0) : R(0.7);
  if (model == RADIAL) {
    cam[6] = fmin(fmax(cam[6] + delta[0], -lim), lim);
    cam[7] = fmin(fmax(cam[7] + delta[1], -lim), lim);
  } else {
    cam[6] = fmin(fmax(cam[6] + delta[0], -lim), lim);
    cam[7] = fmin(fmax(cam[7] + delta[0], -lim), lim);
  }
}

/* ---------------- public entry points ---------------- */

typedef struct {
  int model, batch, height, width;
  int est_gravity, est_focal, apply_dist, shared_intrinsics;
  int num_steps, early_stop, fix_lambda, spherical, log_focal;
  int global_semantics; /* 1: batch-global early stop / Cholesky failure as the reference */
  int compute_uncertainty;
  double lambda0, atol, rtol, up_scale, lat_scale;
} oracle_conf;

/* column map: active[] lists which of the NF full columns enter the solve
 * (lm_optimizer.py:335-344: gravity if estimate_gravity, focal if estimate_focal,
 *  k-columns if camera_has_distortion) */
static int active_columns(const oracle_conf *c, int act[MAXF]) {
  int n = 0;
  if (c->est_gravity) { act[n++] = 0; act[n++] = 1; }
  if (c->est_focal) act[n++] = 2;
  for (int k = 0; k < num_dist(c->model); k++) act[n++] = 3 + k;
  return n;
}

/*
 * System for one image at given parameters: full-column G, H and mean costs.
 * basis: 0 = optimisation basis (spherical/log as configured), 1 = as_rpf.
 */
void FN(gclm_oracle_system)(int model, int H, int W, const REAL *up, const REAL *lat,
                            const REAL *up_conf, const REAL *lat_conf, double up_scale,
                            double lat_scale, const REAL *cam, const REAL *grav, int rpf_basis,
                            int spherical, int log_focal, REAL *G_out, REAL *H_out, REAL *cost_out) {
  image_in im = {model, H, W, up, lat, up_conf, lat_conf, (REAL)up_scale, (REAL)lat_scale};
  image_sys s;
  REAL g[3] = {grav[0], grav[1], grav[2]};
  normalize3(g); /* Gravity.__init__ always re-normalises (gravity.py:26) */
  eval_image(&im, cam, g, rpf_basis, spherical, log_focal, 1, &s);
  const int NF = 3 + num_dist(model);
  for (int a = 0; a < NF; a++) {
    G_out[a] = s.G[a];
    for (int b = 0; b < NF; b++) H_out[a * NF + b] = s.H[a][b];
  }
  cost_out[0] = s.cost_up; cost_out[1] = s.cost_lat;
}

static int allclose1(REAL a, REAL b, REAL atol, REAL rtol) { /* torch.allclose(a, b) */
  if (isnan(a) || isnan(b)) return 0;
  if (a == b) return 1;
  return fabs(a - b) <= atol + rtol * fabs(b);
}

/*
 * Full LM loop (lm_optimizer.py:551-644).  Buffers: up (B,2,H,W), lat (B,1,H,W), confs (B,H,W),
 * cam (B,8) and grav (B,3) are in/out (initial estimate -> optimised).
 * out_costs (B,6): initial total/up/lat, final total/up/lat.  out_stop_at (B) ints.
 * out_cov (B,n,n), out_unc (B,5: roll,pitch,gravity,focal,vfov) when compute_uncertainty.
 * trace (optional, may be NULL): num_steps x B x (8 + 3 + 1 + 1) = cam, grav, lambda, new_cost.
 * Returns number of LM iterations executed.
 */
int FN(gclm_oracle_lm)(const oracle_conf *c, const REAL *up, const REAL *lat, const REAL *up_conf,
                       const REAL *lat_conf, REAL *cam, REAL *grav, REAL *out_costs,
                       int *out_stop_at, REAL *out_cov, REAL *out_unc, int *out_status,
                       REAL *trace) {
  const int B = c->batch, N = c->height * c->width, K = num_dist(c->model);
  int act[MAXF];
  const int n = active_columns(c, act);
  const int n_intr = (c->est_focal ? 1 : 0) + K; /* lm_optimizer.py:237-239 */
  const int n_grav = c->est_gravity ? 2 : 0;
  image_in *ims = (image_in *)malloc(sizeof(image_in) * B);
  for (int b = 0; b < B; b++) {
    ims[b].model = c->model; ims[b].H = c->height; ims[b].W = c->width;
    ims[b].up = up ? up + (size_t)b * 2 * N : NULL;
    ims[b].lat = lat ? lat + (size_t)b * N : NULL;
    ims[b].up_conf = up_conf ? up_conf + (size_t)b * N : NULL;
    ims[b].lat_conf = lat_conf ? lat_conf + (size_t)b * N : NULL;
    ims[b].up_scale = (REAL)c->up_scale; ims[b].lat_scale = (REAL)c->lat_scale;
  }
  const int nl = c->shared_intrinsics ? 1 : B;
  REAL *lamb = (REAL *)malloc(sizeof(REAL) * nl);
  for (int i = 0; i < nl; i++) lamb[i] = (REAL)c->lambda0;
  REAL *prev_cost = (REAL *)calloc(B, sizeof(REAL));
  REAL *new_cost = (REAL *)calloc(B, sizeof(REAL));
  image_sys *sys = (image_sys *)malloc(sizeof(image_sys) * B);
  image_sys *sys2 = (image_sys *)malloc(sizeof(image_sys) * B);
  REAL *delta = (REAL *)calloc((size_t)B * MAXF, sizeof(REAL)); /* per-image active-order */
  char *frozen = (char *)calloc(B, 1);
  int stop_global = c->num_steps, iters = 0;
  for (int b = 0; b < B; b++) { out_stop_at[b] = c->num_steps; if (out_status) out_status[b] = 0; }

  for (int it = 0; it < c->num_steps; it++) {
    iters = it + 1;
    /* residuals + costs + system at the current estimate */
    eval_batch(ims, B, frozen, cam, grav, 0, c->spherical, c->log_focal, 1, sys);
    if (it == 0)
      for (int b = 0; b < B; b++) {
        prev_cost[b] = sys[b].cost_up + sys[b].cost_lat;
        out_costs[6 * b + 0] = prev_cost[b];
        out_costs[6 * b + 1] = sys[b].cost_up;
        out_costs[6 * b + 2] = sys[b].cost_lat;
      }
    /* ---- optimizer_step (lm_optimizer.py:109-137) ---- */
    memset(delta, 0, sizeof(REAL) * (size_t)B * MAXF);
    if (!c->shared_intrinsics) {
      int any_fail = 0;
      for (int b = 0; b < B; b++) {
        if (frozen[b]) continue;
        REAL A[MAXF * MAXF], rhs[MAXF];
        for (int i = 0; i < n; i++) {
          rhs[i] = sys[b].G[act[i]];
          for (int j = 0; j < n; j++) A[i * n + j] = sys[b].H[act[i]][act[j]];
        }
        for (int i = 0; i < n; i++) A[i * n + i] += fmax(A[i * n + i] * lamb[b], R(1e-6));
        if (cholesky(A, n, n)) { any_fail = 1; if (out_status) out_status[b] |= 1; continue; }
        cholesky_solve(A, n, n, rhs);
        for (int i = 0; i < n; i++) delta[b * MAXF + i] = rhs[i];
      }
      if (any_fail && c->global_semantics) memset(delta, 0, sizeof(REAL) * (size_t)B * MAXF);
    } else {
      /* dense arrow system (lm_optimizer.py:350-383): [gravity blocks (2B) | intrinsics (n_intr)] */
      const int Mdim = n_grav * B + n_intr;
      REAL *A = (REAL *)calloc((size_t)Mdim * Mdim, sizeof(REAL));
      REAL *rhs = (REAL *)calloc(Mdim, sizeof(REAL));
      for (int b = 0; b < B; b++) {
        for (int i = 0; i < n_grav; i++) {
          rhs[n_grav * b + i] = sys[b].G[act[i]];
          for (int j = 0; j < n_grav; j++)
            A[(size_t)(n_grav * b + i) * Mdim + n_grav * b + j] = sys[b].H[act[i]][act[j]];
          for (int j = 0; j < n_intr; j++) {
            const REAL e = sys[b].H[act[i]][act[n_grav + j]];
            A[(size_t)(n_grav * b + i) * Mdim + n_grav * B + j] = e;
            A[(size_t)(n_grav * B + j) * Mdim + n_grav * b + i] = e;
          }
        }
        for (int i = 0; i < n_intr; i++) {
          rhs[n_grav * B + i] += sys[b].G[act[n_grav + i]];
          for (int j = 0; j < n_intr; j++)
            A[(size_t)(n_grav * B + i) * Mdim + n_grav * B + j] += sys[b].H[act[n_grav + i]][act[n_grav + j]];
        }
      }
      for (int i = 0; i < Mdim; i++) A[(size_t)i * Mdim + i] += fmax(A[(size_t)i * Mdim + i] * lamb[0], R(1e-6));
      if (!cholesky(A, Mdim, Mdim)) {
        cholesky_solve(A, Mdim, Mdim, rhs);
        for (int b = 0; b < B; b++) {
          for (int i = 0; i < n_grav; i++) delta[b * MAXF + i] = rhs[n_grav * b + i];
          for (int i = 0; i < n_intr; i++) delta[b * MAXF + n_grav + i] = rhs[n_grav * B + i];
        }
      } else if (out_status) for (int b = 0; b < B; b++) out_status[b] |= 1;
      free(A); free(rhs);
    }
    /* ---- update_estimate (lm_optimizer.py:518-549) ---- */
    for (int b = 0; b < B; b++) {
      if (frozen[b]) continue;
      const REAL *d = delta + b * MAXF;
      int o = 0;
      REAL dg[2] = {0, 0};
      if (c->est_gravity) { dg[0] = d[0]; dg[1] = d[1]; o = 2; }
      if (c->spherical) sph_plus(grav + 3 * b, dg); else rp_plus(grav + 3 * b, dg);
      REAL df = R(0);
      if (c->est_focal) df = d[o++];
      update_focal(cam + 8 * b, df, c->log_focal);
      if (K > 0 && c->apply_dist) update_dist(c->model, cam + 8 * b, d + o);
    }
    /* ---- new cost (lm_optimizer.py:607-610) ---- */
    eval_batch(ims, B, frozen, cam, grav, 0, c->spherical, c->log_focal, 0, sys2);
    for (int b = 0; b < B; b++)
      if (!frozen[b]) new_cost[b] = sys2[b].cost_up + sys2[b].cost_lat;
    /* ---- lambda (lm_optimizer.py:95-106, 612-613) ---- */
    if (!c->fix_lambda && !c->shared_intrinsics)
      for (int b = 0; b < B; b++) {
        if (frozen[b]) continue;
        REAL nlm = lamb[b] * (new_cost[b] > prev_cost[b] ? R(10) : R(0.1));
        lamb[b] = fmin(fmax(nlm, R(1e-6)), R(1e2));
      }
    if (trace)
      for (int b = 0; b < B; b++) {
        REAL *t = trace + ((size_t)it * B + b) * 13;
        memcpy(t, cam + 8 * b, sizeof(REAL) * 8);
        memcpy(t + 8, grav + 3 * b, sizeof(REAL) * 3);
        t[11] = lamb[c->shared_intrinsics ? 0 : b];
        t[12] = new_cost[b];
      }
    /* ---- early stop (lm_optimizer.py:90-92, 619-625) ---- */
    int all_close = 1, stop_now = 0;
    for (int b = 0; b < B; b++) {
      if (frozen[b]) continue;
      const int cl = allclose1(new_cost[b], prev_cost[b], (REAL)c->atol, (REAL)c->rtol);
      if (!cl) all_close = 0;
      if (!c->global_semantics && cl) {
        if (it + 1 < out_stop_at[b]) out_stop_at[b] = it + 1;
        if (c->early_stop) frozen[b] = 1;
      }
    }
    if (c->global_semantics && all_close) {
      if (it + 1 < stop_global) stop_global = it + 1;
      if (c->early_stop) stop_now = 1;
    }
    for (int b = 0; b < B; b++) if (!frozen[b]) prev_cost[b] = new_cost[b]; /* :627 */
    if (stop_now) break;
    if (!c->global_semantics) {
      int all_frozen = 1;
      for (int b = 0; b < B; b++) if (!frozen[b]) all_frozen = 0;
      if (all_frozen) break;
    }
  }
  if (c->global_semantics) for (int b = 0; b < B; b++) out_stop_at[b] = stop_global;

  /* ---- final costs + uncertainty (lm_optimizer.py:632-642, 463-516) ---- */
  eval_batch(ims, B, NULL, cam, grav, 1, c->spherical, c->log_focal, c->compute_uncertainty, sys2);
  for (int b = 0; b < B; b++) {
    const image_sys s = sys2[b];
    out_costs[6 * b + 3] = s.cost_up + s.cost_lat;
    out_costs[6 * b + 4] = s.cost_up;
    out_costs[6 * b + 5] = s.cost_lat;
    if (!c->compute_uncertainty) continue;
    REAL A[MAXF * MAXF], Cv[MAXF * MAXF];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) A[i * n + j] = s.H[act[i]][act[j]];
    REAL *unc = out_unc + 5 * b;
    for (int i = 0; i < 5; i++) unc[i] = R(0);
    if (invert(A, n, Cv)) { for (int i = 0; i < n * n; i++) Cv[i] = R(NAN); }
    memcpy(out_cov + (size_t)b * n * n, Cv, sizeof(REAL) * n * n);
    if (c->est_gravity) {
      unc[0] = sqrt(Cv[0]);
      unc[1] = sqrt(Cv[n + 1]);
      /* largest eigenvalue of the symmetric 2x2 block (eigvalsh uses the lower triangle) */
      const REAL a = Cv[0], bb = Cv[n], d = Cv[n + 1];
      const REAL tr = (a + d) / R(2), df = (a - d) / R(2);
      unc[2] = sqrt(tr + sqrt(df * df + bb * bb));
    }
    if (c->est_focal) {
      const int fi = n_grav;
      const REAL cff = Cv[fi * n + fi];
      const REAL fy = cam[8 * b + 3], h = cam[8 * b + 1];
      const REAL jf = -R(4) * h / (R(4) * fy * fy + h * h); /* misc.py:284-287 */
      unc[3] = sqrt(cff) / R(2);
      unc[4] = sqrt(jf * jf * cff / R(2));
    }
  }
  free(ims); free(lamb); free(prev_cost); free(new_cost); free(sys); free(sys2); free(delta); free(frozen);
  return iters;
}

/* Render noise-free fields from parameters (get_perspective_field, perspective_fields.py:278-320).
 * up_out (2,H,W), lat_out (1,H,W) radians. Used by tests to build known-answer inputs. */
void FN(gclm_oracle_render)(int model, int H, int W, const REAL *cam, const REAL *grav, REAL *up_out,
                            REAL *lat_out) {
  const size_t N = (size_t)H * W;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const REAL u = ((REAL)x - cam[4]) / cam[2], v = ((REAL)y - cam[5]) / cam[3];
      model_out mo;
      model_eval(model, u, v, cam[6], cam[7], &mo);
      const REAL p0[2] = {grav[0] - grav[2] * u, grav[1] - grav[2] * v};
      REAL p[2] = {p0[0], p0[1]};
      if (model != PINHOLE) {
        const REAL in = u * p0[0] + v * p0[1];
        p[0] = mo.s * p0[0] + mo.off[0] * in;
        p[1] = mo.s * p0[1] + mo.off[1] * in;
      }
      const REAL rho = fmax(sqrt(p[0] * p[0] + p[1] * p[1]), R(1e-12));
      const size_t n = (size_t)y * W + x;
      up_out[n] = p[0] / rho; up_out[N + n] = p[1] / rho;
      const REAL m = fmax(sqrt(mo.und[0] * mo.und[0] + mo.und[1] * mo.und[1] + R(1)), R(1e-12));
      const REAL ell = (mo.und[0] * grav[0] + mo.und[1] * grav[1] + grav[2]) / m;
      lat_out[n] = asin(fmin(fmax(ell, -R(1) + R(1e-6)), R(1) - R(1e-6)));
    }
}
