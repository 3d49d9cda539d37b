// gclm_engine.cu -- fused residual + Jacobian + reduce + solve pass for sm_100a and the C ABI.
//
// One LM iteration == one launch of lm_pass_kernel over the whole batch:
//   grid  = B * tiles_per_image CTAs; a tile is `rows_per_tile` image rows of ONE image
//   warp  = one 128-pixel-wide column strip (32 lanes x float4) walked down the rows, so every
//           warp-level load is a fully coalesced 512 B segment of each of the 5 planes and the
//           per-column terms (u, ...) stay in registers for the whole strip
//   CTA   = warp-shuffle + shared-memory reduction of the NACC accumulators -> one partial per
//           tile in global scratch -> atomic ticket; the LAST CTA of an image sums the partials in
//           tile order (deterministic), then one thread runs the damped Cholesky solve, the
//           manifold update, the lambda schedule and the early-stop bookkeeping and writes the new
//           parameters for the next launch.  No host round trip anywhere (the reference crosses
//           the device boundary twice per iteration: lm_optimizer.py:128,137).
//   pass num_steps (the final one) re-evaluates costs at the final estimate and accumulates the
//   Hessian in the (roll,pitch,f,k) basis for the covariance (lm_optimizer.py:632-636, 463-516).
//
// Passes per solve: num_steps + 1 (the reference makes 2*num_steps + 2 logical passes).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/gclm.h"
#include "gclm_device.cuh"

namespace gclm {

// ---------------------------------------------------------------------------------------------
// device-side state
// ---------------------------------------------------------------------------------------------
constexpr int PSTRIDE = 12;  // cam[8] + g[3] + pad
constexpr unsigned FLAG_FROZEN = 0x100u;
constexpr int MAX_WARPS = 10;
#ifndef GCLM_RING_STAGES
#define GCLM_RING_STAGES 4
#endif
constexpr int RING_STAGES = GCLM_RING_STAGES;  // rows prefetched per lane by the cp.async ring
// shared-intrinsics Schur partials per image (floats):
//   [0,9) C  [9,18) E^T D~^-1 E  [18,21) g_c  [21,24) E^T D~^-1 g_g  [24] not-converged flag
//   [25,31) W = D~^-1 E (2 x ni)  [31,33) wg = D~^-1 g_g
constexpr int SCHUR_STRIDE = 48;
constexpr int SCHUR_NSUM = 25;

// Which pass a CTA is working on: a kernel parameter for the one-launch-per-pass schedule, a loop variable in
// the persistent small-batch kernel.
struct PassId { int pass, is_final; };
#ifdef GCLM_PHASE_TIMERS
#define GCLM_STAMP(k) do { if (threadIdx.x == 0 && a.dbg) a.dbg[((size_t)pid.pass * gridDim.x + blockIdx.x) * 8 + (k)] = clock64(); } while (0)
#define GCLM_SOLVER_STAMP(k) do { if (a.dbg) a.dbg[400000 + pid.pass * 8 + (k)] = clock64(); } while (0)
#else
#define GCLM_STAMP(k) do {} while (0)
#define GCLM_SOLVER_STAMP(k) do {} while (0)
#endif

struct Ctrl {
  int stop;            // batch-global early stop fired: later LM passes are no-ops
  int stop_at;         // first iteration (1-based) at which all images were converged
  unsigned long long vote[2];  // batch stop vote of the current pass (ping-pong by pass parity):
                               // low word = images that voted, high word = images NOT converged
  unsigned img_done;   // shared intrinsics: images whose Schur partials of this pass are written
  int comm_error;      // sticky: a peer's contribution never arrived (fused exchange timed out)
  unsigned bar_count;  // persistent kernel: arrivals at the current grid barrier (0 between barriers)
  unsigned bar_gen;    // ... and its generation (sense reversal: safe across solves and CUDA-graph replays)
};

// One-shot all-reduce over NVLink peer memory, fused into the pass kernel's epilogue: every rank stores its
// 25 partial sums straight into a slot of EVERY peer's exchange buffer (P2P stores through the
// NVSwitch), publishes a sequence number with system-scope release semantics, waits until all
// peers' sequence numbers have arrived in its own buffer and adds the contributions in rank order
// (bit-identical on every rank).  Two slots (seq & 1): a rank can never be two exchanges ahead.
constexpr int P2P_MAX_RANKS = 8;
struct P2PBuf {  // one per rank, mapped into every peer with CUDA IPC
  float data[2][P2P_MAX_RANKS][32];
  unsigned flag[2][P2P_MAX_RANKS];
};
struct P2PArgs {
  P2PBuf* peer[P2P_MAX_RANKS];  // peer[r] = rank r's buffer as seen from this device (peer[rank] = local)
  int world, rank;
  unsigned seq;  // > 0, identical on all ranks for this exchange; world == 0 disables the exchange
};

struct PassArgs {
  const float *up, *lat, *uconf, *lconf;
  int B, H, W;          // W is the row pitch the kernels walk (padded to a multiple of 4 when repacked)
  int Wt;               // true image width (cost normalisation)
  int rows_per_tile, tiles_per_image, rsub;
  int pass, is_final, shared;
  int conf16;           // uconf / lconf point at fp16 planes and `lat` already holds sin(latitude) (gclm_solve_packed)
  int model, num_steps, early_stop, fix_lambda, spherical, log_focal, stop_mode;
  int est_gravity, est_focal, est_dist, compute_unc;
  float lambda0, atol, rtol, up_scale, lat_scale;
  // scratch state
  float* P;            // (B,2,PSTRIDE) double-buffered parameters
  float* lam;          // (B)
  float* prev_cost;    // (B)
  float* cost0;        // (B,3) initial costs
  int* slot;           // (B) which P buffer is current
  unsigned* flags;     // (B) status bits | FLAG_FROZEN
  int* stop_at;        // (B) per-image mode
  float* partials;     // (B,S,NACC)
  unsigned* tickets;   // (B)
  Ctrl* ctrl;
  // shared-intrinsics scratch
  float* schur;        // (B,SCHUR_STRIDE)
  gclm_result res;
  // gclm_system (diagnostic): raw J^T W r / J^T W J / mean costs of the single pass, full columns
  // shared intrinsics: 1 = the last CTA of the pass reduces the Schur partials, exchanges them with the peers
  // (pp) and applies the update; 0 = separate kernels around an NCCL all-reduce (fallback)
  int fused_shared;
  float* shared_sums;  // (SCHUR_NSUM + 1) all-reduced sums of the current step
  long long* dbg;      // GCLM_PHASE_TIMERS build: clock64 stamps [pass][cta][8]
  // persistent schedule with the set-up folded in (no lm_init_kernel / sin_plane_kernel launches)
  int fused_init;
  const float *cam0, *g0;   // (B,8), (B,3) initial estimate (gclm_problem.camera_init / gravity_init)
  const float* lat_raw;     // (B,1,H,W) latitude in radians; a.lat is the sin() plane in scratch
  P2PArgs pp;
  int final_rpf;       // basis of the final pass: 1 = (roll,pitch) + linear focal (covariance), 0 = LM basis
  float *sys_g, *sys_h, *sys_cost;
};

__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void cp_async16(unsigned dst_smem, const float* src) {
#if defined(GCLM_CPASYNC_L2_128)
  asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
#elif defined(GCLM_CPASYNC_L2_NONE)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
#else
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8(unsigned dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers (GCLM_TMA build) ----
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Bounded wait: a mis-programmed transaction count must never hang the GPU (returns false instead).
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

#ifdef GCLM_TMA2D
#ifndef GCLM_TMA_ROWS
#define GCLM_TMA_ROWS 2
#endif
constexpr int TMA_ROWS = GCLM_TMA_ROWS;  // image rows per TMA box (box = 128 pixels x TMA_ROWS rows)
constexpr int TMA_STAGES = 2;            // boxes in flight per warp and plane
struct TmaMaps { CUtensorMap up, lat, uconf, lconf; };  // 2-D views (width W) of the planes
__device__ __forceinline__ void tma_box_g2s(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(bar) : "memory");
}
#endif

__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}

// four fp16 confidences (8 bytes of shared memory) -> float4
__device__ __forceinline__ float4 lds_half4(unsigned addr) {
  unsigned lo, hi;
  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ int active_columns(int model, int est_gravity, int est_focal, int* act) {
  int n = 0;  // lm_optimizer.py:335-344
  if (est_gravity) { act[n++] = 0; act[n++] = 1; }
  if (est_focal) act[n++] = 2;
  for (int k = 0; k < num_dist(model); k++) act[n++] = 3 + k;
  return n;
}

// What thread 0 of every CTA reads about its image in the prologue (one round trip to L2, all loads
// independent) and the solver of the image's last CTA re-uses from shared memory instead of reading again.
struct ImgState {
  float P[PSTRIDE];  // current parameters: cam[8], g[3]
  float lam, prev_cost;
  int slot;
  unsigned flags;
};

// ---------------------------------------------------------------------------------------------
// per-image epilogue, executed by ONE thread of the last CTA of the image
// ---------------------------------------------------------------------------------------------
template <int MODEL>
__device__ __noinline__ void finalize_image(const PassArgs& a, int b, const float* sums, const ImgState& st) {
  constexpr int K = Dims<MODEL>::K, NF = Dims<MODEL>::NF, NH = Dims<MODEL>::NH;
  const float invN = 1.f / ((float)a.H * (float)a.Wt);
  const float cu = sums[NH + NF] * (a.up_scale * a.up_scale) * invN;
  const float cl = sums[NH + NF + 1] * (a.lat_scale * a.lat_scale) * invN;
  const float* Pb = st.P;
  unsigned fl = __ldcg(&a.flags[b]) & 0xffu;
  int act[NF];
  const int n = active_columns(MODEL, a.est_gravity, a.est_focal, act);
  float Hf[NF][NF];
  {
    int idx = 0;
    for (int i = 0; i < NF; i++)
      for (int j = i; j < NF; j++) { Hf[i][j] = sums[idx]; Hf[j][i] = sums[idx]; idx++; }
  }
  float Ha[MAX_NF * MAX_NF];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Ha[i * n + j] = Hf[act[i]][act[j]];
  if (a.res.hessian_rpf)
    for (int i = 0; i < n * n; i++) a.res.hessian_rpf[(size_t)b * n * n + i] = Ha[i];
  if (a.compute_unc && (a.res.covariance || a.res.uncertainty)) {
    float Cv[MAX_NF * MAX_NF];
    float unc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    bool inverted = false;
    if (n == NF) {  // every block free: symmetric positive definite in all but degenerate cases -> registers only
      float Hs[NF][NF], Ci[NF][NF];
#pragma unroll
      for (int i = 0; i < NF; i++)
#pragma unroll
        for (int j = 0; j < NF; j++) Hs[i][j] = Hf[i][j];
      inverted = spd_inverse_fixed<NF>(Hs, Ci);
      if (inverted) {
#pragma unroll
        for (int i = 0; i < NF; i++)
#pragma unroll
          for (int j = 0; j < NF; j++) Cv[i * NF + j] = Ci[i][j];
      }
    }
    if (!inverted && !invert(Ha, n, Cv)) {
      fl |= GCLM_STATUS_COV_SINGULAR;
      for (int i = 0; i < n * n; i++) Cv[i] = __int_as_float(0x7fc00000);
    }
    if (a.est_gravity) {  // lm_optimizer.py:489-499
      unc[0] = sqrtf(Cv[0]);
      unc[1] = sqrtf(Cv[n + 1]);
      const float p = Cv[0], q = Cv[n], r = Cv[n + 1];  // eigvalsh reads the lower triangle
      const float tr = 0.5f * (p + r), df = 0.5f * (p - r);
      unc[2] = sqrtf(tr + sqrtf(df * df + q * q));
    }
    if (a.est_focal) {  // :501-515, misc.py:284-287
      const int fi = a.est_gravity ? 2 : 0;
      const float cff = Cv[fi * n + fi];
      const float fy = Pb[3], h = Pb[1];
      const float jf = -4.f * h / (4.f * fy * fy + h * h);
      unc[3] = sqrtf(cff) / 2.f;
      unc[4] = sqrtf(jf * jf * cff / 2.f);
    }
    if (a.res.covariance)
      for (int i = 0; i < n * n; i++) a.res.covariance[(size_t)b * n * n + i] = Cv[i];
    if (a.res.uncertainty)
      for (int i = 0; i < 5; i++) a.res.uncertainty[(size_t)b * 5 + i] = unc[i];
  }
  if (a.res.costs) {
    float* c = a.res.costs + (size_t)b * 6;
    if (a.num_steps == 0) { c[0] = cu + cl; c[1] = cu; c[2] = cl; }
    else { c[0] = a.cost0[b * 3]; c[1] = a.cost0[b * 3 + 1]; c[2] = a.cost0[b * 3 + 2]; }
    c[3] = cu + cl; c[4] = cu; c[5] = cl;
  }
  if (a.res.camera) for (int i = 0; i < 8; i++) a.res.camera[(size_t)b * 8 + i] = Pb[i];
  if (a.res.gravity) for (int i = 0; i < 3; i++) a.res.gravity[(size_t)b * 3 + i] = Pb[8 + i];
  if (a.res.stop_at)
    a.res.stop_at[b] = (a.stop_mode == GCLM_STOP_PER_IMAGE && !a.shared) ? a.stop_at[b] : a.ctrl->stop_at;
  if (!isfinite(cu + cl)) fl |= GCLM_STATUS_NONFINITE;  // NaN/Inf in the fields or a diverged estimate
  if (a.res.status) a.res.status[b] = fl;
  (void)K;
}

__device__ __forceinline__ bool allclose1(float nw, float prev, float atol, float rtol) {
  if (nw == prev) return true;  // torch.allclose(new, prev): |new - prev| <= atol + rtol |prev|
  return fabsf(nw - prev) <= atol + rtol * fabsf(prev);  // NaN -> false
}

// Batch-global early-stop bookkeeping (lm_optimizer.py:90-92, 619-625): the last image to finish
// this pass knows whether ALL images were converged.
__device__ __forceinline__ void batch_stop_vote(const PassArgs& a, const PassId pid, bool converged) {
  Ctrl* c = a.ctrl;
  const int par = pid.pass & 1;
  // ONE atomic carries both counts; its return value tells the last voter everything (no fences, no second
  // round trip: what the vote decides is read after the next kernel boundary / grid barrier only)
  const unsigned long long mine = 1ull + (converged ? 0ull : (1ull << 32));
  const unsigned long long before = atomicAdd(&c->vote[par], mine);
  if ((unsigned)(before & 0xffffffffull) == (unsigned)a.B - 1u) {
    const unsigned long long un = (before + mine) >> 32;
    if (un == 0ull) {
      if (pid.pass < c->stop_at) c->stop_at = pid.pass;  // infos["stop_at"] = min(i + 1, ...)
      if (a.early_stop) c->stop = 1;
    }
    c->vote[par ^ 1] = 0ull;  // arm the counter of the next pass
  }
}

template <int MODEL>
__device__ __noinline__ void lm_step_image(const PassArgs& a, const PassId pid, int b, const float* sums, const ImgState& st) {
  constexpr int K = Dims<MODEL>::K, NF = Dims<MODEL>::NF, NH = Dims<MODEL>::NH;
  const int sl = st.slot;
  if (pid.is_final) {
    if (a.sys_h) {  // gclm_system: hand the reduced system out as it is (full columns [g0 g1 f k..])
      const float invN = 1.f / ((float)a.H * (float)a.Wt);
      int idx = 0;
      for (int i = 0; i < NF; i++)
        for (int j = i; j < NF; j++) {
          a.sys_h[((size_t)b * NF + i) * NF + j] = sums[idx];
          a.sys_h[((size_t)b * NF + j) * NF + i] = sums[idx];
          idx++;
        }
      if (a.sys_g) for (int i = 0; i < NF; i++) a.sys_g[(size_t)b * NF + i] = sums[NH + i];
      if (a.sys_cost) {
        a.sys_cost[b * 2] = sums[NH + NF] * (a.up_scale * a.up_scale) * invN;
        a.sys_cost[b * 2 + 1] = sums[NH + NF + 1] * (a.lat_scale * a.lat_scale) * invN;
      }
      return;
    }
    finalize_image<MODEL>(a, b, sums, st);  // st.P: the final estimate (the prologue applied the stop rule)
    return;
  }
  GCLM_SOLVER_STAMP(0);
  const float invN = 1.f / ((float)a.H * (float)a.Wt);
  const float cu = sums[NH + NF] * (a.up_scale * a.up_scale) * invN;
  const float cl = sums[NH + NF + 1] * (a.lat_scale * a.lat_scale) * invN;
  const float cost = cu + cl;  // lm_optimizer.py:584,610
  float lam = st.lam;
  bool converged = false;
  if (pid.pass == 0) {
    if (a.fused_init) { a.flags[b] = 0u; a.stop_at[b] = a.num_steps; }  // what lm_init_kernel would have set
    a.cost0[b * 3] = cost; a.cost0[b * 3 + 1] = cu; a.cost0[b * 3 + 2] = cl;
  } else {
    // this pass evaluated cost(theta_i) == "new_cost" of iteration i-1 (lm_optimizer.py:607-613)
    const float prev = st.prev_cost;
    if (!a.fix_lambda) lam = fminf(fmaxf(lam * (cost > prev ? 10.f : 0.1f), 1e-6f), 1e2f);
    converged = allclose1(cost, prev, a.atol, a.rtol);
    if (a.stop_mode == GCLM_STOP_PER_IMAGE && converged) {
      if (pid.pass < a.stop_at[b]) a.stop_at[b] = pid.pass;
      if (a.early_stop) {  // freeze: theta_i is final, no further work on this image
        a.flags[b] |= FLAG_FROZEN;
        a.prev_cost[b] = cost;
        return;
      }
    }
  }
  a.prev_cost[b] = cost;
  a.lam[b] = lam;

  // ---- optimizer_step (lm_optimizer.py:109-137) on the active columns ----
  int act[NF];
  const int n = active_columns(MODEL, a.est_gravity, a.est_focal, act);
  // the step, by block: gravity (d0, d1), focal (dfo), distortion (dk)
  float d0 = 0.f, d1 = 0.f, dfo = 0.f, dk[2] = {0.f, 0.f};
  bool solved;
  if (n == NF) {  // every block is free (the common case): compile-time size and indices, all in registers
    float Af[NF][NF], df[NF];
    {
      int idx = 0;
#pragma unroll
      for (int i = 0; i < NF; i++) {
        df[i] = sums[NH + i];
#pragma unroll
        for (int j = i; j < NF; j++) { Af[i][j] = sums[idx]; Af[j][i] = sums[idx]; idx++; }
      }
    }
#pragma unroll
    for (int i = 0; i < NF; i++) Af[i][i] += fmaxf(Af[i][i] * lam, 1e-6f);
    solved = cholesky_solve_fixed<NF>(Af, df);
    if (solved) {
      d0 = df[0]; d1 = df[1]; dfo = df[2];
      if constexpr (K > 0) dk[0] = df[3];
      if constexpr (K > 1) dk[1] = df[4];
    }
  } else {
    float Hf[NF][NF];
    {
      int idx = 0;
      for (int i = 0; i < NF; i++)
        for (int j = i; j < NF; j++) { Hf[i][j] = sums[idx]; Hf[j][i] = sums[idx]; idx++; }
    }
    float A[MAX_NF * MAX_NF], d[MAX_NF];
    for (int i = 0; i < n; i++) {
      d[i] = sums[NH + act[i]];
      for (int j = 0; j < n; j++) A[i * n + j] = Hf[act[i]][act[j]];
    }
    for (int i = 0; i < n; i++) A[i * n + i] += fmaxf(A[i * n + i] * lam, 1e-6f);
    solved = cholesky(A, n);
    if (solved) {
      cholesky_solve(A, n, d);
      int o = 0;
      if (a.est_gravity) { d0 = d[0]; d1 = d[1]; o = 2; }
      if (a.est_focal) dfo = d[o++];
      for (int k = 0; k < K; k++) dk[k] = d[o + k];
    }
  }
  GCLM_SOLVER_STAMP(1);
  if (!solved) a.flags[b] |= GCLM_STATUS_CHOLESKY_FAILED;  // zero step for this image (lm_optimizer.py:129-133)
  // ---- update_estimate (lm_optimizer.py:518-549) into the other parameter buffer ----
  const float* Pc = st.P;
  float* Pn = a.P + ((size_t)b * 2 + (sl ^ 1)) * PSTRIDE;
  float cam[8], g[3];
#pragma unroll
  for (int i = 0; i < 8; i++) cam[i] = Pc[i];
#pragma unroll
  for (int i = 0; i < 3; i++) g[i] = Pc[8 + i];
  gravity_update(g, d0, d1, a.spherical);
  GCLM_SOLVER_STAMP(2);
  focal_update(cam, dfo, a.log_focal);
  if (K > 0 && a.est_dist) dist_update(MODEL, cam, dk);
  GCLM_SOLVER_STAMP(3);
  for (int i = 0; i < 8; i++) Pn[i] = cam[i];
  for (int i = 0; i < 3; i++) Pn[8 + i] = g[i];
  a.slot[b] = sl ^ 1;
  GCLM_SOLVER_STAMP(4);
  if (a.stop_mode == GCLM_STOP_BATCH) {
    if (pid.pass >= 1) batch_stop_vote(a, pid, converged);
  }
  GCLM_SOLVER_STAMP(5);
}

// ---------------------------------------------------------------------------------------------
// shared intrinsics (lm_optimizer.py:350-383, 599-603).  The reference assembles the dense
// (2B + ni)^2 arrow matrix [[blockdiag(D_i), E],[E^T, sum C_i]] and Choleskys it on the CPU.
// Eliminating the per-image 2x2 gravity blocks gives the ni x ni Schur system
//     (C~ - sum_i E_i^T D~_i^-1 E_i) dc = sum_i g_c,i - sum_i E_i^T D~_i^-1 g_g,i
//     dg_i = D~_i^-1 (g_g,i - E_i dc)
// with the same damping (diag += max(lambda diag, 1e-6), lambda fixed, :572-573,612).  Only the
// <= 25 sums cross images (and GPUs): one tiny all-reduce per LM step.
// ---------------------------------------------------------------------------------------------
template <int MODEL>
__device__ __noinline__ void schur_image(const PassArgs& a, const PassId pid, int b, const float* sums, const ImgState& st) {
  constexpr int NF = Dims<MODEL>::NF, NH = Dims<MODEL>::NH;
  const float invN = 1.f / ((float)a.H * (float)a.Wt);
  const float cu = sums[NH + NF] * (a.up_scale * a.up_scale) * invN;
  const float cl = sums[NH + NF + 1] * (a.lat_scale * a.lat_scale) * invN;
  const float cost = cu + cl;
  bool converged = false;
  if (pid.pass == 0) { a.cost0[b * 3] = cost; a.cost0[b * 3 + 1] = cu; a.cost0[b * 3 + 2] = cl; }
  else converged = allclose1(cost, st.prev_cost, a.atol, a.rtol);
  a.prev_cost[b] = cost;
  int act[NF];
  const int n = active_columns(MODEL, a.est_gravity, a.est_focal, act);
  const int ng = a.est_gravity ? 2 : 0, ni = n - ng;
  float Hf[NF][NF];
  {
    int idx = 0;
    for (int i = 0; i < NF; i++)
      for (int j = i; j < NF; j++) { Hf[i][j] = sums[idx]; Hf[j][i] = sums[idx]; idx++; }
  }
  float* out = a.schur + (size_t)b * SCHUR_STRIDE;
  for (int i = 0; i < SCHUR_STRIDE; i++) out[i] = 0.f;
  const float lam = a.lambda0;
  float Di[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // D~^-1
  bool ok = true;
  if (ng == 2) {
    float d00 = Hf[0][0], d01 = Hf[0][1], d11 = Hf[1][1];
    d00 += fmaxf(d00 * lam, 1e-6f);
    d11 += fmaxf(d11 * lam, 1e-6f);
    const float det = d00 * d11 - d01 * d01;
    ok = (d00 > 0.f) && (det > 0.f);
    if (ok) { Di[0][0] = d11 / det; Di[1][1] = d00 / det; Di[0][1] = Di[1][0] = -d01 / det; }
    else a.flags[b] |= GCLM_STATUS_CHOLESKY_FAILED;
  }
  float W[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, wg[2] = {0.f, 0.f};
  if (ng == 2 && ok) {
    for (int j = 0; j < ni; j++) {
      const float e0 = Hf[0][act[ng + j]], e1 = Hf[1][act[ng + j]];
      W[0][j] = Di[0][0] * e0 + Di[0][1] * e1;
      W[1][j] = Di[1][0] * e0 + Di[1][1] * e1;
    }
    const float g0 = sums[NH + 0], g1 = sums[NH + 1];
    wg[0] = Di[0][0] * g0 + Di[0][1] * g1;
    wg[1] = Di[1][0] * g0 + Di[1][1] * g1;
  }
  for (int i = 0; i < ni; i++) {
    for (int j = 0; j < ni; j++) {
      out[i * 3 + j] = Hf[act[ng + i]][act[ng + j]];
      float ewe = 0.f;
      if (ng == 2 && ok) ewe = Hf[0][act[ng + i]] * W[0][j] + Hf[1][act[ng + i]] * W[1][j];
      out[9 + i * 3 + j] = ewe;
    }
    out[18 + i] = sums[NH + act[ng + i]];
    out[21 + i] = (ng == 2 && ok) ? Hf[0][act[ng + i]] * wg[0] + Hf[1][act[ng + i]] * wg[1] : 0.f;
  }
  out[24] = (pid.pass >= 1 && converged) ? 0.f : 1.f;
  for (int j = 0; j < 3; j++) { out[25 + j] = W[0][j]; out[28 + j] = W[1][j]; }
  out[31] = wg[0]; out[32] = wg[1];
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sum of this rank's per-image Schur partials, in an order that depends on the block size only: thread t
// adds images t, t + nthreads, ... (independent 128-bit loads), a warp shuffle tree, then the warps in
// order.  Block-wide call; the result is valid in threads j < SCHUR_NSUM.
__device__ __forceinline__ float schur_block_sum(const PassArgs& a, float (*s_part)[SCHUR_NSUM + 3]) {
  float acc[SCHUR_NSUM + 3];
#pragma unroll
  for (int j = 0; j < SCHUR_NSUM + 3; j++) acc[j] = 0.f;
  for (int b = threadIdx.x; b < a.B; b += blockDim.x) {
    const float4* p = reinterpret_cast<const float4*>(a.schur + (size_t)b * SCHUR_STRIDE);
#pragma unroll
    for (int q = 0; q < (SCHUR_NSUM + 3) / 4; q++) {
      const float4 x = __ldcg(p + q);
      acc[4 * q] += x.x; acc[4 * q + 1] += x.y; acc[4 * q + 2] += x.z; acc[4 * q + 3] += x.w;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < SCHUR_NSUM; j++) {
    float x = acc[j];
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    if (lane == 0) s_part[warp][j] = x;
  }
  __syncthreads();
  float mine = 0.f;
  if (threadIdx.x < SCHUR_NSUM)
    for (int w = 0; w < nwarps; w++) mine += s_part[w][threadIdx.x];
  return mine;
}

// Exchange of the SCHUR_NSUM sums with the peers (see P2PBuf); executed by warp 0, `mine` valid in lanes
// j < SCHUR_NSUM.  Returns the total over ranks in those lanes.  A peer that never shows up (bounded wait,
// ~1 minute) latches ctrl->comm_error and ctrl->stop on THIS rank: its solve ends flagged, it stops
// contributing, and the peers run into the same time-out one step later -- no rank keeps iterating alone.
__device__ __forceinline__ float p2p_exchange(const PassArgs& a, float mine) {
  const P2PArgs& pp = a.pp;
  const int slot = pp.seq & 1u;
  if (threadIdx.x < SCHUR_NSUM)
    for (int r = 0; r < pp.world; r++) pp.peer[r]->data[slot][pp.rank][threadIdx.x] = mine;  // P2P stores
  __threadfence_system();
  __syncwarp();
  if (threadIdx.x < pp.world) st_release_sys(&pp.peer[threadIdx.x]->flag[slot][pp.rank], pp.seq);
  bool ok = true;
  if (threadIdx.x < pp.world) {  // wait for rank `threadIdx.x`'s contribution in MY buffer (bounded)
    const unsigned* f = &pp.peer[pp.rank]->flag[slot][threadIdx.x];
    unsigned spins = 0;
    while (ld_acquire_sys(f) != pp.seq) {
      if (++spins > (1u << 28)) { ok = false; break; }
      __nanosleep(spins < 4096u ? 32 : 256);
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  float tot = 0.f;
  if (threadIdx.x < SCHUR_NSUM) {
    const P2PBuf* me = pp.peer[pp.rank];
    for (int r = 0; r < pp.world; r++) tot += *(volatile const float*)&me->data[slot][r][threadIdx.x];
    if (!ok) tot = __int_as_float(0x7fc00000);
  }
  if (!ok && threadIdx.x == 0) { a.ctrl->comm_error = 1; a.ctrl->stop = 1; }
  return tot;
}

// One image: solve the (tiny, identical on every image and rank) Schur system and apply the update.
template <int MODEL>
__device__ __forceinline__ void shared_update_image(const PassArgs& a, int b, const float* sum_in) {
  constexpr int K = Dims<MODEL>::K;
  const int ng = a.est_gravity ? 2 : 0;
  const int ni = (a.est_focal ? 1 : 0) + K;
  float S[9], rhs[3];
  for (int i = 0; i < ni; i++) {
    for (int j = 0; j < ni; j++) S[i * ni + j] = sum_in[i * 3 + j];
    rhs[i] = sum_in[18 + i] - sum_in[21 + i];
  }
  for (int i = 0; i < ni; i++) S[i * ni + i] += fmaxf(S[i * ni + i] * a.lambda0, 1e-6f);
  for (int i = 0; i < ni; i++)
    for (int j = 0; j < ni; j++) S[i * ni + j] -= sum_in[9 + i * 3 + j];
  bool ok = true;
  if (ni > 0) {
    ok = cholesky(S, ni);
    if (ok) cholesky_solve(S, ni, rhs);
  }
  if (!ok) for (int i = 0; i < 3; i++) rhs[i] = 0.f;  // zero step for everybody (:129-133)
  const float* sc = a.schur + (size_t)b * SCHUR_STRIDE;  // written by another CTA of THIS launch: read through L2
  float d0 = 0.f, d1 = 0.f;
  if (ng == 2 && ok) {
    d0 = __ldcg(sc + 31); d1 = __ldcg(sc + 32);
    for (int j = 0; j < ni; j++) { d0 -= __ldcg(sc + 25 + j) * rhs[j]; d1 -= __ldcg(sc + 28 + j) * rhs[j]; }
  }
  if (!ok) a.flags[b] |= GCLM_STATUS_CHOLESKY_FAILED;
  const int sl = __ldcg(&a.slot[b]);
  const float* Pc = a.P + ((size_t)b * 2 + sl) * PSTRIDE;
  float* Pn = a.P + ((size_t)b * 2 + (sl ^ 1)) * PSTRIDE;
  float cam[8], g[3];
  for (int i = 0; i < 8; i++) cam[i] = __ldcg(Pc + i);
  for (int i = 0; i < 3; i++) g[i] = __ldcg(Pc + 8 + i);
  gravity_update(g, d0, d1, a.spherical);
  int o = 0;
  float df = 0.f;
  if (a.est_focal) df = rhs[o++];
  focal_update(cam, df, a.log_focal);
  if (K > 0 && a.est_dist) dist_update(MODEL, cam, rhs + o);
  for (int i = 0; i < 8; i++) Pn[i] = cam[i];
  for (int i = 0; i < 3; i++) Pn[8 + i] = g[i];
  a.slot[b] = sl ^ 1;
}

// Batch-global stop of the shared solve (all images of all ranks converged): one thread, after the update.
__device__ __forceinline__ void shared_stop_latch(const PassArgs& a, const PassId pid, const float* sum_in) {
  if (pid.pass >= 1 && sum_in[24] == 0.f) {
    if (pid.pass < a.ctrl->stop_at) a.ctrl->stop_at = pid.pass;
    if (a.early_stop) a.ctrl->stop = 1;
  }
}

// The tail of one shared-intrinsics LM step, run by the LAST CTA of the pass (block-wide call): ordered sum
// of the per-image partials -> exchange with the peers over NVLink -> every image's update -> stop latch.
// Replaces three kernel launches per step; the grid has drained by now, so this is the pass's critical tail.
template <int MODEL>
__device__ __noinline__ void shared_tail(const PassArgs& a, const PassId pid, float (*s_part)[SCHUR_NSUM + 3], float* s_tot) {
  __threadfence();  // the other CTAs' Schur partials (released before their image ticket)
  float mine = schur_block_sum(a, s_part);
  if (threadIdx.x < 32) {
    if (a.pp.world > 1) mine = p2p_exchange(a, mine);
    if (threadIdx.x < SCHUR_NSUM) { s_tot[threadIdx.x] = mine; a.shared_sums[threadIdx.x] = mine; }
  }
  __syncthreads();
  if (*(volatile int*)&a.ctrl->comm_error) {
    for (int b = threadIdx.x; b < a.B; b += blockDim.x) a.flags[b] |= GCLM_STATUS_COMM_TIMEOUT;
    return;
  }
  for (int b = threadIdx.x; b < a.B; b += blockDim.x) shared_update_image<MODEL>(a, b, s_tot);
  if (threadIdx.x == 0) shared_stop_latch(a, pid, s_tot);
}

// ---- NCCL fallback (no peer memory): the same three steps as separate kernels around ncclAllReduce ----
__global__ void shared_reduce_kernel(const __grid_constant__ PassArgs a) {  // launched with the pass kernel's block size
  __shared__ float s_part[MAX_WARPS][SCHUR_NSUM + 3];
  if (*(volatile int*)&a.ctrl->stop) return;
  const float mine = schur_block_sum(a, s_part);
  if (threadIdx.x < SCHUR_NSUM) a.shared_sums[threadIdx.x] = mine;
}
template <int MODEL>
__global__ void shared_update_kernel(const __grid_constant__ PassArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (*(volatile int*)&a.ctrl->stop) return;
  if (b < a.B) shared_update_image<MODEL>(a, b, a.shared_sums);
}
// Runs after shared_update_kernel of the same pass: latch the batch-global stop.
__global__ void shared_stop_kernel(const __grid_constant__ PassArgs a) { shared_stop_latch(a, PassId{a.pass, a.is_final}, a.shared_sums); }

// ---------------------------------------------------------------------------------------------
// the fused pass kernel
// ---------------------------------------------------------------------------------------------
// Register / occupancy budget per instantiation.  The pass is FP32-issue bound, not latency bound
// (the cp.async ring hides HBM), so registers for the instruction scheduler beat resident warps;
// at least two CTAs per SM so that one CTA's prologue / reduction / solver epilogue overlaps the
// others' streaming loops.  A/B on B200 (profiles/r1_ab_cta_shapes.txt):
//   pinhole            : 2 CTAs x 8 warps, 128 registers (126 used; 10 warps at 96 was 6 % slower)
//   distortion models  : 3 CTAs x 4 warps, 168 registers (packed accumulators need ~165)
#ifndef GCLM_PIN_WARPS
#define GCLM_PIN_WARPS 8
#endif
#ifndef GCLM_PIN_MINB
#define GCLM_PIN_MINB 2
#endif
#ifndef GCLM_DIST_WARPS
#define GCLM_DIST_WARPS 4
#endif
#ifndef GCLM_DIST_MINB
#define GCLM_DIST_MINB 3
#endif
#ifndef GCLM_RING_UNROLL
#define GCLM_RING_UNROLL 1
#endif
#ifndef GCLM_RING_UNROLL_DIST
#define GCLM_RING_UNROLL_DIST 1
#endif
template <int MODEL, bool PACK>
struct Tune {
  static constexpr int MAX_WARPS_CTA = (MODEL == PINHOLE || !PACK) ? GCLM_PIN_WARPS : GCLM_DIST_WARPS;
  static constexpr int MIN_BLOCKS = (MODEL == PINHOLE || !PACK) ? GCLM_PIN_MINB : GCLM_DIST_MINB;
  static constexpr int UNROLL = 2;  // direct-load loop (optional-input path)
  static constexpr int RING_UNROLL = (MODEL == PINHOLE) ? GCLM_RING_UNROLL : GCLM_RING_UNROLL_DIST;
};
inline int max_warps_for(int model, bool pack) { return (model == PINHOLE || !pack) ? GCLM_PIN_WARPS : GCLM_DIST_WARPS; }

// All static shared memory of a CTA in one block, so that the cold parts of a pass (prologue, epilogue) can live
// in their own __noinline__ functions: the streaming loop's register allocation and instruction schedule then
// do not depend on what the epilogue happens to contain (measured: unrelated epilogue edits moved the pinhole
// pass by 6 % through the schedule ptxas picked for the loop).
template <int NACC>
struct CtaSmem {
  PassConst pc;
  RobustConst rc;
  ImgState img;
  float red[MAX_WARPS][NACC];
  float sum[NACC];
  float grp[4][NACC];
  int last, last_img, skip;
  float part[MAX_WARPS][SCHUR_NSUM + 3];
  float tot[SCHUR_NSUM + 3];
};

// Thread 0 of every CTA: the image's state and the constants of this pass.
template <int MODEL>
__device__ __noinline__ void cta_prologue(const PassArgs& a, const PassId pid, int b, CtaSmem<Dims<MODEL>::NACC>& sm) {
    // everything this CTA (and, if it turns out to be the image's last one, its solver) needs to know about the
    // image: independent L2 loads, ONE round trip (both parameter buffers are fetched, the slot picks one)
    int skip = 0;
    if (a.fused_init && pid.pass == 0 && !pid.is_final) {
      // first pass of the persistent schedule with the set-up folded in: the state lm_init_kernel would have
      // written, straight from the caller's initial estimate (Gravity.__init__ normalises, gravity.py:26)
      for (int i = 0; i < 8; i++) sm.img.P[i] = a.cam0[(size_t)b * 8 + i];
      float g[3] = {a.g0[(size_t)b * 3], a.g0[(size_t)b * 3 + 1], a.g0[(size_t)b * 3 + 2]};
      normalize3(g);
      sm.img.P[8] = g[0]; sm.img.P[9] = g[1]; sm.img.P[10] = g[2]; sm.img.P[11] = 0.f;
      sm.img.lam = a.lambda0; sm.img.prev_cost = 0.f; sm.img.slot = 0; sm.img.flags = 0u;
    } else {
      const int stop = __ldcg(&a.ctrl->stop);
      const int sl0 = __ldcg(&a.slot[b]);
      const unsigned fl = __ldcg(&a.flags[b]);
      const float lam = __ldcg(&a.lam[b]), prev = __ldcg(&a.prev_cost[b]);
      const float4* P4 = reinterpret_cast<const float4*>(a.P + (size_t)b * 2 * PSTRIDE);
      float4 q[6];
#pragma unroll
      for (int i = 0; i < 6; i++) q[i] = __ldcg(P4 + i);
      int sl = sl0;
      if (pid.is_final && stop) sl ^= 1;  // batch stop fired: the tentative update of that pass is dropped
      const float4 p0 = sl ? q[3] : q[0], p1 = sl ? q[4] : q[1], p2 = sl ? q[5] : q[2];
      sm.img.P[0] = p0.x; sm.img.P[1] = p0.y; sm.img.P[2] = p0.z; sm.img.P[3] = p0.w;
      sm.img.P[4] = p1.x; sm.img.P[5] = p1.y; sm.img.P[6] = p1.z; sm.img.P[7] = p1.w;
      sm.img.P[8] = p2.x; sm.img.P[9] = p2.y; sm.img.P[10] = p2.z; sm.img.P[11] = p2.w;
      sm.img.lam = lam; sm.img.prev_cost = prev; sm.img.slot = sl0; sm.img.flags = fl;
      skip = !pid.is_final && (stop || (fl & FLAG_FROZEN));
    }
    sm.skip = skip;
    if (!skip) {
      make_pass_const(sm.img.P, sm.img.P + 8, a.spherical != 0, a.log_focal != 0, pid.is_final != 0 && a.final_rpf != 0, sm.pc);
      sm.rc = make_robust_const(a.up_scale, a.lat_scale);
    }
  }

// Everything after the warps have left their per-warp sums in sm.red: tile sum, partials, ticket; in the image's
// last CTA the ordered sum over the tiles and the solver; in the batch's last CTA the shared-intrinsics tail.
template <int MODEL, bool GEN_D, int IN>
__device__ __noinline__ void cta_finish(const PassArgs& a, const PassId pid, int b, int tile, CtaSmem<Dims<MODEL>::NACC>& sm) {
  constexpr bool FULL = IN != 0, C16 = IN == 2;
  constexpr int NACC = Dims<MODEL>::NACC;
  const int nwarps = blockDim.x >> 5;
  __syncthreads();
  const int S = a.tiles_per_image;
  if (threadIdx.x < NACC) {
    float s = 0.f;
    for (int w = 0; w < nwarps; w++) s += sm.red[w][threadIdx.x];
    if (S == 1) sm.sum[threadIdx.x] = s;
    else { a.partials[((size_t)b * S + tile) * NACC + threadIdx.x] = s; __threadfence(); }
  }
  __syncthreads();
  GCLM_STAMP(3);
  if (S > 1) {
    if (threadIdx.x == 0) {
      const unsigned t = atomicAdd(&a.tickets[b], 1u);
      sm.last = (t == (unsigned)S - 1u);
    }
    __syncthreads();
    if (!sm.last) return;
    GCLM_STAMP(4);
    __threadfence();
    // deterministic: tile order, independent of arrival order.  All S x NACC partials are fetched by the whole
    // CTA at once (one L2 round trip) into the idle ring memory when they fit, then added in order.
    const float* p = a.partials + (size_t)b * S * NACC;
    extern __shared__ __align__(16) unsigned char dyn_smem_f[];
    float* stage = reinterpret_cast<float*>(dyn_smem_f);
    const bool staged = FULL && (size_t)S * NACC * sizeof(float) <= (size_t)RING_STAGES * 5 * blockDim.x * (C16 ? 12 : 16);
    if (staged) {
      const int total = S * NACC;
      for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {  // 8 loads in flight per thread, then the stores
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = i0 + k * blockDim.x; x[k] = i < total ? __ldcg(p + i) : 0.f; }
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = i0 + k * blockDim.x; if (i < total) stage[i] = x[k]; }
      }
      __syncthreads();
    }
    // four groups of consecutive tiles are summed concurrently (each in tile order), then the four group sums in
    // order: a fixed association, identical for every schedule; for S <= 4 it IS the plain sequential sum
    constexpr int NG = 4;
    if (threadIdx.x < NG * NACC) {
      const int g = threadIdx.x / NACC, j = threadIdx.x - g * NACC;
      const int chunk = (S + NG - 1) / NG;
      const int t0 = g * chunk, t1 = min(S, t0 + chunk);
      float s = 0.f;
      int t = t0;
      if (staged) {
        for (; t + 8 <= t1; t += 8) {  // loads first (in-order core: a load per add would serialise on LDS latency)
          float x[8];
#pragma unroll
          for (int k = 0; k < 8; k++) x[k] = stage[(t + k) * NACC + j];
#pragma unroll
          for (int k = 0; k < 8; k++) s += x[k];
        }
        for (; t < t1; t++) s += stage[t * NACC + j];
      } else {
        for (; t + 8 <= t1; t += 8) {
          float x[8];
#pragma unroll
          for (int k = 0; k < 8; k++) x[k] = __ldcg(p + (size_t)(t + k) * NACC + j);
#pragma unroll
          for (int k = 0; k < 8; k++) s += x[k];
        }
        for (; t < t1; t++) s += __ldcg(p + (size_t)t * NACC + j);
      }
      sm.grp[g][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < NACC) sm.sum[threadIdx.x] = ((sm.grp[0][threadIdx.x] + sm.grp[1][threadIdx.x]) + sm.grp[2][threadIdx.x]) + sm.grp[3][threadIdx.x];
    __syncthreads();
  }
  GCLM_STAMP(5);
  if (threadIdx.x == 0) {
    a.tickets[b] = 0u;
    finish_sums<MODEL, GEN_D>(sm.sum, sm.pc.k1, sm.pc.k2, a.up_scale);  // [g0 g1 X k..] x 1/a_up -> [g0 g1 f k..]
    sm.last_img = 0;
    if (a.shared && !pid.is_final) {
      schur_image<MODEL>(a, pid, b, sm.sum, sm.img);
      if (a.fused_shared) {  // image-level ticket: the last image's CTA finishes the LM step for the batch
        __threadfence();
        if (atomicAdd(&a.ctrl->img_done, 1u) == (unsigned)a.B - 1u) { a.ctrl->img_done = 0u; sm.last_img = 1; }
      }
    } else {
      lm_step_image<MODEL>(a, pid, b, sm.sum, sm.img);
    }
  }
  GCLM_STAMP(6);
  if (a.shared && a.fused_shared && !pid.is_final) {  // uniform over the grid
    __syncthreads();
    if (sm.last_img) shared_tail<MODEL>(a, pid, sm.part, sm.tot);
  }
}

// Tiny tiles (small batches: the calibrate() case): every job of a warp is ONE row of one strip.  The whole
// streaming phase of such a CTA lives in this separate function -- its own accumulators, its own warp reduction
// into sm.red -- so that the throughput kernel's loop is compiled without it.
template <int MODEL, bool GEN_D, bool PACK>
__device__ __noinline__ void stream_single_rows(const PassArgs& a, int b, int tile, CtaSmem<Dims<MODEL>::NACC>& sm) {
  constexpr int NACC = Dims<MODEL>::NACC;
  constexpr bool FULL = true;
  using V = typename std::conditional<PACK, f2, float>::type;
  const PassConst pc = sm.pc;
  const RobustConst rc = sm.rc;
  const FieldMask fm = {true, true, true, true};
  extern __shared__ __align__(16) unsigned char dyn_smem_s[];
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(dyn_smem_s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int W = a.W, H = a.H;
  const size_t N = (size_t)H * W;
  const int NS = (W + 127) >> 7;
  const int yt0 = tile * a.rows_per_tile;
  const int yt1 = min(H, yt0 + a.rows_per_tile);
  V acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = splat<V>(0.f);
  {
      // Tiny tiles (small batches: the calibrate() case): every job is ONE row of one strip, so the row ring has
      // nothing to prefetch within a job.  Instead the rows of up to RING_STAGES jobs of this warp are fetched
      // together (one L2 / HBM round trip instead of one per job) and then computed.  Same pixels, same
      // per-lane accumulation order as the generic loop below (jobs in increasing order) -> same sums.
      const unsigned plane_stride = blockDim.x * 16u, stage_stride = 5u * plane_stride;
      const unsigned slot0 = smem_u32 + threadIdx.x * 16u;
      const int njobs = NS * a.rsub;
      for (int j0 = warp; j0 < njobs; j0 += nwarps * RING_STAGES) {
#pragma unroll
        for (int st = 0; st < RING_STAGES; st++) {
          const int job = j0 + st * nwarps;
          const int strip = job % NS, y = yt0 + job / NS, x0 = (strip << 7) + (lane << 2);
          if (job < njobs && x0 < W && y < yt1) {
            const size_t off = (size_t)y * W + x0;
            const unsigned d = slot0 + st * stage_stride;
            cp_async16(d, a.up + (size_t)b * 2 * N + off); cp_async16(d + plane_stride, a.up + (size_t)b * 2 * N + N + off);
            cp_async16(d + 2u * plane_stride, a.lat + (size_t)b * N + off);
            cp_async16(d + 3u * plane_stride, a.uconf + (size_t)b * N + off);
            cp_async16(d + 4u * plane_stride, a.lconf + (size_t)b * N + off);
          }
        }
        cp_async_commit();
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int st = 0; st < RING_STAGES; st++) {
          const int job = j0 + st * nwarps;
          const int strip = job % NS, y = yt0 + job / NS, x0 = (strip << 7) + (lane << 2);
          if (job < njobs && x0 < W && y < yt1) {
            float u[4];
#pragma unroll
            for (int i = 0; i < 4; i++) u[i] = pixel_u(pc, x0 + i);
            const unsigned d = slot0 + st * stage_stride;
            const float4 ox = lds128(d), oy = lds128(d + plane_stride), ol = lds128(d + 2u * plane_stride);
            const float4 cu = lds128(d + 3u * plane_stride), cl = lds128(d + 4u * plane_stride);
            row_accumulate<MODEL, GEN_D, FULL, PACK>(pc, fm, make_row_cols<PACK>(pc, u), pixel_v(pc, y), ox, oy, ol, cu, cl,
                                                     rc, acc, &sm.pc, &sm.rc);
          }
        }
      }
  }
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    float x = hsum(acc[i]);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    if (lane == 0) sm.red[warp][i] = x;
  }
}

// One pass of one CTA (a tile of one image): stream, reduce, and -- in the last CTA of the image -- solve.
// `a` lives in kernel parameter space (never copied); `pid` says which pass this is.
// IN: how the inputs arrive -- 0 = any subset of the planes (direct 128-bit loads), 1 = all five fp32 planes
// (cp.async ring), 2 = pre-transformed fields with fp16 confidences (16 B/pixel: gclm_solve_packed).
template <int MODEL, bool GEN_D, int IN, bool PACK>
__device__ __forceinline__ void
#ifdef GCLM_TMA2D
    pass_body(const PassArgs& a, const PassId pid, const TmaMaps& tm) {
#else
    pass_body(const PassArgs& a, const PassId pid) {
#endif
  constexpr bool FULL = IN != 0, C16 = IN == 2;
#if defined(GCLM_TMA) || defined(GCLM_TMA2D)
  static_assert(!C16, "fp16 confidences: default cp.async ring only");
#endif
  constexpr int NACC = Dims<MODEL>::NACC;
  using V = typename std::conditional<PACK, f2, float>::type;  // 2 pixels per op when packed
  const int b = blockIdx.x / a.tiles_per_image;
  const int tile = blockIdx.x - b * a.tiles_per_image;
  GCLM_STAMP(0);
  __shared__ CtaSmem<NACC> sm;
  if (threadIdx.x == 0) cta_prologue<MODEL>(a, pid, b, sm);
  __syncthreads();
  if (sm.skip) return;  // batch stop latched / image frozen: this pass is a no-op for the CTA
  GCLM_STAMP(1);
  const PassConst pc = sm.pc;
  const FieldMask fm = {a.up != nullptr, a.lat != nullptr, a.uconf != nullptr, a.lconf != nullptr};
  const RobustConst rc = sm.rc;

  extern __shared__ __align__(16) unsigned char dyn_smem[];
#ifdef GCLM_TMA2D
  const unsigned smem_u32 = ((unsigned)__cvta_generic_to_shared(dyn_smem) + 127u) & ~127u;  // TMA boxes need 128 B
#else
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(dyn_smem);
#endif
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int W = a.W, H = a.H;
  const size_t N = (size_t)H * W;
  const int NS = (W + 127) >> 7;
  const int yt0 = tile * a.rows_per_tile;
  const int yt1 = min(H, yt0 + a.rows_per_tile);
  const int RJ = (a.rows_per_tile + a.rsub - 1) / a.rsub;

  V acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i] = splat<V>(0.f);

#ifdef GCLM_TMA2D
  unsigned tma_it = 0;  // stages issued == consumed so far by this warp (ring position / phase)
  const unsigned t2_wbase = smem_u32 + (unsigned)warp * (TMA_STAGES * 5u * TMA_ROWS * 512u);
  const unsigned t2_bars = smem_u32 + (unsigned)nwarps * (TMA_STAGES * 5u * TMA_ROWS * 512u) + (unsigned)warp * (TMA_STAGES * 8u);
  if (FULL) {
    if (lane == 0) {
      for (int k = 0; k < TMA_STAGES; k++) mbar_init(t2_bars + k * 8u, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
#endif
#ifdef GCLM_TMA
  unsigned tma_it = 0;  // rows issued == rows consumed so far by this warp (ring position / phase)
  if (FULL && RING_STAGES > 0) {
    if (lane == 0) {
      const unsigned bars0 = smem_u32 + (unsigned)nwarps * (RING_STAGES * 5u * 512u) + (unsigned)warp * (RING_STAGES * 8u);
      for (int k = 0; k < RING_STAGES; k++) mbar_init(bars0 + k * 8u, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
#endif
#if !defined(GCLM_TMA) && !defined(GCLM_TMA2D)
  if constexpr (IN == 1 && RING_STAGES > 0) {
    if (RJ == 1) {  // tiny tiles: single-row jobs, streamed by their own function (see stream_single_rows)
      stream_single_rows<MODEL, GEN_D, PACK>(a, b, tile, sm);
      cta_finish<MODEL, GEN_D, IN>(a, pid, b, tile, sm);
      return;
    }
  }
#endif
  for (int job = warp; job < NS * a.rsub; job += nwarps) {
    const int strip = job % NS, sub = job / NS;
    const int x0 = (strip << 7) + (lane << 2);
    const int y0 = yt0 + sub * RJ, y1 = min(yt1, y0 + RJ);
#if defined(GCLM_TMA) || defined(GCLM_TMA2D)
    if (!FULL && x0 >= W) continue;
#else
    if (x0 >= W) continue;  // W % 4 == 0 is enforced on the host
#endif
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; i++) u[i] = pixel_u(pc, x0 + i);
#ifdef GCLM_HOIST
    const RowCols<PACK> cols = make_row_cols<PACK>(pc, u);  // column-invariant terms, once per strip
#endif
    const size_t off = (size_t)y0 * W + x0;
    const float* p_ux = (FULL || fm.up) ? a.up + (size_t)b * 2 * N + off : nullptr;
    const float* p_lat = (FULL || fm.lat) ? a.lat + (size_t)b * N + off : nullptr;
    // fp16 confidences: the same plane walk in units of half (4 pixels = 8 bytes per lane and row)
    const float* p_cu = C16 ? nullptr : ((FULL || fm.up_conf) ? a.uconf + (size_t)b * N + off : nullptr);
    const float* p_cl = C16 ? nullptr : ((FULL || fm.lat_conf) ? a.lconf + (size_t)b * N + off : nullptr);
    const __half* h_cu = C16 ? reinterpret_cast<const __half*>(a.uconf) + (size_t)b * N + off : nullptr;
    const __half* h_cl = C16 ? reinterpret_cast<const __half*>(a.lconf) + (size_t)b * N + off : nullptr;
    auto row = [&](const float4 ox, const float4 oy, const float4 ol, const float4 cu, const float4 cl, int y) {
#ifdef GCLM_NOCOMPUTE  // experiment: pure streaming ceiling of this access pattern
      acc[0] = acc[0] + splat<V>(ox.x + ox.y + ox.z + ox.w + oy.x + oy.y + oy.z + oy.w + ol.x + ol.y + ol.z + ol.w +
                                 cu.x + cu.y + cu.z + cu.w + cl.x + cl.y + cl.z + cl.w + (float)y);
      return;
#endif
      // `ol` already holds target_lat = sin(latitude_field): sin_plane_kernel ran once before pass 0
#ifdef GCLM_HOIST
      row_accumulate<MODEL, GEN_D, FULL, PACK>(pc, fm, cols, pixel_v(pc, y), ox, oy, ol, cu, cl, rc, acc, &sm.pc, &sm.rc);
#else
      row_accumulate<MODEL, GEN_D, FULL, PACK>(pc, fm, make_row_cols<PACK>(pc, u), pixel_v(pc, y), ox, oy, ol, cu, cl, rc, acc, &sm.pc, &sm.rc);
#endif
    };
#ifdef GCLM_TMA2D
    if constexpr (FULL) {
      // ---- 2-D TMA ring: lane 0 issues ONE tensor copy per plane per stage (box = this warp's
      // 128-pixel strip x TMA_ROWS rows, cp.async.bulk.tensor.2d, SASS UTMALDG), TMA_STAGES stages
      // ahead; completion on a per-stage mbarrier; every lane reads its float4 with LDS.128.
      constexpr unsigned STAGE_BYTES = 5u * TMA_ROWS * 512u, PLANE_BYTES = TMA_ROWS * 512u;  // smem slots
      const unsigned row_pitch = (unsigned)min(128, W) * 4u;            // dense box rows in shared memory
      const unsigned tx_bytes = 5u * TMA_ROWS * row_pitch;              // bytes one stage delivers
      const int nst = (y1 - y0 + TMA_ROWS - 1) / TMA_ROWS;
      const int xs = strip << 7;
      auto issue = [&](unsigned st, int yy) {  // lane 0 only
        const unsigned d = t2_wbase + st * STAGE_BYTES, bar = t2_bars + st * 8u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, tx_bytes);
        tma_box_g2s(d, &tm.up, xs, (b * 2) * H + yy, bar);
        tma_box_g2s(d + PLANE_BYTES, &tm.up, xs, (b * 2 + 1) * H + yy, bar);
        tma_box_g2s(d + 2u * PLANE_BYTES, &tm.lat, xs, b * H + yy, bar);
        tma_box_g2s(d + 3u * PLANE_BYTES, &tm.uconf, xs, b * H + yy, bar);
        tma_box_g2s(d + 4u * PLANE_BYTES, &tm.lconf, xs, b * H + yy, bar);
      };
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TMA_STAGES; k++)
          if (k < nst) issue((tma_it + k) % TMA_STAGES, y0 + k * TMA_ROWS);
      }
      const bool active = x0 < W;
#pragma unroll 1
      for (int i = 0; i < nst; ++i) {
        const unsigned st = tma_it % TMA_STAGES, par = (tma_it / TMA_STAGES) & 1u;
        if (!mbar_wait(t2_bars + st * 8u, par)) a.flags[b] |= GCLM_STATUS_NONFINITE;
        const unsigned d0 = t2_wbase + st * STAGE_BYTES + (unsigned)lane * 16u;
#pragma unroll
        for (int r = 0; r < TMA_ROWS; r++) {
          const int y = y0 + i * TMA_ROWS + r;
          if (active && y < y1) {
            const unsigned d = d0 + r * row_pitch;
            const float4 ox = lds128(d), oy = lds128(d + PLANE_BYTES), ol = lds128(d + 2u * PLANE_BYTES);
            const float4 cu = lds128(d + 3u * PLANE_BYTES), cl = lds128(d + 4u * PLANE_BYTES);
            row(ox, oy, ol, cu, cl, y);
          }
        }
        __syncwarp();  // every lane is done with this stage before it is refilled
        if (lane == 0 && i + TMA_STAGES < nst) issue(st, y0 + (i + TMA_STAGES) * TMA_ROWS);
        ++tma_it;
      }
    } else
#endif
#ifdef GCLM_TMA
    if constexpr (FULL && RING_STAGES > 0) {
      // ---- TMA ring: lane 0 of the warp issues ONE bulk copy per plane per row (the strip's 512 B,
      // cp.async.bulk global->shared, completion on a per-stage mbarrier), RING_STAGES rows ahead;
      // all lanes wait on the stage's mbarrier phase and read their float4 with LDS.128.
      const unsigned nb = (unsigned)min(128, W - (strip << 7)) * 4u;  // bytes of this strip per plane-row
      const unsigned wbase = smem_u32 + (unsigned)warp * (RING_STAGES * 5u * 512u);
      const unsigned bars = smem_u32 + (unsigned)nwarps * (RING_STAGES * 5u * 512u) + (unsigned)warp * (RING_STAGES * 8u);
      const float* g_ux = a.up + (size_t)b * 2 * N + (size_t)y0 * W + (strip << 7);
      const float* g_lat = a.lat + (size_t)b * N + (size_t)y0 * W + (strip << 7);
      const float* g_cu = a.uconf + (size_t)b * N + (size_t)y0 * W + (strip << 7);
      const float* g_cl = a.lconf + (size_t)b * N + (size_t)y0 * W + (strip << 7);
      auto issue = [&](unsigned st, int yy) {  // executed by lane 0 only
        const size_t o = (size_t)(yy - y0) * W;
        const unsigned d = wbase + st * (5u * 512u), bar = bars + st * 8u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, 5u * nb);
        bulk_g2s(d, g_ux + o, nb, bar); bulk_g2s(d + 512u, g_ux + N + o, nb, bar);
        bulk_g2s(d + 1024u, g_lat + o, nb, bar);
        bulk_g2s(d + 1536u, g_cu + o, nb, bar); bulk_g2s(d + 2048u, g_cl + o, nb, bar);
      };
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < RING_STAGES; k++)
          if (y0 + k < y1) issue((tma_it + k) % RING_STAGES, y0 + k);
      }
      const bool active = x0 < W;
#pragma unroll 1
      for (int y = y0; y < y1; ++y) {
        const unsigned st = tma_it % RING_STAGES, par = (tma_it / RING_STAGES) & 1u;
        if (!mbar_wait(bars + st * 8u, par)) { a.flags[b] |= GCLM_STATUS_NONFINITE; }
        const unsigned d = wbase + st * (5u * 512u) + (unsigned)lane * 16u;
        if (active) {
          const float4 ox = lds128(d), oy = lds128(d + 512u), ol = lds128(d + 1024u);
          const float4 cu = lds128(d + 1536u), cl = lds128(d + 2048u);
          row(ox, oy, ol, cu, cl, y);
        }
        __syncwarp();  // every lane is done with this stage before it is refilled
        if (lane == 0 && y + RING_STAGES < y1) issue(st, y + RING_STAGES);
        ++tma_it;
      }
    } else
#endif
    if constexpr (FULL && RING_STAGES > 0) {
      // ---- cp.async (LDGSTS) ring: every lane prefetches ITS float4 of each of the 5 planes
      // STAGES rows ahead into a private shared-memory slot (no barriers: a lane only ever reads
      // what it copied), so STAGES - 1 rows of loads are in flight while a row is computed and no
      // registers are held across the HBM latency.
      // slot layout per stage: [up_x | up_y | sin_lat] x 16 B per thread, then the two confidence planes
      // x 16 B (fp32) or 8 B (fp16) per thread
      const unsigned plane_stride = blockDim.x * 16u, conf_stride = C16 ? blockDim.x * 8u : plane_stride;
      const unsigned stage_stride = 3u * plane_stride + 2u * conf_stride;
      const unsigned slot0 = smem_u32 + threadIdx.x * 16u;  // [stage][plane][thread]
      const unsigned slot_end = slot0 + (unsigned)RING_STAGES * stage_stride;
      const unsigned conf_off = 3u * plane_stride - (C16 ? threadIdx.x * 8u : 0u);  // from a thread's 16-B slot
      auto issue = [&](unsigned d) {
        cp_async16(d, p_ux); cp_async16(d + plane_stride, p_ux + N);
        cp_async16(d + 2u * plane_stride, p_lat);
        if constexpr (C16) {
          cp_async8(d + conf_off, h_cu); cp_async8(d + conf_off + conf_stride, h_cl);
          h_cu += W; h_cl += W;
        } else {
          cp_async16(d + 3u * plane_stride, p_cu); cp_async16(d + 4u * plane_stride, p_cl);
          p_cu += W; p_cl += W;
        }
        p_ux += W; p_lat += W;
      };
#pragma unroll
      for (int st = 0; st < RING_STAGES; st++) {
        if (y0 + st < y1) issue(slot0 + st * stage_stride);
        cp_async_commit();
      }
      unsigned d = slot0;
#pragma unroll(Tune<MODEL, PACK>::RING_UNROLL)
      for (int y = y0; y < y1; ++y) {
        asm volatile("cp.async.wait_group %0;" ::"n"(RING_STAGES > 0 ? RING_STAGES - 1 : 0) : "memory");
        const float4 ox = lds128(d), oy = lds128(d + plane_stride), ol = lds128(d + 2u * plane_stride);
        float4 cu, cl;
        if constexpr (C16) { cu = lds_half4(d + conf_off); cl = lds_half4(d + conf_off + conf_stride); }
        else { cu = lds128(d + 3u * plane_stride); cl = lds128(d + 4u * plane_stride); }
        row(ox, oy, ol, cu, cl, y);
        if (y + RING_STAGES < y1) issue(d);  // refill the slot just consumed (after its values were used)
        cp_async_commit();
        d += stage_stride;
        if (d == slot_end) d = slot0;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
#pragma unroll(Tune<MODEL, PACK>::UNROLL)
      for (int y = y0; y < y1; ++y) {
        float4 ox = make_float4(0.f, 0.f, 0.f, 0.f), oy = ox, ol = ox;
        float4 cu = make_float4(1.f, 1.f, 1.f, 1.f), cl = cu;
        if (FULL || fm.up) { ox = ldg_stream(p_ux); oy = ldg_stream(p_ux + N); p_ux += W; }
        if (FULL || fm.lat) { ol = ldg_stream(p_lat); p_lat += W; }
        if (FULL || fm.up_conf) { cu = ldg_stream(p_cu); p_cu += W; }
        if (FULL || fm.lat_conf) { cl = ldg_stream(p_cl); p_cl += W; }
        row(ox, oy, ol, cu, cl, y);
      }
    }
  }

  // ---- CTA reduction: shuffle tree, then fixed-order sum over warps ----
  GCLM_STAMP(2);
#pragma unroll
  for (int i = 0; i < NACC; i++) {
    float x = hsum(acc[i]);
    x += __shfl_xor_sync(0xffffffffu, x, 16);
    x += __shfl_xor_sync(0xffffffffu, x, 8);
    x += __shfl_xor_sync(0xffffffffu, x, 4);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    if (lane == 0) sm.red[warp][i] = x;
  }
  cta_finish<MODEL, GEN_D, IN>(a, pid, b, tile, sm);
}

template <int MODEL, bool GEN_D, int IN, bool PACK>
__global__ void __launch_bounds__(Tune<MODEL, PACK>::MAX_WARPS_CTA * 32, Tune<MODEL, PACK>::MIN_BLOCKS)
#ifdef GCLM_TMA2D
    lm_pass_kernel(const __grid_constant__ PassArgs a, const __grid_constant__ TmaMaps tm) {
  pass_body<MODEL, GEN_D, IN, PACK>(a, PassId{a.pass, a.is_final}, tm);
}
#else
    lm_pass_kernel(const __grid_constant__ PassArgs a) {
  pass_body<MODEL, GEN_D, IN, PACK>(a, PassId{a.pass, a.is_final});
}
#endif

// Grid-wide barrier of the persistent kernel (all CTAs are co-resident: cooperative launch).  Sense reversal:
// the last arriver zeroes the count and bumps the generation, the others wait for the generation to move.
// No per-solve reset is needed (count is 0 between barriers), so a captured launch can be replayed.
__device__ __forceinline__ void grid_barrier(Ctrl* c, unsigned& my_gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&c->bar_count, 1u) == gridDim.x - 1u) {
      c->bar_count = 0u;
      __threadfence();
      atomicAdd(&c->bar_gen, 1u);
    } else {
      unsigned v;
      unsigned long long spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&c->bar_gen) : "memory");
        if (++spins > (1ull << 27)) { c->comm_error = 2; break; }  // never hang the device: flag and fall through
      } while (v == my_gen);
    }
    __threadfence();
  }
  my_gen++;
  __syncthreads();
}

// Small batches (the calibrate() case: B = 1, a handful of images): ALL num_steps + 1 passes in ONE launch.
// Every CTA keeps its tile; a grid barrier takes the place of the kernel boundary between passes, so the
// schedule, the tiling and every floating-point sum are those of the one-launch-per-pass path (bit-identical
// results); what disappears is num_steps launches and their gaps.  Launched cooperatively, only when the
// whole grid is resident (gclm_solve checks the occupancy).
#ifndef GCLM_TMA2D
template <int MODEL, bool GEN_D, bool FULL, bool PACK>
__global__ void __launch_bounds__(Tune<MODEL, PACK>::MAX_WARPS_CTA * 32, Tune<MODEL, PACK>::MIN_BLOCKS)
    lm_persistent_kernel(const __grid_constant__ PassArgs a) {
  // every CTA reads the generation before it can possibly move (no barrier completes without this CTA)
  unsigned gen = 0;
  if (threadIdx.x == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&a.ctrl->bar_gen) : "memory");
  if (a.fused_init) {
    // set-up folded in: batch control block (read by others only after the first barrier) ...
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.ctrl->stop = 0; a.ctrl->stop_at = a.num_steps;
      a.ctrl->vote[0] = a.ctrl->vote[1] = 0ull;
      a.ctrl->img_done = 0u; a.ctrl->comm_error = 0;
    }
    // ... and sin(latitude) of THIS CTA's tile (the only part of the plane it will ever read)
    const int b = blockIdx.x / a.tiles_per_image, tile = blockIdx.x - b * a.tiles_per_image;
    const int y0 = tile * a.rows_per_tile, y1 = min(a.H, y0 + a.rows_per_tile);
    const size_t base = ((size_t)b * a.H + y0) * a.W;
    const int n4 = (y1 - y0) * a.W / 4;
    const float4* in = reinterpret_cast<const float4*>(a.lat_raw + base);
    float4* out = reinterpret_cast<float4*>(const_cast<float*>(a.lat) + base);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 x = in[i];
      x.x = sinf(x.x); x.y = sinf(x.y); x.z = sinf(x.z); x.w = sinf(x.w);
      out[i] = x;
    }
    __threadfence();
    __syncthreads();
  }
  for (int p = 0; p < a.num_steps; ++p) {
    pass_body<MODEL, GEN_D, FULL ? 1 : 0, PACK>(a, PassId{p, 0});
    grid_barrier(a.ctrl, gen);
#ifdef GCLM_PHASE_TIMERS
    if (threadIdx.x == 0 && a.dbg) a.dbg[((size_t)p * gridDim.x + blockIdx.x) * 8 + 7] = clock64();
#endif
  }
  pass_body<MODEL, true, FULL ? 1 : 0, PACK>(a, PassId{a.num_steps, 1});
}
#endif

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void lm_init_kernel(PassArgs a, const float* cam0, const float* g0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) {
    a.ctrl->stop = 0; a.ctrl->stop_at = a.num_steps;
    a.ctrl->vote[0] = a.ctrl->vote[1] = 0ull;
    a.ctrl->img_done = 0u; a.ctrl->comm_error = 0;
  }
  if (b >= a.B) return;
  float* P0 = a.P + (size_t)b * 2 * PSTRIDE;
  for (int i = 0; i < 8; i++) P0[i] = cam0[(size_t)b * 8 + i];
  float g[3] = {g0[(size_t)b * 3], g0[(size_t)b * 3 + 1], g0[(size_t)b * 3 + 2]};
  normalize3(g);  // Gravity.__init__ (gravity.py:26)
  for (int i = 0; i < 3; i++) P0[8 + i] = g[i];
  for (int i = 0; i < PSTRIDE; i++) P0[PSTRIDE + i] = P0[i];
  a.lam[b] = a.lambda0;
  a.prev_cost[b] = 0.f;
  a.slot[b] = 0;
  a.flags[b] = 0u;
  a.stop_at[b] = a.num_steps;
  a.tickets[b] = 0u;
}

// target_lat = sin(latitude_field) (lm_optimizer.py:270) is iteration-invariant: computed ONCE per
// solve into engine scratch with the exact sinf (any input range), then streamed by every pass
// instead of the latitude plane -- same 4 B/pixel, ~10 fewer FP32 ops per pixel per pass.
__global__ void __launch_bounds__(256) sin_plane_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = in[i];
    x.x = sinf(x.x); x.y = sinf(x.y); x.z = sinf(x.z); x.w = sinf(x.w);
    out[i] = x;
  }
}

// Widths that are not a multiple of 4 (or planes that are not 16-byte aligned) cannot be walked with
// float4: the fields are copied ONCE per solve into engine scratch with the row pitch padded to a
// multiple of 4.  Padding pixels get confidence 0 (weight and cost contributions vanish exactly);
// missing confidence planes become 1, a missing field becomes 0 with confidence 0; sin() of the
// latitude is folded in.  After this every pass runs the aligned 5-plane fast path.
__global__ void __launch_bounds__(256) repack_kernel(int B, int H, int W, int Wp, const float* up, const float* lat,
                                                     const float* uconf, const float* lconf, float* o_up,
                                                     float* o_sinlat, float* o_uconf, float* o_lconf) {
  const size_t total = (size_t)B * H * Wp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wp);
    const size_t by = i / Wp;
    const int y = (int)(by % H);
    const size_t b = by / H;
    const bool in = x < W;
    const size_t src = ((size_t)b * H + y) * W + x;       // (b, y, x) in a (B, H, W) plane
    const size_t Np = (size_t)H * Wp, N = (size_t)H * W;
    const size_t dst = (size_t)y * Wp + x;
    o_up[b * 2 * Np + dst] = (in && up) ? up[b * 2 * N + (size_t)y * W + x] : 0.f;
    o_up[b * 2 * Np + Np + dst] = (in && up) ? up[b * 2 * N + N + (size_t)y * W + x] : 0.f;
    o_sinlat[b * Np + dst] = (in && lat) ? sinf(lat[src]) : 0.f;
    o_uconf[b * Np + dst] = (in && up) ? (uconf ? uconf[src] : 1.f) : 0.f;
    o_lconf[b * Np + dst] = (in && lat) ? (lconf ? lconf[src] : 1.f) : 0.f;
  }
}

// ---- pre-transformed field layout: [up_x, up_y, sin_lat, w_up, w_lat], confidences fp32 or fp16 ----
__device__ __forceinline__ void store_conf4(void* base, size_t i4, float4 c, bool f16) {
  if (f16) {
    const __half2 a = __floats2half2_rn(c.x, c.y), b = __floats2half2_rn(c.z, c.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned*>(&a); u.y = *reinterpret_cast<const unsigned*>(&b);
    reinterpret_cast<uint2*>(base)[i4] = u;
  } else {
    reinterpret_cast<float4*>(base)[i4] = c;
  }
}
// From the five activated planes the CNN heads write today (geocalib.py:44-89): sin(latitude) and the
// confidences in the requested format; the up planes are already what the pass kernel reads.
__global__ void __launch_bounds__(256) pack_fields_kernel(size_t n4, const float4* __restrict__ lat, const float4* __restrict__ uconf,
                                                          const float4* __restrict__ lconf, float4* __restrict__ sin_lat,
                                                          void* __restrict__ uc_out, void* __restrict__ lc_out, int f16) {
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = lat[i];
    x.x = sinf(x.x); x.y = sinf(x.y); x.z = sinf(x.z); x.w = sinf(x.w);
    sin_lat[i] = x;
    store_conf4(uc_out, i, uconf ? uconf[i] : one, f16 != 0);
    store_conf4(lc_out, i, lconf ? lconf[i] : one, f16 != 0);
  }
}
// The decoder tail fused into the layout the engine streams (geocalib.py:57,73-75,86-88): from the heads'
// PRE-activation outputs -- F.normalize(up), sin(asin(clamp(tanh(l)))) = clamp(tanh(l)), sigmoid(conf) --
// written once; no asin/sin round trip, no separate sin plane, optionally half the confidence bytes.
__global__ void __launch_bounds__(256) pack_decoder_kernel(int B, size_t n4_img, const float4* __restrict__ up_logits,
                                                           const float4* __restrict__ lat_logits, const float4* __restrict__ ucl,
                                                           const float4* __restrict__ lcl, float4* __restrict__ up,
                                                           float4* __restrict__ sin_lat, void* __restrict__ uc_out,
                                                           void* __restrict__ lc_out, int f16) {
  const size_t total = (size_t)B * n4_img;
  auto sig = [](float x) { return 1.f / (1.f + __expf(-x)); };
  auto tl = [](float x) { return fminf(fmaxf(tanhf(x), -1.f + 1e-5f), 1.f - 1e-5f); };
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / n4_img, j = i - b * n4_img;
    const float4 x = up_logits[(b * 2) * n4_img + j], y = up_logits[(b * 2 + 1) * n4_img + j];
    float4 ox, oy;
    { const float n = fmaxf(sqrtf(x.x * x.x + y.x * y.x), 1e-12f); ox.x = x.x / n; oy.x = y.x / n; }
    { const float n = fmaxf(sqrtf(x.y * x.y + y.y * y.y), 1e-12f); ox.y = x.y / n; oy.y = y.y / n; }
    { const float n = fmaxf(sqrtf(x.z * x.z + y.z * y.z), 1e-12f); ox.z = x.z / n; oy.z = y.z / n; }
    { const float n = fmaxf(sqrtf(x.w * x.w + y.w * y.w), 1e-12f); ox.w = x.w / n; oy.w = y.w / n; }
    up[(b * 2) * n4_img + j] = ox; up[(b * 2 + 1) * n4_img + j] = oy;
    const float4 l = lat_logits[i];
    sin_lat[i] = make_float4(tl(l.x), tl(l.y), tl(l.z), tl(l.w));
    const float4 cu = ucl[i], cl = lcl[i];
    store_conf4(uc_out, i, make_float4(sig(cu.x), sig(cu.y), sig(cu.z), sig(cu.w)), f16 != 0);
    store_conf4(lc_out, i, make_float4(sig(cl.x), sig(cl.y), sig(cl.z), sig(cl.w)), f16 != 0);
  }
}

// ---- differentiable LM: the adjoint pass of the implicit-function backward ----------------------
// At the LM fixed point G(theta*, x) = J^T W r = 0, so dtheta*/dx = -(dG/dtheta)^-1 dG/dx and for a loss L(theta*)
//   dL/dx = -d(lambda . G)/dx,   lambda = (dG/dtheta)^-T dL/dtheta          (x = the observed fields / confidences).
// This kernel evaluates d(lambda . G)/dx per pixel (gclm_device.cuh: adjoint_pixel): one streaming pass, no reduction.
// Launch: grid (chunks of the image, B), 256 threads; the per-image constants (tangent basis, focal, distortion,
// lambda) are built once per CTA in shared memory, every thread then walks pixels with a CTA-wide stride (coalesced
// 4-byte accesses of the 5 input and 5 output planes: 40 B per pixel).
template <int MODEL>
__global__ void __launch_bounds__(256) lm_adjoint_kernel(int H, int W, const float* __restrict__ up,
                                                         const float* __restrict__ lat, const float* __restrict__ uconf,
                                                         const float* __restrict__ lconf, const float* __restrict__ cam,
                                                         const float* __restrict__ grav, const float* __restrict__ lam,
                                                         float up_scale, float lat_scale, int spherical, int log_focal,
                                                         float* __restrict__ g_up, float* __restrict__ g_lat,
                                                         float* __restrict__ g_uconf, float* __restrict__ g_lconf) {
  constexpr int NF = Dims<MODEL>::NF;
  const size_t N = (size_t)H * W;
  const int b = blockIdx.y;
  __shared__ PassConst pc_s;
  __shared__ float lam_s[NF];
  if (threadIdx.x == 0) {
    float g[3] = {grav[(size_t)b * 3], grav[(size_t)b * 3 + 1], grav[(size_t)b * 3 + 2]};
    normalize3(g);
    make_pass_const(cam + (size_t)b * 8, g, spherical != 0, log_focal != 0, false, pc_s);
  }
  if (threadIdx.x < NF) lam_s[threadIdx.x] = lam[(size_t)b * NF + threadIdx.x];
  __syncthreads();
  const PassConst pc = pc_s;
  float lm[NF];
#pragma unroll
  for (int i = 0; i < NF; i++) lm[i] = lam_s[i];
  const float BIG = 1e18f;
  const RobustConst rc_probe = make_robust_const(BIG, BIG);
  const float unscale = Forms<MODEL>::HUBER ? BIG : 1.f;
  const float iu2 = 1.f / (up_scale * up_scale), il2 = 1.f / (lat_scale * lat_scale);
  up += (size_t)b * 2 * N; lat += (size_t)b * N;
  if (uconf) uconf += (size_t)b * N;
  if (lconf) lconf += (size_t)b * N;
  if (g_up) g_up += (size_t)b * 2 * N;
  if (g_lat) g_lat += (size_t)b * N;
  if (g_uconf) g_uconf += (size_t)b * N;
  if (g_lconf) g_lconf += (size_t)b * N;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(n / W), x = (int)(n - (size_t)y * W);
    float d[5];
    adjoint_pixel<MODEL>(pc, lm, rc_probe, unscale, iu2, il2, pixel_u(pc, x), pixel_v(pc, y), __ldg(up + n), __ldg(up + N + n),
                         __ldg(lat + n), uconf ? __ldg(uconf + n) : 1.f, lconf ? __ldg(lconf + n) : 1.f, d);
    if (g_up) { g_up[n] = d[0]; g_up[N + n] = d[1]; }
    if (g_lat) g_lat[n] = d[2];
    if (g_uconf) g_uconf[n] = d[3];
    if (g_lconf) g_lconf[n] = d[4];
  }
}

// counter-based RNG for the synthetic generator (splitmix64 finaliser)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t r) { return ((r >> 40) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ void normal2(uint64_t r, float& n0, float& n1) {
  const float a = u01(r), b = u01(r << 24 | (r >> 40));
  const float rad = sqrtf(-2.f * logf(a));
  float s, c;
  sincosf(6.28318530717958647692f * b, &s, &c);
  n0 = rad * c; n1 = rad * s;
}

template <int MODEL>
__global__ void render_kernel(int B, int H, int W, const float* cam, const float* grav, float sigma,
                              float conf_lo, uint64_t seed, float* up, float* lat, float* uconf,
                              float* lconf) {
  const size_t N = (size_t)H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * N) return;
  const int b = (int)(idx / N);
  const size_t n = idx - (size_t)b * N;
  const int y = (int)(n / W), x = (int)(n - (size_t)y * W);
  const float* c = cam + (size_t)b * 8;
  float g[3] = {grav[(size_t)b * 3], grav[(size_t)b * 3 + 1], grav[(size_t)b * 3 + 2]};
  PassConst pc;
  pc.inv_fx = 1.f / c[2]; pc.inv_fy = 1.f / c[3]; pc.cx = c[4]; pc.cy = c[5];
  pc.k1 = c[6]; pc.k2 = c[7]; pc.a = g[0]; pc.b = g[1]; pc.c = g[2];
  const float u = ((float)x - pc.cx) / c[2], v = ((float)y - pc.cy) / c[3];
  const float r2 = u * u + v * v;
  ModelScalars<num_dist(MODEL), float> m;
  if (MODEL == SIMPLE_DIVISIONAL && divisional_is_cold(pc, u, v)) model_scalars<MODEL, float, MODEL == SIMPLE_DIVISIONAL>(pc, r2, m);
  else model_scalars<MODEL, float>(pc, r2, m);
  const float p0x = g[0] - g[2] * u, p0y = g[1] - g[2] * v;
  const float iota = u * p0x + v * p0y;
  float px = m.s * p0x + m.phi * iota * u, py = m.s * p0y + m.phi * iota * v;
  const float rho = fmaxf(sqrtf(px * px + py * py), 1e-12f);
  px /= rho; py /= rho;
  const float U = m.su * u, V = m.su * v;
  const float mm = sqrtf(U * U + V * V + 1.f);
  const float l = (U * g[0] + V * g[1] + g[2]) / mm;
  float la = asinf(fminf(fmaxf(l, -1.f + 1e-6f), 1.f - 1e-6f));
  const uint64_t r0 = mix64(seed ^ (idx * 4 + 0)), r1 = mix64(seed ^ (idx * 4 + 1));
  const uint64_t r2_ = mix64(seed ^ (idx * 4 + 2));
  if (sigma > 0.f) {
    float n0, n1, n2, n3;
    normal2(r0, n0, n1);
    normal2(r1, n2, n3);
    px += sigma * n0; py += sigma * n1; la += sigma * n2;
    const float nr = fmaxf(sqrtf(px * px + py * py), 1e-12f);
    px /= nr; py /= nr;
  }
  if (up) { up[(size_t)b * 2 * N + n] = px; up[(size_t)b * 2 * N + N + n] = py; }
  if (lat) lat[(size_t)b * N + n] = la;
  if (uconf) uconf[(size_t)b * N + n] = conf_lo + (1.f - conf_lo) * u01(r2_);
  if (lconf) lconf[(size_t)b * N + n] = conf_lo + (1.f - conf_lo) * u01(r2_ << 24 | (r2_ >> 40));
}

}  // namespace gclm

// =============================================================================================
// host side
// =============================================================================================
using namespace gclm;

struct gclm_handle {
  int device = 0;
  int num_sms = 148;
  int64_t launches = 0;
  int64_t persistent_solves = 0;  // solves that ran as one cooperative launch
  // scratch
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  // host-path staging
  void* stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  cudaStream_t hstream[2] = {nullptr, nullptr};
  cudaEvent_t hevent[2] = {nullptr, nullptr};
  cudaEvent_t hready[2] = {nullptr, nullptr};
  // One handle = one scratch block: solves are serialised.  `mu` orders host threads, `last_done`
  // (recorded at the end of every solve) orders solves enqueued on DIFFERENT streams (gclm.h contract).
  std::mutex mu;
  cudaEvent_t last_done = nullptr;
  cudaStream_t last_stream = nullptr;
  bool have_last = false;
  // > 48 KB dynamic shared memory opt-in per kernel instantiation: a per-DEVICE attribute, so it lives in
  // the per-device handle ([model][gen_d][full][pack])
  size_t smem_configured[4][2][3][2] = {};
  size_t smem_configured_persistent[4][2][2][2] = {};
  int persistent_blocks_per_sm[4][2][2][2] = {};  // occupancy of lm_persistent_kernel (0 = not queried yet) ...
  unsigned persistent_block_threads[4][2][2][2] = {};  // ... for this block size
#ifdef GCLM_TMA2D
  gclm::TmaMaps maps;  // tensor maps of the current solve
#endif
  // opt-in timing of the pass launches (gclm_pass_timing): one event pair per solve, ring of TIMED
  static constexpr int TIMED = 64;
  bool timing = false;
  cudaEvent_t tev[TIMED][2] = {};
  int tpasses[TIMED] = {};
  int64_t tcount = 0;
  char last_error[256] = {0};
  long long* dbg = nullptr;  // GCLM_PHASE_TIMERS build only
  size_t ctrl_offset = 0;    // where the control block of the last solve lives in scratch (gclm_debug_ctrl)
  // The persistent schedule relies on tickets[] and the barrier words being 0 when it starts (every pass and
  // every barrier leaves them 0).  Their place in the scratch block moves with the batch size, so whenever the
  // layout differs from the one last zeroed, that region is cleared on the stream first.
  size_t zeroed_tickets_off = (size_t)-1, zeroed_ctrl_off = (size_t)-1;
  int zeroed_batch = -1;
};

// Event pair around the pass launches of one solve (only when timing is on; never while capturing).
struct PassTimer {
  gclm_handle* h; cudaStream_t s; int slot = -1;
  PassTimer(gclm_handle* h_, cudaStream_t s_, int passes) : h(h_), s(s_) {
    if (!h->timing) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    slot = (int)(h->tcount % gclm_handle::TIMED);
    for (int k = 0; k < 2; k++)
      if (!h->tev[slot][k] && cudaEventCreate(&h->tev[slot][k]) != cudaSuccess) { slot = -1; return; }
    h->tpasses[slot] = passes;
    cudaEventRecord(h->tev[slot][0], s);
  }
  void stop() {
    if (slot < 0) return;
    cudaEventRecord(h->tev[slot][1], s);
    h->tcount++;
    slot = -1;
  }
};

// NCCL is bound at run time (dlopen) so that libgclm.so loads on boxes without NCCL / without a GPU.
struct NcclUniqueId { char internal[GCLM_COMM_ID_BYTES]; };
typedef int (*nccl_get_unique_id_t)(NcclUniqueId*);
typedef int (*nccl_comm_init_rank_t)(void**, int, NcclUniqueId, int);
typedef int (*nccl_all_reduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(void*);
typedef const char* (*nccl_get_error_string_t)(int);

struct NcclApi {
  void* lib = nullptr;
  nccl_get_unique_id_t get_unique_id = nullptr;
  nccl_comm_init_rank_t comm_init_rank = nullptr;
  nccl_all_reduce_t all_reduce = nullptr;
  nccl_comm_destroy_t comm_destroy = nullptr;
  nccl_get_error_string_t get_error_string = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // torch's bundled libnccl.so.2 is already mapped when the host framework is PyTorch
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) {
      api.get_unique_id = (nccl_get_unique_id_t)dlsym(api.lib, "ncclGetUniqueId");
      api.comm_init_rank = (nccl_comm_init_rank_t)dlsym(api.lib, "ncclCommInitRank");
      api.all_reduce = (nccl_all_reduce_t)dlsym(api.lib, "ncclAllReduce");
      api.comm_destroy = (nccl_comm_destroy_t)dlsym(api.lib, "ncclCommDestroy");
      api.get_error_string = (nccl_get_error_string_t)dlsym(api.lib, "ncclGetErrorString");
    }
  }
  const bool ok = api.lib && api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy;
  return ok ? &api : nullptr;
}

struct gclm_comm {
  void* comm = nullptr;  // ncclComm_t
  int world = 1, rank = 0, device = 0;
  // peer-memory exchange (CUDA IPC): local buffer + every peer's buffer mapped here
  gclm::P2PBuf* local = nullptr;
  gclm::P2PBuf* peer[gclm::P2P_MAX_RANKS] = {nullptr};
  bool p2p_ready = false;
  unsigned seq = 0;
};

#define CUDA_TRY(h, expr)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      if (h) snprintf((h)->last_error, sizeof((h)->last_error), "%s: %s", #expr, cudaGetErrorString(e_)); \
      return e_ == cudaErrorMemoryAllocation ? GCLM_ERR_OUT_OF_MEMORY : GCLM_ERR_CUDA;    \
    }                                                                                      \
  } while (0)

namespace {

// Every entry point runs on the handle's device and leaves the caller's current device untouched.
struct DeviceGuard {
  int prev = -1; bool changed = false, good = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
    if (prev != dev) { good = cudaSetDevice(dev) == cudaSuccess; changed = good; }
  }
  ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
  bool ok() const { return good; }
};

// Serialises the solves of one handle: host mutex + (when the stream changes) a wait on the event the
// previous solve recorded, so two solves never share the scratch block concurrently.
struct SolveOrder {
  gclm_handle* h; cudaStream_t s; std::unique_lock<std::mutex> lock;
  SolveOrder(gclm_handle* h_, cudaStream_t s_) : h(h_), s(s_), lock(h_->mu) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
    if (h->have_last && h->last_stream != s && !capturing) cudaStreamWaitEvent(s, h->last_done, 0);
  }
  void done() {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return;
    if (!h->last_done && cudaEventCreateWithFlags(&h->last_done, cudaEventDisableTiming) != cudaSuccess) return;
    cudaEventRecord(h->last_done, s);
    h->last_stream = s; h->have_last = true;
  }
};

struct Layout {  // carve the scratch buffer
  size_t P, lam, prev, cost0, slot, flags, stop_at, tickets, ctrl, partials, schur, sinlat, repack, total;
};

struct TileCfg { int nwarps, rsub, rows_per_tile, tiles_per_image; };

TileCfg tile_config(int B, int H, int W, int num_sms, int rows_override, int max_warps) {
  // A tile is a band of rows of one image, one CTA each.  It is cut into NS 128-pixel column strips
  // x rsub row bands = NS * rsub jobs; warp w walks jobs w, w + nwarps, ...  rsub is the smallest
  // count that gives every warp the same number of jobs (NS / gcd), so all max_warps warps are busy
  // for any image width.
  TileCfg t;
  const int NS = (W + 127) / 128;
  int nw = max_warps;
  // tiny batches (a handful of images: the calibrate() case) are latency-bound and leave most SMs idle: a warp
  // count that is a multiple of the strip count gives every warp ONE job, so that a tile can be as small as
  // rsub rows and each warp computes a single row per pass
  const bool tiny = rows_override <= 0 && (int64_t)B * ((H + 7) / 8) < num_sms && NS <= max_warps;
  if (tiny) nw = (max_warps / NS) * NS;
  int g = NS, r = nw;
  while (r) { const int q = g % r; g = r; r = q; }  // gcd(NS, nw)
  t.nwarps = nw;
  t.rsub = nw / g;
  const int jobs_per_warp = NS / g;
  // ~120 rows walked by a warp per tile (long tiles amortise the CTA prologue/epilogue), tiles of
  // equal height
  const int rj_target = std::max(4, 120 / jobs_per_warp);
  const int tiles = std::max(1, (H + rj_target * t.rsub / 2) / (rj_target * t.rsub));
  int rj = std::max(1, ((H + tiles - 1) / tiles + t.rsub - 1) / t.rsub);
  if (rows_override > 0) rj = std::max(1, rows_override / t.rsub);
  else
  {
    while (rj >= 8 && (int64_t)B * ((H + rj * t.rsub - 1) / (rj * t.rsub)) < 12ll * num_sms) rj /= 2;  // stays >= 4
    // tiny batches (a handful of images: the calibrate() case) are latency-bound: as many CTAs as the image
    // allows, down to single-row jobs (rows_per_tile stays >= 4, which is what gclm_reserve sizes for)
    while (rj >= 2 && (rj / 2) * t.rsub >= (tiny ? 2 : 4) && (int64_t)B * ((H + rj * t.rsub - 1) / (rj * t.rsub)) < num_sms / 2) rj /= 2;
  }
  t.rows_per_tile = rj * t.rsub;
  t.tiles_per_image = (H + t.rows_per_tile - 1) / t.rows_per_tile;
  return t;
}

Layout make_layout(int B, int max_tiles, size_t pixels, bool repack) {
  Layout L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) & ~size_t(255); return r; };
  L.P = take(sizeof(float) * (size_t)B * 2 * PSTRIDE);
  L.lam = take(sizeof(float) * B);
  L.prev = take(sizeof(float) * B);
  L.cost0 = take(sizeof(float) * (size_t)B * 3);
  L.slot = take(sizeof(int) * B);
  L.flags = take(sizeof(unsigned) * B);
  L.stop_at = take(sizeof(int) * B);
  L.tickets = take(sizeof(unsigned) * B);
  L.ctrl = take(sizeof(Ctrl));
  L.partials = take(sizeof(float) * (size_t)B * max_tiles * MAX_NACC);
  L.schur = take(sizeof(float) * (size_t)B * SCHUR_STRIDE + 4096);
  L.sinlat = take(sizeof(float) * (size_t)B * pixels);
  L.repack = take(repack ? sizeof(float) * (size_t)B * pixels * 4 : 0);  // up (2 planes) + 2 confidences
  L.total = o;
  return L;
}

int ensure_scratch(gclm_handle* h, size_t bytes) {
  if (bytes <= h->scratch_bytes) return GCLM_OK;
  if (h->have_last) cudaEventSynchronize(h->last_done);  // a queued solve may still use the old block
  if (h->scratch) cudaFree(h->scratch);
  h->scratch = nullptr; h->scratch_bytes = 0;
  CUDA_TRY(h, cudaMalloc(&h->scratch, bytes));
  // tickets and the persistent kernel's barrier words rely on starting at 0 (every pass / barrier leaves them 0)
  CUDA_TRY(h, cudaMemset(h->scratch, 0, std::min<size_t>(bytes, (size_t)64 << 20)));
  h->scratch_bytes = bytes;
  h->zeroed_batch = -1;
  return GCLM_OK;
}

int max_tiles_for(int B, int H, int num_sms) {
  // smallest tile tile_config can choose or the caller can ask for: 2 rows for tiny batches, 4 rows otherwise
  const bool tiny = (int64_t)B * ((H + 7) / 8) < num_sms;
  return tiny ? (H + 1) / 2 + 1 : (H + 3) / 4 + 1;
}

#ifdef GCLM_TMA2D
// 2-D tensor maps (width W, all rows of all images of a plane stacked) for the TMA build.
bool encode_maps(gclm_handle* h, const PassArgs& a) {
  TmaMaps& g_maps = h->maps;
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
    enc = reinterpret_cast<encode_fn>(fn);
  }
  auto one = [&](CUtensorMap* m, const float* base, size_t rows) {
    cuuint64_t dims[2] = {(cuuint64_t)a.W, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)a.W * 4};
    cuuint32_t box[2] = {(cuuint32_t)std::min(128, a.W), (cuuint32_t)TMA_ROWS};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const size_t rows = (size_t)a.B * a.H;
  return one(&g_maps.up, a.up, rows * 2) && one(&g_maps.lat, a.lat, rows) && one(&g_maps.uconf, a.uconf, rows) &&
         one(&g_maps.lconf, a.lconf, rows);
}
#endif

template <int MODEL, bool GEN_D, int IN, bool PACK>
void launch_one(gclm_handle* h, const PassArgs& a, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
  size_t& configured = h->smem_configured[MODEL][GEN_D][IN][PACK];  // per device (the handle's)
  if (smem > configured) {
    cudaFuncSetAttribute(lm_pass_kernel<MODEL, GEN_D, IN, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
#ifdef GCLM_TMA2D
  lm_pass_kernel<MODEL, GEN_D, IN, PACK><<<grid, block, smem, s>>>(a, h->maps);
#else
  lm_pass_kernel<MODEL, GEN_D, IN, PACK><<<grid, block, smem, s>>>(a);
#endif
}

template <int MODEL, bool PACK>
void launch_pass_model(gclm_handle* h, const PassArgs& a, bool gen_d, bool full, dim3 grid, dim3 block, cudaStream_t s) {
#ifdef GCLM_TMA2D
  const size_t smem = full ? (size_t)(block.x / 32) * (TMA_STAGES * 5 * TMA_ROWS * 512 + TMA_STAGES * 8) + 128 : 0;
  if (full && !encode_maps(h, a)) return;  // leaves the sticky error for cudaGetLastError-less callers: results stay unset
#else
  const size_t smem = full ? (size_t)RING_STAGES * 5 * block.x * 16 + (size_t)(block.x / 32) * RING_STAGES * 8 : 0;
  if (a.conf16) {  // pre-transformed fields with fp16 confidences: 3 x 16 B + 2 x 8 B per thread and stage
    if constexpr (PACK) {
      const size_t smem16 = (size_t)RING_STAGES * block.x * (3 * 16 + 2 * 8);
      if (gen_d) launch_one<MODEL, true, 2, PACK>(h, a, grid, block, smem16, s);
      else launch_one<MODEL, false, 2, PACK>(h, a, grid, block, smem16, s);
    }
    return;
  }
#endif
  if (gen_d) {
    if (full) launch_one<MODEL, true, 1, PACK>(h, a, grid, block, smem, s);
    else launch_one<MODEL, true, 0, PACK>(h, a, grid, block, 0, s);
  } else {
    if (full) launch_one<MODEL, false, 1, PACK>(h, a, grid, block, smem, s);
    else launch_one<MODEL, false, 0, PACK>(h, a, grid, block, 0, s);
  }
}

bool use_packed();
#ifndef GCLM_TMA2D
// The persistent small-batch schedule: one cooperative launch runs every pass.  Returns false (nothing launched)
// when the grid would not be co-resident -- the caller then uses one launch per pass.
template <int MODEL, bool GEN_D, bool FULL, bool PACK>
bool launch_persistent_one(gclm_handle* h, const PassArgs& a, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
  auto kernel = lm_persistent_kernel<MODEL, GEN_D, FULL, PACK>;
  size_t& configured = h->smem_configured_persistent[MODEL][GEN_D][FULL][PACK];
  if (smem > configured) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    configured = smem;
    h->persistent_blocks_per_sm[MODEL][GEN_D][FULL][PACK] = 0;
  }
  int& per_sm = h->persistent_blocks_per_sm[MODEL][GEN_D][FULL][PACK];
  unsigned& for_threads = h->persistent_block_threads[MODEL][GEN_D][FULL][PACK];
  if (for_threads != block.x) { per_sm = 0; for_threads = block.x; }  // tiny batches use other warp counts
  if (per_sm == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, (int)block.x, smem) != cudaSuccess) { cudaGetLastError(); nb = -1; }
    per_sm = nb > 0 ? nb : -1;
  }
  if (per_sm <= 0 || (int64_t)grid.x > (int64_t)per_sm * h->num_sms) return false;
  PassArgs args = a;
  void* params[] = {&args};
  if (cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kernel), grid, block, params, smem, s) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

template <int MODEL, bool PACK>
bool launch_persistent_model(gclm_handle* h, const PassArgs& a, bool gen_d, bool full, dim3 grid, dim3 block, cudaStream_t s) {
  const size_t smem = full ? (size_t)RING_STAGES * 5 * block.x * 16 + (size_t)(block.x / 32) * RING_STAGES * 8 : 0;
  if (gen_d) return full ? launch_persistent_one<MODEL, true, true, PACK>(h, a, grid, block, smem, s)
                         : launch_persistent_one<MODEL, true, false, PACK>(h, a, grid, block, 0, s);
  return full ? launch_persistent_one<MODEL, false, true, PACK>(h, a, grid, block, smem, s)
              : launch_persistent_one<MODEL, false, false, PACK>(h, a, grid, block, 0, s);
}
#endif

// tickets[] .. end of the control block start at 0 for the persistent schedule (see gclm_handle)
void ensure_persistent_state_zero(gclm_handle* h, const Layout& L, int batch, cudaStream_t s) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
  const bool known = h->zeroed_tickets_off == L.tickets && h->zeroed_ctrl_off == L.ctrl && h->zeroed_batch == batch;
  if (known && !capturing) return;
  cudaMemsetAsync(static_cast<char*>(h->scratch) + L.tickets, 0, L.ctrl + sizeof(Ctrl) - L.tickets, s);
  if (!capturing) { h->zeroed_tickets_off = L.tickets; h->zeroed_ctrl_off = L.ctrl; h->zeroed_batch = batch; }
}

// Small grids only (GCLM_PERSISTENT=0 disables): a few hundred CTAs that are all resident at once.
bool launch_persistent(gclm_handle* h, const PassArgs& a, bool gen_d, bool full, dim3 grid, dim3 block, cudaStream_t s) {
#ifdef GCLM_TMA2D
  return false;
#else
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GCLM_PERSISTENT"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled || !use_packed() || a.conf16 || (int64_t)grid.x > 4ll * h->num_sms) return false;
  switch (a.model) {
    case PINHOLE: return launch_persistent_model<PINHOLE, true>(h, a, gen_d, full, grid, block, s);
    case SIMPLE_RADIAL: return launch_persistent_model<SIMPLE_RADIAL, true>(h, a, gen_d, full, grid, block, s);
    case RADIAL: return launch_persistent_model<RADIAL, true>(h, a, gen_d, full, grid, block, s);
    default: return launch_persistent_model<SIMPLE_DIVISIONAL, true>(h, a, gen_d, full, grid, block, s);
  }
#endif
}

// GCLM_PACK=0 selects the scalar (one pixel per instruction) build of the pass kernel: A/B knob.
bool use_packed() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GCLM_PACK"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

void launch_pass(gclm_handle* h, const PassArgs& a, bool gen_d, bool full, dim3 grid, dim3 block, cudaStream_t s) {
  if (use_packed()) {
    switch (a.model) {
      case PINHOLE: launch_pass_model<PINHOLE, true>(h, a, gen_d, full, grid, block, s); break;
      case SIMPLE_RADIAL: launch_pass_model<SIMPLE_RADIAL, true>(h, a, gen_d, full, grid, block, s); break;
      case RADIAL: launch_pass_model<RADIAL, true>(h, a, gen_d, full, grid, block, s); break;
      default: launch_pass_model<SIMPLE_DIVISIONAL, true>(h, a, gen_d, full, grid, block, s); break;
    }
  } else {
    switch (a.model) {
      case PINHOLE: launch_pass_model<PINHOLE, false>(h, a, gen_d, full, grid, block, s); break;
      case SIMPLE_RADIAL: launch_pass_model<SIMPLE_RADIAL, false>(h, a, gen_d, full, grid, block, s); break;
      case RADIAL: launch_pass_model<RADIAL, false>(h, a, gen_d, full, grid, block, s); break;
      default: launch_pass_model<SIMPLE_DIVISIONAL, false>(h, a, gen_d, full, grid, block, s); break;
    }
  }
}

int validate(const gclm_problem* p, const gclm_options* o, bool packed = false) {
  if (!p || !o) return GCLM_ERR_INVALID_ARGUMENT;
  if (p->batch <= 0 || p->height <= 0 || p->width <= 0) return GCLM_ERR_INVALID_ARGUMENT;
  if (p->camera_model < 0 || p->camera_model > 3) return GCLM_ERR_INVALID_ARGUMENT;  // :179-181
  if (!packed && !p->up_field && !p->latitude_field) return GCLM_ERR_INVALID_ARGUMENT;
  if (!p->camera_init || !p->gravity_init) return GCLM_ERR_INVALID_ARGUMENT;
  if (o->num_steps < 0 || o->num_steps > 10000) return GCLM_ERR_INVALID_ARGUMENT;
  if (gclm_num_params(p->camera_model, p->estimate_gravity, p->estimate_focal) <= 0)
    return GCLM_ERR_INVALID_ARGUMENT;  // "No parameters to optimize" (:342)
  return GCLM_OK;
}

bool needs_repack(const gclm_problem* p) {
  auto mis = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) != 0; };
  return p->width % 4 != 0 || mis(p->up_field) || mis(p->latitude_field) || mis(p->up_confidence) ||
         mis(p->latitude_confidence);
}
inline int padded_width(int w) { return (w + 3) & ~3; }

PassArgs make_args(gclm_handle* h, const gclm_problem* p, const gclm_options* o, const gclm_result* r,
                   const TileCfg& t, const Layout& L, const gclm_packed_fields* packed = nullptr) {
  PassArgs a;
  memset(&a, 0, sizeof(a));
  const bool rp = !packed && needs_repack(p);
  const int Wp = padded_width(p->width);
  char* sb = static_cast<char*>(h->scratch);
  if (rp) {  // all five planes live in scratch with pitch Wp (repack_kernel)
    const size_t plane = (size_t)p->batch * p->height * Wp;
    float* r = reinterpret_cast<float*>(sb + L.repack);
    a.up = r; a.uconf = r + 2 * plane; a.lconf = r + 3 * plane;
    a.lat = reinterpret_cast<const float*>(sb + L.sinlat);
  } else if (packed) {  // pre-transformed fields: streamed as they are
    a.up = packed->up; a.lat = packed->sin_lat;
    a.uconf = static_cast<const float*>(packed->up_conf); a.lconf = static_cast<const float*>(packed->lat_conf);
    a.conf16 = packed->conf_format == GCLM_CONF_F16;
  } else {
    a.up = p->up_field;
    a.lat = p->latitude_field ? reinterpret_cast<const float*>(sb + L.sinlat) : nullptr;
    a.uconf = p->up_field ? p->up_confidence : nullptr;
    a.lconf = p->latitude_field ? p->latitude_confidence : nullptr;
  }
  a.B = p->batch; a.H = p->height; a.W = rp ? Wp : p->width; a.Wt = p->width;
  a.rows_per_tile = t.rows_per_tile; a.tiles_per_image = t.tiles_per_image; a.rsub = t.rsub;
  a.model = p->camera_model; a.num_steps = o->num_steps; a.early_stop = o->early_stop;
  a.fix_lambda = o->fix_lambda; a.spherical = o->use_spherical_manifold; a.log_focal = o->use_log_focal;
  a.stop_mode = o->stop_mode; a.est_gravity = p->estimate_gravity != 0; a.est_focal = p->estimate_focal != 0;
  a.est_dist = p->estimate_dist != 0; a.compute_unc = o->compute_uncertainty;
  a.lambda0 = o->lambda_; a.atol = o->atol; a.rtol = o->rtol;
  a.up_scale = o->up_loss_fn_scale; a.lat_scale = o->lat_loss_fn_scale;
  char* base = static_cast<char*>(h->scratch);
  a.P = reinterpret_cast<float*>(base + L.P);
  a.lam = reinterpret_cast<float*>(base + L.lam);
  a.prev_cost = reinterpret_cast<float*>(base + L.prev);
  a.cost0 = reinterpret_cast<float*>(base + L.cost0);
  a.slot = reinterpret_cast<int*>(base + L.slot);
  a.flags = reinterpret_cast<unsigned*>(base + L.flags);
  a.stop_at = reinterpret_cast<int*>(base + L.stop_at);
  a.tickets = reinterpret_cast<unsigned*>(base + L.tickets);
  a.ctrl = reinterpret_cast<Ctrl*>(base + L.ctrl);
  h->ctrl_offset = L.ctrl;
  if (p->batch != h->zeroed_batch) h->zeroed_batch = -1;  // another layout is about to write over that region
  a.partials = reinterpret_cast<float*>(base + L.partials);
  a.schur = reinterpret_cast<float*>(base + L.schur);
  a.res = *r;
  a.final_rpf = 1;
#ifdef GCLM_PHASE_TIMERS
  if (!h->dbg) { cudaMalloc(reinterpret_cast<void**>(&h->dbg), sizeof(long long) * 8 * 64 * 1024); }
  cudaMemsetAsync(h->dbg, 0, sizeof(long long) * 8 * 64 * 1024);
  a.dbg = h->dbg;
#endif
  return a;
}

// Enqueue sin(latitude_field) -> scratch (once per solve); a.lat already points at the scratch plane.
void launch_sin_plane(gclm_handle* h, const gclm_problem* p, const PassArgs& a, cudaStream_t s) {
  if (needs_repack(p)) {
    const size_t total = (size_t)p->batch * p->height * a.W;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 32);
    repack_kernel<<<blocks, 256, 0, s>>>(p->batch, p->height, p->width, a.W, p->up_field, p->latitude_field,
                                         p->up_confidence, p->latitude_confidence, const_cast<float*>(a.up),
                                         const_cast<float*>(a.lat), const_cast<float*>(a.uconf),
                                         const_cast<float*>(a.lconf));
    h->launches++;
    return;
  }
  if (!p->latitude_field) return;
  const size_t n4 = (size_t)p->batch * p->height * p->width / 4;
  const unsigned blocks = (unsigned)std::min<size_t>((n4 + 255) / 256, (size_t)h->num_sms * 16);
  sin_plane_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(p->latitude_field),
                                          reinterpret_cast<float4*>(const_cast<float*>(a.lat)), n4);
  h->launches++;
}

}  // namespace

extern "C" {

int gclm_version(void) { return GCLM_VERSION; }

const char* gclm_strerror(int status) {
  switch (status) {
    case GCLM_OK: return "ok";
    case GCLM_ERR_INVALID_ARGUMENT: return "invalid argument";
    case GCLM_ERR_CUDA: return "CUDA runtime error";
    case GCLM_ERR_OUT_OF_MEMORY: return "out of device memory";
    case GCLM_ERR_UNSUPPORTED: return "unsupported configuration";
    case GCLM_ERR_NCCL: return "NCCL error";
    case GCLM_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

void gclm_default_options(gclm_options* o) {  // lm_optimizer.py:144-162
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->num_steps = 30; o->lambda_ = 0.1f; o->fix_lambda = 0; o->early_stop = 1;
  o->atol = 1e-8f; o->rtol = 1e-8f; o->use_spherical_manifold = 1; o->use_log_focal = 1;
  o->up_loss_fn_scale = 1e-2f; o->lat_loss_fn_scale = 1e-2f;
  o->stop_mode = GCLM_STOP_BATCH; o->compute_uncertainty = 1; o->rows_per_tile = 0;
  o->schedule = GCLM_SCHEDULE_AUTO;
}

int gclm_num_params(int camera_model, int estimate_gravity, int estimate_focal) {
  if (camera_model < 0 || camera_model > 3) return GCLM_ERR_INVALID_ARGUMENT;
  return 2 * (estimate_gravity != 0) + (estimate_focal != 0) + num_dist(camera_model);
}

int gclm_create(int device, gclm_handle** out) {
  if (!out) return GCLM_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return GCLM_ERR_NO_DEVICE;
  if (device < 0 || device >= count) return GCLM_ERR_INVALID_ARGUMENT;
  gclm_handle* h = new (std::nothrow) gclm_handle();
  if (!h) return GCLM_ERR_OUT_OF_MEMORY;
  h->device = device;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
    h->num_sms = sms;
  *out = h;
  return GCLM_OK;
}

int gclm_destroy(gclm_handle* h) {
  if (!h) return GCLM_OK;
  DeviceGuard guard(h->device);
  if (h->scratch) cudaFree(h->scratch);
  for (int i = 0; i < 2; i++) {
    if (h->stage[i]) cudaFree(h->stage[i]);
    if (h->hstream[i]) cudaStreamDestroy(h->hstream[i]);
    if (h->hevent[i]) cudaEventDestroy(h->hevent[i]);
    if (h->hready[i]) cudaEventDestroy(h->hready[i]);
  }
  if (h->last_done) cudaEventDestroy(h->last_done);
  for (int i = 0; i < gclm_handle::TIMED; i++)
    for (int k = 0; k < 2; k++)
      if (h->tev[i][k]) cudaEventDestroy(h->tev[i][k]);
  delete h;
  return GCLM_OK;
}

int gclm_reserve(gclm_handle* h, int batch, int height, int width) {
  if (!h || batch <= 0 || height <= 0 || width <= 0) return GCLM_ERR_INVALID_ARGUMENT;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->have_last) cudaEventSynchronize(h->last_done);  // never free scratch a queued solve still uses
  const Layout L = make_layout(batch, max_tiles_for(batch, height, h->num_sms), (size_t)height * padded_width(width), width % 4 != 0);
  return ensure_scratch(h, L.total);
}

int64_t gclm_launch_count(const gclm_handle* h) { return h ? h->launches : 0; }
int64_t gclm_persistent_count(const gclm_handle* h) { return h ? h->persistent_solves : 0; }
// debug: raw copy of the device control block (stop, stop_at, vote[2] as 4 words, img_done, comm_error, bar_count, bar_gen)
int gclm_debug_ctrl(gclm_handle* h, int32_t out[10]) {
  if (!h || !h->scratch || !out) return GCLM_ERR_INVALID_ARGUMENT;
  cudaDeviceSynchronize();
  const size_t off = h->ctrl_offset;
  return cudaMemcpy(out, static_cast<char*>(h->scratch) + off, sizeof(int32_t) * 10, cudaMemcpyDeviceToHost) == cudaSuccess ? GCLM_OK : GCLM_ERR_CUDA;
}
#ifdef GCLM_PHASE_TIMERS
int gclm_debug_stamps(gclm_handle* h, long long* out, int n) {  // debug build: copy the first n stamps to the host
  if (!h || !h->dbg) return -1;
  cudaDeviceSynchronize();
  return cudaMemcpy(out, h->dbg, sizeof(long long) * n, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
#endif

int gclm_tile_config(int batch, int height, int width, int camera_model, int num_sms, int rows_per_tile,
                     int32_t out[5]) {
  if (!out || batch <= 0 || height <= 0 || width <= 0 || num_sms <= 0 || camera_model < 0 || camera_model > 3)
    return GCLM_ERR_INVALID_ARGUMENT;
  const TileCfg t = tile_config(batch, height, padded_width(width), num_sms, rows_per_tile,
                                max_warps_for(camera_model, use_packed()));
  out[0] = t.nwarps; out[1] = t.rsub; out[2] = t.rows_per_tile; out[3] = t.tiles_per_image;
  out[4] = max_tiles_for(batch, height, num_sms);
  return GCLM_OK;
}

int gclm_pass_timing(gclm_handle* h, int enable) {
  if (!h) return GCLM_ERR_INVALID_ARGUMENT;
  h->timing = enable != 0;
  h->tcount = 0;
  return GCLM_OK;
}

int gclm_pass_time(gclm_handle* h, double* total_ms, int64_t* passes) {
  if (!h || !total_ms || !passes) return GCLM_ERR_INVALID_ARGUMENT;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  *total_ms = 0.0; *passes = 0;
  const int64_t n = std::min<int64_t>(h->tcount, gclm_handle::TIMED);
  for (int64_t i = 0; i < n; i++) {
    float ms = 0.f;
    CUDA_TRY(h, cudaEventSynchronize(h->tev[i][1]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->tev[i][0], h->tev[i][1]));
    *total_ms += ms; *passes += h->tpasses[i];
  }
  return GCLM_OK;
}

static int solve_impl(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt, const gclm_result* res,
                      void* cuda_stream, const gclm_packed_fields* packed) {
  if (!h || !res) return GCLM_ERR_INVALID_ARGUMENT;
  int st = validate(prob, opt, packed != nullptr);
  if (st != GCLM_OK) return st;
  if (prob->shared_intrinsics) return GCLM_ERR_INVALID_ARGUMENT;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  SolveOrder order(h, s);
  const TileCfg t = tile_config(prob->batch, prob->height, padded_width(prob->width), h->num_sms, opt->rows_per_tile,
                                max_warps_for(prob->camera_model, use_packed()));
  const Layout L = make_layout(prob->batch, std::max(t.tiles_per_image, max_tiles_for(prob->batch, prob->height, h->num_sms)),
                               (size_t)prob->height * padded_width(prob->width), !packed && needs_repack(prob));
  st = ensure_scratch(h, L.total);
  if (st != GCLM_OK) return st;
  PassArgs a = make_args(h, prob, opt, res, t, L, packed);
  const bool full = a.up && a.lat && a.uconf && a.lconf;
  const dim3 grid((unsigned)((size_t)prob->batch * t.tiles_per_image)), block(t.nwarps * 32);
  // small batch (the calibrate() case): every pass in ONE cooperative launch, a grid barrier between passes;
  // with all five planes in place even the set-up (initial state, sin(latitude)) happens inside that launch
  const bool want_persistent = opt->schedule != GCLM_SCHEDULE_PER_PASS && opt->num_steps >= 1 && !a.conf16;
  bool launched = false;
  static int fuse_setup = -1;  // GCLM_FUSED_INIT=0: keep lm_init_kernel / sin_plane_kernel as separate launches (debug knob)
  if (fuse_setup < 0) { const char* e = getenv("GCLM_FUSED_INIT"); fuse_setup = (e && e[0] == '0') ? 0 : 1; }
  if (want_persistent && (int64_t)grid.x <= 4ll * h->num_sms) ensure_persistent_state_zero(h, L, prob->batch, s);
  if (want_persistent && fuse_setup && full && !packed && !needs_repack(prob)) {
    a.fused_init = 1; a.cam0 = prob->camera_init; a.g0 = prob->gravity_init; a.lat_raw = prob->latitude_field;
    PassTimer timer(h, s, opt->num_steps + 1);
    launched = launch_persistent(h, a, !opt->use_log_focal, full, grid, block, s);
    if (launched) { h->launches++; h->persistent_solves++; timer.stop(); order.done(); CUDA_TRY(h, cudaGetLastError()); return GCLM_OK; }
    a.fused_init = 0;
  }
  lm_init_kernel<<<(prob->batch + 127) / 128, 128, 0, s>>>(a, prob->camera_init, prob->gravity_init);
  h->launches++;
  if (!packed) launch_sin_plane(h, prob, a, s);  // packed fields already hold sin(latitude)
  PassTimer timer(h, s, opt->num_steps + 1);
  if (want_persistent && launch_persistent(h, a, !opt->use_log_focal, full, grid, block, s)) {
    h->launches++;
    h->persistent_solves++;
  } else if (opt->schedule == GCLM_SCHEDULE_PERSISTENT) {
    snprintf(h->last_error, sizeof(h->last_error), "persistent schedule: %u CTAs are not co-resident on this device", grid.x);
    return GCLM_ERR_UNSUPPORTED;
  } else {
    for (int it = 0; it < opt->num_steps; it++) {
      a.pass = it; a.is_final = 0;
      launch_pass(h, a, !opt->use_log_focal, full, grid, block, s);
      h->launches++;
    }
    a.pass = opt->num_steps; a.is_final = 1;
    launch_pass(h, a, true, full, grid, block, s);
    h->launches++;
  }
  timer.stop();
  order.done();
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

int gclm_solve(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt,
               const gclm_result* res, void* cuda_stream) {
  return solve_impl(h, prob, opt, res, cuda_stream, nullptr);
}

static bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int gclm_solve_packed(gclm_handle* h, const gclm_packed_fields* f, const gclm_problem* prob, const gclm_options* opt,
                      const gclm_result* res, void* cuda_stream) {
  if (!h || !f || !prob || !res) return GCLM_ERR_INVALID_ARGUMENT;
  if (!f->up || !f->sin_lat || !f->up_conf || !f->lat_conf) return GCLM_ERR_INVALID_ARGUMENT;
  if (f->batch != prob->batch || f->height != prob->height || f->width != prob->width) return GCLM_ERR_INVALID_ARGUMENT;
  if (f->conf_format != GCLM_CONF_F32 && f->conf_format != GCLM_CONF_F16) return GCLM_ERR_INVALID_ARGUMENT;
  const size_t ca = f->conf_format == GCLM_CONF_F16 ? 8 : 16;
  if (f->width % 4 != 0 || !aligned_to(f->up, 16) || !aligned_to(f->sin_lat, 16) || !aligned_to(f->up_conf, ca) ||
      !aligned_to(f->lat_conf, ca)) {
    snprintf(h->last_error, sizeof(h->last_error), "packed fields: width must be a multiple of 4 and the planes aligned");
    return GCLM_ERR_UNSUPPORTED;
  }
  if (f->conf_format == GCLM_CONF_F16 && !use_packed()) return GCLM_ERR_UNSUPPORTED;
  return solve_impl(h, prob, opt, res, cuda_stream, f);
}

int gclm_pack_fields(gclm_handle* h, const gclm_problem* prob, int conf_format, float* sin_lat_out, void* up_conf_out,
                     void* lat_conf_out, void* cuda_stream) {
  if (!h || !prob || !sin_lat_out || !up_conf_out || !lat_conf_out) return GCLM_ERR_INVALID_ARGUMENT;
  if (prob->batch <= 0 || prob->height <= 0 || prob->width <= 0 || !prob->latitude_field) return GCLM_ERR_INVALID_ARGUMENT;
  if (conf_format != GCLM_CONF_F32 && conf_format != GCLM_CONF_F16) return GCLM_ERR_INVALID_ARGUMENT;
  const size_t total = (size_t)prob->batch * prob->height * prob->width;
  if (total % 4 != 0 || !aligned_to(prob->latitude_field, 16) || !aligned_to(sin_lat_out, 16) ||
      (prob->up_confidence && !aligned_to(prob->up_confidence, 16)) ||
      (prob->latitude_confidence && !aligned_to(prob->latitude_confidence, 16)) || !aligned_to(up_conf_out, 16) ||
      !aligned_to(lat_conf_out, 16))
    return GCLM_ERR_UNSUPPORTED;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t n4 = total / 4;
  const unsigned blocks = (unsigned)std::min<size_t>((n4 + 255) / 256, (size_t)h->num_sms * 16);
  pack_fields_kernel<<<blocks, 256, 0, s>>>(n4, reinterpret_cast<const float4*>(prob->latitude_field),
                                            reinterpret_cast<const float4*>(prob->up_confidence),
                                            reinterpret_cast<const float4*>(prob->latitude_confidence),
                                            reinterpret_cast<float4*>(sin_lat_out), up_conf_out, lat_conf_out,
                                            conf_format == GCLM_CONF_F16);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

int gclm_pack_decoder_outputs(gclm_handle* h, int batch, int height, int width, const float* up_logits,
                              const float* latitude_logits, const float* up_conf_logits, const float* lat_conf_logits,
                              int conf_format, float* up_out, float* sin_lat_out, void* up_conf_out, void* lat_conf_out,
                              void* cuda_stream) {
  if (!h || batch <= 0 || height <= 0 || width <= 0) return GCLM_ERR_INVALID_ARGUMENT;
  if (!up_logits || !latitude_logits || !up_conf_logits || !lat_conf_logits || !up_out || !sin_lat_out || !up_conf_out ||
      !lat_conf_out)
    return GCLM_ERR_INVALID_ARGUMENT;
  if (conf_format != GCLM_CONF_F32 && conf_format != GCLM_CONF_F16) return GCLM_ERR_INVALID_ARGUMENT;
  const size_t N = (size_t)height * width;
  if (N % 4 != 0 || !aligned_to(up_logits, 16) || !aligned_to(latitude_logits, 16) || !aligned_to(up_conf_logits, 16) ||
      !aligned_to(lat_conf_logits, 16) || !aligned_to(up_out, 16) || !aligned_to(sin_lat_out, 16) ||
      !aligned_to(up_conf_out, 16) || !aligned_to(lat_conf_out, 16))
    return GCLM_ERR_UNSUPPORTED;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t total4 = (size_t)batch * N / 4;
  const unsigned blocks = (unsigned)std::min<size_t>((total4 + 255) / 256, (size_t)h->num_sms * 16);
  pack_decoder_kernel<<<blocks, 256, 0, s>>>(batch, N / 4, reinterpret_cast<const float4*>(up_logits),
                                             reinterpret_cast<const float4*>(latitude_logits),
                                             reinterpret_cast<const float4*>(up_conf_logits),
                                             reinterpret_cast<const float4*>(lat_conf_logits),
                                             reinterpret_cast<float4*>(up_out), reinterpret_cast<float4*>(sin_lat_out),
                                             up_conf_out, lat_conf_out, conf_format == GCLM_CONF_F16);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

int gclm_system(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt, int rpf_basis,
                float* gradient, float* hessian, float* costs, void* cuda_stream) {
  if (!h || !hessian) return GCLM_ERR_INVALID_ARGUMENT;
  int st = validate(prob, opt);
  if (st != GCLM_OK) return st;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  SolveOrder order(h, s);
  const TileCfg t = tile_config(prob->batch, prob->height, padded_width(prob->width), h->num_sms, opt->rows_per_tile,
                                max_warps_for(prob->camera_model, use_packed()));
  const Layout L = make_layout(prob->batch, std::max(t.tiles_per_image, max_tiles_for(prob->batch, prob->height, h->num_sms)),
                               (size_t)prob->height * padded_width(prob->width), needs_repack(prob));
  st = ensure_scratch(h, L.total);
  if (st != GCLM_OK) return st;
  gclm_result none;
  memset(&none, 0, sizeof(none));
  PassArgs a = make_args(h, prob, opt, &none, t, L);
  a.num_steps = 0; a.pass = 0; a.is_final = 1;
  a.final_rpf = rpf_basis != 0;
  a.sys_g = gradient; a.sys_h = hessian; a.sys_cost = costs;
  const bool full = a.up && a.lat && a.uconf && a.lconf;
  lm_init_kernel<<<(prob->batch + 127) / 128, 128, 0, s>>>(a, prob->camera_init, prob->gravity_init);
  launch_sin_plane(h, prob, a, s);
  const dim3 grid((unsigned)((size_t)prob->batch * t.tiles_per_image)), block(t.nwarps * 32);
  launch_pass(h, a, rpf_basis ? true : !opt->use_log_focal, full, grid, block, s);
  h->launches += 2;
  order.done();
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

int gclm_adjoint(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt, const float* camera,
                 const float* gravity, const float* lambda, float* grad_up, float* grad_latitude, float* grad_up_conf,
                 float* grad_lat_conf, void* cuda_stream) {
  if (!h || !prob || !opt || !camera || !gravity || !lambda) return GCLM_ERR_INVALID_ARGUMENT;
  if (prob->batch <= 0 || prob->height <= 0 || prob->width <= 0 || prob->camera_model < 0 || prob->camera_model > 3)
    return GCLM_ERR_INVALID_ARGUMENT;
  if (!prob->up_field || !prob->latitude_field) return GCLM_ERR_UNSUPPORTED;  // both fields (the training configuration)
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  std::lock_guard<std::mutex> lk(h->mu);
  // >= 8 pixels per thread (amortises the per-CTA constants), <= 65535 images per launch (grid.y)
  const size_t px = (size_t)prob->height * prob->width;
  const unsigned chunks = (unsigned)std::max<size_t>(1, std::min<size_t>((px + 2047) / 2048, 1u << 16));
  if (prob->batch > 65535) return GCLM_ERR_UNSUPPORTED;
  const dim3 grid(chunks, (unsigned)prob->batch);
#define ADJ(M) lm_adjoint_kernel<M><<<grid, 256, 0, s>>>(prob->height, prob->width, prob->up_field, prob->latitude_field, \
    prob->up_confidence, prob->latitude_confidence, camera, gravity, lambda, opt->up_loss_fn_scale, opt->lat_loss_fn_scale, \
    opt->use_spherical_manifold, opt->use_log_focal, grad_up, grad_latitude, grad_up_conf, grad_lat_conf)
  switch (prob->camera_model) {
    case PINHOLE: ADJ(PINHOLE); break;
    case SIMPLE_RADIAL: ADJ(SIMPLE_RADIAL); break;
    case RADIAL: ADJ(RADIAL); break;
    default: ADJ(SIMPLE_DIVISIONAL); break;
  }
#undef ADJ
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

int gclm_solve_shared(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt,
                      const gclm_result* res, gclm_comm* comm, void* cuda_stream) {
  if (!h || !res) return GCLM_ERR_INVALID_ARGUMENT;
  int st = validate(prob, opt);
  if (st != GCLM_OK) return st;
  if (!prob->shared_intrinsics) return GCLM_ERR_INVALID_ARGUMENT;
  NcclApi* nccl = nullptr;
  if (comm && comm->world > 1 && !comm->p2p_ready) {
    nccl = nccl_api();
    if (!nccl || !comm->comm) return GCLM_ERR_NCCL;
  }
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  SolveOrder order(h, s);
  const TileCfg t = tile_config(prob->batch, prob->height, padded_width(prob->width), h->num_sms, opt->rows_per_tile,
                                max_warps_for(prob->camera_model, use_packed()));
  const Layout L = make_layout(prob->batch, std::max(t.tiles_per_image, max_tiles_for(prob->batch, prob->height, h->num_sms)),
                               (size_t)prob->height * padded_width(prob->width), needs_repack(prob));
  st = ensure_scratch(h, L.total);
  if (st != GCLM_OK) return st;
  PassArgs a = make_args(h, prob, opt, res, t, L);
  a.shared = 1;
  a.stop_mode = GCLM_STOP_BATCH;  // shared intrinsics couple all images: the criterion is global
  a.shared_sums = a.schur + (size_t)prob->batch * SCHUR_STRIDE;  // SCHUR_NSUM + 1 floats after the partials
  const bool multi = comm && comm->world > 1;
  const bool p2p = multi && comm->p2p_ready;
  a.fused_shared = (!multi || p2p) ? 1 : 0;  // the LM step ends inside the pass kernel (one launch per step)
  memset(&a.pp, 0, sizeof(a.pp));
  if (p2p) {
    a.pp.world = comm->world; a.pp.rank = comm->rank;
    for (int r = 0; r < comm->world; r++) a.pp.peer[r] = comm->peer[r];
  }
  const bool full = a.up && a.lat && a.uconf && a.lconf;
  const unsigned bblocks = (unsigned)((prob->batch + 127) / 128);
  lm_init_kernel<<<bblocks, 128, 0, s>>>(a, prob->camera_init, prob->gravity_init);
  h->launches++;
  launch_sin_plane(h, prob, a, s);
  const dim3 grid((unsigned)((size_t)prob->batch * t.tiles_per_image)), block(t.nwarps * 32);
  PassTimer timer(h, s, opt->num_steps + 1);  // here the bracket also holds the per-step tail (reduce, exchange, update)
  for (int it = 0; it < opt->num_steps; it++) {
    a.pass = it; a.is_final = 0;
    if (p2p) a.pp.seq = ++comm->seq;
    launch_pass(h, a, !opt->use_log_focal, full, grid, block, s);
    h->launches++;
    if (a.fused_shared) continue;
    // ---- fallback without peer memory: reduce -> ncclAllReduce -> update -> stop latch ----
    shared_reduce_kernel<<<1, block, 0, s>>>(a);
    const int rc = nccl->all_reduce(a.shared_sums, a.shared_sums, SCHUR_NSUM, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm->comm, s);
    if (rc != 0) {
      snprintf(h->last_error, sizeof(h->last_error), "ncclAllReduce: %s",
               nccl->get_error_string ? nccl->get_error_string(rc) : "error");
      return GCLM_ERR_NCCL;
    }
    switch (a.model) {
      case PINHOLE: shared_update_kernel<PINHOLE><<<bblocks, 128, 0, s>>>(a); break;
      case SIMPLE_RADIAL: shared_update_kernel<SIMPLE_RADIAL><<<bblocks, 128, 0, s>>>(a); break;
      case RADIAL: shared_update_kernel<RADIAL><<<bblocks, 128, 0, s>>>(a); break;
      default: shared_update_kernel<SIMPLE_DIVISIONAL><<<bblocks, 128, 0, s>>>(a); break;
    }
    shared_stop_kernel<<<1, 1, 0, s>>>(a);
    h->launches += 4;
  }
  a.pass = opt->num_steps; a.is_final = 1;
  launch_pass(h, a, true, full, grid, block, s);
  h->launches++;
  timer.stop();
  order.done();
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

int gclm_comm_unique_id(void* id_out) {
  if (!id_out) return GCLM_ERR_INVALID_ARGUMENT;
  NcclApi* nccl = nccl_api();
  if (!nccl) return GCLM_ERR_NCCL;
  NcclUniqueId id;
  if (nccl->get_unique_id(&id) != 0) return GCLM_ERR_NCCL;
  memcpy(id_out, id.internal, GCLM_COMM_ID_BYTES);
  return GCLM_OK;
}

int gclm_comm_create(const void* unique_id, int world_size, int rank, int device, gclm_comm** out) {
  if (!unique_id || !out || world_size < 1 || rank < 0 || rank >= world_size) return GCLM_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  NcclApi* nccl = nccl_api();
  if (!nccl) return GCLM_ERR_NCCL;
  DeviceGuard guard(device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  gclm_comm* c = new (std::nothrow) gclm_comm();
  if (!c) return GCLM_ERR_OUT_OF_MEMORY;
  c->world = world_size; c->rank = rank; c->device = device;
  NcclUniqueId id;
  memcpy(id.internal, unique_id, GCLM_COMM_ID_BYTES);
  if (nccl->comm_init_rank(&c->comm, world_size, id, rank) != 0) { delete c; return GCLM_ERR_NCCL; }
  *out = c;
  return GCLM_OK;
}

int gclm_comm_destroy(gclm_comm* c) {
  if (!c) return GCLM_OK;
  DeviceGuard guard(c->device);
  for (int r = 0; r < c->world && r < P2P_MAX_RANKS; r++)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->local) cudaFree(c->local);
  NcclApi* nccl = nccl_api();
  if (nccl && c->comm) nccl->comm_destroy(c->comm);
  delete c;
  return GCLM_OK;
}

int gclm_comm_p2p_export(gclm_comm* c, void* handle_out) {
  if (!c || !handle_out) return GCLM_ERR_INVALID_ARGUMENT;
  if (c->world > P2P_MAX_RANKS) return GCLM_ERR_UNSUPPORTED;
  static_assert(sizeof(cudaIpcMemHandle_t) <= GCLM_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard guard(c->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  if (!c->local) {
    if (cudaMalloc(reinterpret_cast<void**>(&c->local), sizeof(P2PBuf)) != cudaSuccess) return GCLM_ERR_OUT_OF_MEMORY;
    if (cudaMemset(c->local, 0, sizeof(P2PBuf)) != cudaSuccess) return GCLM_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return GCLM_ERR_CUDA;
  }
  cudaIpcMemHandle_t hnd;
  if (cudaIpcGetMemHandle(&hnd, c->local) != cudaSuccess) return GCLM_ERR_CUDA;
  memset(handle_out, 0, GCLM_IPC_HANDLE_BYTES);
  memcpy(handle_out, &hnd, sizeof(hnd));
  return GCLM_OK;
}

int gclm_comm_p2p_disable(gclm_comm* c) {
  if (!c) return GCLM_ERR_INVALID_ARGUMENT;
  c->p2p_ready = false;
  return GCLM_OK;
}

int gclm_comm_p2p_attach(gclm_comm* c, const void* handles) {
  if (!c || !handles || !c->local) return GCLM_ERR_INVALID_ARGUMENT;
  if (c->world > P2P_MAX_RANKS) return GCLM_ERR_UNSUPPORTED;
  DeviceGuard guard(c->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) { c->peer[r] = c->local; continue; }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, static_cast<const char*>(handles) + (size_t)r * GCLM_IPC_HANDLE_BYTES, sizeof(hnd));
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return GCLM_ERR_CUDA;
    }
    c->peer[r] = static_cast<P2PBuf*>(ptr);
  }
  c->p2p_ready = true;
  return GCLM_OK;
}

// ---- host-buffer path: chunked, H2D of chunk i+1 overlaps the solve of chunk i ----
int gclm_solve_host(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt,
                    const gclm_result* res, int chunk_images) {
  if (!h || !res) return GCLM_ERR_INVALID_ARGUMENT;
  int st = validate(prob, opt);
  if (st != GCLM_OK) return st;
  if (prob->shared_intrinsics) return GCLM_ERR_UNSUPPORTED;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  const int B = prob->batch, H = prob->height, W = prob->width;
  const size_t N = (size_t)H * W;
  const int chunk = std::max(1, std::min(chunk_images > 0 ? chunk_images : 128, B));
  const int n = gclm_num_params(prob->camera_model, prob->estimate_gravity, prob->estimate_focal);
  // staging layout per buffer (floats): up 2N | lat N | uconf N | lconf N | cam 8 | grav 3 | results
  const size_t in_f = (size_t)chunk * (5 * N + 11);
  const size_t res_f = (size_t)chunk * (8 + 3 + 2 * n * n + 5 + 6 + 2);
  const size_t buf_bytes = ((in_f + res_f) * sizeof(float) + 255) & ~size_t(255);
  if (h->stage_bytes < buf_bytes) {
    for (int i = 0; i < 2; i++) { if (h->stage[i]) cudaFree(h->stage[i]); h->stage[i] = nullptr; }
    h->stage_bytes = 0;
    for (int i = 0; i < 2; i++) CUDA_TRY(h, cudaMalloc(&h->stage[i], buf_bytes));
    h->stage_bytes = buf_bytes;
  }
  for (int i = 0; i < 2; i++) {
    if (!h->hstream[i]) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->hstream[i], cudaStreamNonBlocking));
    if (!h->hevent[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&h->hevent[i], cudaEventDisableTiming));
    if (!h->hready[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&h->hready[i], cudaEventDisableTiming));
  }
  cudaStream_t s_copy = h->hstream[0], s_comp = h->hstream[1];
  cudaEvent_t ready[2] = {h->hready[0], h->hready[1]}, freed[2] = {h->hevent[0], h->hevent[1]};  // handle-owned
  st = gclm_reserve(h, chunk, H, W);
  int rc = st;
  for (int s0 = 0, i = 0; s0 < B && rc == GCLM_OK; s0 += chunk, i++) {
    const int c = std::min(chunk, B - s0), k = i & 1;
    float* base = static_cast<float*>(h->stage[k]);
    float *d_up = base, *d_lat = d_up + (size_t)chunk * 2 * N, *d_uc = d_lat + (size_t)chunk * N;
    float *d_lc = d_uc + (size_t)chunk * N, *d_cam = d_lc + (size_t)chunk * N, *d_g = d_cam + (size_t)chunk * 8;
    float* d_res = d_g + (size_t)chunk * 3;
    if (i >= 2) cudaStreamWaitEvent(s_copy, freed[k], 0);
    auto h2d = [&](float* dst, const float* src, size_t per_image) {
      if (src) cudaMemcpyAsync(dst, src + (size_t)s0 * per_image, (size_t)c * per_image * sizeof(float),
                               cudaMemcpyHostToDevice, s_copy);
    };
    h2d(d_up, prob->up_field, 2 * N); h2d(d_lat, prob->latitude_field, N);
    h2d(d_uc, prob->up_confidence, N); h2d(d_lc, prob->latitude_confidence, N);
    h2d(d_cam, prob->camera_init, 8); h2d(d_g, prob->gravity_init, 3);
    cudaEventRecord(ready[k], s_copy);
    cudaStreamWaitEvent(s_comp, ready[k], 0);
    gclm_problem p = *prob;
    p.batch = c;
    p.up_field = prob->up_field ? d_up : nullptr; p.latitude_field = prob->latitude_field ? d_lat : nullptr;
    p.up_confidence = prob->up_confidence ? d_uc : nullptr;
    p.latitude_confidence = prob->latitude_confidence ? d_lc : nullptr;
    p.camera_init = d_cam; p.gravity_init = d_g;
    gclm_result r;
    float* q = d_res;
    r.camera = q; q += (size_t)chunk * 8;
    r.gravity = q; q += (size_t)chunk * 3;
    r.covariance = q; q += (size_t)chunk * n * n;
    r.hessian_rpf = q; q += (size_t)chunk * n * n;
    r.uncertainty = q; q += (size_t)chunk * 5;
    r.costs = q; q += (size_t)chunk * 6;
    r.stop_at = reinterpret_cast<int32_t*>(q); q += chunk;
    r.status = reinterpret_cast<uint32_t*>(q);
    rc = gclm_solve(h, &p, opt, &r, s_comp);
    if (rc != GCLM_OK) break;
    auto d2h = [&](void* dst, const void* src, size_t per_image_bytes) {
      if (dst) cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)s0 * per_image_bytes, src,
                               (size_t)c * per_image_bytes, cudaMemcpyDeviceToHost, s_comp);
    };
    d2h(res->camera, r.camera, 8 * sizeof(float)); d2h(res->gravity, r.gravity, 3 * sizeof(float));
    d2h(res->covariance, r.covariance, (size_t)n * n * sizeof(float));
    d2h(res->hessian_rpf, r.hessian_rpf, (size_t)n * n * sizeof(float));
    d2h(res->uncertainty, r.uncertainty, 5 * sizeof(float)); d2h(res->costs, r.costs, 6 * sizeof(float));
    d2h(res->stop_at, r.stop_at, sizeof(int32_t)); d2h(res->status, r.status, sizeof(uint32_t));
    cudaEventRecord(freed[k], s_comp);
  }
  cudaError_t e = cudaStreamSynchronize(s_comp);
  cudaStreamSynchronize(s_copy);
  if (rc != GCLM_OK) return rc;
  if (e != cudaSuccess) { snprintf(h->last_error, sizeof(h->last_error), "%s", cudaGetErrorString(e)); return GCLM_ERR_CUDA; }
  return GCLM_OK;
}

int gclm_render_fields(gclm_handle* h, int batch, int height, int width, int camera_model,
                       const float* camera, const float* gravity, float sigma, float conf_lo,
                       uint64_t seed, float* up_field, float* latitude_field, float* up_confidence,
                       float* latitude_confidence, void* cuda_stream) {
  if (!h || batch <= 0 || height <= 0 || width <= 0 || !camera || !gravity) return GCLM_ERR_INVALID_ARGUMENT;
  if (camera_model < 0 || camera_model > 3) return GCLM_ERR_INVALID_ARGUMENT;
  DeviceGuard guard(h->device);
  if (!guard.ok()) return GCLM_ERR_CUDA;
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const size_t total = (size_t)batch * height * width;
  const unsigned blocks = (unsigned)((total + 255) / 256);
#define RENDER(M) render_kernel<M><<<blocks, 256, 0, s>>>(batch, height, width, camera, gravity, sigma, conf_lo, seed, up_field, latitude_field, up_confidence, latitude_confidence)
  switch (camera_model) {
    case PINHOLE: RENDER(PINHOLE); break;
    case SIMPLE_RADIAL: RENDER(SIMPLE_RADIAL); break;
    case RADIAL: RENDER(RADIAL); break;
    default: RENDER(SIMPLE_DIVISIONAL); break;
  }
#undef RENDER
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GCLM_OK;
}

}  // extern "C"
