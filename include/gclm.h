/*
 * gclm.h -- C ABI of the B200-native Levenberg-Marquardt calibration engine (libgclm.so).
 *
 * The reference (cvg/GeoCalib @ 3ee7ec34) has NO native boundary for this path: the seam is the
 * Python object `geocalib.lm_optimizer.LMOptimizer` owned by `geocalib.geocalib.GeoCalib`
 * (geocalib/geocalib.py:106) and called at geocalib/geocalib.py:119.  This header is therefore the
 * boundary a maintainer would bind (ctypes stub in INTEGRATION.md); every entry point names the
 * reference interface it replaces.
 *
 * Conventions
 *   - C linkage, plain pointers and sizes, no C++/torch types, no exceptions across the boundary.
 *   - Every function returns a gclm_status (0 = ok, <0 = error); gclm_strerror() explains it.
 *   - All `const float*` inputs and all outputs are DEVICE pointers owned by the caller unless the
 *     function name ends in `_host`.  The handle owns only scratch memory.
 *   - Work is enqueued on the caller's CUDA stream (a `cudaStream_t` passed as void*); the library
 *     never synchronises the device in gclm_solve / gclm_solve_shared.
 *   - One handle per device, different handles are independent.  A handle owns ONE scratch block, so
 *     its solves are serialised: calls from several host threads take a per-handle mutex, and a solve
 *     enqueued on a different stream than the previous one first waits (cudaStreamWaitEvent) on the
 *     event the previous solve recorded.  Solves captured into a CUDA graph are exempt: replay graphs
 *     of one handle on one stream.
 *   - Every entry point runs on the handle's device and restores the caller's current device.
 *   - Numerical failure never raises (reference: lm_optimizer.py:129-133, 497-499): a failed
 *     Cholesky gives a zero step for THAT image and sets bit 0 of its status word.
 */
#ifndef GCLM_H_
#define GCLM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCLM_VERSION 100 /* 0.1.0 */

typedef enum {
  GCLM_OK = 0,
  GCLM_ERR_INVALID_ARGUMENT = -1,
  GCLM_ERR_CUDA = -2,
  GCLM_ERR_OUT_OF_MEMORY = -3,
  GCLM_ERR_UNSUPPORTED = -4,
  GCLM_ERR_NCCL = -5,
  GCLM_ERR_NO_DEVICE = -6
} gclm_status;

/* Camera models: geocalib/camera.py:945-950 (`camera_models`). */
typedef enum {
  GCLM_PINHOLE = 0,           /* camera.py:522-562 */
  GCLM_SIMPLE_RADIAL = 1,     /* camera.py:565-660 */
  GCLM_RADIAL = 2,            /* camera.py:663-786 */
  GCLM_SIMPLE_DIVISIONAL = 3  /* camera.py:789-942 */
} gclm_camera_model;

/* Early-stop semantics (lm_optimizer.py:90-92, 619-625). */
typedef enum {
  GCLM_STOP_BATCH = 0,    /* reference: stop when ALL images satisfy allclose() at the same step */
  GCLM_STOP_PER_IMAGE = 1 /* freeze each image individually as soon as it converged           */
} gclm_stop_mode;

/* status bits written to gclm_result.status */
#define GCLM_STATUS_CHOLESKY_FAILED 1u /* lm_optimizer.py:129-133 (per image here) */
#define GCLM_STATUS_COV_SINGULAR 2u    /* torch.inverse would have raised (lm_optimizer.py:484) */
#define GCLM_STATUS_NONFINITE 4u
#define GCLM_STATUS_COMM_TIMEOUT 8u /* shared intrinsics, fused exchange: a peer never contributed; the solve was aborted */

/* How the num_steps + 1 passes of gclm_solve are issued.  Results are bit-identical either way. */
typedef enum {
  GCLM_SCHEDULE_AUTO = 0,      /* persistent when the whole grid is co-resident (small batches), else per pass */
  GCLM_SCHEDULE_PER_PASS = 1,  /* one kernel launch per pass (throughput schedule for large batches)           */
  GCLM_SCHEDULE_PERSISTENT = 2 /* ONE cooperative launch, grid barrier between passes; GCLM_ERR_UNSUPPORTED
                                  when the grid does not fit on the device                                     */
} gclm_schedule;

typedef struct gclm_handle gclm_handle;

/*
 * The optimisation problem: replaces the `data` dict + the LMOptimizer attributes
 * (lm_optimizer.py:144-171, 189-246, 646-664).
 */
typedef struct {
  int32_t batch;  /* B */
  int32_t height; /* H of the fields */
  int32_t width;  /* W of the fields */
  int32_t camera_model; /* gclm_camera_model; LMOptimizer.set_camera_model (:173-187) */

  /* data["up_field"] (B,2,H,W), data["latitude_field"] (B,1,H,W) radians, contiguous NCHW fp32.
   * Either may be NULL (lm_optimizer.py:265,269), not both.  Any H, W: rows that are a multiple of
   * 4 pixels with 16-byte aligned planes are streamed in place, other shapes are repacked once per
   * solve into padded scratch. */
  const float* up_field;
  const float* latitude_field;
  /* data["up_confidence"], data["latitude_confidence"] (B,H,W) or NULL (:295,307). */
  const float* up_confidence;
  const float* latitude_confidence;

  /* Initial estimate = get_trivial_estimation() incl. priors/scales (lm_optimizer.py:20-58).
   * camera_init (B,8) = [w,h,fx,fy,cx,cy,k1,k2] (camera.py:24-41); gravity_init (B,3) unit. */
  const float* camera_init;
  const float* gravity_init;

  /* setup_optimization_and_priors (:189-246): which blocks are free. */
  int32_t estimate_gravity; /* 0 when data has "prior_gravity" */
  int32_t estimate_focal;   /* 0 when data has "prior_focal"   */
  int32_t estimate_dist;    /* 0 when data has "prior_dist" (k columns are still solved, :340) */
  int32_t shared_intrinsics; /* must be 0 for gclm_solve, 1 for gclm_solve_shared */
} gclm_problem;

/* LMOptimizer.default_conf (lm_optimizer.py:144-162); gclm_default_options() fills the defaults. */
typedef struct {
  int32_t num_steps;              /* 30   */
  float lambda_;                  /* 0.1  */
  int32_t fix_lambda;             /* 0    */
  int32_t early_stop;             /* 1    */
  float atol;                     /* 1e-8 */
  float rtol;                     /* 1e-8 */
  int32_t use_spherical_manifold; /* 1    */
  int32_t use_log_focal;          /* 1    */
  float up_loss_fn_scale;         /* 1e-2 */
  float lat_loss_fn_scale;        /* 1e-2 */
  int32_t stop_mode;              /* gclm_stop_mode; GCLM_STOP_BATCH */
  int32_t compute_uncertainty;    /* 1 = estimate_uncertainty() (:463-516), eval mode */
  int32_t rows_per_tile;          /* 0 = auto; tuning knob: image rows per CTA */
  int32_t schedule;               /* gclm_schedule; GCLM_SCHEDULE_AUTO */
  int32_t reserved[2];
} gclm_options;

/*
 * Outputs: replaces the dict returned by LMOptimizer.forward (:509-516, 575-588, 638-644, 664).
 * All device pointers, caller allocated; any may be NULL to skip that output.
 * n = 2*estimate_gravity + estimate_focal + num_dist(camera_model) (:421-425).
 */
typedef struct {
  float* camera;      /* (B,8)  out["camera"]._data                                        */
  float* gravity;     /* (B,3)  out["gravity"].vec3d                                       */
  float* covariance;  /* (B,n,n) out["covariance"]                                         */
  float* uncertainty; /* (B,5)  roll, pitch, gravity, focal, vfov _uncertainty              */
  float* costs;       /* (B,6)  initial_cost, initial_up_cost, initial_latitude_cost,
                                final_cost, final_up_cost, final_latitude_cost             */
  int32_t* stop_at;   /* (B)    out["stop_at"]                                             */
  uint32_t* status;   /* (B)    GCLM_STATUS_* bits                                         */
  float* hessian_rpf; /* (B,n,n) undamped J^T W J in the (roll,pitch,f,k) basis (:481-483) */
} gclm_result;

/* Library / build info. */
int gclm_version(void);
const char* gclm_strerror(int status);
void gclm_default_options(gclm_options* opt);
/* n as defined above, or <0 on invalid model. */
int gclm_num_params(int camera_model, int estimate_gravity, int estimate_focal);

/* Create / destroy the per-device engine (scratch allocator, SM count query). */
int gclm_create(int device, gclm_handle** out);
int gclm_destroy(gclm_handle* h);
/* Pre-allocate scratch for problems up to (batch,height,width); makes gclm_solve allocation-free
 * (required before capturing gclm_solve into a CUDA graph). */
int gclm_reserve(gclm_handle* h, int batch, int height, int width);

/*
 * Per-image intrinsics: replaces LMOptimizer.forward -> optimize() with shared_intrinsics=False
 * (lm_optimizer.py:551-644).  Enqueues num_steps+1 fused passes on `cuda_stream`; asynchronous.
 * Large batches: one kernel launch per pass.  Small batches (the calibrate() case, extractor.py:93-95:
 * B = 1): when every CTA of the grid is resident at once the passes run inside ONE cooperative launch
 * with a grid barrier in place of the kernel boundary (opt->schedule).
 */
int gclm_solve(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt,
               const gclm_result* res, void* cuda_stream);

/*
 * Pre-transformed field layout.  The CNN heads end in F.normalize / asin(clamp(tanh)) /
 * sigmoid and write five fp32 planes (geocalib.py:44-89); the LM immediately takes sin() of the latitude
 * (lm_optimizer.py:270) -- the asin cancels.  This layout is what the engine actually streams:
 *   up (B,2,H,W) unit vectors, sin_lat (B,1,H,W) = sin(latitude), up_conf / lat_conf (B,H,W) in fp32
 *   (20 B/pixel/pass) or fp16 (16 B/pixel/pass).
 * gclm_pack_decoder_outputs builds it from the heads' PRE-activation outputs (the fused decoder tail: one
 * write instead of five planes + a sin pass); gclm_pack_fields builds it from the five activated planes
 * (up is used in place).  All buffers are caller-owned device memory; width % 4 == 0 and 16-byte aligned
 * planes are required (GCLM_ERR_UNSUPPORTED otherwise).  gclm_solve_packed is gclm_solve on such fields:
 * `prob` supplies everything except the field pointers (ignored).  fp16 confidences perturb the weights
 * by <= 2^-11 relative, zero-mean: measured parity in tests/test_gpu_packed.py.
 */
typedef enum { GCLM_CONF_F32 = 0, GCLM_CONF_F16 = 1 } gclm_conf_format;
typedef struct {
  int32_t batch, height, width;
  int32_t conf_format;  /* gclm_conf_format */
  const float* up;      /* (B,2,H,W) */
  const float* sin_lat; /* (B,1,H,W) */
  const void* up_conf;  /* (B,H,W) float or half */
  const void* lat_conf; /* (B,H,W) float or half */
} gclm_packed_fields;
int gclm_pack_fields(gclm_handle* h, const gclm_problem* prob, int conf_format, float* sin_lat_out,
                     void* up_conf_out, void* lat_conf_out, void* cuda_stream);
int gclm_pack_decoder_outputs(gclm_handle* h, int batch, int height, int width, const float* up_logits,
                              const float* latitude_logits, const float* up_conf_logits,
                              const float* lat_conf_logits, int conf_format, float* up_out, float* sin_lat_out,
                              void* up_conf_out, void* lat_conf_out, void* cuda_stream);
int gclm_solve_packed(gclm_handle* h, const gclm_packed_fields* fields, const gclm_problem* prob,
                      const gclm_options* opt, const gclm_result* res, void* cuda_stream);

/*
 * Diagnostic: the reduced Gauss-Newton system of ONE pass at the given parameters, i.e. what
 * LMOptimizer.setup_system + calculate_costs return (lm_optimizer.py:276-315, 387-461), with the full
 * column set [g0 g1 | f | k..] (NF = 3 + num_dist): gradient (B,NF) = J^T W r, hessian (B,NF,NF) =
 * J^T W J (undamped), costs (B,2) = mean up / latitude cost.  rpf_basis = 0: the LM basis
 * (spherical tangent or roll/pitch per opt->use_spherical_manifold, log focal per opt->use_log_focal);
 * 1: the covariance basis (roll, pitch, linear focal; as_rpf=True, :411-416).  Parameters are
 * prob->camera_init / gravity_init.  Used by the parity tests (iteration-0 G/H/cost).
 * gradient / costs may be NULL.  Asynchronous on `cuda_stream`.
 */
int gclm_system(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt, int rpf_basis,
                float* gradient, float* hessian, float* costs, void* cuda_stream);

/*
 * Differentiable LM (reference: the siclib training variant unrolls the LM with autograd,
 * siclib/models/optimization/lm_optimizer.py:38-57,577-625).  Here the backward pass is the implicit-function
 * theorem at the converged estimate: G(theta*, x) = J^T W r = 0  =>  dL/dx = -d(lambda . G)/dx with
 * lambda = (dG/dtheta)^-T dL/dtheta.  gclm_adjoint is that last factor: ONE streaming pass that writes
 * d(lambda . G)/d(up_field), /d(latitude_field), /d(up_confidence), /d(latitude_confidence) per pixel at the
 * parameters (camera (B,8), gravity (B,3)) and multipliers lambda (B, 3 + num_dist; full columns [g0 g1 f k..] in
 * the LM basis selected by opt).  dG/dtheta itself comes from central differences of gclm_system (2 n passes);
 * geocalib_b200/autograd.py assembles the torch.autograd.Function.  Any output may be NULL.
 */
int gclm_adjoint(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt, const float* camera,
                 const float* gravity, const float* lambda, float* grad_up, float* grad_latitude,
                 float* grad_up_conf, float* grad_lat_conf, void* cuda_stream);

/*
 * Shared intrinsics: replaces optimize() with shared_intrinsics=True (lm_optimizer.py:350-383,
 * 599-603): one focal / distortion for the whole batch, per-image gravity.  The dense
 * (2B+n_i)^2 arrow system of the reference is solved through its Schur complement.
 * `comm` is NULL for a single GPU, else a handle from gclm_comm_create: the batch is then the
 * union of all ranks' batches and one all-reduce of <= 2 n_i^2 + 2 n_i + 2 floats is done per LM step.
 * With peer memory attached (gclm_comm_p2p_*) or on a single GPU the whole LM step -- ordered sum of
 * the per-image partials, exchange over NVLink, every image's update, stop latch -- runs in the
 * epilogue of the pass kernel's last CTA: num_steps + 1 pass launches per solve.  Without peer memory
 * three small kernels and one ncclAllReduce follow every pass.
 */
typedef struct gclm_comm gclm_comm;
int gclm_solve_shared(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt,
                      const gclm_result* res, gclm_comm* comm, void* cuda_stream);

/*
 * Multi-GPU communicator for gclm_solve_shared (one process per GPU, NCCL over NVLink).
 * unique_id is GCLM_COMM_ID_BYTES bytes produced on rank 0 by gclm_comm_unique_id and broadcast
 * by the host framework (torch.distributed in the Python binding).
 */
#define GCLM_COMM_ID_BYTES 128
int gclm_comm_unique_id(void* id_out);
int gclm_comm_create(const void* unique_id, int world_size, int rank, int device, gclm_comm** out);
int gclm_comm_destroy(gclm_comm* c);
/*
 * Optional peer-memory exchange (one node, NVLink / NVSwitch): instead of an NCCL all-reduce, the
 * per-step reduction kernel stores its partial sums directly into every peer's exchange buffer and
 * waits for the peers' sums in its own -- one kernel does the reduction AND the collective.
 * Each rank exports a CUDA-IPC handle of its buffer (GCLM_IPC_HANDLE_BYTES bytes); the host
 * framework all-gathers the handles (rank order) and every rank attaches them, followed by a
 * barrier.  Without these two calls gclm_solve_shared uses NCCL.  At most 8 ranks.
 */
#define GCLM_IPC_HANDLE_BYTES 64
int gclm_comm_p2p_export(gclm_comm* c, void* handle_out);
int gclm_comm_p2p_attach(gclm_comm* c, const void* handles /* world x GCLM_IPC_HANDLE_BYTES */);
/* Back to NCCL for this communicator (call on every rank, e.g. when one rank failed to attach). */
int gclm_comm_p2p_disable(gclm_comm* c);

/*
 * Host-buffer convenience path (end-to-end measurement; the call a non-CUDA host makes):
 * all pointers in `prob` and `res` are HOST pointers (pinned for full overlap).  The batch is
 * processed in chunks of `chunk_images`; H2D copies of chunk k+1 overlap the solve of chunk k on
 * two internal streams; results are copied back.  Synchronous: returns when `res` is complete.
 * Per-image intrinsics only; GCLM_STOP_BATCH applies per chunk.
 */
int gclm_solve_host(gclm_handle* h, const gclm_problem* prob, const gclm_options* opt,
                    const gclm_result* res, int chunk_images);

/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t gclm_launch_count(const gclm_handle* h);
/* Number of solves that ran under the persistent schedule (one cooperative launch for all passes). */
int64_t gclm_persistent_count(const gclm_handle* h);

/*
 * How the pass kernel would tile a problem (host-only, no device needed): out = {warps per CTA,
 * row bands per tile, rows per tile, tiles per image, tiles per image that gclm_reserve sizes for}.
 * rows_per_tile as in gclm_options (0 = auto).  For tests and tuning.
 */
int gclm_tile_config(int batch, int height, int width, int camera_model, int num_sms, int rows_per_tile,
                     int32_t out[5]);

/*
 * Opt-in timing of the fused pass kernel (bench.py's roofline): while enabled, every gclm_solve /
 * gclm_solve_shared records a CUDA event pair on the caller's stream around its num_steps + 1 pass
 * launches (the once-per-solve set-up kernels stay outside the bracket).  gclm_pass_time waits for
 * the solves enqueued since gclm_pass_timing(h, 1) (the last 64 at most) and returns the summed
 * bracket time and the number of pass launches it covers.  Off by default; solves that are being
 * captured into a CUDA graph are not timed.
 */
int gclm_pass_timing(gclm_handle* h, int enable);
int gclm_pass_time(gclm_handle* h, double* total_ms, int64_t* passes);

/*
 * Synthetic-field generator used by the benchmark: renders the perspective field of
 * the given cameras/gravities on device (get_perspective_field, perspective_fields.py:278-320),
 * adds N(0, sigma^2) noise, re-normalises the up vectors, draws confidences ~ U(conf_lo, 1).
 * camera (B,8), gravity (B,3) device pointers; outputs as in gclm_problem (any may be NULL).
 */
int gclm_render_fields(gclm_handle* h, int batch, int height, int width, int camera_model,
                       const float* camera, const float* gravity, float sigma, float conf_lo,
                       uint64_t seed, float* up_field, float* latitude_field, float* up_confidence,
                       float* latitude_confidence, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* GCLM_H_ */
