/*
 * lightpath.h — C ABI of liblightpath.so, the B200 (sm_100a) implementation of
 * the per-pixel Schwarzschild null-geodesic ray-tracing hot path of
 * dhg14n9/Light-path-tracer.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point cites the reference interface it replaces (file:line into
 * the reference tree).  Conventions:
 *
 *   - every function returns an int: LP_OK (0) or a negative LP_ERR_* code;
 *     nothing throws, nothing prints;
 *   - all array pointers are DEVICE pointers owned by the caller unless the
 *     parameter name starts with `h_` (host); the library never allocates or
 *     frees device memory and keeps no mutable global state, so calls are
 *     re-entrant and may be issued from several host threads / processes,
 *     one or more per GPU;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) and the call returns without synchronising;
 *   - per-ray failures are reported the way the reference does it: integer
 *     status (1 escaped, -1 captured, 0 invalid) and NaN in `final_alpha`
 *     (metrics.py:69, :124-125, :667), never through the return code.
 *
 * There is no CPU fallback: on a machine without a usable CUDA device every
 * compute entry point returns LP_ERR_CUDA.
 */
#ifndef LIGHTPATH_H_
#define LIGHTPATH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LP_ABI_VERSION 2

/* ---- return codes ------------------------------------------------------- */
#define LP_OK               0
#define LP_ERR_INVALID_ARG (-1)   /* null pointer, negative size, bad enum      */
#define LP_ERR_CUDA        (-2)   /* a CUDA runtime call / launch failed        */
#define LP_ERR_UNSUPPORTED (-3)   /* legal request this build cannot serve      */

/* ---- flags for the tracer entry points ---------------------------------- */
#define LP_TRACE_STRICT     0u    /* default: every fp64 op separately rounded, in the
                                     reference's order (bit-identical u, w, phi)   */
#define LP_TRACE_FUSED      1u    /* the same RK4 step in its second-order form, in the scaled
                                     variable 3Mu, FMA-contracted: 14 instead of 34 FP64
                                     instructions (ulp-level different trajectories)      */
#define LP_TRACE_REPACK     2u    /* lp_render_frame: opt-in lane re-packing schedule (persistent
                                     warps; a lane whose ray has left the integration band hands it to
                                     the warp's result queue and takes the next prepared ray, so
                                     captured / escaped rays stop wasting lanes).  Same frame bit for
                                     bit.  NOT the default: measured on B200 it is slower than the
                                     one-ray-per-thread kernel on every frame tried, including the
                                     divergent ones (profiles/r2b_repack_perf_refill_sweep.log) — the
                                     default keeps divergence down with 8x4-pixel warp tiles instead */
#define LP_TRACE_HYBRID     4u    /* FMA-contracted loop for rays that sweep at most
                                     11.5 + ln max(final_alpha, 1e-3) rad inside r < 6M
                                     (where rounding differences can grow: csrc/lp_trace.cu),
                                     strict re-trace of the few longer ones
                                     (near-critical rays, where rounding differences are
                                     amplified): same classification and winding as
                                     LP_TRACE_STRICT, final_alpha within 1e-9 relative  */

#define LP_RENDER_STAGED_STORES 8u /* lp_render_frame, float32 RGB: stage each warp's 32 pixels in
                                     shared memory and store them as 16-byte vectors (full
                                     sectors).  For tiles that live in a PEER GPU's memory
                                     (NVLink): ~1 % slower than plain stores into local HBM  */

#define LP_RENDER_OUT_FRAME_ROWS 16u /* lp_render_frame_bands: `out` addresses frame row `row0` of a
                                     FULL frame (row pitch = width*channels elements) and every pixel
                                     is stored at its frame row, instead of into a compact tile       */

#define LP_RENDER_ROW_RUNS 32u     /* lp_render_frame_bands, 8-bit tiles: 32 x 1 warp tiles (96-byte runs, 16-byte
                                     stores into whole 32-byte sectors) instead of 8 x 4 (24-byte runs, 8-byte
                                     stores).  For frames assembled in ONE GPU's memory by many peers: at 8 GPUs
                                     the root's NVLink ingress takes ~200 GB/s of 8-byte partial-sector stores
                                     but > 600 GB/s of full sectors; with up to four writers the 8 x 4 tile's
                                     better lane efficiency wins (3-4 %).  Same pixels either way             */

/* ---- ray status codes (metrics.py:69, :125) ------------------------------ */
#define LP_RAY_ESCAPED    1
#define LP_RAY_CAPTURED (-1)
#define LP_RAY_INVALID    0

/* ---- source image element types for lp_remap ----------------------------- */
#define LP_DTYPE_U8   0
#define LP_DTYPE_F32  1
#define LP_DTYPE_F64  2
#define LP_DTYPE_U8_UNIT 3        /* uint8 storage standing for float32 value/255: the image as
                                     image_lens.main handles it end to end (imread uint8 ->
                                     float32/255, image_lens.py:448-450; imsave float -> 8 bit,
                                     :510).  Output bytes = trunc(255 * v) of the float32
                                     pipeline's result v (gathered pixels are the source byte:
                                     trunc(255 * (k/255.0f)) == k for every k)              */

/* ---- sampling modes for lp_remap ------------------------------------------ */
#define LP_SAMPLE_NEAREST   0     /* np.rint nearest neighbour: what the reference does
                                     (image_lens.py:367-375)                         */
#define LP_SAMPLE_BILINEAR  1     /* opt-in extension (not in the reference)          */

/* Pinhole camera + black-hole screen frame of one image.
 * Replaces the per-call recomputation at image_lens.py:138-143 / :304-308:
 *   fx = (W/2)/tan(hfov/2), fy = (H/2)/tan(vfov/2); (d, e_x, e_y) = _psi_frame(psi)
 *   (image_lens.py:38-61).  Coordinates are (y, x); +x right, +y down, +z forward. */
typedef struct lp_camera {
    int32_t height, width;      /* full frame size in pixels                     */
    double  fx, fy;             /* focal lengths in pixels                        */
    double  d[3];               /* unit vector camera -> black hole               */
    double  e_x[3], e_y[3];     /* tangent basis around d                          */
} lp_camera;

/* Per-frame reductions ("kernel 3"): the counts the reference
 * obtains with np.count_nonzero / np.any over boolean masks (image_lens.py:319-337,
 * :178) plus the step totals the roofline numerator needs.  Lives in DEVICE memory,
 * zero-initialised by lp_frame_stats_reset, accumulated by the tracer kernels. */
typedef struct lp_frame_stats {
    uint64_t n_rays;
    uint64_t n_escaped;          /* status == 1 (finite final_alpha)                */
    uint64_t n_captured;         /* status == -1                                    */
    uint64_t n_invalid;          /* status == 0                                     */
    uint64_t n_winding;          /* escaped and float32(final_alpha) > float32(pi/2)
                                    (image_lens.py:322, NEP-50 float32 compare)     */
    uint64_t sum_steps;          /* RK4 steps actually integrated                    */
    uint64_t sum_warp_steps;     /* sum over warps of 32 * (steps the warp ran):
                                    lane efficiency = sum_steps / sum_warp_steps     */
    uint32_t max_steps;
    uint32_t max_winding;        /* max n_half_orbits over all rays (clipped to u16) */
    double   min_final_alpha;    /* over escaped rays; +inf / 0.0 when there are none */
    double   max_final_alpha;
} lp_frame_stats;

/* ---- library / device ---------------------------------------------------- */

int         lp_abi_version(void);
const char *lp_error_string(int code);
/* Number of visible CUDA devices (0 without a GPU); never fails. */
int         lp_device_count(void);
/* SM count and SM clock (kHz) of the current device. */
int         lp_device_props(int32_t *sm_count, int32_t *clock_khz);

/* Host-side helper: fill an lp_camera exactly as image_lens.py:38-61, :138-139 do.
 * psi_y = pitch up, psi_x = yaw right (radians).  Pure host arithmetic (libm). */
int lp_camera_init(int32_t height, int32_t width, double hfov, double vfov,
                   double psi_y, double psi_x, lp_camera *h_cam);

/* Diagnostics: whether the kernels may form the pixel coordinates (i - n/2)/f
 * (image_lens.py:141-142) with a multiplication by RN(1/f) and two fused corrections
 * instead of a division.  The library checks, per call, that this is bit-identical to
 * the division for every column / row of the frame and otherwise divides. */
int lp_camera_fast_coords(const lp_camera *h_cam, int32_t *fast_x, int32_t *fast_y);

/* ---- kernel (1a): Schwarzschild Binet-equation RK4 tracer ---------------- */

/* Replaces Schwarzschild.trace_rays_batch / _trace_rays_batch_schwarzschild
 * (metrics.py:831-833, :661-668):  for every i in [0, n)
 *     out_fa[i] = final_alpha if the ray escapes else NaN
 *     out_w[i]  = n_half_orbits (written for every status, metrics.py:668)
 * with _schwarzschild_trace_ray_numba (metrics.py:120-145) semantics, fixed step
 * h_max up to phi_max (the reference hard-codes 50.0 / 0.05 at metrics.py:833).
 * out_status (int8: 1/-1/0) and out_steps (int32) are optional (NULL to skip).
 * stats (device, optional) is accumulated into. */
int lp_schw_trace_batch_f64(const double *alphas, int64_t n,
                            double M, double R_S, double r_obs,
                            double phi_max, double h_max,
                            double *out_fa, int64_t *out_w,
                            int8_t *out_status, int32_t *out_steps,
                            lp_frame_stats *stats, uint32_t flags, void *stream);

/* Replaces image_lens.precompute_final_alpha_lookup (image_lens.py:155-178) for a
 * Schwarzschild metric: float32 alpha table in (widened to fp64 per ray,
 * image_lens.py:157), float32 final_alpha (NaN unless escaped) and uint16
 * clipped winding out (image_lens.py:176-177).  n = H*W; phi_max/h_max as above. */
int lp_schw_trace_alpha32(const float *alpha32, int64_t n,
                          double M, double R_S, double r_obs,
                          double phi_max, double h_max,
                          float *out_fa32, uint16_t *out_w16,
                          int8_t *out_status, int32_t *out_steps,
                          lp_frame_stats *stats, uint32_t flags, void *stream);

/* Fused build_alpha_lookup + precompute_final_alpha_lookup for rows
 * [row0, row0+rows) of the frame described by cam (image_lens.py:133-178 in one
 * launch; row tiles are what multi-GPU sharding hands each rank).  Output
 * pointers address the FIRST ROW OF THE TILE (rows*width elements each).
 * out_alpha32 is optional.  The alpha value is rounded to float32 and widened
 * again before tracing, exactly as the reference's two-stage pipeline does. */
int lp_schw_trace_frame(const lp_camera *h_cam, int32_t row0, int32_t rows,
                        double M, double R_S, double r_obs,
                        double phi_max, double h_max,
                        float *out_alpha32, float *out_fa32, uint16_t *out_w16,
                        int8_t *out_status, int32_t *out_steps,
                        lp_frame_stats *stats, uint32_t flags, void *stream);

/* Replaces image_lens.build_alpha_lookup (image_lens.py:133-152) for rows
 * [row0, row0+rows).  decimals < 0 means "no rounding" (decimals=None). */
int lp_build_alpha_lookup(const lp_camera *h_cam, int32_t row0, int32_t rows,
                          int32_t decimals, float *out_alpha32, void *stream);

/* ---- kernel (2): deflection -> background remap --------------------------- */

/* Replaces image_lens.render_lensed_image (image_lens.py:296-397) for output rows
 * [row0, row0+rows).  src is the FULL source image [H, W, channels] (channels = 1
 * for a 2-D image), element type src_dtype; out addresses the first row of the
 * tile and has the same element type / channel count.  fa32 / w16 address the
 * tile's first row too; w16 may be NULL (winding_lookup=None).  Captured/invalid
 * pixels are written as 0, winding pixels as WINDING_COLORS[clip(w,0,4)]
 * (luma for channels == 1), out-of-frame samples as magenta.  channels == 4 with
 * winding pixels present raises in the reference (shape mismatch): here the
 * colour is written to the first 3 channels and the 4th is left 0. */
int lp_remap(const void *src, int32_t src_dtype, int32_t channels,
             const lp_camera *h_cam, const float *fa32, const uint16_t *w16,
             int32_t render_loop_around, int32_t sampling,
             int32_t row0, int32_t rows, void *out, void *stream);

/* Fully fused frame: pixel -> alpha(f32) -> trace -> fa(f32), w(u16) -> remap, one
 * launch, nothing but the finished pixels (and optional lookups) written. */
int lp_render_frame(const void *src, int32_t src_dtype, int32_t channels,
                    const lp_camera *h_cam, int32_t row0, int32_t rows,
                    double M, double R_S, double r_obs, double phi_max, double h_max,
                    int32_t render_loop_around, int32_t sampling,
                    void *out, float *out_fa32, uint16_t *out_w16,
                    lp_frame_stats *stats, uint32_t flags, void *stream);

/* lp_render_frame over an INTERLEAVED set of rows (multi-GPU load balance: the black hole
 * sits in the centre rows, so contiguous row tiles are unevenly expensive; rank g of G renders
 * bands g, g+G, g+2G, ... of band_rows rows each).  Tile-local row r is frame row
 *     row0 + (r / band_rows) * band_stride + (r % band_rows),        r in [0, rows)
 * band_rows == 0 means contiguous rows (then this IS lp_render_frame).  out / out_fa32 /
 * out_w16 are compact tiles of `rows` rows unless flags has LP_RENDER_OUT_FRAME_ROWS, in which
 * case `out` (only) is frame-addressed: it points at frame row row0 of a full frame and must
 * reach the tile's last frame row.  Replaces the same reference lines as lp_render_frame
 * (image_lens.py:133-178, :296-397); the reference has no sharding. */
int lp_render_frame_bands(const void *src, int32_t src_dtype, int32_t channels,
                          const lp_camera *h_cam, int32_t row0, int32_t rows,
                          int32_t band_rows, int32_t band_stride,
                          double M, double R_S, double r_obs, double phi_max, double h_max,
                          int32_t render_loop_around, int32_t sampling,
                          void *out, float *out_fa32, uint16_t *out_w16,
                          lp_frame_stats *stats, uint32_t flags, void *stream);

/* ---- peer-memory completion flags (multi-GPU frame assembly) -------------
 * Row tiles are stored by every rank's render kernel straight into the root GPU's frame
 * through NVLink peer mappings; these two stream-ordered calls order "all tiles of frame e
 * have landed" and "the root has consumed frame e" without a collective.  Flags are uint64
 * epoch counters in (peer-mapped) device memory, only ever increased.
 *   lp_peer_signal: after all prior work of `stream` has completed and its writes are visible
 *                   system-wide, store `value` (release, system scope) to each of the n_flags
 *                   device addresses in the HOST array h_flags (n_flags <= 16).
 *   lp_peer_wait:   block `stream` until each of flags[0..n_flags) (local device memory) is
 *                   >= value (acquire, system scope).  Gives up after timeout_ms (0 = 10 s)
 *                   and then writes 1 to *timed_out (device int32, optional) so that a dead
 *                   peer cannot hang the GPU. */
int lp_peer_signal(uint64_t *const *h_flags, int32_t n_flags, uint64_t value, void *stream);
int lp_peer_wait(const uint64_t *flags, int32_t n_flags, uint64_t value,
                 uint32_t timeout_ms, int32_t *timed_out, void *stream);

/* ---- kernel (3): shadow classification and frame reductions --------------- */

/* Replaces the pixel loop of black_hole_shadow.main (black_hole_shadow.py:30-37):
 * image[i*height + j] = 0.0 if arccos(cos(ax_i)*cos(ay_j)) < alpha_crit else 1.0,
 * ax_i = arctan(((i - width/2)/(width/2)) * tan(fov/2)) (black_hole_shadow.py:7-9).
 * image is float64 [width][height] (indexed [x][y] like the reference).
 * n_shadow (device uint64, optional) receives the number of 0.0 pixels. */
int lp_shadow_classify(int32_t width, int32_t height, double fov, double alpha_crit,
                       double *image, uint64_t *n_shadow, void *stream);

int lp_frame_stats_reset(lp_frame_stats *stats, void *stream);
/* Stand-alone reduction over finished lookups (np.count_nonzero / np.any of
 * image_lens.py:319-337): status and steps are optional. */
int lp_frame_stats_reduce(const float *fa32, const uint16_t *w16,
                          const int8_t *status, const int32_t *steps, int64_t n,
                          lp_frame_stats *stats, void *stream);

/* ---- kernel (1b): generic 8-D Hamiltonian tracer (scipy RK45 semantics) ---- */

/* Replaces geodesic_tracer.trace_ray / integrate_geodesic (geodesic_tracer.py:22-82)
 * for a Schwarzschild metric, batched over viewing angles: initial conditions as
 * metrics.py:794-809, RHS as metrics.py:763-790, Dormand-Prince 5(4) with scipy's
 * step controller (rtol, atol, max_step, first-step selection — the reference calls
 * solve_ivp(method='RK45', max_step=1.0, rtol=1e-8, atol=1e-10, dense_output=True),
 * geodesic_tracer.py:57-67), terminal events at r_stop_inner (falling) / r_stop_outer
 * (rising) located with brentq on the quartic dense output, outcome = captured if
 * r_final <= 1.1*r_stop_inner (geodesic_tracer.py:69-70).
 *   out_state  : [n][8] final state (t, r, theta, phi, p_t, p_r, p_theta, p_phi)
 *                = OdeResult.y[:, -1]; NaN for invalid rays
 *   out_lambda : [n] final affine parameter = OdeResult.t[-1]
 *   out_outcome: [n] 1 escaped / -1 captured / 0 invalid (initial_conditions -> None,
 *                geodesic_tracer.py:79-81)
 *   out_nsteps : [n][2] optional: len(OdeResult.t) (1 + accepted steps), OdeResult.nfev
 *   out_status : [n] optional: OdeResult.status (1 event, 0 reached lambda_max,
 *                -1 step size too small), -2 for invalid rays
 * r_stop_inner / r_stop_outer <= 0 select the reference's defaults
 * (capture_radius() = 1.01 R_S, 2*r_obs). */
int lp_schw_rk45_trace_batch(const double *alphas, int64_t n,
                             double M, double R_S, double r_obs,
                             double lambda_max, double rtol, double atol, double max_step,
                             double r_stop_inner, double r_stop_outer,
                             double *out_state, double *out_lambda,
                             int8_t *out_outcome, int32_t *out_nsteps, int8_t *out_status,
                             void *stream);

/* Same, also recording every accepted point (what OdeResult.t / .y hold and
 * main.py:30-31 / plot_trajectories, geodesic_tracer.py:89-142, read):
 * traj is [n][max_points][9] rows (lambda, state[8]); n_points[i] receives
 * len(OdeResult.t) of ray i (rows beyond max_points are dropped, the count is not
 * clipped).  The last row is the event point. */
int lp_schw_rk45_trace_paths(const double *alphas, int64_t n,
                             double M, double R_S, double r_obs,
                             double lambda_max, double rtol, double atol, double max_step,
                             double r_stop_inner, double r_stop_outer,
                             double *traj, int32_t max_points, int32_t *n_points,
                             double *out_state, double *out_lambda,
                             int8_t *out_outcome, int32_t *out_nsteps, int8_t *out_status,
                             void *stream);

/* integrate_geodesic(metric, state0, lambda_max, r_stop_inner, r_stop_outer)
 * (geodesic_tracer.py:22-71) for explicit initial states: state0 is [n][8]
 * (t, r, theta, phi, p_t, p_r, p_theta, p_phi), not necessarily equatorial or null.
 * traj / n_points may be NULL (no trajectory).  r_stop_outer <= 0 -> 2*state0[1]
 * per ray (geodesic_tracer.py:44-45). */
int lp_schw_rk45_integrate_paths(const double *state0, int64_t n,
                                 double M, double R_S,
                                 double lambda_max, double rtol, double atol, double max_step,
                                 double r_stop_inner, double r_stop_outer,
                                 double *traj, int32_t max_points, int32_t *n_points,
                                 double *out_state, double *out_lambda,
                                 int8_t *out_outcome, int32_t *out_nsteps, int8_t *out_status,
                                 void *stream);

/* integrate_geodesic for a Kerr metric (Kerr.geodesic_equations, metrics.py:946-1029): same
 * stepper, events and outputs as lp_schw_rk45_integrate_paths.  r_stop_inner <= 0 selects
 * capture_radius() = 1.01 r_plus (metrics.py:861-862). */
int lp_kerr_rk45_integrate_paths(const double *state0, int64_t n,
                                 double M, double a, double r_plus,
                                 double lambda_max, double rtol, double atol, double max_step,
                                 double r_stop_inner, double r_stop_outer,
                                 double *traj, int32_t max_points, int32_t *n_points,
                                 double *out_state, double *out_lambda,
                                 int8_t *out_outcome, int32_t *out_nsteps, int8_t *out_status,
                                 void *stream);

/* Trajectories WITH their dense output — what solve_ivp(dense_output=True) keeps for
 * OdeResult.sol (geodesic_tracer.py:57-67; scipy rk.py:178-180, :715-737).  As the *_paths entry
 * points above plus dense[n][max_points][25]: row k >= 1 holds (h, Q[6][4]) of the accepted step
 * that ended in point k, Q = K^T P for the six moving components (t, r, theta, phi, p_r, p_theta;
 * p_t and p_phi are constants of the motion, their rows of Q are exactly zero):
 *     sol(t) = y_old + h * Q @ (x, x^2, x^3, x^4),  x = (t - t_old) / h.
 * Exactly one of alphas / state0 is given.  metric 0 = Schwarzschild (a ignored, R_S_or_r_plus =
 * R_S), 1 = Kerr (R_S_or_r_plus = r_plus; state0 only). */
int lp_rk45_paths_dense(int32_t metric, const double *alphas, const double *state0, int64_t n,
                        double M, double a, double R_S_or_r_plus, double r_obs,
                        double lambda_max, double rtol, double atol, double max_step,
                        double r_stop_inner, double r_stop_outer,
                        double *traj, int32_t max_points, int32_t *n_points, double *dense,
                        double *out_state, double *out_lambda,
                        int8_t *out_outcome, int32_t *out_nsteps, int8_t *out_status,
                        void *stream);

/* ---- Kerr tracer (next row after the Schwarzschild path) ------ */

/* Replaces Kerr.trace_rays_batch / _trace_rays_batch_kerr (metrics.py:1128-1132,
 * :671-679): for every i, _kerr_trace_ray_numba (metrics.py:419-567; Dormand-Prince
 * 4(5) on the 5-D reduced Hamiltonian state, atol/rtol = 1e-8/1e-6, or 1e-10/1e-8
 * where axis_refines[i] != 0) from viewing angle alphas[i] and screen angle thetas[i]
 * for an observer at (r_obs, theta_obs):
 *     out_fa[i] = final_alpha if the ray escapes else NaN,  out_w[i] = n_half_orbits.
 * r_plus = M + sqrt(M^2 - a^2) (metrics.py:852); lambda_max = max(5000, 6 r_obs) in the
 * reference's calls (metrics.py:1120, :1131).  axis_refines (uint8), out_status
 * (1 / -1 / 0) and out_steps ([n][2]: accepted steps, attempts) are optional. */
int lp_kerr_trace_batch_f64(const double *alphas, const double *thetas, const uint8_t *axis_refines,
                            int64_t n, double M, double a, double r_plus, double r_obs,
                            double theta_obs, double lambda_max,
                            double *out_fa, int64_t *out_w, int8_t *out_status, int32_t *out_steps,
                            void *stream);

/* Replaces the tracing part of image_lens.precompute_final_alpha_lookup_2d
 * (image_lens.py:185-280) for rows [row0, row0+rows): alpha from the float32 table
 * (tile-relative, rows*width entries), the per-pixel screen angle theta_pixel
 * (image_lens.py:194-208) evaluated on the device, axis_refine per column
 * (image_lens.py:210-216; uint8[width], optional), float32 final_alpha and clipped
 * uint16 winding out (image_lens.py:261-262).  The top/bottom mirror of
 * image_lens.py:218-220, :272-276 is the caller's (trace the top rows, flip). */
int lp_kerr_trace_alpha32(const float *alpha32, const lp_camera *h_cam, int32_t row0, int32_t rows,
                          const uint8_t *axis_refine_cols,
                          double M, double a, double r_plus, double r_obs, double theta_obs,
                          double lambda_max, float *out_fa32, uint16_t *out_w16,
                          int8_t *out_status, int32_t *out_steps, void *stream);

/* ---- introspection ---------------------------------------------------------- */

/* The LP_TRACE_HYBRID rule of one configuration (host arithmetic only, no device needed; csrc/lp_trace.cu):
 * a ray of the FMA loop is re-traced strictly when steps > *steps_all, or when steps > *steps_none and
 * max(final_alpha, 1e-3) < exp(steps * h_max - *exp_offset).  *phi_outside = the angle a critical ray sweeps
 * outside r = 6M between r_obs and 2 r_obs (where rounding differences cannot grow).  Any output may be NULL. */
int lp_hybrid_retrace_rule(double M, double r_obs, double h_max,
                           int32_t *steps_all, int32_t *steps_none, double *exp_offset, double *phi_outside);

/* ---- measurement helpers --------------------------------------------------- */

/* FP64 pipe micro-benchmark: every thread runs `iters` rounds of 8 independent
 * dependent-chain DFMAs (16*iters flop per thread).  Used by bench.py to MEASURE the
 * FP64 roofline denominator on the box (it is not in MEASURED_PEAKS.json).
 * sink: device double[blocks*threads]. */
int lp_bench_dfma(int32_t blocks, int32_t threads, int32_t iters, double *sink, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTPATH_H_ */
