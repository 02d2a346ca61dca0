// lp_torch.cpp — thin PyTorch C++ extension over the C ABI (include/lightpath.h).
// PyTorch is plumbing here: it owns device memory and the current stream; every compute
// call below only validates tensors and forwards raw pointers to liblightpath.so.
#include <torch/extension.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGuard.h>

#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "lightpath.h"

namespace {

using torch::Tensor;
using OptTensor = std::optional<Tensor>;

void check(int rc, const char *what)
{
    if (rc != LP_OK)
        throw std::runtime_error(std::string(what) + ": " + lp_error_string(rc) + " (code " + std::to_string(rc) + ")");
}

void *ptr(const Tensor &t, c10::ScalarType dt, const char *name, int64_t min_numel)
{
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.scalar_type() == dt, name, " has dtype ", t.scalar_type(), ", expected ", dt);
    TORCH_CHECK(t.numel() >= min_numel, name, " has ", t.numel(), " elements, need ", min_numel);
    return t.data_ptr();
}

void *optptr(const OptTensor &t, c10::ScalarType dt, const char *name, int64_t min_numel)
{
    if (!t.has_value() || !t->defined()) return nullptr;
    return ptr(*t, dt, name, min_numel);
}

void *stream_of(const Tensor &t)
{
    return (void *)c10::cuda::getCurrentCUDAStream(t.get_device()).stream();
}

lp_camera make_cam(const std::vector<double> &v)
{
    TORCH_CHECK(v.size() == 13, "camera must be (H, W, fx, fy, d[3], e_x[3], e_y[3])");
    lp_camera c;
    c.height = (int32_t)v[0];
    c.width = (int32_t)v[1];
    c.fx = v[2];
    c.fy = v[3];
    for (int i = 0; i < 3; ++i) { c.d[i] = v[4 + i]; c.e_x[i] = v[7 + i]; c.e_y[i] = v[10 + i]; }
    return c;
}

int dtype_code(const Tensor &t)
{
    switch (t.scalar_type()) {
    case c10::ScalarType::Byte: return LP_DTYPE_U8;
    case c10::ScalarType::Float: return LP_DTYPE_F32;
    case c10::ScalarType::Double: return LP_DTYPE_F64;
    default: TORCH_CHECK(false, "source image dtype must be uint8, float32 or float64, got ", t.scalar_type());
    }
    return -1;
}

constexpr int64_t kStatsWords = (int64_t)(sizeof(lp_frame_stats) / 8);

lp_frame_stats *stats_ptr(const OptTensor &s)
{
    return (lp_frame_stats *)optptr(s, c10::ScalarType::Long, "stats", kStatsWords);
}

void trace_batch_f64(const Tensor &alphas, double M, double R_S, double r_obs, double phi_max, double h_max,
                     Tensor out_fa, Tensor out_w, OptTensor status, OptTensor steps, OptTensor stats, int64_t flags)
{
    const int64_t n = alphas.numel();
    c10::cuda::CUDAGuard g(alphas.device());
    check(lp_schw_trace_batch_f64((const double *)ptr(alphas, c10::ScalarType::Double, "alphas", 0), n, M, R_S, r_obs,
                                  phi_max, h_max, (double *)ptr(out_fa, c10::ScalarType::Double, "out_fa", n),
                                  (int64_t *)ptr(out_w, c10::ScalarType::Long, "out_w", n),
                                  (int8_t *)optptr(status, c10::ScalarType::Char, "status", n),
                                  (int32_t *)optptr(steps, c10::ScalarType::Int, "steps", n), stats_ptr(stats),
                                  (uint32_t)flags, stream_of(alphas)),
          "lp_schw_trace_batch_f64");
}

void trace_alpha32(const Tensor &alpha32, double M, double R_S, double r_obs, double phi_max, double h_max,
                   Tensor out_fa32, Tensor out_w16, OptTensor status, OptTensor steps, OptTensor stats, int64_t flags)
{
    const int64_t n = alpha32.numel();
    c10::cuda::CUDAGuard g(alpha32.device());
    check(lp_schw_trace_alpha32((const float *)ptr(alpha32, c10::ScalarType::Float, "alpha32", 0), n, M, R_S, r_obs,
                                phi_max, h_max, (float *)ptr(out_fa32, c10::ScalarType::Float, "out_fa32", n),
                                (uint16_t *)ptr(out_w16, c10::ScalarType::UInt16, "out_w16", n),
                                (int8_t *)optptr(status, c10::ScalarType::Char, "status", n),
                                (int32_t *)optptr(steps, c10::ScalarType::Int, "steps", n), stats_ptr(stats),
                                (uint32_t)flags, stream_of(alpha32)),
          "lp_schw_trace_alpha32");
}

void trace_frame(const std::vector<double> &camv, int64_t row0, int64_t rows, double M, double R_S, double r_obs,
                 double phi_max, double h_max, OptTensor alpha32, Tensor out_fa32, Tensor out_w16,
                 OptTensor status, OptTensor steps, OptTensor stats, int64_t flags)
{
    lp_camera cam = make_cam(camv);
    const int64_t n = rows * (int64_t)cam.width;
    c10::cuda::CUDAGuard g(out_fa32.device());
    check(lp_schw_trace_frame(&cam, (int32_t)row0, (int32_t)rows, M, R_S, r_obs, phi_max, h_max,
                              (float *)optptr(alpha32, c10::ScalarType::Float, "alpha32", n),
                              (float *)ptr(out_fa32, c10::ScalarType::Float, "out_fa32", n),
                              (uint16_t *)ptr(out_w16, c10::ScalarType::UInt16, "out_w16", n),
                              (int8_t *)optptr(status, c10::ScalarType::Char, "status", n),
                              (int32_t *)optptr(steps, c10::ScalarType::Int, "steps", n), stats_ptr(stats),
                              (uint32_t)flags, stream_of(out_fa32)),
          "lp_schw_trace_frame");
}

void build_alpha_lookup(const std::vector<double> &camv, int64_t row0, int64_t rows, int64_t decimals, Tensor out)
{
    lp_camera cam = make_cam(camv);
    c10::cuda::CUDAGuard g(out.device());
    check(lp_build_alpha_lookup(&cam, (int32_t)row0, (int32_t)rows, (int32_t)decimals,
                                (float *)ptr(out, c10::ScalarType::Float, "out", rows * (int64_t)cam.width),
                                stream_of(out)),
          "lp_build_alpha_lookup");
}

void remap(const Tensor &src, int64_t channels, const std::vector<double> &camv, const Tensor &fa32, OptTensor w16,
           bool loop_around, int64_t sampling, int64_t row0, int64_t rows, Tensor out, bool unit_u8)
{
    lp_camera cam = make_cam(camv);
    const int64_t n = rows * (int64_t)cam.width;
    int code = dtype_code(src);
    if (unit_u8) { TORCH_CHECK(code == LP_DTYPE_U8, "unit_u8 needs a uint8 image"); code = LP_DTYPE_U8_UNIT; }
    TORCH_CHECK(out.scalar_type() == src.scalar_type(), "out dtype must equal source dtype");
    c10::cuda::CUDAGuard g(src.device());
    check(lp_remap(ptr(src, src.scalar_type(), "src", (int64_t)cam.height * cam.width * channels), code,
                   (int32_t)channels, &cam, (const float *)ptr(fa32, c10::ScalarType::Float, "fa32", n),
                   (const uint16_t *)optptr(w16, c10::ScalarType::UInt16, "w16", n), loop_around ? 1 : 0,
                   (int32_t)sampling, (int32_t)row0, (int32_t)rows, ptr(out, out.scalar_type(), "out", n * channels),
                   stream_of(src)),
          "lp_remap");
}

void render_frame(const Tensor &src, int64_t channels, const std::vector<double> &camv, int64_t row0, int64_t rows,
                  double M, double R_S, double r_obs, double phi_max, double h_max, bool loop_around,
                  int64_t sampling, Tensor out, OptTensor fa32, OptTensor w16, OptTensor stats, int64_t flags,
                  bool unit_u8, int64_t band_rows, int64_t band_stride)
{
    lp_camera cam = make_cam(camv);
    const int64_t n = rows * (int64_t)cam.width;
    int code = dtype_code(src);
    if (unit_u8) { TORCH_CHECK(code == LP_DTYPE_U8, "unit_u8 needs a uint8 image"); code = LP_DTYPE_U8_UNIT; }
    TORCH_CHECK(out.scalar_type() == src.scalar_type(), "out dtype must equal source dtype");
    // a frame-addressed tile of interleaved bands must reach its last frame row
    int64_t out_rows = rows;
    if ((flags & LP_RENDER_OUT_FRAME_ROWS) && band_rows > 0 && rows > 0)
        out_rows = ((rows - 1) / band_rows) * band_stride + ((rows - 1) % band_rows) + 1;
    c10::cuda::CUDAGuard g(src.device());
    check(lp_render_frame_bands(ptr(src, src.scalar_type(), "src", (int64_t)cam.height * cam.width * channels), code,
                                (int32_t)channels, &cam, (int32_t)row0, (int32_t)rows, (int32_t)band_rows,
                                (int32_t)band_stride, M, R_S, r_obs, phi_max, h_max,
                                loop_around ? 1 : 0, (int32_t)sampling,
                                ptr(out, out.scalar_type(), "out", out_rows * (int64_t)cam.width * channels),
                                (float *)optptr(fa32, c10::ScalarType::Float, "fa32", n),
                                (uint16_t *)optptr(w16, c10::ScalarType::UInt16, "w16", n), stats_ptr(stats),
                                (uint32_t)flags, stream_of(src)),
          "lp_render_frame_bands");
}

// flag_ptrs: raw device addresses (possibly peer-mapped) as integers; stream = current stream of `device_of`
void peer_signal(const std::vector<int64_t> &flag_ptrs, int64_t value, const Tensor &device_of)
{
    std::vector<uint64_t *> p;
    for (int64_t v : flag_ptrs) p.push_back((uint64_t *)(uintptr_t)v);
    c10::cuda::CUDAGuard g(device_of.device());
    check(lp_peer_signal(p.data(), (int32_t)p.size(), (uint64_t)value, stream_of(device_of)), "lp_peer_signal");
}

void peer_wait(const Tensor &flags, int64_t n_flags, int64_t value, int64_t timeout_ms, OptTensor timed_out)
{
    c10::cuda::CUDAGuard g(flags.device());
    check(lp_peer_wait((const uint64_t *)ptr(flags, c10::ScalarType::Long, "flags", n_flags), (int32_t)n_flags,
                       (uint64_t)value, (uint32_t)timeout_ms,
                       (int32_t *)optptr(timed_out, c10::ScalarType::Int, "timed_out", 1), stream_of(flags)),
          "lp_peer_wait");
}

void shadow_classify(int64_t width, int64_t height, double fov, double alpha_crit, Tensor image, OptTensor n_shadow)
{
    c10::cuda::CUDAGuard g(image.device());
    check(lp_shadow_classify((int32_t)width, (int32_t)height, fov, alpha_crit,
                             (double *)ptr(image, c10::ScalarType::Double, "image", width * height),
                             (uint64_t *)optptr(n_shadow, c10::ScalarType::Long, "n_shadow", 1), stream_of(image)),
          "lp_shadow_classify");
}

void stats_reset(Tensor stats)
{
    c10::cuda::CUDAGuard g(stats.device());
    check(lp_frame_stats_reset(stats_ptr(stats), stream_of(stats)), "lp_frame_stats_reset");
}

void stats_reduce(const Tensor &fa32, OptTensor w16, OptTensor status, OptTensor steps, Tensor stats)
{
    const int64_t n = fa32.numel();
    c10::cuda::CUDAGuard g(fa32.device());
    check(lp_frame_stats_reduce((const float *)ptr(fa32, c10::ScalarType::Float, "fa32", 0),
                                (const uint16_t *)optptr(w16, c10::ScalarType::UInt16, "w16", n),
                                (const int8_t *)optptr(status, c10::ScalarType::Char, "status", n),
                                (const int32_t *)optptr(steps, c10::ScalarType::Int, "steps", n), n, stats_ptr(stats),
                                stream_of(fa32)),
          "lp_frame_stats_reduce");
}

void rk45_trace_batch(const Tensor &alphas, double M, double R_S, double r_obs, double lambda_max, double rtol,
                      double atol, double max_step, double r_in, double r_out, Tensor out_state, Tensor out_lambda,
                      Tensor out_outcome, OptTensor out_nsteps, OptTensor out_status)
{
    const int64_t n = alphas.numel();
    c10::cuda::CUDAGuard g(alphas.device());
    check(lp_schw_rk45_trace_batch((const double *)ptr(alphas, c10::ScalarType::Double, "alphas", 0), n, M, R_S, r_obs,
                                   lambda_max, rtol, atol, max_step, r_in, r_out,
                                   (double *)ptr(out_state, c10::ScalarType::Double, "out_state", 8 * n),
                                   (double *)ptr(out_lambda, c10::ScalarType::Double, "out_lambda", n),
                                   (int8_t *)ptr(out_outcome, c10::ScalarType::Char, "out_outcome", n),
                                   (int32_t *)optptr(out_nsteps, c10::ScalarType::Int, "out_nsteps", 2 * n),
                                   (int8_t *)optptr(out_status, c10::ScalarType::Char, "out_status", n),
                                   stream_of(alphas)),
          "lp_schw_rk45_trace_batch");
}

void rk45_trace_paths(const Tensor &alphas, double M, double R_S, double r_obs, double lambda_max, double rtol,
                      double atol, double max_step, double r_in, double r_out, Tensor traj, int64_t max_points,
                      Tensor n_points, Tensor out_state, Tensor out_lambda, Tensor out_outcome, Tensor out_nsteps,
                      Tensor out_status)
{
    const int64_t n = alphas.numel();
    c10::cuda::CUDAGuard g(alphas.device());
    check(lp_schw_rk45_trace_paths((const double *)ptr(alphas, c10::ScalarType::Double, "alphas", 0), n, M, R_S, r_obs,
                                   lambda_max, rtol, atol, max_step, r_in, r_out,
                                   (double *)ptr(traj, c10::ScalarType::Double, "traj", n * max_points * 9),
                                   (int32_t)max_points, (int32_t *)ptr(n_points, c10::ScalarType::Int, "n_points", n),
                                   (double *)ptr(out_state, c10::ScalarType::Double, "out_state", 8 * n),
                                   (double *)ptr(out_lambda, c10::ScalarType::Double, "out_lambda", n),
                                   (int8_t *)ptr(out_outcome, c10::ScalarType::Char, "out_outcome", n),
                                   (int32_t *)ptr(out_nsteps, c10::ScalarType::Int, "out_nsteps", 2 * n),
                                   (int8_t *)ptr(out_status, c10::ScalarType::Char, "out_status", n),
                                   stream_of(alphas)),
          "lp_schw_rk45_trace_paths");
}

void rk45_integrate_paths(const Tensor &state0, double M, double R_S, double lambda_max, double rtol, double atol,
                          double max_step, double r_in, double r_out, OptTensor traj, int64_t max_points,
                          OptTensor n_points, Tensor out_state, Tensor out_lambda, Tensor out_outcome,
                          Tensor out_nsteps, Tensor out_status)
{
    const int64_t n = state0.numel() / 8;
    c10::cuda::CUDAGuard g(state0.device());
    check(lp_schw_rk45_integrate_paths((const double *)ptr(state0, c10::ScalarType::Double, "state0", 0), n, M, R_S,
                                       lambda_max, rtol, atol, max_step, r_in, r_out,
                                       (double *)optptr(traj, c10::ScalarType::Double, "traj", n * max_points * 9),
                                       (int32_t)max_points,
                                       (int32_t *)optptr(n_points, c10::ScalarType::Int, "n_points", n),
                                       (double *)ptr(out_state, c10::ScalarType::Double, "out_state", 8 * n),
                                       (double *)ptr(out_lambda, c10::ScalarType::Double, "out_lambda", n),
                                       (int8_t *)ptr(out_outcome, c10::ScalarType::Char, "out_outcome", n),
                                       (int32_t *)ptr(out_nsteps, c10::ScalarType::Int, "out_nsteps", 2 * n),
                                       (int8_t *)ptr(out_status, c10::ScalarType::Char, "out_status", n),
                                       stream_of(state0)),
          "lp_schw_rk45_integrate_paths");
}

void kerr_rk45_integrate_paths(const Tensor &state0, double M, double a, double r_plus, double lambda_max, double rtol,
                               double atol, double max_step, double r_in, double r_out, OptTensor traj,
                               int64_t max_points, OptTensor n_points, Tensor out_state, Tensor out_lambda,
                               Tensor out_outcome, Tensor out_nsteps, Tensor out_status)
{
    const int64_t n = state0.numel() / 8;
    c10::cuda::CUDAGuard g(state0.device());
    check(lp_kerr_rk45_integrate_paths((const double *)ptr(state0, c10::ScalarType::Double, "state0", 0), n, M, a, r_plus,
                                       lambda_max, rtol, atol, max_step, r_in, r_out,
                                       (double *)optptr(traj, c10::ScalarType::Double, "traj", n * max_points * 9),
                                       (int32_t)max_points,
                                       (int32_t *)optptr(n_points, c10::ScalarType::Int, "n_points", n),
                                       (double *)ptr(out_state, c10::ScalarType::Double, "out_state", 8 * n),
                                       (double *)ptr(out_lambda, c10::ScalarType::Double, "out_lambda", n),
                                       (int8_t *)ptr(out_outcome, c10::ScalarType::Char, "out_outcome", n),
                                       (int32_t *)ptr(out_nsteps, c10::ScalarType::Int, "out_nsteps", 2 * n),
                                       (int8_t *)ptr(out_status, c10::ScalarType::Char, "out_status", n),
                                       stream_of(state0)),
          "lp_kerr_rk45_integrate_paths");
}

void rk45_paths_dense(int64_t metric, OptTensor alphas, OptTensor state0, double M, double a, double rs, double r_obs,
                      double lambda_max, double rtol, double atol, double max_step, double r_in, double r_out,
                      Tensor traj, int64_t max_points, Tensor n_points, Tensor dense, Tensor out_state,
                      Tensor out_lambda, Tensor out_outcome, Tensor out_nsteps, Tensor out_status)
{
    const bool by_alpha = alphas.has_value() && alphas->defined();
    const Tensor &in = by_alpha ? *alphas : *state0;
    const int64_t n = by_alpha ? in.numel() : in.numel() / 8;
    c10::cuda::CUDAGuard g(in.device());
    check(lp_rk45_paths_dense((int32_t)metric, (const double *)optptr(alphas, c10::ScalarType::Double, "alphas", 0),
                              (const double *)optptr(state0, c10::ScalarType::Double, "state0", 0), n, M, a, rs, r_obs,
                              lambda_max, rtol, atol, max_step, r_in, r_out,
                              (double *)ptr(traj, c10::ScalarType::Double, "traj", n * max_points * 9),
                              (int32_t)max_points, (int32_t *)ptr(n_points, c10::ScalarType::Int, "n_points", n),
                              (double *)ptr(dense, c10::ScalarType::Double, "dense", n * max_points * 25),
                              (double *)ptr(out_state, c10::ScalarType::Double, "out_state", 8 * n),
                              (double *)ptr(out_lambda, c10::ScalarType::Double, "out_lambda", n),
                              (int8_t *)ptr(out_outcome, c10::ScalarType::Char, "out_outcome", n),
                              (int32_t *)ptr(out_nsteps, c10::ScalarType::Int, "out_nsteps", 2 * n),
                              (int8_t *)ptr(out_status, c10::ScalarType::Char, "out_status", n), stream_of(in)),
          "lp_rk45_paths_dense");
}

void kerr_trace_batch(const Tensor &alphas, const Tensor &thetas, OptTensor refine, double M, double a, double r_plus,
                      double r_obs, double theta_obs, double lambda_max, Tensor out_fa, Tensor out_w,
                      OptTensor status, OptTensor steps)
{
    const int64_t n = alphas.numel();
    c10::cuda::CUDAGuard g(alphas.device());
    check(lp_kerr_trace_batch_f64((const double *)ptr(alphas, c10::ScalarType::Double, "alphas", 0),
                                  (const double *)ptr(thetas, c10::ScalarType::Double, "thetas", n),
                                  (const uint8_t *)optptr(refine, c10::ScalarType::Byte, "axis_refines", n), n, M, a,
                                  r_plus, r_obs, theta_obs, lambda_max,
                                  (double *)ptr(out_fa, c10::ScalarType::Double, "out_fa", n),
                                  (int64_t *)ptr(out_w, c10::ScalarType::Long, "out_w", n),
                                  (int8_t *)optptr(status, c10::ScalarType::Char, "status", n),
                                  (int32_t *)optptr(steps, c10::ScalarType::Int, "steps", 2 * n), stream_of(alphas)),
          "lp_kerr_trace_batch_f64");
}

void kerr_trace_alpha32(const Tensor &alpha32, const std::vector<double> &camv, int64_t row0, int64_t rows,
                        OptTensor refine_cols, double M, double a, double r_plus, double r_obs, double theta_obs,
                        double lambda_max, Tensor out_fa32, Tensor out_w16, OptTensor status, OptTensor steps)
{
    lp_camera cam = make_cam(camv);
    const int64_t n = rows * (int64_t)cam.width;
    c10::cuda::CUDAGuard g(alpha32.device());
    check(lp_kerr_trace_alpha32((const float *)ptr(alpha32, c10::ScalarType::Float, "alpha32", n), &cam, (int32_t)row0,
                                (int32_t)rows,
                                (const uint8_t *)optptr(refine_cols, c10::ScalarType::Byte, "axis_refine_cols", cam.width),
                                M, a, r_plus, r_obs, theta_obs, lambda_max,
                                (float *)ptr(out_fa32, c10::ScalarType::Float, "out_fa32", n),
                                (uint16_t *)ptr(out_w16, c10::ScalarType::UInt16, "out_w16", n),
                                (int8_t *)optptr(status, c10::ScalarType::Char, "status", n),
                                (int32_t *)optptr(steps, c10::ScalarType::Int, "steps", 2 * n), stream_of(alpha32)),
          "lp_kerr_trace_alpha32");
}

void bench_dfma(int64_t blocks, int64_t threads, int64_t iters, Tensor sink)
{
    c10::cuda::CUDAGuard g(sink.device());
    check(lp_bench_dfma((int32_t)blocks, (int32_t)threads, (int32_t)iters,
                        (double *)ptr(sink, c10::ScalarType::Double, "sink", blocks * threads), stream_of(sink)),
          "lp_bench_dfma");
}

std::vector<double> camera_init(int64_t height, int64_t width, double hfov, double vfov, double psi_y, double psi_x)
{
    lp_camera c;
    check(lp_camera_init((int32_t)height, (int32_t)width, hfov, vfov, psi_y, psi_x, &c), "lp_camera_init");
    std::vector<double> v = {(double)c.height, (double)c.width, c.fx, c.fy};
    for (int i = 0; i < 3; ++i) v.push_back(c.d[i]);
    for (int i = 0; i < 3; ++i) v.push_back(c.e_x[i]);
    for (int i = 0; i < 3; ++i) v.push_back(c.e_y[i]);
    return v;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.doc() = "thin PyTorch binding of liblightpath.so (C ABI in include/lightpath.h)";
    m.attr("STATS_WORDS") = kStatsWords;
    m.def("abi_version", []() { return lp_abi_version(); });
    m.def("device_count", []() { return lp_device_count(); });
    m.def("camera_init", &camera_init);
    m.def("trace_batch_f64", &trace_batch_f64);
    m.def("trace_alpha32", &trace_alpha32);
    m.def("trace_frame", &trace_frame);
    m.def("build_alpha_lookup", &build_alpha_lookup);
    m.def("remap", &remap);
    m.def("render_frame", &render_frame, py::arg("src"), py::arg("channels"), py::arg("cam"), py::arg("row0"),
          py::arg("rows"), py::arg("M"), py::arg("R_S"), py::arg("r_obs"), py::arg("phi_max"), py::arg("h_max"),
          py::arg("loop_around"), py::arg("sampling"), py::arg("out"), py::arg("fa32"), py::arg("w16"),
          py::arg("stats"), py::arg("flags"), py::arg("unit_u8"), py::arg("band_rows") = 0,
          py::arg("band_stride") = 0);
    m.def("peer_signal", &peer_signal);
    m.def("peer_wait", &peer_wait);
    m.def("shadow_classify", &shadow_classify);
    m.def("stats_reset", &stats_reset);
    m.def("stats_reduce", &stats_reduce);
    m.def("rk45_trace_batch", &rk45_trace_batch);
    m.def("rk45_trace_paths", &rk45_trace_paths);
    m.def("rk45_integrate_paths", &rk45_integrate_paths);
    m.def("kerr_rk45_integrate_paths", &kerr_rk45_integrate_paths);
    m.def("rk45_paths_dense", &rk45_paths_dense);
    m.def("kerr_trace_batch", &kerr_trace_batch);
    m.def("kerr_trace_alpha32", &kerr_trace_alpha32);
    m.def("bench_dfma", &bench_dfma);
}
