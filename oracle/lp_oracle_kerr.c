/*
 * lp_oracle_kerr.c — CPU restatement of the reference's Kerr tracer:
 *   kerr_initial_conditions   metrics.py:148-224   _kerr_initial_conditions_numba
 *   kerr_rhs                  metrics.py:227-306   _kerr_geodesic_equations_numba
 *   kerr_extract_angle        metrics.py:362-416   _kerr_extract_angle
 *   lp_oracle_kerr_trace_ray  metrics.py:419-567   _kerr_trace_ray_numba (Dormand-Prince 4(5),
 *                                                  FSAL, the reference's own step controller)
 *   lp_oracle_kerr_trace_batch metrics.py:671-679  _trace_rays_batch_kerr
 *
 * TEST INFRASTRUCTURE ONLY (see lp_oracle.c).  Every expression keeps the reference's
 * left-to-right evaluation order; build with -ffp-contract=off (numba emits no FMA).  Pinned
 * bit-for-bit against the unmodified reference in tests/golden/kerr_rays.npz
 * (tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

#define LP_PI 3.141592653589793

static double clip_scalar(double x, double lo, double hi)
{
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

static double py_floordiv(double vx, double wx)   /* numba / CPython float floor division */
{
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    double floordiv;
    if (mod != 0.0) {
        if ((wx < 0.0) != (mod < 0.0)) div -= 1.0;
    }
    if (div != 0.0) {
        floordiv = floor(div);
        if (div - floordiv > 0.5) floordiv += 1.0;
    } else {
        div *= div;
        floordiv = div * vx / wx;
    }
    return floordiv;
}

/* metrics.py:148-224.  state = [r, theta, phi, p_r, p_theta].  Returns 0 when not ok. */
static int kerr_initial_conditions(double M, double a, double r_obs, double alpha, double theta,
                                   double theta_obs, double *state, double *p_t_out, double *p_phi_out)
{
    const double r = r_obs, th = theta_obs;
    const double sin_th = sin(th), cos_th = cos(th);
    double sin_th_sq = sin_th * sin_th;
    if (sin_th_sq < 1e-15) sin_th_sq = 1e-15;
    const double Sigma = r * r + a * a * cos_th * cos_th;
    const double Delta = r * r - 2.0 * M * r + a * a;
    if (Delta <= 0.0 || Sigma <= 0.0) return 0;
    const double sin_alpha = sin(alpha);
    const double sin_screen = sin(theta), cos_screen = cos(theta);
    const double E = 1.0;
    const double sqrt_Delta = sqrt(Delta), sqrt_Sigma = sqrt(Sigma);
    const double rho = r * sin_alpha * sqrt_Sigma / sqrt_Delta;
    const double alpha_screen = -rho * sin_screen;
    const double beta_screen = -rho * cos_screen;
    const double xi = -alpha_screen * sin_th;
    const double eta = beta_screen * beta_screen + cos_th * cos_th * (alpha_screen * alpha_screen - a * a);
    const double L = xi * E;
    const double Q = eta * E * E;
    const double p_t = -E, p_phi = L;
    double Theta = Q - cos_th * cos_th * (L * L / sin_th_sq - a * a * E * E);
    if (Theta < 0.0) Theta = 0.0;
    const double p_th_sign = (cos_screen > 0.0) ? -1.0 : 1.0;
    const double p_theta = p_th_sign * sqrt(Theta);
    const double A_val = (r * r + a * a) * (r * r + a * a) - a * a * Delta * sin_th_sq;
    const double g_tt_inv = -A_val / (Sigma * Delta);
    const double g_tphi_inv = -2.0 * M * a * r / (Sigma * Delta);
    const double g_rr_inv = Delta / Sigma;
    const double g_thth_inv = 1.0 / Sigma;
    const double g_phiphi_inv = (Delta - a * a * sin_th_sq) / (Sigma * Delta * sin_th_sq);
    const double other = (g_tt_inv * p_t * p_t
                          + 2.0 * g_tphi_inv * p_t * p_phi
                          + g_thth_inv * p_theta * p_theta
                          + g_phiphi_inv * p_phi * p_phi);
    double p_r_sq = -other / g_rr_inv;
    if (p_r_sq < 0.0) p_r_sq = 0.0;
    state[0] = r; state[1] = th; state[2] = 0.0; state[3] = -sqrt(p_r_sq); state[4] = p_theta;
    *p_t_out = p_t; *p_phi_out = p_phi;
    return 1;
}

/* Sensitivity probe (NOT a reference path): every sin / cos of theta in the right-hand side moved
 * by g_trig_shift ulps.  The reference calls the host libm there; another libm (or CUDA's) differs
 * from it by an ulp in a few percent of the calls, and the adaptive controller turns such ulps
 * into ~1e-9..1e-8 of final_alpha for some rays (rtol 1e-8 makes the error norm a 9-digit
 * cancellation).  The tests use this to measure, per ray, how far the REFERENCE's own result
 * moves under that kind of perturbation. */
static __thread int g_trig_shift = 0;      /* probe pattern, see trig_probe() */
static __thread unsigned g_trig_calls = 0;

static double nudge(double x, int k)
{
    for (; k > 0; --k) x = nextafter(x, INFINITY);
    for (; k < 0; ++k) x = nextafter(x, -INFINITY);
    return x;
}

/* probe patterns: 1 / 2: sin and cos one ulp apart in opposite directions, every call;
 * 3: sin + 1 ulp; 4: cos + 1 ulp; 5: both, sign alternating from call to call;
 * 6: every 7th call only (a sparse error pattern, like a libm that is off in a few % of calls) */
static void trig_probe(double *s, double *c)
{
    const unsigned k = g_trig_calls++;
    switch (g_trig_shift) {
    case 1: *s = nudge(*s, 1); *c = nudge(*c, -1); break;
    case 2: *s = nudge(*s, -1); *c = nudge(*c, 1); break;
    case 3: *s = nudge(*s, 1); break;
    case 4: *c = nudge(*c, 1); break;
    case 5: { const int d = (k & 1) ? 1 : -1; *s = nudge(*s, d); *c = nudge(*c, d); break; }
    case 6: if (k % 7 == 3) { *s = nudge(*s, 1); *c = nudge(*c, 1); } break;
    default: break;
    }
}

/* metrics.py:227-306 */
static void kerr_rhs(const double *s, double p_t, double p_phi, double M, double a, double r_plus, double *out)
{
    const double r = s[0], th = s[1], p_r = s[3], p_th = s[4];
    if (r <= r_plus * 1.001) { for (int i = 0; i < 5; ++i) out[i] = 0.0; return; }
    double sin_th = sin(th), cos_th = cos(th);
    if (g_trig_shift) trig_probe(&sin_th, &cos_th);
    double sin_th_sq = sin_th * sin_th;
    if (sin_th_sq < 1e-15) sin_th_sq = 1e-15;
    const double Sigma = r * r + a * a * cos_th * cos_th;
    const double Delta = r * r - 2.0 * M * r + a * a;
    const double A = (r * r + a * a) * (r * r + a * a) - a * a * Delta * sin_th_sq;
    const double g_tphi_inv = -2.0 * M * a * r / (Sigma * Delta);
    const double g_rr_inv = Delta / Sigma;
    const double g_thth_inv = 1.0 / Sigma;
    const double g_phiphi_inv = (Delta - a * a * sin_th_sq) / (Sigma * Delta * sin_th_sq);
    const double dr = g_rr_inv * p_r;
    const double dth = g_thth_inv * p_th;
    const double dphi = g_tphi_inv * p_t + g_phiphi_inv * p_phi;
    const double dSigma_dr = 2.0 * r;
    const double dDelta_dr = 2.0 * r - 2.0 * M;
    const double dA_dr = 4.0 * r * (r * r + a * a) - a * a * dDelta_dr * sin_th_sq;
    const double sigma_delta = Sigma * Delta;
    const double sigma_delta_sq = sigma_delta * sigma_delta;
    const double dg_tt_inv_dr = (-(dA_dr * sigma_delta - A * (dSigma_dr * Delta + Sigma * dDelta_dr))
                                 / sigma_delta_sq);
    const double dg_tphi_inv_dr = (-(2.0 * M * a * (sigma_delta - r * (dSigma_dr * Delta + Sigma * dDelta_dr)))
                                   / sigma_delta_sq);
    const double dg_rr_inv_dr = (dDelta_dr * Sigma - Delta * dSigma_dr) / (Sigma * Sigma);
    const double dg_thth_inv_dr = -dSigma_dr / (Sigma * Sigma);
    const double den_phi_dr = Sigma * Delta * sin_th_sq;
    const double dg_phiphi_inv_dr = ((dDelta_dr * den_phi_dr
                                      - (Delta - a * a * sin_th_sq)
                                      * (dSigma_dr * Delta + Sigma * dDelta_dr) * sin_th_sq)
                                     / (den_phi_dr * den_phi_dr));
    const double dp_r = -0.5 * (dg_tt_inv_dr * p_t * p_t
                                + 2.0 * dg_tphi_inv_dr * p_t * p_phi
                                + dg_rr_inv_dr * p_r * p_r
                                + dg_thth_inv_dr * p_th * p_th
                                + dg_phiphi_inv_dr * p_phi * p_phi);
    const double dSigma_dth = -2.0 * a * a * sin_th * cos_th;
    const double dA_dth = -a * a * Delta * 2.0 * sin_th * cos_th;
    const double dg_tt_inv_dth = (-(dA_dth * Sigma * Delta - A * dSigma_dth * Delta) / sigma_delta_sq);
    const double dg_tphi_inv_dth = 2.0 * M * a * r * dSigma_dth / (Sigma * Sigma * Delta);
    const double dg_rr_inv_dth = -Delta * dSigma_dth / (Sigma * Sigma);
    const double dg_thth_inv_dth = -dSigma_dth / (Sigma * Sigma);
    const double num = Delta - a * a * sin_th_sq;
    const double den = Sigma * Delta * sin_th_sq;
    const double dnum_dth = -a * a * 2.0 * sin_th * cos_th;
    const double dden_dth = dSigma_dth * Delta * sin_th_sq + Sigma * Delta * 2.0 * sin_th * cos_th;
    const double dg_phiphi_inv_dth = (dnum_dth * den - num * dden_dth) / (den * den);
    const double dp_th = -0.5 * (dg_tt_inv_dth * p_t * p_t
                                 + 2.0 * dg_tphi_inv_dth * p_t * p_phi
                                 + dg_rr_inv_dth * p_r * p_r
                                 + dg_thth_inv_dth * p_th * p_th
                                 + dg_phiphi_inv_dth * p_phi * p_phi);
    out[0] = dr; out[1] = dth; out[2] = dphi; out[3] = dp_r; out[4] = dp_th;
}

static int all_finite5(const double *x)
{
    for (int i = 0; i < 5; ++i) if (!isfinite(x[i])) return 0;
    return 1;
}

/* metrics.py:362-416 -> status (1 / -1 / 0), *fa, *nh */
static int kerr_extract_angle(const double *state, double p_t, double p_phi, double M, double a,
                              double r_capture, int event_status, double *fa, int64_t *nh)
{
    const double r_f = state[0], th_f = state[1], phi_f = state[2], p_r_f = state[3], p_th_f = state[4];
    const int64_t n_half = (int64_t)py_floordiv(fabs(phi_f), LP_PI);
    *fa = NAN;
    if (r_f <= r_capture * 1.1 || event_status == -1) { *nh = n_half; return -1; }
    if (!isfinite(r_f) || !isfinite(th_f) || !isfinite(phi_f)) { *nh = 0; return 0; }
    const double sin_th = sin(th_f), cos_th = cos(th_f);
    double sin_th_sq = sin_th * sin_th;
    if (sin_th_sq < 1e-15) sin_th_sq = 1e-15;
    const double Sigma_f = r_f * r_f + a * a * cos_th * cos_th;
    const double Delta_f = r_f * r_f - 2.0 * M * r_f + a * a;
    *nh = n_half;
    if (Sigma_f <= 1e-15 || fabs(Delta_f) <= 1e-15) return 0;
    const double dr_dl = Delta_f / Sigma_f * p_r_f;
    const double dth_dl = p_th_f / Sigma_f;
    const double dphi_dl = (-2.0 * M * a * r_f / (Sigma_f * Delta_f) * p_t
                            + (Delta_f - a * a * sin_th_sq) / (Sigma_f * Delta_f * sin_th_sq) * p_phi);
    const double sin_phi = sin(phi_f), cos_phi = cos(phi_f);
    const double vx = (sin_th * cos_phi * dr_dl + r_f * cos_th * cos_phi * dth_dl - r_f * sin_th * sin_phi * dphi_dl);
    const double vy = (sin_th * sin_phi * dr_dl + r_f * cos_th * sin_phi * dth_dl + r_f * sin_th * cos_phi * dphi_dl);
    const double vz = cos_th * dr_dl - r_f * sin_th * dth_dl;
    if (!isfinite(vx) || !isfinite(vy) || !isfinite(vz)) return 0;
    const double v_mag = sqrt(vx * vx + vy * vy + vz * vz);
    if (v_mag < 1e-30) return 1;                       /* status 1 with NaN angle */
    *fa = acos(clip_scalar(-vx / v_mag, -1.0, 1.0));
    return 1;
}

static const double A21 = 1.0 / 5.0, A31 = 3.0 / 40.0, A32 = 9.0 / 40.0, A41 = 44.0 / 45.0, A42 = -56.0 / 15.0,
                    A43 = 32.0 / 9.0, A51 = 19372.0 / 6561.0, A52 = -25360.0 / 2187.0, A53 = 64448.0 / 6561.0,
                    A54 = -212.0 / 729.0, A61 = 9017.0 / 3168.0, A62 = -355.0 / 33.0, A63 = 46732.0 / 5247.0,
                    A64 = 49.0 / 176.0, A65 = -5103.0 / 18656.0, B1 = 35.0 / 384.0, B3 = 500.0 / 1113.0,
                    B4 = 125.0 / 192.0, B5 = -2187.0 / 6784.0, B6 = 11.0 / 84.0, E1 = 71.0 / 57600.0,
                    E3 = -71.0 / 16695.0, E4 = 71.0 / 1920.0, E5 = -17253.0 / 339200.0, E6 = 22.0 / 525.0,
                    E7 = -1.0 / 40.0;

/* metrics.py:419-567.  *steps_out = [accepted steps, step attempts] (not in the reference's
 * return value). */
int lp_oracle_kerr_trace_ray(double M, double a, double r_plus, double r_obs, double alpha, double theta,
                             double theta_obs, double lambda_max, int axis_refine,
                             double *fa_out, int64_t *nh_out, int32_t *steps_out)
{
    double state[5], p_t, p_phi;
    if (steps_out) { steps_out[0] = 0; steps_out[1] = 0; }
    if (!kerr_initial_conditions(M, a, r_obs, alpha, theta, theta_obs, state, &p_t, &p_phi)) {
        *fa_out = NAN; *nh_out = 0; return 0;
    }
    const double r_capture = r_plus * 1.01, r_escape = r_obs * 2.0;
    const double atol = axis_refine ? 1e-10 : 1e-8, rtol = axis_refine ? 1e-8 : 1e-6;
    double k1[5], k2[5], k3[5], k4[5], k5[5], k6[5], k7[5], tmp[5], next_state[5];
    kerr_rhs(state, p_t, p_phi, M, a, r_plus, k1);
    double lam = 0.0;
    double h = fmax(1.0, 0.01 * r_obs);
    const double h_min = 1e-12;
    int event_status = 2;
    int accepted = 0, attempts = 0;
    for (int step = 0; step < 200000; ++step) {
        if (lam >= lambda_max) break;
        const double remaining = lambda_max - lam;
        if (h > remaining) h = remaining;
        if (h <= 0.0) break;
        attempts++;
        for (int i = 0; i < 5; ++i) tmp[i] = state[i] + h * A21 * k1[i];
        kerr_rhs(tmp, p_t, p_phi, M, a, r_plus, k2);
        for (int i = 0; i < 5; ++i) tmp[i] = state[i] + h * (A31 * k1[i] + A32 * k2[i]);
        kerr_rhs(tmp, p_t, p_phi, M, a, r_plus, k3);
        for (int i = 0; i < 5; ++i) tmp[i] = state[i] + h * (A41 * k1[i] + A42 * k2[i] + A43 * k3[i]);
        kerr_rhs(tmp, p_t, p_phi, M, a, r_plus, k4);
        for (int i = 0; i < 5; ++i)
            tmp[i] = state[i] + h * (A51 * k1[i] + A52 * k2[i] + A53 * k3[i] + A54 * k4[i]);
        kerr_rhs(tmp, p_t, p_phi, M, a, r_plus, k5);
        for (int i = 0; i < 5; ++i)
            tmp[i] = state[i] + h * (A61 * k1[i] + A62 * k2[i] + A63 * k3[i] + A64 * k4[i] + A65 * k5[i]);
        kerr_rhs(tmp, p_t, p_phi, M, a, r_plus, k6);
        for (int i = 0; i < 5; ++i)
            next_state[i] = state[i] + h * (B1 * k1[i] + B3 * k3[i] + B4 * k4[i] + B5 * k5[i] + B6 * k6[i]);
        kerr_rhs(next_state, p_t, p_phi, M, a, r_plus, k7);
        if (!all_finite5(next_state) || next_state[0] <= 0.0) {
            h *= 0.25;
            if (h < h_min) { *fa_out = NAN; *nh_out = 0; goto done_invalid; }
            continue;
        }
        double err_sq = 0.0;
        for (int i = 0; i < 5; ++i) {
            const double ei = h * (E1 * k1[i] + E3 * k3[i] + E4 * k4[i] + E5 * k5[i] + E6 * k6[i] + E7 * k7[i]);
            const double sc = atol + rtol * fmax(fabs(state[i]), fabs(next_state[i]));
            const double q = ei / sc;
            err_sq += q * q;
        }
        const double err_norm = sqrt(err_sq / 5.0);
        if (err_norm > 1.0) {
            const double factor = fmax(0.2, 0.9 * pow(err_norm, -0.2));
            h *= factor;
            if (h < h_min) { *fa_out = NAN; *nh_out = 0; goto done_invalid; }
            continue;
        }
        accepted++;
        const double r_prev = state[0], r_next = next_state[0];
        if (r_prev > r_capture && r_next <= r_capture) {
            const double denom = r_next - r_prev;
            double frac = (denom == 0.0) ? 1.0 : (r_capture - r_prev) / denom;
            frac = clip_scalar(frac, 0.0, 1.0);
            for (int i = 0; i < 5; ++i) state[i] = state[i] + frac * (next_state[i] - state[i]);
            lam += frac * h;
            event_status = -1;
            break;
        }
        if (r_prev < r_escape && r_next >= r_escape) {
            const double denom = r_next - r_prev;
            double frac = (denom == 0.0) ? 1.0 : (r_escape - r_prev) / denom;
            frac = clip_scalar(frac, 0.0, 1.0);
            for (int i = 0; i < 5; ++i) state[i] = state[i] + frac * (next_state[i] - state[i]);
            lam += frac * h;
            event_status = 1;
            break;
        }
        for (int i = 0; i < 5; ++i) state[i] = next_state[i];
        for (int i = 0; i < 5; ++i) k1[i] = k7[i];
        lam += h;
        if (!all_finite5(state)) { *fa_out = NAN; *nh_out = 0; goto done_invalid; }
        if (err_norm < 1e-10) h *= 5.0;
        else h *= fmin(5.0, 0.9 * pow(err_norm, -0.2));
    }
    if (steps_out) { steps_out[0] = accepted; steps_out[1] = attempts; }
    return kerr_extract_angle(state, p_t, p_phi, M, a, r_capture, event_status, fa_out, nh_out);
done_invalid:
    if (steps_out) { steps_out[0] = accepted; steps_out[1] = attempts; }
    return 0;
}

/* metrics.py:671-679 (+ status / steps for the tests).  lambda_max as Kerr.trace_rays_batch
 * passes it: max(5000, 6 r_obs) (metrics.py:1131). */
void lp_oracle_kerr_trace_batch(double M, double a, double r_plus, double r_obs,
                                const double *alphas, const double *thetas, double theta_obs,
                                double lambda_max, const uint8_t *axis_refines, int64_t n,
                                double *out_fa, int64_t *out_w, int8_t *out_status, int32_t *out_steps)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        double fa; int64_t nh; int32_t st[2];
        const int s = lp_oracle_kerr_trace_ray(M, a, r_plus, r_obs, alphas[i], thetas[i], theta_obs, lambda_max,
                                               axis_refines ? axis_refines[i] : 0, &fa, &nh, st);
        out_fa[i] = (s == 1) ? fa : NAN;
        out_w[i] = nh;
        if (out_status) out_status[i] = (int8_t)s;
        if (out_steps) { out_steps[2 * i] = st[0]; out_steps[2 * i + 1] = st[1]; }
    }
}

/* lp_oracle_kerr_trace_batch with the trig-shift probe switched on (see g_trig_shift). */
void lp_oracle_kerr_trace_batch_trigshift(double M, double a, double r_plus, double r_obs,
                                          const double *alphas, const double *thetas, double theta_obs,
                                          double lambda_max, const uint8_t *axis_refines, int64_t n, int shift,
                                          double *out_fa, int64_t *out_w, int8_t *out_status, int32_t *out_steps)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        double fa; int64_t nh; int32_t st[2];
        g_trig_shift = shift; g_trig_calls = 0;
        const int s = lp_oracle_kerr_trace_ray(M, a, r_plus, r_obs, alphas[i], thetas[i], theta_obs, lambda_max,
                                               axis_refines ? axis_refines[i] : 0, &fa, &nh, st);
        g_trig_shift = 0;
        out_fa[i] = (s == 1) ? fa : NAN;
        out_w[i] = nh;
        if (out_status) out_status[i] = (int8_t)s;
        if (out_steps) { out_steps[2 * i] = st[0]; out_steps[2 * i + 1] = st[1]; }
    }
}
