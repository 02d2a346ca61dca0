/*
 * lp_oracle_rk45.c — CPU restatement of the reference's GENERIC path:
 * geodesic_tracer.trace_ray / integrate_geodesic (geodesic_tracer.py:22-82) on a
 * Schwarzschild metric (metrics.py:763-809).
 *
 * TEST INFRASTRUCTURE ONLY (see lp_oracle.c): only tests/, __graft_entry__.smoke() and
 * bench.py's CPU legs may load this.
 *
 * The stepper itself is NOT in the reference tree: it is scipy.integrate.solve_ivp
 * (method='RK45'), a third-party dependency the reference does not pin
 * (requirements.txt:2 says just `scipy`; the build container has scipy 1.18.1).  This file
 * restates scipy's published algorithm, citing scipy/integrate/_ivp/<file>:<line> of that
 * version:
 *   Dormand-Prince 5(4) tableau C, A, B, E, P          rk.py:538-566
 *   rk_step (6 stages + FSAL)                          rk.py:14-73
 *   step controller (_step_impl)                       rk.py:111-176
 *   select_initial_step, RMS norm                      common.py:63-134
 *   quartic dense output                               rk.py:178-180, :715-737
 *   event detection / terminal handling                ivp.py:134-157, :676-697
 *   brentq(xtol = rtol = 4 EPS, maxiter 100)           ivp.py:52-77 -> scipy/optimize/Zeros/brentq.c
 *
 * Pinning: against outputs of the UNMODIFIED reference (scipy 1.18.1) in
 * tests/golden/rk45_rays.npz — number of accepted points and nfev EXACTLY, the whole
 * accepted-step trajectory and the event point to ~1e-12 (tests/test_oracle_golden.py).
 * Bit equality is not attainable: scipy forms the stage sums with np.dot (BLAS), whose
 * summation order / FMA use is build specific; here they are plain
 * left-to-right sums.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define N 8
#define SAFETY 0.9
#define MIN_FACTOR 0.2
#define MAX_FACTOR 10.0

/* (the nodes C are not needed: the right-hand sides are autonomous) */
static const double A_[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double B_[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double E_[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525,
                             1.0 / 40};
static const double P_[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

typedef struct { double M, R_S; int nfev; int kerr; double a, r_plus; } rhs_ctx;

/* Kerr.geodesic_equations, metrics.py:946-1029 (the generic path's 8-D form: no sin^2 floor) */
static void rhs_kerr(rhs_ctx *c, const double *s, double *d)
{
    const double r = s[1], th = s[2], p_t = s[4], p_r = s[5], p_th = s[6], p_phi = s[7];
    const double M = c->M, a = c->a;
    if (r <= c->r_plus * 1.001) { for (int i = 0; i < N; ++i) d[i] = 0.0; return; }
    const double sin_th = sin(th), cos_th = cos(th);
    const double r2 = r * r, a2 = a * a, s2 = sin_th * sin_th;
    const double Sigma = r2 + a2 * (cos_th * cos_th);
    const double Delta = r2 - 2 * M * r + a2;
    const double A = (r2 + a2) * (r2 + a2) - a2 * Delta * s2;
    const double g_tt_inv = -A / (Sigma * Delta);
    const double g_tphi_inv = -2 * M * a * r / (Sigma * Delta);
    const double g_rr_inv = Delta / Sigma;
    const double g_thth_inv = 1.0 / Sigma;
    const double g_phiphi_inv = (Delta - a2 * s2) / (Sigma * Delta * s2);
    d[0] = g_tt_inv * p_t + g_tphi_inv * p_phi;
    d[1] = g_rr_inv * p_r;
    d[2] = g_thth_inv * p_th;
    d[3] = g_tphi_inv * p_t + g_phiphi_inv * p_phi;
    const double dSigma_dr = 2 * r, dDelta_dr = 2 * r - 2 * M;
    const double dA_dr = 4 * r * (r2 + a2) - a2 * dDelta_dr * s2;
    const double SD = Sigma * Delta;
    const double dg_tt_inv_dr = (-(dA_dr * Sigma * Delta - A * (dSigma_dr * Delta + Sigma * dDelta_dr)) / (SD * SD));
    const double dg_tphi_inv_dr = (-(2 * M * a * (Sigma * Delta - r * (dSigma_dr * Delta + Sigma * dDelta_dr)))
                                   / (SD * SD));
    const double dg_rr_inv_dr = (dDelta_dr * Sigma - Delta * dSigma_dr) / (Sigma * Sigma);
    const double dg_thth_inv_dr = -dSigma_dr / (Sigma * Sigma);
    const double den_r = Sigma * Delta * s2;
    const double dg_phiphi_inv_dr = ((dDelta_dr * Sigma * Delta * s2
                                      - (Delta - a2 * s2) * (dSigma_dr * Delta + Sigma * dDelta_dr) * s2)
                                     / (den_r * den_r));
    d[4] = 0.0;
    d[5] = -0.5 * (dg_tt_inv_dr * (p_t * p_t) + 2 * dg_tphi_inv_dr * p_t * p_phi + dg_rr_inv_dr * (p_r * p_r)
                   + dg_thth_inv_dr * (p_th * p_th) + dg_phiphi_inv_dr * (p_phi * p_phi));
    const double dSigma_dth = -2 * a2 * sin_th * cos_th;
    const double dA_dth = -a2 * Delta * 2 * sin_th * cos_th;
    const double dg_tt_inv_dth = (-(dA_dth * Sigma * Delta - A * dSigma_dth * Delta) / (SD * SD));
    const double dg_tphi_inv_dth = 2 * M * a * r * dSigma_dth / ((Sigma * Sigma) * Delta);
    const double dg_rr_inv_dth = -Delta * dSigma_dth / (Sigma * Sigma);
    const double dg_thth_inv_dth = -dSigma_dth / (Sigma * Sigma);
    const double num = Delta - a2 * s2, den = Sigma * Delta * s2;
    const double dnum_dth = -a2 * 2 * sin_th * cos_th;
    const double dden_dth = (dSigma_dth * Delta * s2 + Sigma * Delta * 2 * sin_th * cos_th);
    const double dg_phiphi_inv_dth = (dnum_dth * den - num * dden_dth) / (den * den);
    d[6] = -0.5 * (dg_tt_inv_dth * (p_t * p_t) + 2 * dg_tphi_inv_dth * p_t * p_phi + dg_rr_inv_dth * (p_r * p_r)
                   + dg_thth_inv_dth * (p_th * p_th) + dg_phiphi_inv_dth * (p_phi * p_phi));
    d[7] = 0.0;
}

/* Schwarzschild.geodesic_equations, metrics.py:763-790 */
static void rhs(rhs_ctx *c, const double *s, double *d)
{
    c->nfev++;
    if (c->kerr) { rhs_kerr(c, s, d); return; }
    const double r = s[1], th = s[2], p_t = s[4], p_r = s[5], p_th = s[6], p_phi = s[7];
    const double R_S = c->R_S;
    if (r <= R_S * 1.001) { for (int i = 0; i < N; ++i) d[i] = 0.0; return; }
    const double f = 1 - R_S / r;
    const double sin_th = sin(th), cos_th = cos(th);
    double sin_th_sq = sin_th * sin_th;
    if (sin_th_sq < 1e-15) sin_th_sq = 1e-15;
    const double r2 = r * r, r3 = r * r * r;
    d[0] = -p_t / f;
    d[1] = f * p_r;
    d[2] = p_th / r2;
    d[3] = p_phi / (r2 * sin_th_sq);
    d[4] = 0.0;
    d[5] = (-(R_S / (2 * r2)) * ((p_t * p_t) / (f * f))
            - (R_S / (2 * r2)) * (p_r * p_r)
            + ((p_th * p_th) + (p_phi * p_phi) / sin_th_sq) / r3);
    d[6] = cos_th * (p_phi * p_phi) / (r2 * sin_th_sq * sin_th);
    d[7] = 0.0;
}

static double rms_norm(const double *x) /* common.py:63-65 */
{
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += x[i] * x[i];
    return sqrt(s) / sqrt((double)N);
}

/* common.py:68-134 */
static double select_initial_step(rhs_ctx *c, double t0, const double *y0, double t_bound, double max_step,
                                  const double *f0, double rtol, double atol)
{
    (void)t0;
    const double interval_length = fabs(t_bound - t0);
    if (interval_length == 0.0) return 0.0;
    double scale[N], v[N], y1[N], f1[N];
    for (int i = 0; i < N; ++i) scale[i] = atol + fabs(y0[i]) * rtol;
    for (int i = 0; i < N; ++i) v[i] = y0[i] / scale[i];
    const double d0 = rms_norm(v);
    for (int i = 0; i < N; ++i) v[i] = f0[i] / scale[i];
    const double d1 = rms_norm(v);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (interval_length < h0) h0 = interval_length;
    for (int i = 0; i < N; ++i) y1[i] = y0[i] + h0 * 1.0 * f0[i];
    rhs(c, y1, f1);
    for (int i = 0; i < N; ++i) v[i] = (f1[i] - f0[i]) / scale[i];
    const double d2 = rms_norm(v) / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
    double h = 100 * h0;
    if (h1 < h) h = h1;
    if (interval_length < h) h = interval_length;
    if (max_step < h) h = max_step;
    return h;
}

typedef struct { double t_old, h, y_old[N], Q[N][4]; } dense_t;

static void dense_eval(const dense_t *d, double t, double *y) /* rk.py:723-737 */
{
    const double x = (t - d->t_old) / d->h;
    double p[4];
    p[0] = x; p[1] = p[0] * x; p[2] = p[1] * x; p[3] = p[2] * x;
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int m = 0; m < 4; ++m) s += d->Q[i][m] * p[m];
        y[i] = d->h * s + d->y_old[i];
    }
}

static double event_r(const dense_t *d, double t, double r_stop)
{
    double y[N];
    dense_eval(d, t, y);
    return y[1] - r_stop;
}

/* scipy/optimize/Zeros/brentq.c, as called by ivp.py:52-77 (xtol = rtol = 4 EPS, 100 iterations) */
static double brentq(const dense_t *d, double r_stop, double xa, double xb)
{
    const double xtol = 4 * 2.220446049250313e-16, rtol = 4 * 2.220446049250313e-16;
    double xpre = xa, xcur = xb, xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    fpre = event_r(d, xpre, r_stop);
    fcur = event_r(d, xcur, r_stop);
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) return NAN;   /* scipy raises ValueError */
    for (int i = 0; i < 100; ++i) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre; spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            const double lim = fmin(fabs(spre), 3 * fabs(sbis) - delta);
            if (2 * fabs(stry) < lim) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = event_r(d, xcur, r_stop);
    }
    return xcur;
}

/* Schwarzschild.initial_conditions, metrics.py:794-809.  Returns 0 for None. */
int lp_oracle_rk45_initial_conditions(double M, double r_obs, double alpha, double *state0)
{
    const double R_S = 2 * M;
    const double f0 = 1 - R_S / r_obs;
    const double b = r_obs * sin(alpha) / sqrt(f0);
    const double E = 1.0, L = b * E;
    const double p_r_sq = (E * E / f0 - (L * L) / (r_obs * r_obs)) / f0;
    if (!(p_r_sq >= 0)) return 0;   /* `p_r_sq < 0 -> None`; a NaN (alpha = NaN) would hang solve_ivp, treated as invalid */
    state0[0] = 0.0; state0[1] = r_obs; state0[2] = 3.141592653589793 / 2; state0[3] = 0.0;
    state0[4] = -E; state0[5] = -sqrt(p_r_sq); state0[6] = 0.0; state0[7] = L;
    return 1;
}

/* Metric selection for lp_oracle_rk45_integrate: thread-local so that the C signature of the
 * Schwarzschild path stays as it was (test infrastructure, single caller). */
static __thread int g_kerr_a_set = 0;
static __thread double g_kerr_a = 0.0, g_kerr_r_plus = 0.0;

/*
 * integrate_geodesic (geodesic_tracer.py:22-71).
 *   traj: optional [max_points][9] rows (t, y[8]) = OdeResult.t / .y columns; *n_points = the
 *   number of points solve_ivp would return (1 + accepted steps), even beyond max_points.
 *   *status: scipy's 1 (terminal event), 0 (reached lambda_max), -1 (step size too small).
 * Returns outcome: -1 captured, 1 escaped (geodesic_tracer.py:69-70).
 */
int lp_oracle_rk45_integrate(double M, double R_S, const double *state0, double lambda_max,
                             double rtol, double atol, double max_step,
                             double r_stop_inner, double r_stop_outer,
                             double *traj, int32_t max_points, int32_t *n_points, int32_t *nfev,
                             int32_t *status_out, double *t_final, double *y_final)
{
    rhs_ctx ctx = {M, R_S, 0, g_kerr_a_set, g_kerr_a, g_kerr_r_plus};
    double t = 0.0, y[N], f[N], K[7][N];
    const double t_bound = lambda_max;
    memcpy(y, state0, sizeof y);
    rhs(&ctx, y, f);                                                 /* rk.py:95 */
    double h_abs = select_initial_step(&ctx, t, y, t_bound, max_step, f, rtol, atol);
    int npts = 0;
    if (traj && npts < max_points) { traj[0] = t; memcpy(traj + 1, y, sizeof y); }
    npts = 1;
    double g0 = y[1] - r_stop_inner, g1 = y[1] - r_stop_outer;       /* ivp.py:650 */
    int status = 2;                                                  /* 2 = running */
    while (status == 2) {
        /* OdeSolver.step (base.py): already at the bound -> finished */
        if (t == t_bound) { status = 0; break; }
        /* ---- RungeKutta._step_impl, rk.py:111-176 ---- */
        const double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        double ha = h_abs;
        if (ha > max_step) ha = max_step; else if (ha < min_step) ha = min_step;
        int accepted = 0, rejected = 0, failed = 0;
        double t_new = t, h = 0, y_new[N], f_new[N];
        while (!accepted) {
            if (ha < min_step) { failed = 1; break; }
            h = ha;
            t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t;
            ha = fabs(h);
            /* rk_step, rk.py:14-73 */
            memcpy(K[0], f, sizeof f);
            for (int s = 1; s < 6; ++s) {
                double ys[N];
                for (int i = 0; i < N; ++i) {
                    double dy = 0.0;
                    for (int j = 0; j < s; ++j) dy += K[j][i] * A_[s][j];
                    ys[i] = y[i] + dy * h;
                }
                rhs(&ctx, ys, K[s]);
            }
            for (int i = 0; i < N; ++i) {
                double acc = 0.0;
                for (int j = 0; j < 6; ++j) acc += K[j][i] * B_[j];
                y_new[i] = y[i] + h * acc;
            }
            rhs(&ctx, y_new, f_new);
            memcpy(K[6], f_new, sizeof f_new);
            double e[N];
            for (int i = 0; i < N; ++i) {
                const double scale = atol + fmax(fabs(y[i]), fabs(y_new[i])) * rtol;
                double acc = 0.0;
                for (int j = 0; j < 7; ++j) acc += K[j][i] * E_[j];
                e[i] = acc * h / scale;
            }
            const double error_norm = rms_norm(e);
            if (error_norm < 1) {
                double factor = (error_norm == 0) ? MAX_FACTOR : fmin(MAX_FACTOR, SAFETY * pow(error_norm, -0.2));
                if (rejected) factor = fmin(1.0, factor);
                ha *= factor;
                accepted = 1;
            } else {
                ha *= fmax(MIN_FACTOR, SAFETY * pow(error_norm, -0.2));
                rejected = 1;
            }
        }
        if (failed) { status = -1; break; }                          /* ivp.py:663-665 */
        dense_t d;
        d.t_old = t; d.h = h; memcpy(d.y_old, y, sizeof y);
        for (int i = 0; i < N; ++i)
            for (int m = 0; m < 4; ++m) {
                double acc = 0.0;
                for (int j = 0; j < 7; ++j) acc += K[j][i] * P_[j][m];
                d.Q[i][m] = acc;
            }
        const double t_old = t;
        t = t_new; memcpy(y, y_new, sizeof y); memcpy(f, f_new, sizeof f); h_abs = ha;
        if (t - t_bound >= 0) status = 0;                            /* base.py: finished */
        /* ---- events, ivp.py:676-699 ---- */
        double tt = t, yy[N];
        memcpy(yy, y, sizeof y);
        const double gn0 = y[1] - r_stop_inner, gn1 = y[1] - r_stop_outer;
        const int act0 = (g0 >= 0) && (gn0 <= 0);                    /* direction -1 */
        const int act1 = (g1 <= 0) && (gn1 >= 0);                    /* direction +1 */
        if (act0 || act1) {
            double root0 = 0, root1 = 0;
            if (act0) root0 = brentq(&d, r_stop_inner, t_old, t);
            if (act1) root1 = brentq(&d, r_stop_outer, t_old, t);
            double root = act0 ? root0 : root1;
            if (act0 && act1 && root1 < root0) root = root1;         /* first terminal root */
            status = 1;
            tt = root;
            dense_eval(&d, tt, yy);
        }
        g0 = gn0; g1 = gn1;
        if (traj && npts < max_points) { traj[9 * npts] = tt; memcpy(traj + 9 * npts + 1, yy, sizeof yy); }
        npts++;
        if (status != 2) { t = tt; memcpy(y, yy, sizeof y); }
    }
    *n_points = npts; *nfev = ctx.nfev; *status_out = status; *t_final = t;
    memcpy(y_final, y, sizeof y);
    return (y[1] <= r_stop_inner * 1.1) ? -1 : 1;                    /* geodesic_tracer.py:69-70 */
}

/* geodesic_tracer.trace_ray over a batch of viewing angles (OpenMP over rays).
 * out_state [n][8], out_lambda [n], out_outcome [n] (1 / -1 / 0 = invalid), out_nsteps [n][2]
 * = (accepted points, nfev), out_status [n] (scipy status). r_stop_* <= 0 -> reference defaults. */
void lp_oracle_rk45_trace_batch(double M, double r_obs, const double *alphas, int64_t n,
                                double lambda_max, double rtol, double atol, double max_step,
                                double r_stop_inner, double r_stop_outer,
                                double *out_state, double *out_lambda, int8_t *out_outcome,
                                int32_t *out_nsteps, int8_t *out_status)
{
    const double R_S = 2 * M;
    const double r_in = r_stop_inner > 0 ? r_stop_inner : R_S * 1.01;   /* metrics.py:750-751 */
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        double s0[N], yf[N], tf = 0.0;
        int32_t np_ = 0, nf = 0, st = 0;
        if (!lp_oracle_rk45_initial_conditions(M, r_obs, alphas[i], s0)) {
            for (int k = 0; k < N; ++k) out_state[i * N + k] = NAN;
            out_lambda[i] = NAN; out_outcome[i] = 0;
            if (out_nsteps) { out_nsteps[2 * i] = 0; out_nsteps[2 * i + 1] = 0; }
            if (out_status) out_status[i] = -2;
            continue;
        }
        const double r_out = r_stop_outer > 0 ? r_stop_outer : s0[1] * 2.0;   /* geodesic_tracer.py:44-45 */
        const int oc = lp_oracle_rk45_integrate(M, R_S, s0, lambda_max, rtol, atol, max_step, r_in, r_out,
                                                NULL, 0, &np_, &nf, &st, &tf, yf);
        memcpy(out_state + i * N, yf, sizeof yf);
        out_lambda[i] = tf; out_outcome[i] = (int8_t)oc;
        if (out_nsteps) { out_nsteps[2 * i] = np_; out_nsteps[2 * i + 1] = nf; }
        if (out_status) out_status[i] = (int8_t)st;
    }
}

/* integrate_geodesic with a Kerr metric (metrics.py:946-1029); r_stop_inner default
 * capture_radius() = 1.01 r_plus (metrics.py:861-862) is the caller's to pass. */
int lp_oracle_rk45_integrate_kerr(double M, double a, double r_plus, const double *state0, double lambda_max,
                                  double rtol, double atol, double max_step,
                                  double r_stop_inner, double r_stop_outer,
                                  double *traj, int32_t max_points, int32_t *n_points, int32_t *nfev,
                                  int32_t *status_out, double *t_final, double *y_final)
{
    g_kerr_a_set = 1; g_kerr_a = a; g_kerr_r_plus = r_plus;
    const int oc = lp_oracle_rk45_integrate(M, 2 * M, state0, lambda_max, rtol, atol, max_step, r_stop_inner,
                                            r_stop_outer, traj, max_points, n_points, nfev, status_out, t_final,
                                            y_final);
    g_kerr_a_set = 0;
    return oc;
}
