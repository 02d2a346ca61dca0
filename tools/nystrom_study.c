// nystrom_study.c — CPU study behind LP_TRACE_HYBRID's FMA loop (host tool, not product code).
//
// Compares three arithmetics of the SAME classical RK4 step of the Binet equation
// u'' = -u + 3 M u^2 (metrics.py:44-46, :83-92), h = 0.05:
//   S  strict   — every operation separately rounded, the reference's order (what numba emits)
//   F  fused    — the same stage structure with FMA contraction (22 fp64 operations per step)
//   N  Nystrom  — the algebraically identical second-order form: the stage slopes of u are the
//                 stage values of w, so w_a, w_b, w_c never have to be formed:
//                   k1 = g(u); ua = u + hh w;        k2 = g(ua); ub = ua + hh^2 k1;  k3 = g(ub)
//                   uhw = u + h w; uc = uhw + (h^2/2) k2;  k4 = g(uc)
//                   u' = uhw + (h^2/6)(k1 + k2 + k3);  w' = w + (h/6)(k1 + 2 k2 + 2 k3 + k4)
//                 18 fp64 operations per step.
// For rays binned by the number of steps they take it prints the largest relative difference
// of final_alpha against S, and the shortest ray whose status / half-orbit count differs.
//   gcc -O2 -fopenmp -ffp-contract=off tools/nystrom_study.c -lm -o /tmp/nystrom_study
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#define PI 3.141592653589793
static inline double gF(double u, double M3) { return fma(M3 * u, u, -u); }
static inline double gS(double u, double M3) { return -u + M3 * u * u; }
static int orbit(int mode, double M, double R_S, double r_obs, double alpha, double *phi_o, double *u_o, double *w_o, int *steps_o)
{
    double f0 = 1.0 - R_S / r_obs; double b = r_obs * sin(alpha) / sqrt(f0); if (b == 0.0) return 0;
    double u = 1.0 / r_obs; double w0 = 1.0 / (b * b) - u * u + 2.0 * M * u * u * u; if (w0 < 0) return 0; double w = sqrt(w0);
    double phi = 0, uc = 1.0 / (R_S * 1.01), ue = 1.0 / (2.0 * r_obs); int status = 2; int steps = 0;
    const double h = 0.05, hh = 0.5 * h, h6 = h / 6.0, M3 = 3.0 * M;
    const double hh2 = hh * hh, h2_2 = h * hh, h2_6 = h * h6;
    while (phi < 50.0) {
        double rem = 50.0 - phi; double hs = h; if (rem < hs) hs = rem; if (hs <= 0) break;
        double up = u, wp = w;
        if (mode == 2 && hs == h) {
            const double k1 = gF(u, M3);
            const double ua = fma(hh, w, u);
            const double k2 = gF(ua, M3);
            const double ub = fma(hh2, k1, ua);
            const double k3 = gF(ub, M3);
            const double uhw = fma(h, w, u);
            const double ucc = fma(h2_2, k2, uhw);
            const double k4 = gF(ucc, M3);
            const double p = k2 + k3;
            const double s3 = k1 + p;
            u = fma(h2_6, s3, uhw);
            w = fma(h6, (s3 + p) + k4, wp);
        } else if (mode == 3 && hs == h) {
            const double k1 = gF(u, M3);
            const double ua = fma(hh, w, u);
            const double k2 = gF(ua, M3);
            const double ub = fma(hh2, k1, ua);
            const double k3 = gF(ub, M3);
            const double uhw = fma(h, w, u);
            const double ucc = fma(h2_2, k2, uhw);
            const double k4 = gF(ucc, M3);
            const double p = k2 + k3;
            const double s3 = k1 + p;
            u = fma(h, fma(h6, s3, w), up);
            w = fma(h6, (s3 + p) + k4, wp);
        } else if (mode == 1 && hs == h) {
            double k1u = w, k1w = gF(u, M3); double ua = fma(hh, k1u, u), wa = fma(hh, k1w, w);
            double k2u = wa, k2w = gF(ua, M3); double ub = fma(hh, k2u, u), wb = fma(hh, k2w, w);
            double k3u = wb, k3w = gF(ub, M3); double ucc = fma(h, k3u, u), wc = fma(h, k3w, w);
            double k4u = wc, k4w = gF(ucc, M3);
            double su = fma(2.0, k3u, fma(2.0, k2u, k1u)) + k4u, sw = fma(2.0, k3w, fma(2.0, k2w, k1w)) + k4w;
            u = fma(h6, su, up); w = fma(h6, sw, wp);
        } else {
            double H = hs, HH = 0.5 * hs, H6 = hs / 6.0;
            double k1u = w, k1w = gS(u, M3); double ua = up + HH * k1u, wa = wp + HH * k1w;
            double k2u = wa, k2w = gS(ua, M3); double ub = up + HH * k2u, wb = wp + HH * k2w;
            double k3u = wb, k3w = gS(ub, M3); double ucc = up + H * k3u, wc = wp + H * k3w;
            double k4u = wc, k4w = gS(ucc, M3);
            u = up + H6 * (k1u + 2.0 * k2u + 2.0 * k3u + k4u); w = wp + H6 * (k1w + 2.0 * k2w + 2.0 * k3w + k4w);
        }
        steps++;
        if (up < uc && u >= uc) { double d = u - up; double fr = d == 0 ? 1 : (uc - up) / d; if (fr < 0) fr = 0; if (fr > 1) fr = 1; phi += fr * hs; w = wp + fr * (w - wp); u = uc; status = -1; break; }
        if (up > ue && u <= ue) { double d = u - up; double fr = d == 0 ? 1 : (ue - up) / d; if (fr < 0) fr = 0; if (fr > 1) fr = 1; phi += fr * hs; w = wp + fr * (w - wp); u = ue; status = 1; break; }
        phi += hs;
    }
    *phi_o = phi; *u_o = u; *w_o = w; *steps_o = steps; return status;
}

// mode 4: the Nystrom step in the SCALED variable v = 3M u (v'' = -v + v^2): g is ONE fma, 14 fp64
// operations per step.  The loop state, the band test and the exit are in v; the crossing interpolation
// is the reference's, on the un-scaled values.
static int orbit_scaled(double M, double R_S, double r_obs, double alpha, double *phi_o, double *u_o, double *w_o, int *steps_o)
{
    double f0 = 1.0 - R_S / r_obs; double b = r_obs * sin(alpha) / sqrt(f0); if (b == 0.0) return 0;
    double u = 1.0 / r_obs; double w0 = 1.0 / (b * b) - u * u + 2.0 * M * u * u * u; if (w0 < 0) return 0; double w = sqrt(w0);
    double phi = 0, uc = 1.0 / (R_S * 1.01), ue = 1.0 / (2.0 * r_obs); int status = 2; int steps = 0;
    const double h = 0.05, hh = 0.5 * h, h6 = h / 6.0, M3 = 3.0 * M, iM3 = 1.0 / M3;
    const double hh2 = hh * hh, h2_2 = h * hh, h2_6 = h * h6;
    const double VC = M3 * uc, VE = M3 * ue;
    double v = M3 * u, wv = M3 * w;
    while (phi < 50.0) {
        double rem = 50.0 - phi; double hs = h; if (rem < hs) hs = rem; if (hs <= 0) break;
        if (hs != h) { fprintf(stderr, "short step\n"); exit(1); }
        double vp = v, wvp = wv;
        const double k1 = fma(v, v, -v);
        const double va = fma(hh, wv, v);
        const double k2 = fma(va, va, -va);
        const double vb = fma(hh2, k1, va);
        const double k3 = fma(vb, vb, -vb);
        const double vhw = fma(h, wv, v);
        const double vcc = fma(h2_2, k2, vhw);
        const double k4 = fma(vcc, vcc, -vcc);
        const double p = k2 + k3;
        const double s3 = k1 + p;
        v = fma(h2_6, s3, vhw);
        wv = fma(h6, (s3 + p) + k4, wvp);
        steps++;
        if (v >= VC || v <= VE) {
            double up = vp * iM3, wp = wvp * iM3; u = v * iM3; w = wv * iM3;
            double tg = v >= VC ? uc : ue;
            double d = u - up; double fr = d == 0 ? 1 : (tg - up) / d; if (fr < 0) fr = 0; if (fr > 1) fr = 1;
            phi += fr * hs; w = wp + fr * (w - wp); u = tg; status = v >= VC ? -1 : 1; break;
        }
        phi += hs;
    }
    if (status == 2) { u = v * iM3; w = wv * iM3; }
    *phi_o = phi; *u_o = u; *w_o = w; *steps_o = steps; return status;
}
static int ray(int mode, double M, double r_obs, double alpha, double *fa, long *nh, int *steps)
{
    double phi, u, w; int st = mode == 4 ? orbit_scaled(M, 2 * M, r_obs, alpha, &phi, &u, &w, steps) : orbit(mode, M, 2 * M, r_obs, alpha, &phi, &u, &w, steps); if (st == 0) { *fa = NAN; *nh = 0; return 0; }
    double r = 1.0 / u; *nh = (long)floor(fabs(phi) / PI); if (st == -1 || r <= 2.2 * M) { *fa = NAN; return -1; }
    double dr = -w / (u * u); double hy = dr * sin(phi) + r * cos(phi), hx = dr * cos(phi) - r * sin(phi); double c = -cos(atan2(hy, hx)); if (c > 1) c = 1; if (c < -1) c = -1; *fa = acos(c); return 1;
}
int main(int argc, char **argv)
{
    double robs[] = {3.5, 6, 15, 25, 50, 100, 300, 1000};
    long N = argc > 1 ? atol(argv[1]) : 4000000;
    for (int ir = 0; ir < 8; ir++) {
        double r_obs = robs[ir]; double M = 1; double ac = asin(3 * sqrt(3.0) * sqrt(1 - 2 / r_obs) / r_obs);
        double maxrel[4][40] = {{0}}; long cnt[40] = {0}; long flips[4] = {0, 0, 0, 0}, nhdiff[4] = {0, 0, 0, 0}; int flipmin[4] = {100000, 100000, 100000, 100000};
#pragma omp parallel
        {
            double lmax[4][40] = {{0}}; long lcnt[40] = {0}; long lfl[4] = {0, 0, 0, 0}, lnh[4] = {0, 0, 0, 0}; int lmin[4] = {100000, 100000, 100000, 100000};
#pragma omp for schedule(dynamic, 4096)
            for (long i = 0; i < N; i++) {
                double t = (double)i / N; double alpha;
                if (i % 2 == 0) { double e = pow(10, -14 + 13.5 * t); alpha = ac * (1 + ((i / 2) % 2 ? e : -e)); } else alpha = ac * (0.2 + 3.0 * t);
                if (r_obs < 3.0 * M + 1e-9 || !(alpha < PI)) continue;
                double fa0; long n0; int s0; int st0 = ray(0, M, r_obs, alpha, &fa0, &n0, &s0);
                int bin = s0 / 25; if (bin > 39) bin = 39; lcnt[bin]++;
                for (int m = 1; m <= 4; m++) {
                    double fa; long n; int s; int st = ray(m, M, r_obs, alpha, &fa, &n, &s);
                    int sm = s < s0 ? s : s0;
                    if (st != st0) { lfl[m - 1]++; if (sm < lmin[m - 1]) lmin[m - 1] = sm; }
                    else { if (n != n0) { lnh[m - 1]++; if (sm < lmin[m - 1]) lmin[m - 1] = sm; }
                           if (st0 == 1) { double d = fabs(fa0 - fa) / fmax(fa0, 1e-3); if (d > lmax[m - 1][bin]) lmax[m - 1][bin] = d; } }
                }
            }
#pragma omp critical
            {
                for (int m = 0; m < 4; m++) { flips[m] += lfl[m]; nhdiff[m] += lnh[m]; if (lmin[m] < flipmin[m]) flipmin[m] = lmin[m];
                    for (int b = 0; b < 40; b++) if (lmax[m][b] > maxrel[m][b]) maxrel[m][b] = lmax[m][b]; }
                for (int b = 0; b < 40; b++) cnt[b] += lcnt[b];
            }
        }
        printf("r_obs=%g  F: flips=%ld nhdiff=%ld shortest=%d | N: flips=%ld nhdiff=%ld shortest=%d | N2: flips=%ld nhdiff=%ld shortest=%d | V: flips=%ld nhdiff=%ld shortest=%d\n", r_obs, flips[0], nhdiff[0], flipmin[0], flips[1], nhdiff[1], flipmin[1], flips[2], nhdiff[2], flipmin[2], flips[3], nhdiff[3], flipmin[3]);
        for (int b = 0; b < 40; b++) if (cnt[b]) printf("  steps %4d-%4d n=%8ld  maxrel F=%.3e  N=%.3e  N2=%.3e  V=%.3e\n", b * 25, b * 25 + 24, cnt[b], maxrel[0][b], maxrel[1][b], maxrel[2][b], maxrel[3][b]);
        fflush(stdout);
    }
    return 0;
}
