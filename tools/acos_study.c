// acos_study.c — host restatement of lp_acos_unit (csrc/lp_internal.cuh) measured against long double
// arithmetic (host tool, not product code; the MUFU.RSQ64H seed is modelled as a ~20-bit reciprocal square root).
//   python tools/acos_fit.py   (writes the coefficient table, needs mpmath)
//   gcc -O2 -ffp-contract=off tools/acos_study.c -lm -o /tmp/acos_study && /tmp/acos_study
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "acos_coef.h"
static double rsq_seed(double x) { // ~20-bit approximation
    double r = 1.0 / sqrt(x); uint64_t b; memcpy(&b, &r, 8); b &= 0xffffffff00000000ull; b |= 0; memcpy(&r, &b, 8);
    // keep 20 mantissa bits, then perturb by up to 2^-21 relative
    return r * (1.0 + ((rand() & 1023) - 512) * (1.0 / 1024) * 1e-6);
}
static double my_acos_pos(double ax) // ax in (0.5625, 1]
{
    const double t = 1.0 - ax;
    const double t2 = t + t;
    double p = ACOS_P[11];
    for (int k = 10; k >= 0; --k) p = fma(p, t, ACOS_P[k]);
    const double tp = t * p;
    const double r0 = rsq_seed(t2);
    const double e = fma(-(t * r0), r0, 0.5);
    const double r1 = fma(r0, e, r0);
    const double s0 = t2 * r1;
    const double hr = 0.5 * r1;
    const double d = fma(-s0, s0, t2);
    const double s1 = fma(d, hr, s0);
    const double d1 = fma(-s1, s1, t2);
    const double corr = d1 * hr;
    const double res = s1 + fma(s1, tp, corr);
    return t > 0.0 ? res : 0.0;
}
static double my_acos(double x)
{
    const double ax = fabs(x);
    if (!(ax > 0.5625)) return acos(x);
    const double r = my_acos_pos(ax);
    const double pi_hi = 0x1.921fb54442d18p+1, pi_lo = 0x1.1a62633145c07p-53;
    return x < 0.0 ? (pi_hi - (r - pi_lo)) : r;
}
static double ulp_err(double got, long double ref) {
    double rd = (double)ref; if (rd == 0) return got == 0 ? 0 : 1e9;
    int ex; frexp(rd, &ex); long double u = ldexpl(1.0L, ex - 53);
    return (double)fabsl(((long double)got - ref) / u);
}
int main(int argc, char **argv) {
    const long n_samples = argc > 1 ? atol(argv[1]) : 40000000;
    double worst = 0, worstg = 0, wx = 0; long ndiff = 0, n = 0;
    srand(1);
    for (long i = 0; i < n_samples; i++) {
        double x;
        int m = i % 4;
        double u = (rand() + 0.5) / (RAND_MAX + 1.0), v = (rand() + 0.5) / (RAND_MAX + 1.0);
        if (m == 0) x = 0.5625 + (1 - 0.5625) * (u + v * 1e-9);
        else if (m == 1) x = 1.0 - pow(10, -16 * u) * v;       // near 1
        else if (m == 2) x = cos(0.5 * u + v * 1e-9);              // typical alpha
        else x = -(0.5625 + (1 - 0.5625) * u);
        if (fabs(x) > 1) x = 1;
        long double ref = acosl((long double)x);
        double g = my_acos(x), l = acos(x);
        double e = ulp_err(g, ref), el = ulp_err(l, ref);
        if (e > worst) { worst = e; wx = x; }
        if (el > worstg) worstg = el;
        if (g != l) ndiff++;
        n++;
    }
    printf("n=%ld worst mine=%.4f ulp (x=%.17g) glibc=%.4f ulp, differ from glibc: %ld (%.4f%%)\n", n, worst, wx, worstg, ndiff, 100.0 * ndiff / n);
    double xs[] = {1.0, -1.0, 0.5625, 0.57, 0.999999999999999889, 0.9999, NAN};
    for (int i = 0; i < 7; i++) printf("%.17g -> %.17g (libm %.17g)\n", xs[i], my_acos(xs[i]), acos(xs[i]));
    return 0;
}
