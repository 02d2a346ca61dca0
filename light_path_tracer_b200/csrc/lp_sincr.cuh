// lp_sincr.cuh — sine of the ray's viewing angle, rounded correctly (to within ~2^-16 ulp of
// the half-way points), for the initial conditions b = r_obs * sin(alpha) / sqrt(f0)
// (reference metrics.py:55).
//
// Why: sin(alpha) is the ONLY libm-dependent input of the whole RK4 integration; every other
// operation of the step loop is an IEEE add/mul/div/sqrt that the kernel reproduces bit for
// bit.  The reference calls the host libm (glibc), whose sin is correctly rounded for
// 99.8-99.95 % of arguments (measured against mpmath, 4 x 10^5 samples); CUDA's sin()
// differs from it for 2.6 % of frame angles (measured on B200).  A one-ulp difference in
// sin(alpha) is harmless for ordinary rays but is amplified ~1/|b - b_crit| near the photon
// sphere, so the closer this is to "the correctly rounded value" the more rays have a
// trajectory that is bit-identical to the reference's.
//
// Method (double-double, table driven): for y in [0, pi/2], k = rint(256 y), t = y - k/256
// (exact), sin(y) = S_k cos t + C_k sin t with S_k, C_k = sin, cos(k/256) stored as hi + lo
// pairs; the leading product C_hi * t and the sum S_hi + C_hi t are formed exactly with
// fma / two-sum, everything else is a <= 2^-19 relative correction evaluated in plain double.
// Arguments in (pi/2, pi] are reflected exactly, y = (pi_hi - x) + pi_lo as a double-double.
// Anything else (negative, > pi, non-finite) goes to the CUDA library sin().
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LP_SINCR_FN __device__ __forceinline__
#define LP_SINTAB_QUAL static __device__ const
#define LP_FMA(a, b, c) fma((a), (b), (c))
#define LP_LIBSIN(x) sin(x)
#else
#define LP_SINCR_FN static inline
#define LP_SINTAB_QUAL static const
#define LP_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define LP_LIBSIN(x) sin(x)
#endif
#include "lp_sintab.h"

// Polynomial constants of lp_sin_dd.  On the device they sit in the constant bank: an FP64 instruction of
// sm_100 takes a constant through a uniform register, and an immediate double costs two issue slots
// (UMOV lo, hi) where a table entry costs half a slot (LDCU.128).
#ifdef __CUDACC__
static __constant__ double c_sincr_k[6] = {-0x1.6c16c16c16c17p-10, 0x1.5555555555555p-5, -0x1.a01a01a01a01ap-13,
                                           0x1.1111111111111p-7, -0x1.5555555555555p-3, -0.00390625};
#define LP_SINCR_K(i) c_sincr_k[i]
#else
static const double lp_sincr_k[6] = {-0x1.6c16c16c16c17p-10, 0x1.5555555555555p-5, -0x1.a01a01a01a01ap-13,
                                     0x1.1111111111111p-7, -0x1.5555555555555p-3, -0.00390625};
#define LP_SINCR_K(i) lp_sincr_k[i]
#endif

// sin(yh + yl) for 0 <= yh <= (LP_SINTAB_N-1)/256, |yl| <= ulp(yh)/2; has_lo = 0: yl is zero (the low-word
// term is skipped: the compiler cannot drop a multiplication by a zero it has to treat as IEEE)
LP_SINCR_FN double lp_sin_dd(double yh, double yl, int has_lo)
{
    const double shifter = 6755399441055744.0;           // 1.5 * 2^52: rint via add/sub
    const double kf = (yh * 256.0 + shifter) - shifter;  // rint(256 yh), exact integer
    const int k = (int)kf;
    const double t = LP_FMA(kf, LP_SINCR_K(5), yh);      // yh - k/256, exact
    const double Sh = lp_sintab[k][0], Sl = lp_sintab[k][1];
    const double Ch = lp_sintab[k][2], Cl = lp_sintab[k][3];
    const double t2 = t * t;
    // cos t - 1 and sin t - t, |t| <= 2^-9
    double a = LP_FMA(t2, LP_SINCR_K(0), LP_SINCR_K(1));
    a = LP_FMA(t2, a, -0.5);
    const double pc = t2 * a;
    double b = LP_FMA(t2, LP_SINCR_K(2), LP_SINCR_K(3));
    b = LP_FMA(t2, b, LP_SINCR_K(4));
    const double ps = (t * t2) * b;
    // leading terms, exactly: Sh + Ch*t = s + (err + e)
    const double p = Ch * t;
    const double e = LP_FMA(Ch, t, -p);
    const double s = Sh + p;
    const double bb = s - Sh;
    const double err = (Sh - (s - bb)) + (p - bb);       // two-sum
    double corr = LP_FMA(Cl, t, Sl + e);
    corr = corr + err;
    corr = LP_FMA(Sh, pc, corr);
    corr = LP_FMA(Ch, ps, corr);
    if (has_lo) {
        // cos(yh) to first order, for the low word of the argument
        const double cy = LP_FMA(-Sh, t, Ch);
        corr = LP_FMA(cy, yl, corr);
    }
    return s + corr;
}

LP_SINCR_FN double lp_sin_cr_pos(double x);

LP_SINCR_FN double lp_sin_cr(double x)
{
    return (x < 0.0) ? -lp_sin_cr_pos(-x) : lp_sin_cr_pos(x);     // sin is odd; so is rounding
}

LP_SINCR_FN double lp_sin_cr_pos(double x)
{
    const double pi_hi = 0x1.921fb54442d18p+1, pi_lo = 0x1.1a62633145c07p-53;
    if (x >= 0x1.0p-26 && x <= 1.5707963267948966) return lp_sin_dd(x, 0.0, 0);
    if (x > 1.5707963267948966 && x <= pi_hi) {
        const double d = pi_hi - x;                      // exact (Sterbenz)
        const double yh = d + pi_lo;
        const double bb = yh - d;
        const double yl = (d - (yh - bb)) + (pi_lo - bb);
        return lp_sin_dd(yh, yl, 1);
    }
    if (x >= 0.0 && x < 0x1.0p-26) return x;             // sin x = x(1 - x^2/6), rounds to x
    return LP_LIBSIN(x);
}
