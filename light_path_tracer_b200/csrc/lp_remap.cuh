// lp_remap.cuh — per-pixel device code of the remap (shared by the stand-alone remap
// kernel and the fully fused frame kernel).
#pragma once
#include "lp_internal.cuh"

// WINDING_COLORS (image_lens.py:287-293) as float32, and WINDING_COLORS @ luma
// (image_lens.py:330-331) evaluated in float32 by numpy.
static __constant__ float c_wind_rgb[5][3] = {
    {0.0f, 0.2f, 1.0f}, {0.0f, 0.7f, 1.0f}, {0.0f, 1.0f, 0.4f}, {1.0f, 1.0f, 0.0f}, {1.0f, 0.4f, 0.0f}};
static __constant__ float c_wind_luma[5] = {
    0x1.d9e84p-3f, 0x1.0cbfb2p-1f, 0x1.43e426p-1f, 0x1.c5a1ccp-1f, 0x1.114e3cp-1f};

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ unsigned char from_f32<unsigned char>(float v) { return (unsigned char)v; }  // C cast, like numpy's assignment
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ double from_f32<double>(float v) { return (double)v; }

struct RemapArgs {
    const void *src;
    void *out;
    const float *fa32;
    const unsigned short *w16;   // optional
    long long n;                 // pixels in the tile
    int32_t row0;
    int32_t channels;
    int32_t loop_around;
    int32_t sampling;
};

// python `a % n` for n > 0
__device__ __forceinline__ long long pymod(long long a, long long n)
{
    long long m = a % n;
    return m < 0 ? m + n : m;
}

// Source direction of an escaped pixel -> pinhole plane coordinates (image_lens.py:310-352).
// Returns front (src_vz > 1e-12) and the continuous source pixel coordinates.
//
// The reference normalises the pixel ray v, takes theta = arctan2(v.e_x, v.e_y) and then
// sin(theta), cos(theta).  theta is invariant under scaling of v, and sin/cos of an arctan2
// are the normalised components themselves, so here (st, ct) = (A, B)/hypot(A, B) with
// A, B the dot products of the UN-normalised ray (x_cam, y_cam, 1): no arctan2, no second
// sincos, no normalisation of v.  What has to match the reference is the INTEGER source
// index rint(px), rint(py); the two evaluations differ by a few ulp of px (~1e-13 pixel), so
// they can only disagree on a tie that is that close to a half-integer.
__device__ __forceinline__ bool source_coords(const CamConsts &cam, int row, int col, float fa32,
                                              double &px, double &py)
{
    const double xc = cam_x(cam, col);
    const double yc = cam_y(cam, row);
    const double A = fma(xc, cam.ex0, fma(yc, cam.ex1, cam.ex2));
    const double B = fma(xc, cam.ey0, fma(yc, cam.ey1, cam.ey2));
    const double n2 = fma(A, A, B * B);
    double st = 0.0, ct = 1.0;                              // arctan2(0, 0) = 0
    if (n2 > 0.0) {
        const double inv = rsqrt(n2);
        st = A * inv; ct = B * inv;
    }
    double sf, cf;
    sincos((double)fa32, &sf, &cf);                         // image_lens.py:340-346
    const double tx = fma(st, cam.ex0, ct * cam.ey0);
    const double ty = fma(st, cam.ex1, ct * cam.ey1);
    const double tz = fma(st, cam.ex2, ct * cam.ey2);
    const double sx = fma(cf, cam.d0, sf * tx);
    const double sy = fma(cf, cam.d1, sf * ty);
    const double sz = fma(cf, cam.d2, sf * tz);
    const bool front = sz > 1e-12;
    if (front) {
        px = fma(__ddiv_rn(sx, sz), cam.fx, cam.half_w);          // image_lens.py:374
        py = fma(__ddiv_rn(sy, sz), cam.fy, cam.half_h);
    } else {
        px = cam.half_w;                                          // image_lens.py:356-361 (zeros * f + n/2)
        py = cam.half_h;
    }
    return front;
}

template <typename T>
__device__ __forceinline__ void remap_pixel(const RemapArgs &a, const CamConsts &cam, long long i,
                                            int row, int col, float fa32, unsigned wnd)
{
    const int C = a.channels;
    const T *__restrict__ src = (const T *)a.src;
    T *__restrict__ dst = (T *)a.out + i * C;
    if (!isfinite(fa32)) {                                   // captured / invalid: zeros_like
        for (int ch = 0; ch < C; ++ch) dst[ch] = (T)0;
        return;
    }
    if (fa32 > LP_HALF_PI_F32) {                             // winding false colour, image_lens.py:322-333
        const unsigned k = wnd > 4u ? 4u : wnd;
        if (C == 1) {
            dst[0] = from_f32<T>(c_wind_luma[k]);
        } else {
            for (int ch = 0; ch < C; ++ch) dst[ch] = (ch < 3) ? from_f32<T>(c_wind_rgb[k][ch]) : (T)0;
        }
        return;
    }
    double px, py;
    const bool front = source_coords(cam, row, col, fa32, px, py);
    const long long H = cam.height, W = cam.width;
    long long ix = (long long)rint(px), iy = (long long)rint(py);   // np.rint -> intp
    bool ok;
    if (a.loop_around) {                                     // image_lens.py:354-365
        ix = pymod(ix, W);
        iy = pymod(iy, H);
        ok = true;
    } else {                                                 // image_lens.py:367-379
        ok = front && iy >= 0 && iy < H && ix >= 0 && ix < W;
    }
    if (!ok) {                                               // magenta, image_lens.py:381-393
        if (C == 1) dst[0] = (T)1;
        else for (int ch = 0; ch < C; ++ch) dst[ch] = (ch == 0 || ch == 2) ? (T)1 : (T)0;
        return;
    }
    if (a.sampling == LP_SAMPLE_NEAREST) {
        const T *s = src + (iy * W + ix) * C;
        for (int ch = 0; ch < C; ++ch) dst[ch] = __ldg(s + ch);
        return;
    }
    // bilinear extension: same in/out-of-frame decision as nearest, 4-tap blend around the
    // continuous coordinate, taps clamped (or wrapped when loop_around) at the border.
    const double fxp = floor(px), fyp = floor(py);
    const double tx = px - fxp, ty = py - fyp;
    long long x0 = (long long)fxp, y0 = (long long)fyp, x1 = x0 + 1, y1 = y0 + 1;
    if (a.loop_around) {
        x0 = pymod(x0, W); x1 = pymod(x1, W); y0 = pymod(y0, H); y1 = pymod(y1, H);
    } else {
        x0 = min(max(x0, 0ll), W - 1); x1 = min(max(x1, 0ll), W - 1);
        y0 = min(max(y0, 0ll), H - 1); y1 = min(max(y1, 0ll), H - 1);
    }
    const double w00 = (1.0 - tx) * (1.0 - ty), w01 = tx * (1.0 - ty), w10 = (1.0 - tx) * ty, w11 = tx * ty;
    for (int ch = 0; ch < C; ++ch) {
        const double v = w00 * (double)__ldg(src + (y0 * W + x0) * C + ch)
                       + w01 * (double)__ldg(src + (y0 * W + x1) * C + ch)
                       + w10 * (double)__ldg(src + (y1 * W + x0) * C + ch)
                       + w11 * (double)__ldg(src + (y1 * W + x1) * C + ch);
        if (sizeof(T) == 1) dst[ch] = (T)(int)rint(v);
        else dst[ch] = (T)v;
    }
}

