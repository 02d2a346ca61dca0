// lp_remap_tma.cu — kernel (2), variant with the source image STAGED THROUGH SHARED MEMORY BY TMA
// (north star item 2: "coalesced, vectorised loads of the source image staged through shared
// memory and TMA, with bilinear sampling").  Replaces image_lens.render_lensed_image
// (image_lens.py:296-397) for float32 / uint8 RGB images like lp_remap_f32rgb_x4_kernel does;
// same per-pixel decisions and the same integer source index (remap_pixel's arithmetic), so the
// frame is identical.
//
// A CTA owns a 32 x 8 tile of OUTPUT pixels (one pixel per thread).  Away from the shadow the
// lensing map is smooth — a tile's source pixels lie in a compact, mildly stretched and displaced
// copy of the tile — but it is not bounded a priori (the magnification diverges at the photon
// ring), so the footprint is not predicted, it is MEASURED: every thread first computes its source
// index (the fp64 part of the remap), the CTA reduces the bounding box of the indices with warp
// shuffles, and if the box fits one of four shared-memory box shapes a single elected thread
// issues ONE `cp.async.bulk.tensor.2d` (TMA tiled load of rows x (pixels*3) elements out of the
// image viewed as a 2-D tensor [H][3W]) completing on an mbarrier; all threads then gather their
// texel (or their four bilinear taps) from shared memory.  A tile whose footprint does not fit
// (photon ring, wrap-around seams) gathers from global memory exactly like the other kernels.
// Out-of-image parts of a box are zero-filled by the TMA unit and never read.
#include "lp_remap.cuh"
#include <cuda.h>
#include <stdlib.h>

#define RT_TX 32
#define RT_TY 8
#define RT_NBOX 4

// the four tensor maps travel as four separate __grid_constant__ kernel parameters: the TMA unit
// fetches a descriptor from the address it is given, and only a parameter named directly (not an
// element of a parameter array picked at run time) is guaranteed to be addressed in place
// box shapes in (pixels, rows); pixels * 3 elements <= 256 and (pixels * 3 * elem size) % 16 == 0
// for float32 and uint8 alike
static const int h_box_px[RT_NBOX] = {48, 64, 80, 80};
static const int h_box_rows[RT_NBOX] = {12, 16, 24, 32};
__constant__ int c_box_px[RT_NBOX] = {48, 64, 80, 80};
__constant__ int c_box_rows[RT_NBOX] = {12, 16, 24, 32};

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

// kind of an output pixel: what remap_pixel() would do with it
enum { PX_ZERO = 0, PX_WIND = 1, PX_MAGENTA = 2, PX_SAMPLE = 3 };

template <typename T>
__global__ void __launch_bounds__(RT_TX *RT_TY)
lp_remap_tma_kernel(const RemapArgs a, const CamConsts cam, const __grid_constant__ CUtensorMap map0,
                    const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                    const __grid_constant__ CUtensorMap map3, const int rows, const int debug)
{
    extern __shared__ __align__(128) unsigned char box_smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ int red[4][RT_TY];
    __shared__ int sel[3];                      // box index (-1: none), x0 (pixels), y0 (rows)

    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int col = blockIdx.x * RT_TX + lane;
    const int trow = blockIdx.y * RT_TY + wrp;
    const int H = cam.height, W = cam.width;
    const bool live = col < W && trow < rows;
    const int row = a.row0 + trow;
    const long long i = (long long)trow * W + col;
    const bool bilinear = a.sampling == LP_SAMPLE_BILINEAR;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---- per-pixel decision and source index: the arithmetic of remap_pixel() ----
    int kind = PX_ZERO;
    unsigned wnd = 0u;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;             // nearest: (x0, y0); bilinear: the four taps
    double tx = 0.0, ty = 0.0;
    if (live) {
        const float fa32 = __ldg(a.fa32 + i);
        wnd = a.w16 ? (unsigned)__ldg(a.w16 + i) : 0u;
        if (!isfinite(fa32)) kind = PX_ZERO;
        else if (fa32 > LP_HALF_PI_F32) kind = PX_WIND;
        else {
            double px, py;
            const bool front = source_coords(cam, row, col, fa32, px, py);
            long long ix = (long long)rint(px), iy = (long long)rint(py);
            bool ok;
            if (a.loop_around) { ix = pymod(ix, W); iy = pymod(iy, H); ok = true; }
            else ok = front && iy >= 0 && iy < H && ix >= 0 && ix < W;
            kind = ok ? PX_SAMPLE : PX_MAGENTA;
            if (ok) {
                x0 = (int)ix; y0 = (int)iy; x1 = x0; y1 = y0;
                if (bilinear) {
                    const double fxp = floor(px), fyp = floor(py);
                    tx = px - fxp; ty = py - fyp;
                    long long bx0 = (long long)fxp, by0 = (long long)fyp, bx1 = bx0 + 1, by1 = by0 + 1;
                    if (a.loop_around) {
                        bx0 = pymod(bx0, W); bx1 = pymod(bx1, W); by0 = pymod(by0, H); by1 = pymod(by1, H);
                    } else {
                        bx0 = min(max(bx0, 0ll), (long long)W - 1); bx1 = min(max(bx1, 0ll), (long long)W - 1);
                        by0 = min(max(by0, 0ll), (long long)H - 1); by1 = min(max(by1, 0ll), (long long)H - 1);
                    }
                    x0 = (int)bx0; x1 = (int)bx1; y0 = (int)by0; y1 = (int)by1;
                }
            }
        }
    }

    // ---- bounding box of the tile's source pixels (warp shuffles, then across the 8 warps) ----
    const bool samp = kind == PX_SAMPLE;
    int lo_x = samp ? min(x0, x1) : 0x7fffffff, hi_x = samp ? max(x0, x1) : -0x7fffffff;
    int lo_y = samp ? min(y0, y1) : 0x7fffffff, hi_y = samp ? max(y0, y1) : -0x7fffffff;
    lo_x = __reduce_min_sync(0xffffffffu, lo_x); hi_x = __reduce_max_sync(0xffffffffu, hi_x);
    lo_y = __reduce_min_sync(0xffffffffu, lo_y); hi_y = __reduce_max_sync(0xffffffffu, hi_y);
    if (lane == 0) { red[0][wrp] = lo_x; red[1][wrp] = hi_x; red[2][wrp] = lo_y; red[3][wrp] = hi_y; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bx0 = red[0][0], bx1 = red[1][0], by0 = red[2][0], by1 = red[3][0];
#pragma unroll
        for (int k = 1; k < RT_TY; ++k) {
            bx0 = min(bx0, red[0][k]); bx1 = max(bx1, red[1][k]);
            by0 = min(by0, red[2][k]); by1 = max(by1, red[3][k]);
        }
        int pick = -1;
        // the TMA unit faults ("illegal instruction") unless the box starts on a 16-byte boundary of global
        // memory (tools/tma_probe.cu, profiles/r2g_tma_probe.log): with 3 elements per pixel that is a
        // multiple of 4 pixels for float32 and of 16 pixels for uint8 (the row pitch is a multiple of 16 bytes)
        bx0 &= ~((sizeof(T) == 4 ? 4 : 16) - 1);
        if (bx1 >= bx0) {                                    // at least one sampling pixel
            const int need_w = bx1 - bx0 + 1, need_h = by1 - by0 + 1;
#pragma unroll
            for (int b = RT_NBOX - 1; b >= 0; --b)
                if (need_w <= c_box_px[b] && need_h <= c_box_rows[b]) pick = b;     // smallest that fits
        }
        if (debug & 1) pick = -1;
        sel[0] = pick; sel[1] = bx0; sel[2] = by0;
        if (pick >= 0 && (debug & 4)) pick = -2;      // debug: pretend, issue nothing
        if (pick >= 0) {
            const unsigned bytes = (unsigned)(c_box_px[pick] * 3 * c_box_rows[pick]) * (unsigned)sizeof(T);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes)
                         : "memory");
#define RT_ISSUE(MAP)                                                                                         \
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "                \
                 "[%0], [%1, {%2, %3}], [%4];"                                                                 \
                 ::"r"(smem_u32(box_smem)), "l"((unsigned long long)&(MAP)), "r"(bx0 * 3), "r"(by0),         \
                   "r"(smem_u32(&mbar))                                                                        \
                 : "memory")
            if (pick == 0) RT_ISSUE(map0);
            else if (pick == 1) RT_ISSUE(map1);
            else if (pick == 2) RT_ISSUE(map2);
            else RT_ISSUE(map3);
#undef RT_ISSUE
        }
    }
    __syncthreads();
    const int pick = sel[0];
    bool staged = pick >= 0 && !(debug & 6);
    if (staged) {
        // wait for the box (phase 0 of the barrier); bounded, so that a TMA that never completes
        // (a bad descriptor) degrades to the global-memory gather instead of hanging the GPU
        unsigned done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
        staged = done != 0;
    }
    if (!live) return;

    // ---- this pixel's value ----
    const T *__restrict__ src = (const T *)a.src;
    T *dst = (T *)a.out + i * 3;
    if (kind == PX_ZERO) { dst[0] = (T)0; dst[1] = (T)0; dst[2] = (T)0; return; }
    if (kind == PX_WIND) {
        const unsigned k = wnd > 4u ? 4u : wnd;
        dst[0] = colour<T>(c_wind_rgb[k][0], a.u8_scale); dst[1] = colour<T>(c_wind_rgb[k][1], a.u8_scale);
        dst[2] = colour<T>(c_wind_rgb[k][2], a.u8_scale);
        return;
    }
    if (kind == PX_MAGENTA) {
        const T one = colour<T>(1.0f, a.u8_scale);
        dst[0] = one; dst[1] = (T)0; dst[2] = one;
        return;
    }
    const int bw3 = staged ? c_box_px[pick] * 3 : 0, bx = sel[1], by = sel[2];
    const T *sbox = (const T *)box_smem;
    auto texel = [&](int x, int y, int ch) -> T {
        if (staged) return sbox[(y - by) * bw3 + (x - bx) * 3 + ch];
        return __ldg(src + ((long long)y * W + x) * 3 + ch);
    };
    if (!bilinear) {
        dst[0] = texel(x0, y0, 0); dst[1] = texel(x0, y0, 1); dst[2] = texel(x0, y0, 2);
        return;
    }
    const double w00 = (1.0 - tx) * (1.0 - ty), w01 = tx * (1.0 - ty), w10 = (1.0 - tx) * ty, w11 = tx * ty;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double v = w00 * (double)texel(x0, y0, ch) + w01 * (double)texel(x1, y0, ch)
                       + w10 * (double)texel(x0, y1, ch) + w11 * (double)texel(x1, y1, ch);
        if (sizeof(T) == 1) dst[ch] = (T)(int)rint(v);
        else dst[ch] = (T)v;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// 1 = launched, 0 = not applicable here (caller uses the other kernels), < 0 = error
int lp_remap_try_tma(const RemapArgs &a, const CamConsts &cam, int src_dtype, int rows, cudaStream_t stream)
{
    const int esz = src_dtype == LP_DTYPE_F32 ? 4 : (src_dtype == LP_DTYPE_U8 ? 1 : 0);
    if (!esz || a.channels != 3) return 0;
    const long long pitch = (long long)cam.width * 3 * esz;
    if (pitch % 16 != 0 || ((uintptr_t)a.src % 16) != 0) return 0;
    if ((long long)cam.width * 3 >= (1ll << 31) || cam.height < 1) return 0;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return 0;
    CUtensorMap maps[RT_NBOX];
    const cuuint64_t gdim[2] = {(cuuint64_t)cam.width * 3, (cuuint64_t)cam.height};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch};
    const cuuint32_t estr[2] = {1, 1};
    size_t smem = 0;
    for (int b = 0; b < RT_NBOX; ++b) {
        const cuuint32_t box[2] = {(cuuint32_t)(h_box_px[b] * 3), (cuuint32_t)h_box_rows[b]};
        const CUresult rc = enc(&maps[b], esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
                                2, const_cast<void *>(a.src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return 0;
        const size_t bytes = (size_t)h_box_px[b] * 3 * h_box_rows[b] * esz;
        if (bytes > smem) smem = bytes;
    }
    const char *de = getenv("LP_REMAP_TMA_DEBUG");
    const int dbg = de ? atoi(de) : 0;
    const dim3 grid((unsigned)((cam.width + RT_TX - 1) / RT_TX), (unsigned)((rows + RT_TY - 1) / RT_TY));
    if (grid.y > 65535u) return 0;
    if (esz == 4) {
        auto k = lp_remap_tma_kernel<float>;
        static int attr_f = 0;
        if (!attr_f) {
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
            attr_f = 1;
        }
        k<<<grid, RT_TX * RT_TY, smem, stream>>>(a, cam, maps[0], maps[1], maps[2], maps[3], rows, dbg);
    } else {
        auto k = lp_remap_tma_kernel<unsigned char>;
        k<<<grid, RT_TX * RT_TY, smem, stream>>>(a, cam, maps[0], maps[1], maps[2], maps[3], rows, dbg);
    }
    return lp_check_launch() == LP_OK ? 1 : LP_ERR_CUDA;
}
