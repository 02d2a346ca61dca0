// lp_trace.cuh — launch arguments shared by the Binet tracer kernels (lp_trace.cu) and the
// lane re-packing frame kernel (lp_repack.cu).
#pragma once
#include "lp_internal.cuh"
#include "lp_remap.cuh"

#define LP_TRACE_BLOCK 256          /* launch bound (max threads per CTA) */

enum { SRC_F64 = 0, SRC_F32 = 1, SRC_CAM = 2 };

struct TraceArgs {
    const void *alphas;     // SRC_F64: const double*, SRC_F32: const float*, SRC_CAM: unused
    long long n;            // rays in this launch
    void *out_fa;           // WIDE: double*, else float*
    void *out_w;            // WIDE: int64_t*, else uint16_t*
    float *out_alpha32;     // SRC_CAM only, optional
    int8_t *out_status;     // optional
    int32_t *out_steps;     // optional
    lp_frame_stats *stats;  // optional
    int32_t row0;           // SRC_CAM: first frame row of the tile
    int32_t retrace_steps;  // FUSED kernels: a ray that ran more RK4 steps than this is traced
                            // again with strict arithmetic (LP_TRACE_HYBRID); INT_MAX = never
    // interleaved row bands (lp_render_frame_bands): tile-local row r is frame row
    // row0 + (r / band_rows) * band_stride + (r % band_rows); band_rows == 0: contiguous
    int32_t band_rows, band_stride;
    int32_t out_frame_rows; // the pixel tile is frame-addressed (LP_RENDER_OUT_FRAME_ROWS)
};

// tile-local pixel index -> (frame row, column, element index of the pixel in the output tile)
__device__ __forceinline__ void tile_pixel(const TraceArgs &a, int width, long long i,
                                           int &row, int &col, long long &oi)
{
    int r;
    pixel_row_col(i, a.n, width, 0, r, col);
    if (a.band_rows > 0) {
        const int q = r / a.band_rows;
        r += q * (a.band_stride - a.band_rows);
    }
    row = a.row0 + r;
    oi = a.out_frame_rows ? (long long)r * width + col : i;
}

int lp_retrace_steps_for(uint32_t flags, double h_max);
int lp_launch_render_repack(const TraceArgs &a, const RemapArgs &ra, const BinetConsts &c, const CamConsts &cam,
                            int src_dtype, uint32_t flags, cudaStream_t stream);
