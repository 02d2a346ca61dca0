// lp_remap.cu — kernel (2): deflection -> background remap
// (replaces image_lens.render_lensed_image, image_lens.py:296-397).
//
// One thread per output pixel; a warp covers 32 consecutive pixels of a row, so the
// fa / winding loads and the pixel stores are coalesced; the source gather follows the
// lensing map, which is smooth away from the photon ring, so neighbouring lanes hit
// neighbouring source texels (L1/L2 resident).  All direction math is fp64 like the
// reference's; what must match is the INTEGER source index.
#include "lp_remap.cuh"

template <typename T>
__global__ void __launch_bounds__(256)
lp_remap_kernel(const RemapArgs a, const CamConsts cam)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        const float fa32 = __ldg(a.fa32 + i);
        const unsigned wnd = a.w16 ? (unsigned)__ldg(a.w16 + i) : 0u;
        int row, col;
        pixel_row_col(i, a.n, cam.width, a.row0, row, col);
        remap_pixel<T>(a, cam, i, row, col, fa32, wnd);
    }
}

extern "C" int lp_remap(const void *src, int32_t src_dtype, int32_t channels,
                        const lp_camera *h_cam, const float *fa32, const uint16_t *w16,
                        int32_t render_loop_around, int32_t sampling,
                        int32_t row0, int32_t rows, void *out, void *stream)
{
    CamConsts cam;
    int rc = lp_make_cam_consts(h_cam, &cam);
    if (rc != LP_OK) return rc;
    if (row0 < 0 || rows < 0 || (long long)row0 + rows > cam.height) return LP_ERR_INVALID_ARG;
    if (channels < 1 || channels > 4) return LP_ERR_INVALID_ARG;
    if (sampling != LP_SAMPLE_NEAREST && sampling != LP_SAMPLE_BILINEAR) return LP_ERR_INVALID_ARG;
    RemapArgs a;
    a.src = src; a.out = out; a.fa32 = fa32; a.w16 = w16;
    a.n = (long long)rows * cam.width;
    a.row0 = row0; a.channels = channels; a.loop_around = render_loop_around; a.sampling = sampling;
    if (a.n == 0) return LP_OK;
    if (!src || !out || !fa32) return LP_ERR_INVALID_ARG;
    const void *fn;
    switch (src_dtype) {
    case LP_DTYPE_U8: fn = (const void *)lp_remap_kernel<unsigned char>; break;
    case LP_DTYPE_F32: fn = (const void *)lp_remap_kernel<float>; break;
    case LP_DTYPE_F64: fn = (const void *)lp_remap_kernel<double>; break;
    default: return LP_ERR_INVALID_ARG;
    }
    int grid = 0;
    rc = lp_grid_for(fn, 256, &grid);
    if (rc != LP_OK) return rc;
    const long long chunks = (a.n + 255) / 256;
    if (chunks < grid) grid = (int)chunks;
    cudaStream_t st = (cudaStream_t)stream;
    switch (src_dtype) {
    case LP_DTYPE_U8: lp_remap_kernel<unsigned char><<<grid, 256, 0, st>>>(a, cam); break;
    case LP_DTYPE_F32: lp_remap_kernel<float><<<grid, 256, 0, st>>>(a, cam); break;
    default: lp_remap_kernel<double><<<grid, 256, 0, st>>>(a, cam); break;
    }
    return lp_check_launch();
}
