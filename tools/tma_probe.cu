// tma_probe.cu — stand-alone probe of cp.async.bulk.tensor.2d on this box (no torch): which
// combination of (descriptor location, box shape, issuing thread) works.  Each case runs in a
// child process, so a faulting case does not take the others down.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: descriptor as __grid_constant__ param, 1: descriptor in global memory
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, float *out, int bw, int bh,
                      int c0, int c1, int elect)
{
    extern __shared__ __align__(128) unsigned char box[];
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long desc = MODE == 0 ? (unsigned long long)&pmap : (unsigned long long)gmap;
    bool issuer = threadIdx.x == 0;
    if (elect) {                      // warp 0 converged, one elected lane
        issuer = false;
        if (threadIdx.x < 32) {
            unsigned pred = 0;
            asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
            issuer = pred != 0;
        }
    }
    if (issuer) {
        const unsigned bytes = (unsigned)(bw * bh * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(box)), "l"(desc), "r"(c0), "r"(c1), "r"(smem_u32(&mbar)) : "memory");
    }
    __syncthreads();
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 20) && !done; ++spin)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    const float *b = (const float *)box;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = done ? b[i] : -1.0f;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int run_case(int mode, int bw, int bh, int c0, int c1, int elect, int W, int H)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 10;
    EncodeFn enc = (EncodeFn)p;
    float *src, *out, *h = (float *)malloc((size_t)W * H * 4);
    for (int i = 0; i < W * H; ++i) h[i] = (float)i;
    cudaMalloc(&src, (size_t)W * H * 4);
    cudaMalloc(&out, (size_t)bw * bh * 4);
    cudaMemcpy(src, h, (size_t)W * H * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstr[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 11; }
    CUtensorMap *gmap;
    cudaMalloc(&gmap, sizeof(map));
    cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)bw * bh * 4;
    if (mode == 0) {
        cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<0><<<1, 256, smem>>>(map, gmap, out, bw, bh, c0, c1, elect);
    } else {
        cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<1><<<1, 256, smem>>>(map, gmap, out, bw, bh, c0, c1, elect);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 12; }
    float *o = (float *)malloc(smem);
    cudaMemcpy(o, out, smem, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x) {
            const int gx = c0 + x, gy = c1 + y;
            const float want = (gx < W && gy < H && gx >= 0 && gy >= 0) ? (float)(gy * W + gx) : 0.0f;
            if (o[y * bw + x] != want) bad++;
        }
    printf("%s\n", bad ? "DATA MISMATCH" : "ok");
    return bad ? 13 : 0;
}

int main(int argc, char **argv)
{
    if (argc == 9) return run_case(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]),
                                   atoi(argv[7]), atoi(argv[8]));
    const int cases[][8] = {
        // mode bw  bh  c0  c1 elect  W     H
        {1, 32, 8, 0, 0, 1, 1440, 270},  {0, 32, 8, 0, 0, 1, 1440, 270},  {0, 32, 8, 0, 0, 0, 1440, 270},
        {0, 64, 8, 0, 0, 1, 1440, 270},  {0, 128, 8, 0, 0, 1, 1440, 270}, {0, 144, 12, 0, 0, 1, 1440, 270},
        {0, 144, 12, 33, 7, 1, 1440, 270}, {0, 144, 12, 33, 7, 0, 1440, 270}, {0, 192, 16, 300, 100, 0, 1440, 270},
        {0, 240, 32, 1300, 250, 0, 1440, 270}, {1, 240, 32, 1300, 250, 0, 1440, 270}, {0, 256, 16, 5, 5, 0, 1440, 270},
        {0, 144, 12, 33, 7, 0, 11520, 2160}};
    for (unsigned k = 0; k < sizeof(cases) / sizeof(cases[0]); ++k) {
        const int *c = cases[k];
        printf("mode=%d box=%dx%d at (%d,%d) elect=%d tensor %dx%d : ", c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
        fflush(stdout);
        pid_t pid = fork();
        if (pid == 0) {
            char a[8][16];
            for (int i = 0; i < 8; ++i) snprintf(a[i], 16, "%d", c[i]);
            execl(argv[0], argv[0], a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], (char *)0);
            _exit(99);
        }
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) printf("  -> rc %d\n", WIFEXITED(st) ? WEXITSTATUS(st) : -1);
    }
    return 0;
}
