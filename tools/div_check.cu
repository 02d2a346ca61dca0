// tools/div_check.cu — GPU sweep of div_by(x, d, div_rcp(d)) against the IEEE quotient x / d.
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -Iinclude
//        -Ilight_path_tracer_b200/csrc tools/div_check.cu -o /tmp/div_check && /tmp/div_check
// Patterns: random mantissas, mantissas near all-ones / near a power of two (the hard cases of
// reciprocal-based division), exponents over the range the Kerr right-hand side produces and
// well beyond it, numerator exactly 1.  Prints the mismatch count per pattern; exit 1 on any.
#include <cstdio>
#include <cstdint>
#include "lp_internal.cuh"

__device__ __forceinline__ uint64_t mix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double make(uint64_t mant, int e)
{
    return __longlong_as_double((long long)(((uint64_t)(1023 + e) << 52) | (mant & 0xFFFFFFFFFFFFFull)));
}

__global__ void sweep(unsigned long long per_thread, int erange, unsigned long long *bad)
{
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned long long local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned long long it = 0; it < per_thread; ++it) {
        const uint64_t r1 = mix(tid * per_thread * 2 + 2 * it), r2 = mix(tid * per_thread * 2 + 2 * it + 1);
        uint64_t md = r1, mx = r2;
        const int mode = (int)(it & 7);
        if (mode == 1) md |= 0xFFFFFFFFFF000ull;
        if (mode == 2) md &= 0xFFFull;
        if (mode == 3) mx |= 0xFFFFFFFFFF000ull;
        if (mode == 4) mx &= 0xFFFull;
        if (mode == 5) md |= 0xFFFFFFFFFFFF0ull;
        const int ed = (int)((r1 >> 52) % (2 * erange + 1)) - erange;
        const int ex = (int)((r2 >> 52) % (2 * erange + 1)) - erange;
        const double d = make(md, ed);
        double x = make(mx, ex);
        if (mode == 6) x = 1.0;
        if (mode == 7) x = -x;
        const double ref = __ddiv_rn(x, d);
        const double got = div_by(x, d, div_rcp(d));
        if (__double_as_longlong(ref) != __double_as_longlong(got)) local[mode]++;
    }
    for (int m = 0; m < 8; ++m)
        if (local[m]) atomicAdd(bad + m, local[m]);
}

int main()
{
    unsigned long long *bad, host[8];
    cudaMalloc(&bad, sizeof(host));
    int rc = 0;
    const int ranges[3] = {60, 200, 450};
    for (int r = 0; r < 3; ++r) {
        cudaMemset(bad, 0, sizeof(host));
        const unsigned long long per_thread = 1 << 14;
        sweep<<<148 * 8, 256>>>(per_thread, ranges[r], bad);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 2; }
        cudaMemcpy(host, bad, sizeof(host), cudaMemcpyDeviceToHost);
        unsigned long long total = 0;
        for (int m = 0; m < 8; ++m) total += host[m];
        printf("exponents +-%d: %llu quotients, mismatches by pattern:", ranges[r], 148ull * 8 * 256 * per_thread);
        for (int m = 0; m < 8; ++m) printf(" %llu", host[m]);
        printf("\n");
        if (total) rc = 1;
    }
    return rc;
}
