// DMMA.8x8x4 dependent-issue latency and single-warp throughput on sm_100a (clock64 around unrolled chains).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC> __global__ void k(double* out, long long* cyc, int iters) {
    double c[NACC][2];
    for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) dmma(c[j][0], c[j][1], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 256); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    long long h;
#define RUN(NACC, WARPS) { k<NACC><<<1, 32 * WARPS>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("warps %2d nacc %d: %.1f cycles per DMMA per warp (%.1f per dependent step)\n", WARPS, NACC, (double)h / iters / NACC, (double)h / iters); }
    RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(8, 4) RUN(1, 8) RUN(4, 8) RUN(8, 8) RUN(8, 16)
    return 0;
}
