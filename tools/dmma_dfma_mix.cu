// Micro-benchmark: do DMMA.8x8x4 and plain DFMA share one FP64 datapath on sm_100a, or can they be issued side by side?
// Each warp runs NDM independent DMMA accumulator chains and NDF independent DFMA chains in the same loop; a second mode
// gives half of the warps only DMMA and the other half only DFMA.  If the pipes were separate the combined rate would
// approach the sum of the two stand-alone rates (~37 + ~36 TFLOP/s, profiles/r01_dmma_peak.txt).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NDM, int NDF, bool SPLIT>
__global__ void k_mix(double* out, int iters, double seed) {
    double c[NDM > 0 ? NDM : 1][2], f[NDF > 0 ? NDF : 1];
    const double a = seed * (threadIdx.x + 1) * 1e-3, b = seed * 0.5;
    for (int j = 0; j < NDM; ++j) c[j][0] = c[j][1] = 0.0;
    for (int j = 0; j < NDF; ++j) f[j] = j;
    const int warp = threadIdx.x >> 5;
    const bool do_m = !SPLIT || (warp & 1) == 0, do_f = !SPLIT || (warp & 1) == 1;
    for (int it = 0; it < iters; ++it) {
        if (do_m) {
#pragma unroll
            for (int j = 0; j < NDM; ++j) dmma(c[j][0], c[j][1], a, b);
        }
        if (do_f) {
#pragma unroll
            for (int j = 0; j < NDF; ++j) asm volatile("fma.rn.f64 %0, %1, %0, %2;" : "+d"(f[j]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
    for (int j = 0; j < NDM; ++j) s += c[j][0] + c[j][1];
    for (int j = 0; j < NDF; ++j) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s sms %d\n", p.name, sms);
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * sms * 2 * 1024));
    const int iters = 20000, threads = 512, grid = sms;   // 16 warps per SM
#define RUN(NDM, NDF, SPLIT)                                                                                                    \
    {                                                                                                                           \
        float ms = timeit([&] { k_mix<NDM, NDF, SPLIT><<<grid, threads>>>(out, iters, 1.0); });                               \
        const double wm = SPLIT ? 0.5 : 1.0;                                                                                    \
        const double fm = 2.0 * 8 * 8 * 4 * NDM * wm, ff = 2.0 * 32 * NDF * wm;                                                 \
        const double per = (double)iters * grid * (threads / 32) / (ms * 1e-3) / 1e12;                                          \
        printf("%s dmma chains %2d dfma chains %2d : %8.3f ms  DMMA %6.2f + DFMA %6.2f = %6.2f TFLOP/s\n", SPLIT ? "split-warps" : "same-warp  ", \
               NDM, NDF, ms, fm * per, ff * per, (fm + ff) * per);                                                              \
    }
    RUN(8, 0, false)
    RUN(0, 16, false)
    RUN(8, 8, false)
    RUN(8, 16, false)
    RUN(8, 32, false)
    RUN(4, 32, false)
    RUN(8, 64, false)
    RUN(8, 16, true)
    RUN(8, 64, true)
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
