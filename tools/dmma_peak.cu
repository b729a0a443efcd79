// Micro-benchmark: raw FP64 tensor-core (DMMA) issue rate on sm_100a for each mma.sync f64 shape,
// plus plain DFMA, so the roofline denominator for the CSA kernels is a measured number.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int SHAPE> __device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  if constexpr (SHAPE == 0) { // m8n8k4 (two of them to fill 4 acc regs)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n" : "+d"(c[2]), "+d"(c[3]) : "d"(a[1]), "d"(b[0]));
  } else if constexpr (SHAPE == 1) { // m16n8k4
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  } else if constexpr (SHAPE == 2) { // m16n8k8
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  } else { // m16n8k16
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
}

template<int SHAPE, int NACC> __global__ void k_mma(double* out, int iters, double seed) {
  double c[NACC][4]; double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + i) * 1e-3;
  for (int i = 0; i < 4; i++) b[i] = seed * (threadIdx.x - i) * 1e-3;
  for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) c[j][i] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) mma<SHAPE>(c[j], a, b);
  }
  double s = 0; for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int NACC> __global__ void k_fma(double* out, int iters, double seed) {
  double c[NACC]; double a = seed * threadIdx.x * 1e-3, b = seed * 0.5;
  for (int j = 0; j < NACC; j++) c[j] = j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) c[j] = fma(a, c[j], b);
  }
  double s = 0; for (int j = 0; j < NACC; j++) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename F> float timeit(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount; printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 20000;
  const char* names[4] = {"m8n8k4(x2)", "m16n8k4", "m16n8k8", "m16n8k16"};
  const double fl[4] = {2.0 * 2 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
  for (int warps = 4; warps <= 16; warps *= 2) {
    int threads = warps * 32 > 512 ? 512 : warps * 32; int ctas_per_sm = warps * 32 / threads;
    int grid = sms * ctas_per_sm;
#define RUN(S, NACC) { float ms = timeit([&]{ k_mma<S, NACC><<<grid, threads>>>(out, iters, 1.0); }); \
      double tf = fl[S] * NACC * iters * (double)grid * (threads / 32) / (ms * 1e-3) / 1e12; \
      printf("warps/SM %2d shape %-11s nacc %d : %8.3f ms  %7.2f TFLOP/s\n", warps, names[S], NACC, ms, tf); }
    RUN(0, 4) RUN(0, 8) RUN(1, 4) RUN(1, 8) RUN(2, 4) RUN(2, 8) RUN(3, 4) RUN(3, 8)
    { float ms = timeit([&]{ k_fma<16><<<grid, threads>>>(out, iters, 1.0); });
      double tf = 2.0 * 16 * iters * (double)grid * threads / (ms * 1e-3) / 1e12;
      printf("warps/SM %2d DFMA nacc 16          : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, tf); }
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
