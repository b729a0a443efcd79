// lanczos_kernels.cuh -- leading eigenpair of the device-resident residual (sm_100a).
//
// The reference's initial guess for a greedy step, tbt_svd_1st (guesses.jl:51-84), diagonalises the whole
// N^2 x N^2 matrix view of the residual (`eigen(Symmetric(reshape(tbt,(N^2,N^2))))`, O(N^6)) and keeps ONE eigenpair:
// the one of largest |eigenvalue| (guesses.jl:60-63).  Here that pair comes from a Lanczos iteration with full
// re-orthogonalisation on the matrix the engine already holds in HBM (T~: packed M x M in the 8-fold-symmetric
// mode, N^2 x N^2 otherwise).  The only O(R^2) work per step is the symmetric matrix-vector product below, an
// HBM-bound stream over T~ (8 R^2 bytes, 2 R^2 flop -> 0.25 flop/B); everything else is O(R m).
//
//   k_symv     : y = T~ v            one CTA per row, 16-byte streaming loads, v through L1/L2 (HBM roofline)
//   k_dots     : c[k] = V[k] . w     one CTA per Krylov vector (classical Gram-Schmidt, applied twice)
//   k_project  : w -= sum_k c[k] V[k]  (and accumulates the coefficients: alpha_j is c[j])
//   k_normalise: beta = ||w||, V[j+1] = w / beta
//   k_combine  : Ritz vector sum_k s[k] V[k] scattered to the reference's N x N column-major layout
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lanczos {

// y[r] = sum_j T[r][j] v[j] for r in [0, rows).  T row-major with pitch ldT (multiple of 64, zero padded), v has ldT
// entries (zero padded).  One CTA per group of RPC = 4 consecutive rows (grid-stride over groups): the 8 warps take
// interleaved 512-byte segments of each row, up to four 16-byte streaming loads in flight per thread; the four row sums go
// through ONE shuffle + shared-memory combine per group (fixed order: bit-reproducible), so the reduction epilogue and its
// two block barriers are paid once per 160 KB instead of once per 40 KB row.
// GENERAL != 0: the matrix is Symmetric(upper triangle) of a non-symmetric tensor view; T2 holds the transpose, so
// element (r, j<r) is taken from T2 (== T[j][r]).
constexpr int SYMV_RPC = 4;
template <int GENERAL>
__global__ void __launch_bounds__(256) k_symv(const double* __restrict__ T, const double* __restrict__ T2, long long ldT, int rows,
                                              const double* __restrict__ v, double* __restrict__ y) {
    __shared__ double red[SYMV_RPC][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npair = (int)(ldT >> 1);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    const int ngroups = (rows + SYMV_RPC - 1) / SYMV_RPC;
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        double sum[SYMV_RPC];
#pragma unroll
        for (int rr = 0; rr < SYMV_RPC; ++rr) {
            const int r = grp * SYMV_RPC + rr;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (r < rows) {
                const double2* t = reinterpret_cast<const double2*>(T + (long long)r * ldT);
                const double2* u = GENERAL ? reinterpret_cast<const double2*>(T2 + (long long)r * ldT) : nullptr;
                int j = threadIdx.x;
                for (; j + 768 < npair; j += 1024) {
                    double2 x0 = __ldcs(t + j), x1 = __ldcs(t + j + 256), x2 = __ldcs(t + j + 512), x3 = __ldcs(t + j + 768);
                    if (GENERAL) {
                        const double2 z0 = __ldcs(u + j), z1 = __ldcs(u + j + 256), z2 = __ldcs(u + j + 512), z3 = __ldcs(u + j + 768);
                        if (2 * j < r) x0.x = z0.x;
                        if (2 * j + 1 < r) x0.y = z0.y;
                        if (2 * (j + 256) < r) x1.x = z1.x;
                        if (2 * (j + 256) + 1 < r) x1.y = z1.y;
                        if (2 * (j + 512) < r) x2.x = z2.x;
                        if (2 * (j + 512) + 1 < r) x2.y = z2.y;
                        if (2 * (j + 768) < r) x3.x = z3.x;
                        if (2 * (j + 768) + 1 < r) x3.y = z3.y;
                    }
                    const double2 w0 = __ldg(v2 + j), w1 = __ldg(v2 + j + 256), w2 = __ldg(v2 + j + 512), w3 = __ldg(v2 + j + 768);
                    a0 = fma(x0.x, w0.x, fma(x0.y, w0.y, a0));
                    a1 = fma(x1.x, w1.x, fma(x1.y, w1.y, a1));
                    a2 = fma(x2.x, w2.x, fma(x2.y, w2.y, a2));
                    a3 = fma(x3.x, w3.x, fma(x3.y, w3.y, a3));
                }
                for (; j < npair; j += 256) {
                    double2 x0 = __ldcs(t + j);
                    if (GENERAL) {
                        const double2 z0 = __ldcs(u + j);
                        if (2 * j < r) x0.x = z0.x;
                        if (2 * j + 1 < r) x0.y = z0.y;
                    }
                    const double2 w0 = __ldg(v2 + j);
                    a0 = fma(x0.x, w0.x, fma(x0.y, w0.y, a0));
                }
            }
            sum[rr] = (a0 + a1) + (a2 + a3);
        }
#pragma unroll
        for (int rr = 0; rr < SYMV_RPC; ++rr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum[rr] += __shfl_xor_sync(0xffffffffu, sum[rr], o);
            if (lane == 0) red[rr][warp] = sum[rr];
        }
        __syncthreads();
        if (threadIdx.x < SYMV_RPC) {
            const int r = grp * SYMV_RPC + threadIdx.x;
            if (r < rows) {
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < 8; ++q) s += red[threadIdx.x][q];
                y[r] = s;
            }
        }
        __syncthreads();
    }
}

// c[k] = V[k] . w  (k < nv); V vectors have pitch ldv.  One CTA per vector, fixed-order reduction.
static __global__ void __launch_bounds__(256) k_dots(const double* __restrict__ V, long long ldv, int len, const double* __restrict__ w,
                                              double* __restrict__ c) {
    __shared__ double red[8];
    const double* vk = V + (long long)blockIdx.x * ldv;
    double a = 0.0;
    for (int i = threadIdx.x; i < len; i += 256) a = fma(vk[i], w[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q];
        c[blockIdx.x] = s;
    }
}

// w -= sum_k c[k] V[k]; block 0 also accumulates the coefficient of the newest vector: alpha[j] (+)= c[j]
static __global__ void __launch_bounds__(256) k_project(const double* __restrict__ V, long long ldv, int len, int nv, const double* __restrict__ c,
                                                 double* __restrict__ w, double* __restrict__ alpha, int j, int accumulate) {
    extern __shared__ double sc[];
    for (int k = threadIdx.x; k < nv; k += 256) sc[k] = c[k];
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < len) {
        double a = w[i];
        for (int k = 0; k < nv; ++k) a = fma(-sc[k], V[(long long)k * ldv + i], a);
        w[i] = a;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) alpha[j] = accumulate ? alpha[j] + sc[j] : sc[j];
}

// beta[j] = ||w||; vnext = w / beta (zero when w vanished: invariant subspace reached, the host stops on beta)
static __global__ void __launch_bounds__(1024) k_normalise(const double* __restrict__ w, int len, int lenpad, double* __restrict__ vnext,
                                                    double* __restrict__ beta, int j) {
    __shared__ double red[32];
    __shared__ double s_inv;
    double a = 0.0;
    for (int i = threadIdx.x; i < len; i += 1024) a = fma(w[i], w[i], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < 32; ++q) s += red[q];
        const double b = sqrt(s);
        beta[j] = b;
        s_inv = b > 0.0 ? 1.0 / b : 0.0;
    }
    __syncthreads();
    const double inv = s_inv;
    for (int i = threadIdx.x; i < lenpad; i += 1024) vnext[i] = i < len ? w[i] * inv : 0.0;
}

// start vector: deterministic and positive.  Packed mode (rows are the pairs p <= q) needs nothing else.  In the unpacked modes the
// key distinguishes (p,q) from (q,p): a start vector that is symmetric under p <-> q would confine the whole Krylov space to the
// symmetric subspace whenever the tensor commutes with that swap, and a dominant ANTI-symmetric eigenvector -- the anti-Hermitian
// branch of tbt_svd_1st (guesses.jl:67-75) -- would never be found.  normalised by k_normalise afterwards.
static __global__ void k_start(double* __restrict__ w, int len, const int* __restrict__ pa, const int* __restrict__ pq, int n, int packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) {
        const int p = pa[i], q = pq[i];
        const int key = packed ? min(p, q) * n + max(p, q) : p * n + q;
        unsigned long long z = 0x9E3779B97F4A7C15ull * (unsigned long long)(key + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        w[i] = 1.0 + (double)(z >> 11) * (1.0 / 9007199254740992.0);   // in [1, 2): never orthogonal to a positive vector
    }
}

// Ritz vector y = sum_k s[k] V[k] (rows of the matrix) scattered to the N x N column-major operator of the reference:
// packed mode undoes the sqrt(2) row scaling and mirrors (p,q) -> (q,p).
static __global__ void __launch_bounds__(256) k_combine(const double* __restrict__ V, long long ldv, int len, int nv, const double* __restrict__ s,
                                                 const int* __restrict__ pa, const int* __restrict__ pq, const double* __restrict__ sw,
                                                 int n, int packed, double* __restrict__ out) {
    extern __shared__ double sc[];
    for (int k = threadIdx.x; k < nv; k += 256) sc[k] = s[k];
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    double a = 0.0;
    for (int k = 0; k < nv; ++k) a = fma(sc[k], V[(long long)k * ldv + i], a);
    const int p = pa[i], q = pq[i];
    if (packed) {
        a /= sw[i];
        out[p + (long long)n * q] = a;
        out[q + (long long)n * p] = a;
    } else {
        out[p + (long long)n * q] = a;
    }
}

}  // namespace lanczos
