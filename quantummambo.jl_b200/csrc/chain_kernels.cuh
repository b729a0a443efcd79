// chain_kernels.cuh -- the orbital-rotation builder as two persistent kernels (sm_100a).
//
//   k_chain_u : theta -> K, ||K||_1 -> s, X = K/2^s, U = exp(K) by Taylor-16 (Paterson-Stockmeyer) + s squarings
//               (reference: unitaries.jl:31-46, `exp(Ulog)`), keeping every intermediate for the derivative;
//   k_chain_d : Lt = Dexp(K)[G^T] by the product rule over the same sequence, then the gradient in x order
//               (reference: del_u_theta gradient.jl:76-129 contracted as in grad_theta :131-146; `vcat` :148-151).
//
// One CTA owns one 16x16 tile of every N x N matrix of the chain; the ~5+s dependent products are separated by a
// grid-wide barrier (one atomic counter in global memory) instead of kernel boundaries, which removes ~30 launch
// gaps per evaluation.  Products run on the FP64 tensor pipe (DMMA.8x8x4): 8 warps = 4 (8x8 blocks) x 2 (k halves).
// The kernels are launched cooperatively (all CTAs co-resident; grid <= 121 CTAs for N <= 168).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace chain {

constexpr int CT = 16;       // tile edge
constexpr int NTHR = 256;    // 8 warps
constexpr int LDB = 24;      // pitch of a B strip row (16 columns + pad): 192 B == 64 (mod 128) -> conflict-free fragments
constexpr int SMAXC = 14;

__constant__ double c_t[17] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                               1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0, 1.0 / 87178291200.0,
                               1.0 / 1307674368000.0, 1.0 / 20922789888000.0};

struct UParams {
    const double* x;
    int n, lam_off, s_override;
    double *X, *X2, *X4, *P0, *P1, *P2, *Q3, *Q2, *Q1, *R, *Lam;
    int* s_out;
    double* scale_out;
    int* status;
    unsigned* bar;  // [0] barrier counter, [1] exit counter (both zero between launches)
};
struct DParams {
    const double* part;   // [cost, S, G]
    const double* g1;     // CSA_SD one-body gradient (or null)
    int n, sd;
    const double *X, *X2, *X4, *Q3, *Q2, *Q1, *R;
    double *dX2, *dX4, *dP0, *dP1, *dP2, *dQ3, *dQ2, *dQ1, *dR;
    const int* s_in;
    const double* scale_in;
    double* out;          // [cost, grad...]
    unsigned* bar;
};

__host__ __device__ inline int lda_for(int n) { return ((n + 15) / 16) * 16 + 4; }
__host__ __device__ inline int kpad(int n) { return ((n + 3) / 4) * 4; }
inline size_t smem_bytes(int n) { return ((size_t)2 * CT * lda_for(n) + (size_t)4 * kpad(n) * LDB + 4 * 32 * 4) * sizeof(double); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Grid-wide barrier: release-reduction on arrival, relaxed (L2) polling, one acquire fence on the way out.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
    __syncthreads();
    target += gridDim.x * gridDim.y;
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void grid_exit(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(bar + 1, 1u);
        if (old == gridDim.x * gridDim.y - 1) {   // last CTA out: everyone has passed every barrier
            bar[0] = 0;
            bar[1] = 0;
            __threadfence();
        }
    }
}

// rows [i0, i0+16) of a row-major n x n matrix -> As[r*lda + k], zero padded to kpad(n)
__device__ __forceinline__ void load_a(double* As, const double* __restrict__ M, int n, int i0) {
    const int lda = lda_for(n), kp = kpad(n);
    const int r = threadIdx.x >> 4;                 // 16 rows x 16 threads per row: no runtime division
    const bool rok = i0 + r < n;
    const double* src = M + (size_t)(i0 + r) * n;
    for (int k = threadIdx.x & 15; k < kp; k += 16) {
        if (rok && k < n) cp8(As + r * lda + k, src + k);
        else As[r * lda + k] = 0.0;
    }
}
// columns [j0, j0+16) -> Bs[k*LDB + c]
__device__ __forceinline__ void load_b(double* Bs, const double* __restrict__ M, int n, int j0) {
    const int kp = kpad(n);
    for (int idx = threadIdx.x; idx < CT * kp; idx += NTHR) {
        const int k = idx / CT, c = idx % CT;
        if (k < n && j0 + c < n) cp8(Bs + k * LDB + c, M + (size_t)k * n + j0 + c);
        else Bs[k * LDB + c] = 0.0;
    }
}
// this warp's share (k half) of the 8x8 block product: c += As[block rows] . Bs[block cols]
__device__ __forceinline__ void mm_half(const double* As, const double* Bs, int n, double& c0, double& c1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2, b = warp & 3, half = warp >> 2;
    const int lda = lda_for(n), ks = kpad(n) / 4;
    const double* a = As + (8 * (b >> 1) + t1) * lda + t0;
    const double* bb = Bs + t0 * LDB + 8 * (b & 1) + t1;
    double d0 = 0.0, d1 = 0.0;                       // second accumulator: halves the dependent DMMA chain
    int kk = half;
    for (; kk + 2 < ks; kk += 4) {
        dmma(c0, c1, a[4 * kk], bb[4 * kk * LDB]);
        dmma(d0, d1, a[4 * (kk + 2)], bb[4 * (kk + 2) * LDB]);
    }
    if (kk < ks) dmma(c0, c1, a[4 * kk], bb[4 * kk * LDB]);
    c0 += d0;
    c1 += d1;
}
// combine the two k halves; afterwards warps 0..3 hold the full sums (warps 4..7 must not use them)
__device__ __forceinline__ void combine(double* red, double* v, int nv) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= 4)
        for (int i = 0; i < nv; ++i) red[((warp - 4) * 32 + lane) * 4 + i] = v[i];
    __syncthreads();
    if (warp < 4)
        for (int i = 0; i < nv; ++i) v[i] += red[(warp * 32 + lane) * 4 + i];
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR, 1) k_chain_u(UParams p) {
    extern __shared__ double sm[];
    const int n = p.n, lda = lda_for(n), kp = kpad(n), nn = n * n;
    double* As0 = sm;
    double* As1 = As0 + CT * lda;
    double* Bs0 = As1 + CT * lda;
    double* Bs1 = Bs0 + kp * LDB;
    double* red = Bs0 + 4 * kp * LDB;
    __shared__ double s_red[8];
    __shared__ double s_scale;
    __shared__ int s_s;
    const int i0 = blockIdx.y * CT, j0 = blockIdx.x * CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2, b = warp & 3;
    const int ei = i0 + 8 * (b >> 1) + t1, ej = j0 + 8 * (b & 1) + 2 * t0;   // this lane's output elements (ei, ej), (ei, ej+1)
    const int cL = n * (n + 1) / 2;
    const double* lam = p.x + p.lam_off;
    const double* th = p.x + p.lam_off + cL;
    unsigned target = 0;

    // ---- ||K||_1 (== max row sum; every CTA computes the same number in the same order) and s ---------------
    {
        double mx = 0.0;
        for (int i = threadIdx.x; i < n; i += NTHR) {
            double s = 0.0;
            for (int j = 0; j < i; ++j) s += fabs(th[j * n - j * (j + 1) / 2 + (i - j - 1)]);
            const double* row = th + (size_t)i * n - (size_t)i * (i + 1) / 2 - i - 1;
            for (int j = i + 1; j < n; ++j) s += fabs(row[j]);
            if (!(s == s)) s = INFINITY;
            mx = fmax(mx, s);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[warp] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0.0;
            for (int w = 0; w < 8; ++w) m = fmax(m, s_red[w]);
            int s = 0;
            if (m > 1.0) s = (int)ceil(log2(m));
            if (p.s_override >= 0) s = p.s_override;
            if (isinf(m) || s > SMAXC || s < 0) {
                if (blockIdx.x == 0 && blockIdx.y == 0) *p.status = 1;
                s = SMAXC;
            }
            s_s = s;
            s_scale = ldexp(1.0, -s);
            if (blockIdx.x == 0 && blockIdx.y == 0) {
                *p.s_out = s;
                *p.scale_out = s_scale;
            }
        }
        __syncthreads();
    }
    const int s = s_s;
    const double sc = s_scale;
    auto kval = [&](int i, int j) -> double {   // X[i][j]
        if (i >= n || j >= n || i == j) return 0.0;
        return (i < j ? th[(size_t)i * n - (size_t)i * (i + 1) / 2 + (j - i - 1)] : -th[(size_t)j * n - (size_t)j * (j + 1) / 2 + (i - j - 1)]) * sc;
    };
    // ---- step 1: X2 = X.X with the strips of X built straight from theta --------------------------------------
    {
        const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
#pragma unroll 4
        for (int k = c; k < kp; k += 16) As0[r * lda + k] = kval(i0 + r, k);
#pragma unroll 4
        for (int k = r; k < kp; k += 16) Bs0[k * LDB + c] = kval(k, j0 + c);
    }
    __syncthreads();
    double v[4] = {0, 0, 0, 0};
    mm_half(As0, Bs0, n, v[0], v[1]);
    combine(red, v, 2);
    double x1[2] = {0, 0}, x2[2] = {0, 0};
    if (warp < 4) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                const int o = i * n + j;
                x1[e] = kval(i, j);
                x2[e] = v[e];
                p.X[o] = x1[e];
                p.X2[o] = x2[e];
                const int a = i > j ? i : j, bq = i > j ? j : i;
                p.Lam[o] = lam[a * (a + 1) / 2 + bq];
            }
        }
    }
    grid_barrier(p.bar, target);
    // ---- step 2: X3 = X2.X, X4 = X2.X2; P0..P2, Q3 ----------------------------------------------------------------
    load_a(As0, p.X2, n, i0);
    load_b(Bs1, p.X2, n, j0);     // Bs0 still holds the columns of X
    cp_wait();
    __syncthreads();
    v[0] = v[1] = v[2] = v[3] = 0.0;
    mm_half(As0, Bs0, n, v[0], v[1]);
    mm_half(As0, Bs1, n, v[2], v[3]);
    combine(red, v, 4);
    if (warp < 4) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                const int o = i * n + j;
                const double id = (i == j) ? 1.0 : 0.0, x3 = v[e], x4 = v[2 + e];
                p.X4[o] = x4;
                p.P0[o] = c_t[0] * id + c_t[1] * x1[e] + c_t[2] * x2[e] + c_t[3] * x3;
                p.P1[o] = c_t[4] * id + c_t[5] * x1[e] + c_t[6] * x2[e] + c_t[7] * x3;
                p.P2[o] = c_t[8] * id + c_t[9] * x1[e] + c_t[10] * x2[e] + c_t[11] * x3;
                p.Q3[o] = c_t[12] * id + c_t[13] * x1[e] + c_t[14] * x2[e] + c_t[15] * x3 + c_t[16] * x4;
            }
        }
    }
    grid_barrier(p.bar, target);
    // ---- steps 3-5: Q2 = P2 + X4.Q3, Q1 = P1 + X4.Q2, R0 = P0 + X4.Q1 (rows of X4 loaded once) ---------------------
    load_a(As0, p.X4, n, i0);
    {
        const double* Bsrc[3] = {p.Q3, p.Q2, p.Q1};
        const double* Csrc[3] = {p.P2, p.P1, p.P0};
        double* Cdst[3] = {p.Q2, p.Q1, p.R};
        for (int st = 0; st < 3; ++st) {
            load_b(Bs0, Bsrc[st], n, j0);
            cp_wait();
            __syncthreads();
            v[0] = v[1] = 0.0;
            mm_half(As0, Bs0, n, v[0], v[1]);
            combine(red, v, 2);
            if (warp < 4) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = ei, j = ej + e;
                    if (i < n && j < n) Cdst[st][i * n + j] = v[e] + Csrc[st][i * n + j];
                }
            }
            grid_barrier(p.bar, target);
        }
    }
    // ---- s squarings: R[q+1] = R[q].R[q] -------------------------------------------------------------------------
    for (int q = 0; q < s; ++q) {
        const double* Rq = p.R + (size_t)q * nn;
        load_a(As0, Rq, n, i0);
        load_b(Bs0, Rq, n, j0);
        cp_wait();
        __syncthreads();
        v[0] = v[1] = 0.0;
        mm_half(As0, Bs0, n, v[0], v[1]);
        combine(red, v, 2);
        if (warp < 4) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = ei, j = ej + e;
                if (i < n && j < n) p.R[(size_t)(q + 1) * nn + i * n + j] = v[e];
            }
        }
        grid_barrier(p.bar, target);
    }
    grid_exit(p.bar);
}

// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR, 1) k_chain_d(DParams p) {
    extern __shared__ double sm[];
    const int n = p.n, lda = lda_for(n), kp = kpad(n), nn = n * n;
    double* As0 = sm;
    double* As1 = As0 + CT * lda;
    double* Bs0 = As1 + CT * lda;
    double* Bs1 = Bs0 + kp * LDB;
    double* Bs2 = Bs1 + kp * LDB;
    double* Bs3 = Bs2 + kp * LDB;
    double* red = Bs0 + 4 * kp * LDB;
    const int i0 = blockIdx.y * CT, j0 = blockIdx.x * CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2, b = warp & 3;
    const int ei = i0 + 8 * (b >> 1) + t1, ej = j0 + 8 * (b & 1) + 2 * t0;
    const int s = *p.s_in;
    const double sc = *p.scale_in;
    const double* G = p.part + 1 + nn;
    unsigned target = 0;
    auto dxval = [&](int i, int j) -> double { return (i < n && j < n) ? G[(size_t)j * n + i] * sc : 0.0; };   // dX = G^T 2^-s

    // ---- step 1: dX2 = dX.X + X.dX --------------------------------------------------------------------------------
    {
        const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
#pragma unroll 4
        for (int k = c; k < kp; k += 16) As0[r * lda + k] = dxval(i0 + r, k);
#pragma unroll 4
        for (int k = r; k < kp; k += 16) Bs1[k * LDB + c] = dxval(k, j0 + c);
    }
    load_a(As1, p.X, n, i0);
    load_b(Bs0, p.X, n, j0);
    cp_wait();
    __syncthreads();
    double v[4] = {0, 0, 0, 0};
    mm_half(As0, Bs0, n, v[0], v[1]);   // dX.X
    mm_half(As1, Bs1, n, v[0], v[1]);   // X.dX
    combine(red, v, 2);
    double d1[2] = {0, 0}, d2[2] = {0, 0};
    if (warp < 4) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                d1[e] = dxval(i, j);
                d2[e] = v[e];
                p.dX2[i * n + j] = v[e];
            }
        }
    }
    grid_barrier(p.bar, target);
    // ---- step 2: dX3 = dX2.X + X2.dX ; dX4 = dX2.X2 + X2.dX2 ; dP0..dP2, dQ3 -----------------------------------------
    // Bs0 = cols of X (kept), Bs1 = cols of dX (kept); As0 <- rows of dX2, As1 <- rows of X2, Bs2 <- cols X2, Bs3 <- cols dX2
    __syncthreads();
    load_a(As0, p.dX2, n, i0);
    load_a(As1, p.X2, n, i0);
    load_b(Bs2, p.X2, n, j0);
    load_b(Bs3, p.dX2, n, j0);
    cp_wait();
    __syncthreads();
    v[0] = v[1] = v[2] = v[3] = 0.0;
    mm_half(As0, Bs0, n, v[0], v[1]);   // dX2.X
    mm_half(As1, Bs1, n, v[0], v[1]);   // X2.dX
    mm_half(As0, Bs2, n, v[2], v[3]);   // dX2.X2
    mm_half(As1, Bs3, n, v[2], v[3]);   // X2.dX2
    combine(red, v, 4);
    if (warp < 4) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                const int o = i * n + j;
                const double d3 = v[e], d4 = v[2 + e];
                p.dX4[o] = d4;
                p.dP0[o] = c_t[1] * d1[e] + c_t[2] * d2[e] + c_t[3] * d3;
                p.dP1[o] = c_t[5] * d1[e] + c_t[6] * d2[e] + c_t[7] * d3;
                p.dP2[o] = c_t[9] * d1[e] + c_t[10] * d2[e] + c_t[11] * d3;
                p.dQ3[o] = c_t[13] * d1[e] + c_t[14] * d2[e] + c_t[15] * d3 + c_t[16] * d4;
            }
        }
    }
    grid_barrier(p.bar, target);
    // ---- steps 3-5: dQ2 = dP2 + dX4.Q3 + X4.dQ3 ; dQ1 = dP1 + dX4.Q2 + X4.dQ2 ; dR0 = dP0 + dX4.Q1 + X4.dQ1 ----------
    load_a(As0, p.dX4, n, i0);
    load_a(As1, p.X4, n, i0);
    {
        const double* B1[3] = {p.Q3, p.Q2, p.Q1};
        const double* B2[3] = {p.dQ3, p.dQ2, p.dQ1};
        const double* Csrc[3] = {p.dP2, p.dP1, p.dP0};
        double* Cdst[3] = {p.dQ2, p.dQ1, p.dR};
        for (int st = 0; st < 3; ++st) {
            load_b(Bs0, B1[st], n, j0);
            load_b(Bs1, B2[st], n, j0);
            cp_wait();
            __syncthreads();
            v[0] = v[1] = 0.0;
            mm_half(As0, Bs0, n, v[0], v[1]);
            mm_half(As1, Bs1, n, v[0], v[1]);
            combine(red, v, 2);
            if (warp < 4) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = ei, j = ej + e;
                    if (i < n && j < n) Cdst[st][i * n + j] = v[e] + Csrc[st][i * n + j];
                }
            }
            grid_barrier(p.bar, target);
        }
    }
    // ---- s squarings: dR[q+1] = dR[q].R[q] + R[q].dR[q] ---------------------------------------------------------------
    for (int q = 0; q < s; ++q) {
        const double* Rq = p.R + (size_t)q * nn;
        const double* dRq = p.dR + (size_t)q * nn;
        load_a(As0, dRq, n, i0);
        load_a(As1, Rq, n, i0);
        load_b(Bs0, Rq, n, j0);
        load_b(Bs1, dRq, n, j0);
        cp_wait();
        __syncthreads();
        v[0] = v[1] = 0.0;
        mm_half(As0, Bs0, n, v[0], v[1]);
        mm_half(As1, Bs1, n, v[0], v[1]);
        combine(red, v, 2);
        if (warp < 4) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = ei, j = ej + e;
                if (i < n && j < n) p.dR[(size_t)(q + 1) * nn + i * n + j] = v[e];
            }
        }
        grid_barrier(p.bar, target);
    }
    // ---- gradient in x order (gradient.jl:148-151, 287-290): out = [cost, (g1), g_lambda, g_theta] -------------------
    {
        const double* Lt = p.dR + (size_t)s * nn;
        const int cL = n * (n + 1) / 2, off = p.sd ? n : 0;
        for (int idx = threadIdx.x; idx < CT * CT; idx += NTHR) {
            const int i = i0 + idx / CT, j = j0 + idx % CT;
            if (i < n && j < n) {
                if (j <= i) p.out[1 + off + i * (i + 1) / 2 + j] = 2.0 * p.part[1 + i * n + j] * (i != j ? 2.0 : 1.0);
                else p.out[1 + off + cL + i * n - i * (i + 1) / 2 + (j - i - 1)] = 2.0 * (Lt[j * n + i] - Lt[i * n + j]);
            }
        }
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            if (threadIdx.x == 0) p.out[0] = p.part[0];
            if (p.sd)
                for (int k = threadIdx.x; k < n; k += NTHR) p.out[1 + k] = p.g1[k];
        }
    }
    grid_exit(p.bar);
}

}  // namespace chain
