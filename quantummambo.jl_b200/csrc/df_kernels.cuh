// df_kernels.cuh -- double factorisation of the device-resident residual (sm_100a).
//
// Reference: DF_decomposition (/root/reference/src/UTILS/decompose.jl:297-398) diagonalises the whole N^2 x N^2 matrix view of
// the two-body tensor with LAPACK (`eigen`, O(N^6)), keeps every eigenpair with |eigenvalue| >= SVD_tol and then
// diagonalises each reshaped eigenvector (an N x N symmetric matrix) again; DF_based_greedy (:400-436) folds all fragments
// into the orbital frame of the first one.  Here everything runs on the matrix the engine already holds in HBM:
//
//   tridiagonalisation : Lanczos with full re-orthogonalisation run to completion (df.cu) on lanczos::k_symv (HBM-bound);
//                        k_project_split / k_apply_part / k_scale_norm below are the bit-reproducible projection of that driver
//   k_bisect           : all eigenvalues of the (block-diagonal) tridiagonal, one thread per eigenvalue (Sturm counts)
//   k_twisted          : eigenvectors of the tridiagonal by twisted factorisation, one thread per eigenvector
//   k_dgemm<TA>        : FP64 tensor-core GEMM (DMMA.8x8x4) C = op(A) B for the Gram matrix Z^T Z, the Newton-Schulz
//                        re-orthonormalisation Z <- Z (3 I - Z^T Z) / 2 and the back-transformation Y = Z^T V
//   k_frag_matrix      : eigenvector (row space of the engine) -> the N x N operator L of the reference (column-major)
//   k_jacobi_batch     : eigen-decomposition of every fragment's N x N matrix, one CTA per fragment, parallel-order
//                        two-sided Jacobi in shared memory (the `eigen(cur_l)` of decompose.jl:354)
//   k_df_frame / k_df_lmat : DF_based_greedy's Cartan matrix in the frame of the leading fragment
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace df {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- re-orthogonalisation of the Lanczos driver ---------------------------------------------------------------------------
// part[ks][i] = sum_{k in chunk ks} c[k] V[k][i]   (grid: (ceil(len/256), KS); chunk = ceil(nv / KS) vectors)
// A single thread per element with all nv vectors in sequence (lanczos::k_project) keeps only len/256 CTAs busy; the
// split over KS chunks fills the device and the fixed-order combine in k_apply_part keeps the result bit-reproducible.
static __global__ void __launch_bounds__(256) k_project_split(const double* __restrict__ V, long long ldv, int len, int nv, const double* __restrict__ c,
                                                       double* __restrict__ part, int lenpad) {
    __shared__ double sc[512];
    const int KS = gridDim.y, chunk = (nv + KS - 1) / KS;
    const int k0 = blockIdx.y * chunk, k1 = min(nv, k0 + chunk);
    const int i = blockIdx.x * 256 + threadIdx.x;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int kb = k0; kb < k1; kb += 512) {
        const int kn = min(512, k1 - kb);
        __syncthreads();
        for (int k = threadIdx.x; k < kn; k += 256) sc[k] = c[kb + k];
        __syncthreads();
        if (i < len) {
            const double* v = V + (long long)kb * ldv + i;
            int k = 0;
            for (; k + 3 < kn; k += 4) {
                const double v0 = __ldcs(v + (long long)k * ldv), v1 = __ldcs(v + (long long)(k + 1) * ldv);
                const double v2 = __ldcs(v + (long long)(k + 2) * ldv), v3 = __ldcs(v + (long long)(k + 3) * ldv);
                a0 = fma(sc[k], v0, a0);
                a1 = fma(sc[k + 1], v1, a1);
                a2 = fma(sc[k + 2], v2, a2);
                a3 = fma(sc[k + 3], v3, a3);
            }
            for (; k < kn; ++k) a0 = fma(sc[k], __ldcs(v + (long long)k * ldv), a0);
        }
    }
    if (i < len) part[(long long)blockIdx.y * lenpad + i] = (a0 + a1) + (a2 + a3);
}

// w <- w - sum_ks part[ks]  (fixed order) and the per-CTA partial sums of |w|^2  (grid: ceil(len / 256) CTAs of 256 threads)
static __global__ void __launch_bounds__(256) k_apply_part(double* w, int len, int lenpad, const double* __restrict__ part, int KS,
                                                           double* __restrict__ npart) {
    __shared__ double red[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double v = 0.0;
    if (i < len) {
        double s = 0.0;
        for (int ks = 0; ks < KS; ++ks) s += part[(long long)ks * lenpad + i];
        v = w[i] - s;
        w[i] = v;
    }
    double a = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q];
        npart[blockIdx.x] = t;
    }
}
// nrm = sqrt(sum of the partials, fixed order), vnext = w / nrm (zero in the padding); grid: ceil(lenpad / 256) CTAs.
//   mode 0 (first pass of step j):  alpha[j] = c[j];  beta[j] = nrm
//   mode 1 (second pass, on the unit vector): alpha[j] += beta[j] * c[j];  beta[j] *= nrm
//   mode 2 (restart vector): neither alpha nor beta is touched; rnorm[0] = nrm
// w may alias vnext (the second pass works in place on V[j+1]).  Only CTA 0 touches alpha / beta / rnorm.
static __global__ void __launch_bounds__(256) k_scale_norm(const double* w, int len, int lenpad, const double* __restrict__ npart, int nparts,
                                                           double* vnext, const double* __restrict__ c, double* __restrict__ alpha,
                                                           double* __restrict__ beta, int j, int mode, double* __restrict__ rnorm) {
    __shared__ double s_inv;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < nparts; ++q) s += npart[q];
        const double nrm = sqrt(s);
        if (blockIdx.x == 0) {
            if (mode == 0) {
                alpha[j] = c[j];
                beta[j] = nrm;
            } else if (mode == 1) {
                alpha[j] += beta[j] * c[j];
                beta[j] *= nrm;
            } else {
                rnorm[0] = nrm;
            }
        }
        s_inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < lenpad) vnext[i] = i < len ? w[i] * s_inv : 0.0;
}

// pseudo-random restart vector (deterministic in `seed`), entries in (-1, 1)
static __global__ void k_random_vec(double* __restrict__ w, int len, int lenpad, unsigned long long seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lenpad) return;
    unsigned long long z = 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1) + seed * 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    w[i] = i < len ? 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0 : 0.0;
}

// ---- eigenvalues of the block-diagonal tridiagonal --------------------------------------------------------------------------
// The Lanczos recurrence breaks down (beta = 0) whenever a Krylov space is exhausted (degenerate eigenvalues, low rank); every
// restart opens a new unreduced block.  blk[3 b + {0,1,2}] = (first row, rows, first eigenvalue slot) of block b; thread e
// finds the (e - slot)-th smallest eigenvalue of its block by bisection on the Sturm count (LAPACK dlaebz's recurrence and
// pivmin guard).  glo / ghi: Gershgorin bounds of the block.
static __global__ void __launch_bounds__(64) k_bisect(const double* __restrict__ alpha, const double* __restrict__ beta, int m, const int* __restrict__ blk,
                                               int nblk, const double* __restrict__ glo, const double* __restrict__ ghi, double pivmin,
                                               double* __restrict__ theta) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    int b = 0;
    while (b + 1 < nblk && e >= blk[3 * (b + 1) + 2]) ++b;
    const int r0 = blk[3 * b], nr = blk[3 * b + 1], kth = e - blk[3 * b + 2];
    const double* al = alpha + r0;
    const double* be = beta + r0;
    double lo = glo[b], hi = ghi[b];
    for (int it = 0; it < 128; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (!(mid > lo) || !(mid < hi)) break;
        int cnt = 0;
        double d = al[0] - mid;
        if (fabs(d) < pivmin) d = -pivmin;
        cnt += d < 0.0;
        for (int i = 1; i < nr; ++i) {
            const double bb = be[i - 1];
            d = al[i] - mid - bb * bb / d;
            if (fabs(d) < pivmin) d = -pivmin;
            cnt += d < 0.0;
        }
        if (cnt > kth) hi = mid;
        else lo = mid;
    }
    theta[e] = 0.5 * (lo + hi);
}

// ---- eigenvectors of the tridiagonal: twisted factorisation ---------------------------------------------------------------
// For eigenvalue theta of an unreduced block (rows r0 .. r0+nr): forward pivots dp_i of L D L^T = T - theta, backward pivots
// dm_i of U D U^T, twist index r = argmin |gamma_i|, gamma_i = dp_i + dm_i - (alpha_i - theta); z_r = 1,
// z_i = -(beta_i / dp_i) z_{i+1} above the twist, z_{i+1} = -(beta_i / dm_{i+1}) z_i below it.  One thread per selected
// eigenvalue; the pivots live in two m x ne work arrays with the eigenvalue index fastest (coalesced).  Z[k][e] (pitch ldz)
// receives the unit vector in the coordinates of the whole tridiagonal (zero outside the block).
static __global__ void __launch_bounds__(64) k_twisted(const double* __restrict__ alpha, const double* __restrict__ beta, int m, const int* __restrict__ blk,
                                                int nblk, const double* __restrict__ theta, const int* __restrict__ sel_blk, int ne,
                                                double pivmin, double* __restrict__ wp, double* __restrict__ wm, long long ldw,
                                                double* __restrict__ Z, long long ldz) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int b = sel_blk[e];
    const int r0 = blk[3 * b], nr = blk[3 * b + 1];
    const double th = theta[e];
    const double* al = alpha + r0;
    const double* be = beta + r0;
    double* dp = wp + e;
    double* dm = wm + e;
    // forward
    double d = al[0] - th;
    if (fabs(d) < pivmin) d = -pivmin;
    dp[0] = d;
    for (int i = 1; i < nr; ++i) {
        const double bb = be[i - 1];
        d = al[i] - th - bb * bb / d;
        if (fabs(d) < pivmin) d = -pivmin;
        dp[(long long)i * ldw] = d;
    }
    // backward, tracking the twist
    d = al[nr - 1] - th;
    if (fabs(d) < pivmin) d = -pivmin;
    dm[(long long)(nr - 1) * ldw] = d;
    int r = nr - 1;
    double gmin = fabs(dp[(long long)(nr - 1) * ldw] + d - (al[nr - 1] - th));
    for (int i = nr - 2; i >= 0; --i) {
        const double bb = be[i];
        d = al[i] - th - bb * bb / d;
        if (fabs(d) < pivmin) d = -pivmin;
        dm[(long long)i * ldw] = d;
        const double g = fabs(dp[(long long)i * ldw] + d - (al[i] - th));
        if (g < gmin) { gmin = g; r = i; }
    }
    // the vector, un-normalised, written into the pivot array dp (no longer needed entry by entry as we go)
    double nrm2 = 1.0;
    {
        double z = 1.0;
        for (int i = r - 1; i >= 0; --i) {
            z = -(be[i] / dp[(long long)i * ldw]) * z;
            dp[(long long)i * ldw] = z;
            nrm2 = fma(z, z, nrm2);
        }
        z = 1.0;
        for (int i = r; i + 1 < nr; ++i) {
            z = -(be[i] / dm[(long long)(i + 1) * ldw]) * z;
            dm[(long long)(i + 1) * ldw] = z;
            nrm2 = fma(z, z, nrm2);
        }
    }
    const double inv = rsqrt(nrm2);
    double* zc = Z + e;
    for (int k = 0; k < r0; ++k) zc[(long long)k * ldz] = 0.0;
    for (int i = 0; i < r; ++i) zc[(long long)(r0 + i) * ldz] = dp[(long long)i * ldw] * inv;
    zc[(long long)(r0 + r) * ldz] = inv;
    for (int i = r + 1; i < nr; ++i) zc[(long long)(r0 + i) * ldz] = dm[(long long)i * ldw] * inv;
    for (int k = r0 + nr; k < m; ++k) zc[(long long)k * ldz] = 0.0;
}

// ---- FP64 tensor-core GEMM --------------------------------------------------------------------------------------------------
// C[M x N] = op(A) B, row-major B [K][ldb] and C [M][ldc];  TA: A is stored [K][lda] (op(A) = A^T), otherwise [M][lda].
// EPI 0: C = op(A) B;  EPI 1: C = 1.5 I_{MxN} - 0.5 op(A) B  (the Newton-Schulz factor).
// CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA.8x8x4 tiles; operands staged through registers
// into double-buffered shared tiles stored k-major with pitch 132 (== 4 mod 16: the (t * pitch + g) fragment loads of a
// half-warp hit 16 distinct 8-byte banks).  Leading dimensions must be even and the base pointers 16-byte aligned.
constexpr int GBM = 128, GBN = 128, GBK = 16, GLD = 132;
constexpr size_t GEMM_SMEM = sizeof(double) * 4 * GBK * GLD;   // 67.6 KB: opt-in dynamic shared memory
template <bool TA, int EPI>
__global__ void __launch_bounds__(256) k_dgemm(int M, int N, int K, const double* __restrict__ A, long long lda, const double* __restrict__ B,
                                               long long ldb, double* __restrict__ C, long long ldc) {
    extern __shared__ __align__(16) double gsm[];
    double (*As)[GBK][GLD] = reinterpret_cast<double (*)[GBK][GLD]>(gsm);
    double (*Bs)[GBK][GLD] = reinterpret_cast<double (*)[GBK][GLD]>(gsm + 2 * GBK * GLD);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double2 ra[4], rb[4];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + 256 * q;            // 1024 double2 per tile
            {   // B tile: 16 rows (k) x 64 double2 (n)
                const int kr = idx >> 6, c2 = (idx & 63) * 2;
                const int kk = k0 + kr, nn = n0 + c2;
                rb[q] = (kk < K && nn < N) ? *reinterpret_cast<const double2*>(B + (long long)kk * ldb + nn) : make_double2(0.0, 0.0);
                if (nn + 1 >= N) rb[q].y = 0.0;
            }
            if (TA) {   // A stored [K][lda]: 16 rows (k) x 64 double2 (m)
                const int kr = idx >> 6, c2 = (idx & 63) * 2;
                const int kk = k0 + kr, mm = m0 + c2;
                ra[q] = (kk < K && mm < M) ? *reinterpret_cast<const double2*>(A + (long long)kk * lda + mm) : make_double2(0.0, 0.0);
                if (mm + 1 >= M) ra[q].y = 0.0;
            } else {    // A stored [M][lda]: 128 rows (m) x 8 double2 (k)
                const int mr = idx >> 3, c2 = (idx & 7) * 2;
                const int mm = m0 + mr, kk = k0 + c2;
                ra[q] = (mm < M && kk < K) ? *reinterpret_cast<const double2*>(A + (long long)mm * lda + kk) : make_double2(0.0, 0.0);
                if (kk + 1 >= K) ra[q].y = 0.0;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + 256 * q;
            {
                const int kr = idx >> 6, c2 = (idx & 63) * 2;
                *reinterpret_cast<double2*>(&Bs[buf][kr][c2]) = rb[q];
            }
            if (TA) {
                const int kr = idx >> 6, c2 = (idx & 63) * 2;
                *reinterpret_cast<double2*>(&As[buf][kr][c2]) = ra[q];
            } else {
                const int mr = idx >> 3, c2 = (idx & 7) * 2;
                As[buf][c2][mr] = ra[q].x;
                As[buf][c2 + 1][mr] = ra[q].y;
            }
        }
    };
    const int nk = (K + GBK - 1) / GBK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * GBK);
#pragma unroll
        for (int ks = 0; ks < GBK; ks += 4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = As[buf][ks + t][wm + 8 * i + g];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[buf][ks + t][wn + 8 * j + g];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm + 8 * i + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn + 8 * j + 2 * t;
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            if (EPI == 1) {
                v0 = (row == col ? 1.5 : 0.0) - 0.5 * v0;
                v1 = (row == col + 1 ? 1.5 : 0.0) - 0.5 * v1;
            }
            if (col + 1 < N) *reinterpret_cast<double2*>(C + (long long)row * ldc + col) = make_double2(v0, v1);
            else if (col < N) C[(long long)row * ldc + col] = v0;
        }
    }
}

// max |G - I| over an n x n matrix (pitch ld); out[0] receives the maximum (single CTA combine via atomicMax on the bit pattern)
static __global__ void __launch_bounds__(256) k_max_dev_identity(const double* __restrict__ G, int n, long long ld, unsigned long long* __restrict__ out) {
    double mx = 0.0;
    const long long total = (long long)n * n;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int r = (int)(idx / n), c = (int)(idx % n);
        mx = fmax(mx, fabs(G[(long long)r * ld + c] - (r == c ? 1.0 : 0.0)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));   // non-negative doubles order like integers
}

// ---- fragments ------------------------------------------------------------------------------------------------------------
// Eigenvector e of the row space (Y[e][i], pitch ldy) -> the N x N operator L_e of the reference (reshape(U[:, e], (n, n)),
// decompose.jl:353), column-major n x n at Lout + e n^2.  Packed mode undoes the sqrt(2) row scaling and mirrors (p,q) -> (q,p).
static __global__ void __launch_bounds__(256) k_frag_matrix(const double* __restrict__ Y, long long ldy, int len, const int* __restrict__ pa,
                                                     const int* __restrict__ pq, const double* __restrict__ sw, int n, int packed,
                                                     double* __restrict__ Lout) {
    const int e = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    double a = Y[(long long)e * ldy + i];
    double* out = Lout + (long long)e * n * n;
    const int p = pa[i], q = pq[i];
    if (packed) {
        a /= sw[i];
        out[p + (long long)n * q] = a;
        out[q + (long long)n * p] = a;
    } else {
        out[p + (long long)n * q] = a;
    }
}

// Sign convention: an eigenvector is defined up to its sign, and DF_based_greedy (decompose.jl:418-430) depends on it through the
// order of omega.  Here every fragment matrix gets its largest-magnitude entry (the first one in column-major order) positive.
// One CTA per fragment.
static __global__ void __launch_bounds__(256) k_frag_fix_sign(double* __restrict__ Lm, int n) {
    __shared__ double bv[8];
    __shared__ int bi[8];
    __shared__ int s_flip;
    double* L = Lm + (long long)blockIdx.x * n * n;
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int idx = threadIdx.x; idx < n * n; idx += 256) {
        const double a = fabs(L[idx]);
        if (a > best) { best = a; bidx = idx; }          // ascending idx per thread: the first maximum is kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < bidx)) { best = bv[w]; bidx = bi[w]; }
        s_flip = best > 0.0 && L[bidx] < 0.0;
    }
    __syncthreads();
    if (s_flip)
        for (int idx = threadIdx.x; idx < n * n; idx += 256) L[idx] = -L[idx];
}

// Symmetry class of every fragment matrix (decompose.jl:354-361): sym[e] = sum (Symmetric(L) - L)^2 (upper triangle kept),
// asym[e] = sum |L + L'|.  One CTA per fragment.
static __global__ void __launch_bounds__(256) k_frag_symmetry(const double* __restrict__ Lm, int n, double* __restrict__ sym, double* __restrict__ asym) {
    __shared__ double r1[8], r2[8];
    const double* L = Lm + (long long)blockIdx.x * n * n;
    double s = 0.0, a = 0.0;
    for (int idx = threadIdx.x; idx < n * n; idx += 256) {
        const int p = idx % n, q = idx / n;
        const double v = L[p + (long long)n * q], w = L[q + (long long)n * p];
        if (p > q) s += (w - v) * (w - v);        // Symmetric() takes the upper triangle: entry (p > q) is replaced by L[q,p]
        a += fabs(v + w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s; r2[threadIdx.x >> 5] = a; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double S = 0.0, A = 0.0;
        for (int q = 0; q < 8; ++q) { S += r1[q]; A += r2[q]; }
        sym[blockIdx.x] = S;
        asym[blockIdx.x] = A;
    }
}

// ---- batched symmetric eigen-decomposition (two-sided Jacobi, parallel ordering) ----------------------------------------
// One CTA per matrix.  A (n x n, symmetric; the upper triangle of the input is mirrored like Symmetric()) lives in shared
// memory with an odd pitch; the eigenvector accumulator U^T lives in shared memory as well when both fit (USM), otherwise in
// global memory (L2).  A sweep is n' - 1 rounds (n' = n rounded up to even) of n'/2 disjoint pairs from the round-robin
// tournament; every round computes all rotations from the current matrix, then applies them to the columns (and to U), then to
// the rows.  Converged when the off-diagonal mass is below 1e-30 of the total.  Eigenvalues ascending (LAPACK's order),
// eigenvectors in the columns of Uout (column-major n x n); status[e] = sweeps used (negative: not converged).
template <bool USM>
__global__ void __launch_bounds__(1024) k_jacobi_batch(const double* __restrict__ Lm, int n, double* __restrict__ omega, double* __restrict__ Uout,
                                                       double* __restrict__ Ut_glob, int* __restrict__ status, int max_sweeps) {
    extern __shared__ double sm[];
    const int np = n + (n & 1);                 // players of the tournament (a dummy when n is odd)
    const int lda = n | 1;                      // odd pitch: column accesses of a half-warp hit distinct banks
    double* A = sm;                             // [n][lda]
    double* cs = A + (size_t)n * lda;           // [np/2][2] rotations of the round
    int* pr = reinterpret_cast<int*>(cs + np);  // [np/2][2] pairs of the round
    double* red = reinterpret_cast<double*>(pr + np + (np & 1));   // 64 doubles of reduction scratch + the `rotated` flag
    double* Ut = USM ? red + 66 : Ut_glob + (size_t)blockIdx.x * n * n;   // [n][n]: row p = eigenvector p (transposed accumulator)
    const long long e = blockIdx.x;
    const double* L = Lm + e * n * n;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < n * n; idx += nt) {
        const int p = idx % n, q = idx / n;     // column-major input
        A[(size_t)p * lda + q] = p <= q ? L[p + (long long)n * q] : L[q + (long long)n * p];
        Ut[(size_t)q * n + p] = p == q ? 1.0 : 0.0;
    }
    __syncthreads();
    auto block_sum2 = [&](double a, double b, double& ra, double& rb) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if ((tid & 31) == 0) { red[tid >> 5] = a; red[32 + (tid >> 5)] = b; }
        __syncthreads();
        double sa = 0.0, sb = 0.0;
        for (int w = 0; w < (nt + 31) / 32; ++w) { sa += red[w]; sb += red[32 + w]; }
        __syncthreads();
        ra = sa; rb = sb;
    };
    const int half = np / 2;
    int sweep = 0;
    bool converged = false;
    // rotations below thr are skipped; a sweep without any rotation ends the iteration (then every |a_pq| < thr, i.e. the
    // off-diagonal norm is below ~2e-17 sqrt(n) ||A||_F).  The Frobenius norm is invariant under the rotations.
    double thr;
    {
        double tot = 0.0, dummy = 0.0, st, sd;
        for (int idx = tid; idx < n * n; idx += nt) {
            const double v = A[(size_t)(idx / n) * lda + (idx % n)];
            tot = fma(v, v, tot);
        }
        block_sum2(tot, dummy, st, sd);
        thr = 2.5e-17 * sqrt(st / n);
    }
    int* rotated = reinterpret_cast<int*>(red + 64);
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) *rotated = 0;
        __syncthreads();
        for (int round = 0; round < np - 1; ++round) {
            // round-robin tournament: player np-1 is fixed, the others rotate
            if (tid < half) {
                int a = tid == 0 ? np - 1 : (round + tid) % (np - 1);
                int b = (round + np - 1 - tid) % (np - 1);
                int p = min(a, b), q = max(a, b);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = A[(size_t)p * lda + q];
                    if (fabs(apq) > thr) {
                        *rotated = 1;
                        const double app = A[(size_t)p * lda + p], aqq = A[(size_t)q * lda + q];
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + tt * tt);
                        s = tt * c;
                    }
                } else {
                    q = -1;      // dummy pair
                }
                pr[2 * tid] = p; pr[2 * tid + 1] = q;
                cs[2 * tid] = c; cs[2 * tid + 1] = s;
            }
            __syncthreads();
            // columns: A <- A J, U <- U J   (J: [c s; -s c] on columns p, q)
            for (int idx = tid; idx < half * n; idx += nt) {
                const int k = idx / n, r = idx % n;
                const int p = pr[2 * k], q = pr[2 * k + 1];
                if (q < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                if (s == 0.0) continue;
                const double x = A[(size_t)r * lda + p], y = A[(size_t)r * lda + q];
                A[(size_t)r * lda + p] = c * x - s * y;
                A[(size_t)r * lda + q] = s * x + c * y;
                const double u = Ut[(size_t)p * n + r], v = Ut[(size_t)q * n + r];
                Ut[(size_t)p * n + r] = c * u - s * v;
                Ut[(size_t)q * n + r] = s * u + c * v;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int idx = tid; idx < half * n; idx += nt) {
                const int k = idx / n, r = idx % n;
                const int p = pr[2 * k], q = pr[2 * k + 1];
                if (q < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                if (s == 0.0) continue;
                const double x = A[(size_t)p * lda + r], y = A[(size_t)q * lda + r];
                A[(size_t)p * lda + r] = r == q ? 0.0 : c * x - s * y;      // the rotated pair is annihilated exactly
                A[(size_t)q * lda + r] = r == p ? 0.0 : s * x + c * y;
            }
            __syncthreads();
        }
        const int rot = *rotated;          // every thread reads the flag before thread 0 may clear it for the next sweep
        __syncthreads();
        if (rot == 0) { converged = true; ++sweep; break; }
    }
    // ascending order by counting ranks
    double* om = omega + e * n;
    double* Uo = Uout + e * n * n;
    for (int i = tid; i < n; i += nt) {
        const double wi = A[(size_t)i * lda + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double wj = A[(size_t)j * lda + j];
            rank += (wj < wi) || (wj == wi && j < i);
        }
        pr[i] = rank;                 // pairs scratch is free now (np >= n entries x 2)
        om[rank] = wi;
    }
    // sign convention: the largest-magnitude component (the first one) of every eigenvector is positive
    for (int i = tid >> 5; i < n; i += nt >> 5) {
        double best = -1.0;
        int bidx = 0x7fffffff;
        for (int r = tid & 31; r < n; r += 32) {
            const double a = fabs(Ut[(size_t)i * n + r]);
            if (a > best) { best = a; bidx = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if ((tid & 31) == 0) cs[i] = Ut[(size_t)i * n + bidx] < 0.0 ? -1.0 : 1.0;     // rotation scratch is free now
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx / n, r = idx % n;
        Uo[r + (long long)n * pr[i]] = cs[i] * Ut[(size_t)i * n + r];
    }
    if (tid == 0) status[e] = converged ? sweep : -sweep;
}
inline size_t jacobi_smem(int n, bool usm) {
    const int np = n + (n & 1), lda = n | 1;
    size_t doubles = (size_t)n * lda + np + (np + (np & 1)) / 2 + 1 + 66;
    if (usm) doubles += (size_t)n * n;
    return doubles * sizeof(double) + 16;
}

// ---- DF_based_greedy (decompose.jl:400-436) ---------------------------------------------------------------------------------
// vv[m][i] = sum_p omega_m[p] |Vm[i,p]|^2,  Vm = Um U1^T  (Vm[i,p] = sum_k Um[i,k] U1[p,k]); one CTA per fragment.
static __global__ void __launch_bounds__(256) k_df_frame(const double* __restrict__ Uall, const double* __restrict__ omega, int n, int nfrag,
                                                  double* __restrict__ vv) {
    extern __shared__ double sm[];
    double* U1 = sm;                       // [n][n] column-major (broadcast reads); Um is read coalesced from global / L2
    const int m = blockIdx.x;
    const double* Um = Uall + (long long)m * n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += 256) U1[idx] = Uall[idx];
    __syncthreads();
    const double* om = omega + (long long)m * n;
    for (int i = threadIdx.x; i < n; i += 256) {
        double acc = 0.0;
        for (int p = 0; p < n; ++p) {
            double v = 0.0;
            for (int k = 0; k < n; ++k) v = fma(Um[i + (size_t)n * k], U1[p + (size_t)n * k], v);
            acc = fma(om[p], v * v, acc);
        }
        vv[(long long)m * n + i] = acc;
    }
}
// Lmat[i][j] = sum_m coeff[m] vv[m][i] vv[m][j]   (fixed order over m: bit-reproducible); one thread per entry
static __global__ void __launch_bounds__(256) k_df_lmat(const double* __restrict__ vv, const double* __restrict__ coeff, int n, int nfrag,
                                                 double* __restrict__ Lmat) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= n * n) return;
    const int i = idx % n, j = idx / n;
    double acc = 0.0;
    for (int m = 0; m < nfrag; ++m) acc = fma(coeff[m] * vv[(long long)m * n + i], vv[(long long)m * n + j], acc);
    Lmat[i + (long long)n * j] = acc;
}

}  // namespace df
