// tiny_kernels.cuh -- the whole CSA / CSA_SD cost + gradient evaluation in ONE kernel for molecule-sized problems (sm_100a).
//
// BASELINE configs[0..1] (LiH N = 6, H2O N = 7, N2 N = 10) are far below the size where the tensor-core pipeline of the main
// path pays: there one evaluation is 11 launches of which every one is latency-bound (~0.1 ms per evaluation whatever N is).
// For N <= NMAX and an 8-fold symmetric tensor (packed rows, M = N(N+1)/2 <= 136) a single CTA holds every operand in shared
// memory and walks through the same algebra (reference: gradient.jl:1-290, cost.jl:16-55, unitaries.jl:31-46):
//
//   K, Lambda from x;  s: ||K||_1 / 2^s <= 1;  U = exp(K): Taylor-16 of X = K / 2^s, s squarings (every square kept); the powers
//   X^2 | X^3, X^4 | X^5..X^8 | X^9..X^16 come from four passes of independent products (the chain is latency-bound)
//   A~[i,l] = sw_i U[p_i,l] U[q_i,l];  B~ = A~ Lambda
//   D~[i,j] = <B~_i, A~_j> - T~[i,j]  (never stored);  cost = sum D~^2 (explicit: nothing to cancel);  E~ = D~ A~
//   S = A~^T E~ -> grad_lambda (one-sided, gradient.jl:22-26);  F~ = E~ Lambda;  G[a,b] = 4 sum_q U[q,b] F~[(a,q),b] / sw_(a,q)
//   CSA_SD: Dh = U diag(l1) U^T - h,  cost += |Dh|^2,  G += ((Dh + Dh^T) U) diag(l1),  grad_l1 = 2 diag(U^T Dh U)
//   Lt = Dexp(K)[G^T]: the same four passes and the squarings differentiated in direction G^T / 2^s (product rule)
//   grad_theta[(p<q)] = 2 (Lt[q,p] - Lt[p,q])
//
// T~ is read once per evaluation straight from global memory (L2-resident: <= 148 KB), coalesced through its symmetry.
// Every reduction has a fixed order: results are bit-reproducible.
#pragma once
#include <cuda_runtime.h>

namespace tiny {

constexpr int NMAX = 16;                       // largest N the kernel is instantiated for
constexpr int MMAX = NMAX * (NMAX + 1) / 2;    // packed rows
constexpr int NT = 512;                        // threads of the single CTA
constexpr int SMAXT = 15;                      // squarings: ||K||_1 <= 2^14 (the range of the main path) needs 14
// 1 / k!, k = 0..16 (Taylor-16 at ||X||_1 <= 1: truncation 1/17! = 3e-15, like the main path's chain)
__device__ const double kInvFact[17] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                                        1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0, 1.0 / 87178291200.0,
                                        1.0 / 1307674368000.0, 1.0 / 20922789888000.0};

struct Params {
    const double* x;        // [L] the optimiser vector (device)
    const double* T;        // packed residual T~ [M][ldT]
    long long ldT;
    const double* hT;       // one-body residual, column-major n x n (CSA_SD), else nullptr
    const int* pa;          // [M] p of row i (p <= q)
    const int* pq;          // [M] q of row i
    const double* sw;       // [M] 1 or sqrt(2)
    int n, M, sd, lam_off, cL, need_grad;
    double* out;            // need_grad: [1 + L] = cost, gradient in x order
    double* cost_out;       // !need_grad: the cost alone
    double* obc;            // optional: one-body cost on its own (CSA_SD)
    int* status;            // set to 1 when ||K||_1 > 2^14 or x holds a NaN
};

// shared-memory plan (doubles); n2 = n * n, Mn = M * n
__host__ __device__ inline size_t smem_doubles(int n, int M) {
    const size_t n2 = (size_t)n * n, Mn = (size_t)M * n;
    const int C = (NT / M) < M ? (NT / M) : M;
    const size_t stage = (size_t)(C > 0 ? C : 1) * Mn > 8 * n2 ? (size_t)(C > 0 ? C : 1) * Mn : 8 * n2;
    return n2 * (6 + 16 + (SMAXT + 1)) + 4 * Mn + stage + NT + 2 * n + 8;
}
inline size_t smem_bytes(int n, int M) { return smem_doubles(n, M) * sizeof(double) + (size_t)n * n * sizeof(int) + 64; }

__device__ __forceinline__ double block_sum(double v, double* red, int tid) {   // fixed order: shuffle tree, then 16 warp sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(NT, 1) k_eval(Params p) {
    extern __shared__ __align__(16) double sm[];
    const int n = p.n, M = p.M, n2 = n * n, Mn = M * n, tid = threadIdx.x;
    double* K = sm;                    // n2 each
    double* Lam = K + 2 * n2;          // (one n2 slot after K is spare)
    double* Lbuf = Lam + n2;           // the derivative L (ping-pong with Dh during its squarings)
    double* Gt = Lbuf + n2;            // G^T / 2^s
    double* Dh = Gt + n2;              // one-body difference (CSA_SD); scratch of the squarings otherwise
    double* Pw = Dh + n2;              // X^1 .. X^8
    double* dPw = Pw + 8 * n2;         // their derivatives
    double* Esq = dPw + 8 * n2;        // exp(X)^(2^j), j = 0..s
    double* At = Esq + (size_t)(SMAXT + 1) * n2;
    double* Bt = At + Mn;
    double* Et = Bt + Mn;
    double* Ft = Et + Mn;
    const int C = min(NT / M, M);      // column chunks of the main pass: C * M <= NT threads
    double* Epart = Ft + Mn;           // [C][M][n]; also the staging area [8][n2] of the fourth power pass
    double* red = Epart + max((size_t)C * Mn, (size_t)8 * n2);      // NT
    double* l1 = red + NT;             // n
    double* colsum = l1 + n;           // n
    double* scal = colsum + n;         // [0] = 2^-s, [1] = cost
    int* idxmap = reinterpret_cast<int*>(scal + 8);   // [n][n] -> packed row
    __shared__ int s_sh;

    // ---- x -> K, Lambda, lambda_1 -----------------------------------------------------------------------------------------
    const double* lam = p.x + p.lam_off;
    const double* th = lam + p.cL;
    for (int e = tid; e < n2; e += NT) {
        const int a = e / n, b = e % n;
        const int hi = max(a, b), lo = min(a, b);
        Lam[e] = lam[hi * (hi + 1) / 2 + lo];                                   // lower triangle, row-major (fermionic.jl:20-32)
        double k = 0.0;
        if (a != b) {
            const int idx = lo * n - lo * (lo + 1) / 2 + (hi - lo - 1);         // upper triangle, row-major (unitaries.jl:406-415)
            k = (a < b) ? th[idx] : -th[idx];
        }
        K[e] = k;
    }
    if (p.sd) for (int i = tid; i < n; i += NT) l1[i] = p.x[i];
    for (int i = tid; i < M; i += NT) {
        idxmap[p.pa[i] * n + p.pq[i]] = i;
        idxmap[p.pq[i] * n + p.pa[i]] = i;
    }
    __syncthreads();
    if (tid < n) {
        double s = 0.0;
        for (int r = 0; r < n; ++r) s += fabs(K[r * n + tid]);
        colsum[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double nrm = 0.0;
        bool bad = false;
        for (int c = 0; c < n; ++c) {
            if (!(colsum[c] == colsum[c])) bad = true;
            nrm = fmax(nrm, colsum[c]);
        }
        if (bad || !(nrm <= 16384.0)) *p.status = 1;
        int s = 0;
        double sc = 1.0;
        while (s < SMAXT && nrm * sc > 1.0) { ++s; sc *= 0.5; }
        s_sh = s;
        scal[0] = sc;
    }
    __syncthreads();
    const int s = s_sh;
    const double sc = scal[0];

    // ---- U = exp(K) ----------------------------------------------------------------------------------------------------------
    // C[a][b] = sum_r A[a][r] B[r][b]
    auto dot = [&](const double* A, const double* B, int a, int b) {
        double v = 0.0;
#pragma unroll 4
        for (int r = 0; r < n; ++r) v = fma(A[a * n + r], B[r * n + b], v);
        return v;
    };
    // passes with only n2 independent elements (X^2, the squarings): LG lanes share one dot product (n2 * LG <= NT), fixed order
    const int LG = (4 * n2 <= NT) ? 4 : ((2 * n2 <= NT) ? 2 : 1);
    auto mm_split = [&](double* dst, const double* A, const double* B, const double* A2, const double* B2) {   // dst = A B (+ A2 B2)
        const int e = tid / LG, sub = tid % LG;
        double v = 0.0;
        if (e < n2) {
            const int a = e / n, b = e % n;
            for (int r = sub; r < n; r += LG) {
                v = fma(A[a * n + r], B[r * n + b], v);
                if (A2) v = fma(A2[a * n + r], B2[r * n + b], v);
            }
        }
        if (LG >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (LG >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (e < n2 && sub == 0) dst[e] = v;
    };
    for (int e = tid; e < n2; e += NT) Pw[e] = K[e] * sc;
    __syncthreads();
    mm_split(Pw + n2, Pw, Pw, nullptr, nullptr);                                                             // X^2
    __syncthreads();
    for (int t = tid; t < 2 * n2; t += NT) {                                                                   // X^3, X^4
        const int q = t / n2, e = t % n2;
        Pw[(2 + q) * n2 + e] = dot(Pw + n2, Pw + q * n2, e / n, e % n);
    }
    __syncthreads();
    for (int t = tid; t < 4 * n2; t += NT) {                                                                   // X^5 .. X^8
        const int q = t / n2, e = t % n2;
        Pw[(4 + q) * n2 + e] = dot(Pw + 3 * n2, Pw + q * n2, e / n, e % n);
    }
    __syncthreads();
    for (int t = tid; t < 8 * n2; t += NT) {                                                                   // X^9 .. X^16 (staged)
        const int q = t / n2, e = t % n2;
        Epart[q * n2 + e] = dot(Pw + 7 * n2, Pw + q * n2, e / n, e % n);
    }
    __syncthreads();
    for (int e = tid; e < n2; e += NT) {
        double v = (e / n == e % n) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 1; k <= 8; ++k) v = fma(Pw[(k - 1) * n2 + e], kInvFact[k], v);
#pragma unroll
        for (int k = 9; k <= 16; ++k) v = fma(Epart[(k - 9) * n2 + e], kInvFact[k], v);
        Esq[e] = v;
    }
    __syncthreads();
    for (int j = 1; j <= s; ++j) {
        const double* src = Esq + (size_t)(j - 1) * n2;
        mm_split(Esq + (size_t)j * n2, src, src, nullptr, nullptr);
        __syncthreads();
    }
    const double* U = Esq + (size_t)s * n2;            // exp(K)

    // ---- operands A~, B~ ------------------------------------------------------------------------------------------------------
    for (int e = tid; e < Mn; e += NT) {
        const int i = e / n, l = e % n;
        At[e] = p.sw[i] * U[p.pa[i] * n + l] * U[p.pq[i] * n + l];
    }
    __syncthreads();
    for (int e = tid; e < Mn; e += NT) {
        const int i = e / n, l = e % n;
        double v = 0.0;
        for (int r = 0; r < n; ++r) v = fma(At[i * n + r], Lam[r * n + l], v);
        Bt[e] = v;
    }
    __syncthreads();

    // ---- main pass: D~ (on the fly), cost, E~ partials ------------------------------------------------------------------------
    double cpart = 0.0;
    if (tid < C * M) {
        const int i = tid % M, c = tid / M;
        const int j0 = (int)((long long)c * M / C), j1 = (int)((long long)(c + 1) * M / C);
        double a[NMAX], ev[NMAX];
#pragma unroll
        for (int l = 0; l < NMAX; ++l) {
            a[l] = (l < n) ? At[i * n + l] : 0.0;
            ev[l] = 0.0;
        }
        for (int j = j0; j < j1; ++j) {
            double d = -__ldg(p.T + (long long)j * p.ldT + i);      // T~ is symmetric: row j, column i (coalesced over i)
            const double* bj = Bt + j * n;
            const double* aj = At + j * n;
#pragma unroll
            for (int l = 0; l < NMAX; ++l)
                if (l < n) d = fma(bj[l], a[l], d);
            cpart = fma(d, d, cpart);
#pragma unroll
            for (int l = 0; l < NMAX; ++l)
                if (l < n) ev[l] = fma(d, aj[l], ev[l]);
        }
#pragma unroll
        for (int l = 0; l < NMAX; ++l)
            if (l < n) Epart[((size_t)c * M + i) * n + l] = ev[l];
    }
    // one-body difference (CSA_SD)
    double obpart = 0.0;
    if (p.sd) {
        for (int e = tid; e < n2; e += NT) {
            const int r = e / n, q = e % n;
            double h = 0.0;
            for (int i = 0; i < n; ++i) h = fma(U[r * n + i] * l1[i], U[q * n + i], h);
            const double d = h - p.hT[r + n * q];
            Dh[e] = d;
            obpart = fma(d, d, obpart);
        }
    }
    const double c2 = block_sum(cpart, red, tid);
    const double c1 = p.sd ? block_sum(obpart, red, tid) : 0.0;
    const double cost = c2 + c1;
    if (tid == 0 && p.obc) *p.obc = c1;
    if (!p.need_grad) {
        if (tid == 0) *p.cost_out = cost;
        return;
    }
    for (int e = tid; e < Mn; e += NT) {
        double v = 0.0;
        for (int c = 0; c < C; ++c) v += Epart[(size_t)c * Mn + e];
        Et[e] = v;
    }
    __syncthreads();

    // ---- lambda gradient, F~, G -----------------------------------------------------------------------------------------------
    double* gout = p.out + 1;
    for (int e = tid; e < n2; e += NT) {
        const int a = e / n, b = e % n;
        if (b <= a) {
            double v = 0.0;
            for (int i = 0; i < M; ++i) v = fma(At[i * n + a], Et[i * n + b], v);
            gout[p.lam_off + a * (a + 1) / 2 + b] = 2.0 * v * (a != b ? 2.0 : 1.0);
        }
    }
    for (int e = tid; e < Mn; e += NT) {
        const int i = e / n, l = e % n;
        double v = 0.0;
        for (int r = 0; r < n; ++r) v = fma(Et[i * n + r], Lam[r * n + l], v);
        Ft[e] = v;
    }
    __syncthreads();
    for (int e = tid; e < n2; e += NT) {
        const int a = e / n, b = e % n;
        double g = 0.0;
        for (int q = 0; q < n; ++q) {
            const int i = idxmap[a * n + q];
            g = fma(U[q * n + b], Ft[i * n + b] / p.sw[i], g);
        }
        g *= 4.0;
        if (p.sd) {
            double v = 0.0;
            for (int r = 0; r < n; ++r) v = fma(Dh[a * n + r] + Dh[r * n + a], U[r * n + b], v);
            g = fma(v, l1[b], g);
        }
        Gt[b * n + a] = g * sc;                                  // transposed and scaled: the direction of the derivative
    }
    if (p.sd) {
        // DU[r][k] = sum_q Dh[r][q] U[q][k] -> the staging area (free now);  grad_l1[k] = 2 sum_r U[r][k] DU[r][k]
        for (int e = tid; e < n2; e += NT) {
            const int r = e / n, k = e % n;
            double v = 0.0;
            for (int q = 0; q < n; ++q) v = fma(Dh[r * n + q], U[q * n + k], v);
            Epart[e] = v;
        }
        __syncthreads();
        for (int k = tid; k < n; k += NT) {
            double v = 0.0;
            for (int r = 0; r < n; ++r) v = fma(U[r * n + k], Epart[r * n + k], v);
            gout[k] = 2.0 * v;
        }
    }
    __syncthreads();

    // ---- Lt = Dexp(K)[G^T]: the power passes with their derivatives, then the squarings -----------------------------------------
    //   d(X^(a+b)) = d(X^a) X^b + X^a d(X^b), d(X^1) = Gt;  L = sum_k d(X^k) / k!;  L <- E_j L + L E_j for every kept square E_j
    auto dot2 = [&](const double* dA, const double* B, const double* A, const double* dB, int a, int b) {
        double v = 0.0;
#pragma unroll 4
        for (int r = 0; r < n; ++r) {
            v = fma(dA[a * n + r], B[r * n + b], v);
            v = fma(A[a * n + r], dB[r * n + b], v);
        }
        return v;
    };
    double* Ls = Lbuf;
    for (int e = tid; e < n2; e += NT) dPw[e] = Gt[e];
    __syncthreads();
    mm_split(dPw + n2, dPw, Pw, Pw, dPw);
    __syncthreads();
    for (int t = tid; t < 2 * n2; t += NT) {
        const int q = t / n2, e = t % n2;
        dPw[(2 + q) * n2 + e] = dot2(dPw + n2, Pw + q * n2, Pw + n2, dPw + q * n2, e / n, e % n);
    }
    __syncthreads();
    for (int t = tid; t < 4 * n2; t += NT) {
        const int q = t / n2, e = t % n2;
        dPw[(4 + q) * n2 + e] = dot2(dPw + 3 * n2, Pw + q * n2, Pw + 3 * n2, dPw + q * n2, e / n, e % n);
    }
    __syncthreads();
    for (int t = tid; t < 8 * n2; t += NT) {
        const int q = t / n2, e = t % n2;
        Epart[q * n2 + e] = dot2(dPw + 7 * n2, Pw + q * n2, Pw + 7 * n2, dPw + q * n2, e / n, e % n);
    }
    __syncthreads();
    for (int e = tid; e < n2; e += NT) {
        double v = 0.0;
#pragma unroll
        for (int k = 1; k <= 8; ++k) v = fma(dPw[(k - 1) * n2 + e], kInvFact[k], v);
#pragma unroll
        for (int k = 9; k <= 16; ++k) v = fma(Epart[(k - 9) * n2 + e], kInvFact[k], v);
        Ls[e] = v;
    }
    __syncthreads();
    double* Lalt = Dh;                                           // the one-body difference is no longer needed
    for (int j = 0; j < s; ++j) {
        const double* Ej = Esq + (size_t)j * n2;
        mm_split(Lalt, Ej, Ls, Ls, Ej);
        __syncthreads();
        double* t = Ls; Ls = Lalt; Lalt = t;
    }
    double* gth = gout + p.lam_off + p.cL;
    for (int e = tid; e < n2; e += NT) {
        const int a = e / n, b = e % n;
        if (a < b) gth[a * n - a * (a + 1) / 2 + (b - a - 1)] = 2.0 * (Ls[b * n + a] - Ls[a * n + b]);
    }
    if (tid == 0) p.out[0] = cost;
}

}  // namespace tiny
