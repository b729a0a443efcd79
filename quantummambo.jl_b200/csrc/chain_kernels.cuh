// chain_kernels.cuh -- the orbital-rotation builder as two persistent kernels (sm_100a).
//
//   k_chain_u : theta -> K, ||K||_1 -> s, X = K/2^s, U = exp(K) by Taylor-16 (Paterson-Stockmeyer) + s squarings
//               (reference: unitaries.jl:31-46, `exp(Ulog)`), keeping every intermediate for the derivative;
//   k_chain_d : Lt = Dexp(K)[G^T] by the product rule over the same sequence, then the gradient in x order
//               (reference: del_u_theta gradient.jl:76-129 contracted as in grad_theta :131-146; `vcat` :148-151).
//
// One CTA owns one CT x CT tile of every N x N matrix of the chain; the ~5+s dependent products are separated by a
// barrier instead of kernel boundaries.  CT = 16 (16 warps = 4 (8x8 blocks) x 4 (k quarters)) in two flavours:
//   CLUSTER : N <= 64.  At most 4 x 4 tiles, so the whole chain runs in ONE thread-block cluster (<= 16 CTAs) and the
//             steps are separated by the hardware cluster barrier (barrier.cluster release/acquire) instead of a round
//             trip through a global-memory counter (~1.2 us).
//   grid    : larger N.  Cooperative launch, grid-wide barrier on one counter in global memory.
// (CT = 32 also compiles -- 16 warps = 16 blocks, full k -- and would put N <= 128 into one cluster, but a step is then
// bound by the DMMA rate of 16 SMs: measured slower than 49 CTAs with the global barrier at N = 100.)  Every matrix that is later a
// right-hand operand is stored twice (M and M^T), and every matrix has a padded pitch ldm (== 4 mod 16, zero pad), so
// that each operand of a tile product is ONE contiguous 16 x ldm block: it is fetched by the TMA engine with a single
// cp.async.bulk (completion on an mbarrier) and lands in shared memory already in a bank-conflict-free layout.  Products run
// on the FP64 tensor pipe (DMMA.8x8x4) with two accumulators per product term.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace chain {

constexpr int NTHR = 512;    // 16 warps = (CT/8)^2 blocks x KQ k-slices
constexpr int SMAXC = 14;
constexpr int NSTRIP = 6;
template <int CT>
struct Geo {
    static constexpr int BW = CT / 8;           // 8x8 blocks per tile edge
    static constexpr int NBLK = BW * BW;        // blocks per tile (4 or 16)
    static constexpr int KQ = 16 / NBLK;        // k split across warps (4 or 1)
};

__constant__ double c_t[17] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
                               1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0, 1.0 / 87178291200.0,
                               1.0 / 1307674368000.0, 1.0 / 20922789888000.0};

// matrices of the U chain; `*t` = transposed copy
struct UParams {
    const double* x;
    int n, lam_off, s_override;
    double *X, *Xt, *X2, *X2t, *X4, *P0, *P1, *P2, *Q3t, *Q2t, *Q1t, *R, *Rt;   // pitch ldm, zero padded
    double *Lam, *U;                                                           // dense n x n outputs
    double* LamP;                                                              // optional padded copy [Np][ldl] for the panel GEMMs
    int ldl;
    int* s_out;
    double* scale_out;
    int* status;
    unsigned* bar;  // [0] barrier counter, [1] exit counter (both zero between launches)
};
struct DParams {
    const double* part;   // [cost, S, G]
    const double* g1;     // CSA_SD one-body gradient (or null)
    int n, sd;
    const double *X, *Xt, *X2, *X2t, *X4, *Q3t, *Q2t, *Q1t, *R, *Rt;
    double *dX2, *dX2t, *dX4, *dP0, *dP1, *dP2, *dQ3t, *dQ2t, *dQ1t, *dR, *dRt;
    const int* s_in;
    const double* scale_in;
    double* out;          // [cost, grad...]
    unsigned* bar;
    int* status;
    long long* dbg;       // optional clock64 stamps of CTA 0 (development aid; MAMBO_CHAIN_DEBUG=1)
    int write_cost;       // 1: out[0] = part[0]; 0: the cost is written by small::k_cost_select running beside this kernel
    int write_lambda;     // 1: lambda gradient from part's S; 0: small::k_reduce_s (side stream) writes it
};

// pitch of every chain matrix: smallest ldm >= n with ldm == 4 (mod 16)  (100 -> 100, 152 -> 164)
__host__ __device__ inline int ldm_for(int n) { return ((n + 11) / 16) * 16 + 4; }
__host__ __device__ inline int lda_for(int n) { return ldm_for(n); }
__host__ __device__ inline int rows_pad(int n, int ct) { return ((n + ct - 1) / ct) * ct; }
__host__ __device__ inline size_t mat_stride(int n, int ct) { return (size_t)rows_pad(n, ct) * ldm_for(n); }
__host__ __device__ inline int kpad(int n) { return ((n + 3) / 4) * 4; }
inline size_t smem_bytes(int n, int ct) {
    const int kq = 16 / ((ct / 8) * (ct / 8));
    return ((size_t)NSTRIP * ct * lda_for(n) + (size_t)(kq - 1) * 4 * 32 * 4) * sizeof(double) + 64;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Grid-wide barrier: release-reduction on arrival, acquire polling at L2.  Data produced by other CTAs is afterwards
// read only through L2 (TMA bulk copies and ld.global.cg), never through this SM's L1, so no L1 invalidation
// (fence.acq_rel.gpu -> CCTL.IVALL) is paid per barrier.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target, long long* stamps = nullptr) {
    __syncthreads();
    target += gridDim.x * gridDim.y;
    if (threadIdx.x == 0) {
        if (stamps) stamps[0] = clock64();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        if (stamps) stamps[1] = clock64();
        unsigned v;
        int polls = 0;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            ++polls;
        } while (v < target);
        if (stamps) {
            stamps[2] = clock64();
            stamps[3] = polls;
        }
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

// How the dependent steps of a chain are ordered across CTAs:
//   SYNC_GRID     grid-wide barrier on a global counter (cooperative launch), operands re-read by TMA afterwards;
//   SYNC_CLUSTER  hardware cluster barrier (N <= 64: all tiles in one thread-block cluster), TMA afterwards;
constexpr int SYNC_GRID = 0, SYNC_CLUSTER = 1;

template <int SYNC>
__device__ __forceinline__ void step_barrier(unsigned* ctr, unsigned& target, long long* stamps = nullptr) {
    // Every thread orders its generic-proxy accesses to the operand strips (the LDS of this step's products) before the
    // async-proxy refills that follow the barrier.  The barrier itself makes a real overlap implausible, but the fused
    // kernels' operand ring showed in round 2 that a TMA refill may overtake loads that are merely issued; a few cycles per step.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (SYNC == SYNC_CLUSTER) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    } else {
        grid_barrier(ctr, target, stamps);
    }
}
template <int SYNC>
__device__ __forceinline__ void chain_exit(unsigned* bar) {
    if (SYNC != SYNC_CLUSTER) grid_exit(bar);
}

// ---- strip loader: rows [r0, r0+CT) of a chain matrix (pitch ldm, zero padded) -> S, one TMA bulk copy ---------------
template <int CT>
struct Loader {
    uint64_t* mbar;
    uint32_t parity;
    int lda;
    uint32_t pending;   // bytes expected in the current phase
    int nissued;        // strips issued in the current phase (the i-th strip is issued by lane 0 of warp i)

    // proxy_fence: needed only when the strips about to be overwritten were last *written* through the generic
    // proxy (zero fill, operands built in registers); generic reads are ordered by the preceding bar.sync.
    __device__ __forceinline__ void begin(bool proxy_fence = false) {
        pending = 0;
        nissued = 0;
        if (proxy_fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __device__ __forceinline__ void rows(double* S, const double* __restrict__ M, int r0) {
        const uint32_t bytes = (uint32_t)(CT * lda * 8);
        pending += bytes;
        if ((int)threadIdx.x == 32 * nissued)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(S)),
                         "l"(M + (size_t)r0 * lda), "r"(bytes), "r"(s32(mbar))
                         : "memory");
        ++nissued;
    }
    __device__ __forceinline__ void wait() {
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(mbar)), "r"(pending) : "memory");
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "LAB_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
            "@P1 bra DONE;\n"
            "bra LAB_WAIT;\n"
            "DONE:\n"
            "}\n" ::"r"(s32(mbar)), "r"(parity) : "memory");
        parity ^= 1u;
    }
};

// this warp's k quarter of an 8x8 block product, two accumulators: (c, d) += SA[block rows] . SB[block cols]^T
// SA: [16 rows i][lda] (k contiguous), SB: [16 cols j][lda] (k contiguous)
template <int CT>
__device__ __forceinline__ void mm_q(const double* SA, const double* SB, int lda, int ks, double (&c)[2], double (&d)[2]) {
    constexpr int KQ = Geo<CT>::KQ, NBLK = Geo<CT>::NBLK, BW = Geo<CT>::BW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2, b = warp % NBLK, kq = warp / NBLK;
    const double* a = SA + (8 * (b / BW) + t1) * lda + t0;
    const double* bb = SB + (8 * (b % BW) + t1) * lda + t0;
    int kk = kq;
    for (; kk + KQ < ks; kk += 2 * KQ) {
        dmma(c[0], c[1], a[4 * kk], bb[4 * kk]);
        dmma(d[0], d[1], a[4 * (kk + KQ)], bb[4 * (kk + KQ)]);
    }
    if (kk < ks) dmma(c[0], c[1], a[4 * kk], bb[4 * kk]);
}
// combine the k quarters; afterwards warps 0..NBLK-1 hold the full sums (nothing to do when every warp owns its block)
template <int CT>
__device__ __forceinline__ void combine(double* red, double* v, int nv) {
    constexpr int KQ = Geo<CT>::KQ;
    if (KQ == 1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= 4)
        for (int i = 0; i < nv; ++i) red[(((warp >> 2) - 1) * 128 + (warp & 3) * 32 + lane) * 4 + i] = v[i];
    __syncthreads();
    if (warp < 4)
        for (int q = 0; q < KQ - 1; ++q)
            for (int i = 0; i < nv; ++i) v[i] += red[((q * 128) + warp * 32 + lane) * 4 + i];
}

struct Ctx {
    double* S[NSTRIP];
    double* red;
    int n, lda, ks, i0, j0, ei, ej;
};

template <int CT>
__device__ __forceinline__ Ctx make_ctx(double* sm, int n, uint64_t** mbar) {
    constexpr int KQ = Geo<CT>::KQ, NBLK = Geo<CT>::NBLK, BW = Geo<CT>::BW;
    Ctx c;
    c.n = n;
    c.lda = lda_for(n);
    c.ks = kpad(n) / 4;
    for (int i = 0; i < NSTRIP; ++i) c.S[i] = sm + (size_t)i * CT * c.lda;
    c.red = sm + (size_t)NSTRIP * CT * c.lda;
    *mbar = reinterpret_cast<uint64_t*>(c.red + (KQ - 1) * 4 * 32 * 4);
    c.i0 = blockIdx.y * CT;
    c.j0 = blockIdx.x * CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = warp % NBLK;
    c.ei = c.i0 + 8 * (b / BW) + (lane >> 2);
    c.ej = c.j0 + 8 * (b % BW) + 2 * (lane & 3);
    for (int idx = threadIdx.x; idx < NSTRIP * CT * c.lda; idx += NTHR) sm[idx] = 0.0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(*mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return c;
}

// ------------------------------------------------------------------------------------------------------------
template <int CT, int SYNC>
__global__ void __launch_bounds__(NTHR, 2) k_chain_u(UParams p) {
    constexpr int NBLK = Geo<CT>::NBLK;
    extern __shared__ __align__(128) double sm[];
    const int n = p.n;
    const size_t nn = mat_stride(n, CT);
    uint64_t* mbar;
    Ctx c = make_ctx<CT>(sm, n, &mbar);
    Loader<CT> ld{mbar, 0u, c.lda, 0u, 0};
    __shared__ double s_red[NTHR / 32];
    __shared__ double s_scale;
    __shared__ int s_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = c.i0, j0 = c.j0, lda = c.lda, ks = c.ks, ei = c.ei, ej = c.ej;
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
            for (int w = 0; w < NTHR / 32; ++w) m = fmax(m, s_red[w]);
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
    // ---- step 1: X2 = X.X; rows I of X -> S0, columns J of X (as rows of X^T) -> S1, built straight from theta ----
    for (int r = threadIdx.x >> 5; r < CT; r += NTHR / 32) {
#pragma unroll 4
        for (int k = lane; k < n; k += 32) {
            c.S[0][r * lda + k] = kval(i0 + r, k);
            c.S[1][r * lda + k] = kval(k, j0 + r);
        }
    }
    __syncthreads();
    double v[4] = {0, 0, 0, 0};
    {
        double a2[2] = {0, 0}, b2[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[1], lda, ks, a2, b2);
        v[0] = a2[0] + b2[0];
        v[1] = a2[1] + b2[1];
    }
    combine<CT>(c.red, v, 2);
    double x1[2] = {0, 0}, x2[2] = {0, 0};
    if (warp < NBLK) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                const int o = i * lda + j, ot = j * lda + i;
                x1[e] = kval(i, j);
                x2[e] = v[e];
                p.X[o] = x1[e];
                p.Xt[ot] = x1[e];
                p.X2[o] = x2[e];
                p.X2t[ot] = x2[e];
                const int a = i > j ? i : j, bq = i > j ? j : i;
                const double lv = lam[a * (a + 1) / 2 + bq];
                p.Lam[i * n + j] = lv;                               // dense n x n (FMA epilogues)
                if (p.LamP) p.LamP[i * p.ldl + j] = lv;              // padded pitch (tensor-core panel GEMMs)
            }
        }
    }
    step_barrier<SYNC>(p.bar, target);
    // ---- step 2: X3 = X2.X, X4 = X2.X2; P0..P2, Q3 (S1 still holds the columns of X) -----------------------------
    ld.begin(true);
    ld.rows(c.S[0], p.X2, i0);
    ld.rows(c.S[2], p.X2t, j0);
    ld.wait();
    {
        double a3[2] = {0, 0}, b3[2] = {0, 0}, a4[2] = {0, 0}, b4[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[1], lda, ks, a3, b3);
        mm_q<CT>(c.S[0], c.S[2], lda, ks, a4, b4);
        v[0] = a3[0] + b3[0];
        v[1] = a3[1] + b3[1];
        v[2] = a4[0] + b4[0];
        v[3] = a4[1] + b4[1];
    }
    combine<CT>(c.red, v, 4);
    if (warp < NBLK) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                const int o = i * lda + j, ot = j * lda + i;
                const double id = (i == j) ? 1.0 : 0.0, x3 = v[e], x4 = v[2 + e];
                p.X4[o] = x4;
                p.P0[o] = c_t[0] * id + c_t[1] * x1[e] + c_t[2] * x2[e] + c_t[3] * x3;
                p.P1[o] = c_t[4] * id + c_t[5] * x1[e] + c_t[6] * x2[e] + c_t[7] * x3;
                p.P2[o] = c_t[8] * id + c_t[9] * x1[e] + c_t[10] * x2[e] + c_t[11] * x3;
                p.Q3t[ot] = c_t[12] * id + c_t[13] * x1[e] + c_t[14] * x2[e] + c_t[15] * x3 + c_t[16] * x4;
            }
        }
    }
    step_barrier<SYNC>(p.bar, target);
    // ---- steps 3-5: Q2 = P2 + X4.Q3, Q1 = P1 + X4.Q2, R0 = P0 + X4.Q1 (rows of X4 loaded once) ---------------------
    {
        const double* Bsrc[3] = {p.Q3t, p.Q2t, p.Q1t};
        const double* Csrc[3] = {p.P2, p.P1, p.P0};
        for (int st = 0; st < 3; ++st) {
            ld.begin(st == 0);
            if (st == 0) ld.rows(c.S[0], p.X4, i0);
            ld.rows(c.S[1], Bsrc[st], j0);
            double cin[2] = {0, 0};
            if (warp < NBLK) {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (ei < n && ej + e < n) cin[e] = Csrc[st][ei * lda + ej + e];
            }
            ld.wait();
            double a2[2] = {0, 0}, b2[2] = {0, 0};
            mm_q<CT>(c.S[0], c.S[1], lda, ks, a2, b2);
            v[0] = a2[0] + b2[0];
            v[1] = a2[1] + b2[1];
            combine<CT>(c.red, v, 2);
            if (warp < NBLK) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = ei, j = ej + e;
                    if (i < n && j < n) {
                        const double val = v[e] + cin[e];
                        if (st == 0) p.Q2t[j * lda + i] = val;
                        else if (st == 1) p.Q1t[j * lda + i] = val;
                        else {
                            p.R[i * lda + j] = val;
                            p.Rt[j * lda + i] = val;
                            if (s == 0) p.U[i * n + j] = val;
                        }
                    }
                }
            }
            step_barrier<SYNC>(p.bar, target);
        }
    }
    // ---- s squarings: R[q+1] = R[q].R[q] -------------------------------------------------------------------------
    for (int q = 0; q < s; ++q) {
        ld.begin();
        ld.rows(c.S[0], p.R + (size_t)q * nn, i0);
        ld.rows(c.S[1], p.Rt + (size_t)q * nn, j0);
        ld.wait();
        double a2[2] = {0, 0}, b2[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[1], lda, ks, a2, b2);
        v[0] = a2[0] + b2[0];
        v[1] = a2[1] + b2[1];
        combine<CT>(c.red, v, 2);
        if (warp < NBLK) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = ei, j = ej + e;
                if (i < n && j < n) {
                    p.R[(size_t)(q + 1) * nn + i * lda + j] = v[e];
                    p.Rt[(size_t)(q + 1) * nn + j * lda + i] = v[e];
                    if (q == s - 1) p.U[i * n + j] = v[e];            // dense copy of U = exp(K) for the other kernels
                }
            }
        }
        step_barrier<SYNC>(p.bar, target);
    }
    chain_exit<SYNC>(p.bar);
}

// ------------------------------------------------------------------------------------------------------------
template <int CT, int SYNC>
__global__ void __launch_bounds__(NTHR, 2) k_chain_d(DParams p) {
    constexpr int NBLK = Geo<CT>::NBLK;
    extern __shared__ __align__(128) double sm[];
    const int n = p.n;
    const size_t nn = mat_stride(n, CT);
    uint64_t* mbar;
    Ctx c = make_ctx<CT>(sm, n, &mbar);
    Loader<CT> ld{mbar, 0u, c.lda, 0u, 0};
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[91] = clock64();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = c.i0, j0 = c.j0, lda = c.lda, ks = c.ks, ei = c.ei, ej = c.ej;
    const int s = *p.s_in;
    const double sc = *p.scale_in;
    const double* G = p.part + 1 + (size_t)n * n;
    unsigned target = 0;
    auto dxval = [&](int i, int j) -> double { return (i < n && j < n) ? G[(size_t)j * n + i] * sc : 0.0; };   // dX = G^T 2^-s
    double v[4];

    // ---- step 1: dX2 = dX.X + X.dX : S0 = rows of dX, S1 = rows of X, S2 = cols of X, S3 = cols of dX --------------
    ld.begin(true);
    ld.rows(c.S[1], p.X, i0);
    ld.rows(c.S[2], p.Xt, j0);
    for (int r = threadIdx.x >> 5; r < CT; r += NTHR / 32) {
#pragma unroll 4
        for (int k = lane; k < n; k += 32) {
            c.S[0][r * lda + k] = dxval(i0 + r, k);
            c.S[3][r * lda + k] = dxval(k, j0 + r);
        }
    }
    __syncthreads();
    ld.wait();
    {
        double a1[2] = {0, 0}, b1[2] = {0, 0}, a2[2] = {0, 0}, b2[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[2], lda, ks, a1, b1);   // dX.X
        mm_q<CT>(c.S[1], c.S[3], lda, ks, a2, b2);   // X.dX
        v[0] = (a1[0] + b1[0]) + (a2[0] + b2[0]);
        v[1] = (a1[1] + b1[1]) + (a2[1] + b2[1]);
    }
    combine<CT>(c.red, v, 2);
    double d1[2] = {0, 0}, d2[2] = {0, 0};
    if (warp < NBLK) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                d1[e] = dxval(i, j);
                d2[e] = v[e];
                p.dX2[i * lda + j] = v[e];
                p.dX2t[j * lda + i] = v[e];
            }
        }
    }
    step_barrier<SYNC>(p.bar, target);
    // ---- step 2: dX3 = dX2.X + X2.dX ; dX4 = dX2.X2 + X2.dX2 ; dP0..dP2, dQ3 -----------------------------------------
    // kept: S2 = cols X, S3 = cols dX.  S0 <- rows dX2, S1 <- rows X2, S4 <- cols X2, S5 <- cols dX2
    ld.begin(true);
    ld.rows(c.S[0], p.dX2, i0);
    ld.rows(c.S[1], p.X2, i0);
    ld.rows(c.S[4], p.X2t, j0);
    ld.rows(c.S[5], p.dX2t, j0);
    ld.wait();
    {
        double a1[2] = {0, 0}, b1[2] = {0, 0}, a2[2] = {0, 0}, b2[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[2], lda, ks, a1, b1);   // dX2.X
        mm_q<CT>(c.S[1], c.S[3], lda, ks, a2, b2);   // X2.dX
        v[0] = (a1[0] + b1[0]) + (a2[0] + b2[0]);
        v[1] = (a1[1] + b1[1]) + (a2[1] + b2[1]);
        double a3[2] = {0, 0}, b3[2] = {0, 0}, a4[2] = {0, 0}, b4[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[4], lda, ks, a3, b3);   // dX2.X2
        mm_q<CT>(c.S[1], c.S[5], lda, ks, a4, b4);   // X2.dX2
        v[2] = (a3[0] + b3[0]) + (a4[0] + b4[0]);
        v[3] = (a3[1] + b3[1]) + (a4[1] + b4[1]);
    }
    combine<CT>(c.red, v, 4);
    if (warp < NBLK) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = ei, j = ej + e;
            if (i < n && j < n) {
                const int o = i * lda + j, ot = j * lda + i;
                const double d3 = v[e], d4 = v[2 + e];
                p.dX4[o] = d4;
                p.dP0[o] = c_t[1] * d1[e] + c_t[2] * d2[e] + c_t[3] * d3;
                p.dP1[o] = c_t[5] * d1[e] + c_t[6] * d2[e] + c_t[7] * d3;
                p.dP2[o] = c_t[9] * d1[e] + c_t[10] * d2[e] + c_t[11] * d3;
                p.dQ3t[ot] = c_t[13] * d1[e] + c_t[14] * d2[e] + c_t[15] * d3 + c_t[16] * d4;
            }
        }
    }
    step_barrier<SYNC>(p.bar, target);
    // ---- steps 3-5: dQ2 = dP2 + dX4.Q3 + X4.dQ3 ; dQ1 = dP1 + dX4.Q2 + X4.dQ2 ; dR0 = dP0 + dX4.Q1 + X4.dQ1 ----------
    {
        const double* B1[3] = {p.Q3t, p.Q2t, p.Q1t};
        const double* B2[3] = {p.dQ3t, p.dQ2t, p.dQ1t};
        const double* Csrc[3] = {p.dP2, p.dP1, p.dP0};
        for (int st = 0; st < 3; ++st) {
            ld.begin(st == 0);
            if (st == 0) {
                ld.rows(c.S[0], p.dX4, i0);
                ld.rows(c.S[1], p.X4, i0);
            }
            ld.rows(c.S[2], B1[st], j0);
            ld.rows(c.S[3], B2[st], j0);
            double cin[2] = {0, 0};
            if (warp < NBLK) {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (ei < n && ej + e < n) cin[e] = Csrc[st][ei * lda + ej + e];
            }
            ld.wait();
            double a1[2] = {0, 0}, b1[2] = {0, 0}, a2[2] = {0, 0}, b2[2] = {0, 0};
            mm_q<CT>(c.S[0], c.S[2], lda, ks, a1, b1);
            mm_q<CT>(c.S[1], c.S[3], lda, ks, a2, b2);
            v[0] = (a1[0] + b1[0]) + (a2[0] + b2[0]);
            v[1] = (a1[1] + b1[1]) + (a2[1] + b2[1]);
            combine<CT>(c.red, v, 2);
            if (warp < NBLK) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int i = ei, j = ej + e;
                    if (i < n && j < n) {
                        const double val = v[e] + cin[e];
                        if (st == 0) p.dQ2t[j * lda + i] = val;
                        else if (st == 1) p.dQ1t[j * lda + i] = val;
                        else {
                            p.dR[i * lda + j] = val;
                            p.dRt[j * lda + i] = val;
                        }
                    }
                }
            }
            step_barrier<SYNC>(p.bar, target);
        }
    }
    // ---- s squarings: dR[q+1] = dR[q].R[q] + R[q].dR[q] ---------------------------------------------------------------
    const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
    if (dbg) p.dbg[90] = clock64();
    for (int q = 0; q < s; ++q) {
        if (dbg) p.dbg[q * 6 + 0] = clock64();
        ld.begin();
        ld.rows(c.S[0], p.dR + (size_t)q * nn, i0);
        ld.rows(c.S[1], p.R + (size_t)q * nn, i0);
        ld.rows(c.S[2], p.Rt + (size_t)q * nn, j0);
        ld.rows(c.S[3], p.dRt + (size_t)q * nn, j0);
        if (dbg) p.dbg[q * 6 + 1] = clock64();
        ld.wait();
        if (dbg) p.dbg[q * 6 + 2] = clock64();
        double a1[2] = {0, 0}, b1[2] = {0, 0}, a2[2] = {0, 0}, b2[2] = {0, 0};
        mm_q<CT>(c.S[0], c.S[2], lda, ks, a1, b1);
        mm_q<CT>(c.S[1], c.S[3], lda, ks, a2, b2);
        v[0] = (a1[0] + b1[0]) + (a2[0] + b2[0]);
        v[1] = (a1[1] + b1[1]) + (a2[1] + b2[1]);
        if (dbg) p.dbg[q * 6 + 3] = clock64();
        combine<CT>(c.red, v, 2);
        if (warp < NBLK) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = ei, j = ej + e;
                if (i < n && j < n) {
                    p.dR[(size_t)(q + 1) * nn + i * lda + j] = v[e];
                    p.dRt[(size_t)(q + 1) * nn + j * lda + i] = v[e];
                }
            }
        }
        if (dbg) p.dbg[q * 6 + 4] = clock64();
        step_barrier<SYNC>(p.bar, target, dbg ? p.dbg + 56 + 4 * q : nullptr);
        if (dbg) p.dbg[q * 6 + 5] = clock64();
    }
    // ---- gradient in x order (gradient.jl:148-151, 287-290): out = [cost, (g1), g_lambda, g_theta] -------------------
    {
        const double* Lt = p.dR + (size_t)s * nn;
        const double* Ltt = p.dRt + (size_t)s * nn;
        const int cL = n * (n + 1) / 2, off = p.sd ? n : 0;
        for (int idx = threadIdx.x; idx < CT * CT; idx += NTHR) {
            const int i = i0 + idx / CT, j = j0 + idx % CT;
            if (i < n && j < n) {
                if (j <= i) {
                    if (p.write_lambda) p.out[1 + off + i * (i + 1) / 2 + j] = 2.0 * p.part[1 + i * n + j] * (i != j ? 2.0 : 1.0);
                } else p.out[1 + off + cL + i * n - i * (i + 1) / 2 + (j - i - 1)] = 2.0 * (__ldcg(Ltt + i * lda + j) - __ldcg(Lt + i * lda + j));
            }
        }
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            if (threadIdx.x == 0 && p.write_cost) p.out[0] = p.part[0];
            if (p.sd)
                for (int k = threadIdx.x; k < n; k += NTHR) p.out[1 + k] = p.g1[k];
        }
    }
    if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) p.dbg[92] = clock64();
    chain_exit<SYNC>(p.bar);
}

}  // namespace chain
