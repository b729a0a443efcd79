// panel_kernels.cuh -- the O(N^4) operand build and gradient epilogues as FP64 tensor-core panel GEMMs (sm_100a).
//
//   k_prep_mma : A~ panel (64 rows, gathered from U) -> global; B~_I = A~_I Lambda            (fermionic.jl:74-83 operands)
//   k_post_mma : E~_I = fixed-order sum of the fused kernel's pieces; F_I = E~_I Lambda (theta gradient, gradient.jl:57-74);
//                Spart_I = A~_I^T E~_I (lambda gradient, gradient.jl:13-41)
//
// One CTA per 64-row panel, 8 warps, DMMA.8x8x4.  Lambda is symmetric, so it is staged as [column m][k] with the same
// pitch LD = Np + 4 as the panels: every fragment load of the k-contiguous products follows the conflict-free
// (t1 * LD + t0) pattern.  Lambda (padded, written by chain::k_chain_u) and the A~ panel arrive by single TMA bulk copies.
// Used when the panels and Lambda fit in shared memory together (NB <= 13, i.e. N <= 104); larger N keeps the
// FMA kernels of small_kernels.cuh.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace panel {

constexpr int ROWS = 64;
constexpr int NTHR = 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_init(uint64_t* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(b)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
                 "r"(bytes), "r"(s32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}

template <int NB>
struct Geo {
    static constexpr int Np = 8 * NB, LD = Np + 4;
    static constexpr int NBW = (NB + 1) / 2;          // n-blocks per warp column (2 warp columns)
    static constexpr size_t panel_elems = (size_t)ROWS * LD;
    static constexpr size_t lam_elems = (size_t)Np * LD;
};
template <int NB> inline size_t prep_smem() { return (Geo<NB>::panel_elems + Geo<NB>::lam_elems) * 8 + 64; }
template <int NB> inline size_t post_smem() { return (2 * Geo<NB>::panel_elems + Geo<NB>::lam_elems) * 8 + 64; }

// C[16 rows of this warp][its n-blocks] += A[rows][k] . B[cols][k]^T ; both operands k-contiguous with pitch LD
template <int NB>
__device__ __forceinline__ void gemm_kk(const double* __restrict__ sA, const double* __restrict__ sB, double (&acc)[2][Geo<NB>::NBW][2]) {
    constexpr int LD = Geo<NB>::LD, NBW = Geo<NB>::NBW, KS = Geo<NB>::Np / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2, wr = warp >> 1, wc = warp & 1;
    const int nb0 = wc * NBW;                               // first n-block of this warp column
    const double* a = sA + (16 * wr + t1) * LD + t0;
    const double* b = sB + (8 * nb0 + t1) * LD + t0;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
#pragma unroll 2
    for (int ks = 0; ks < KS; ++ks) {
        const double a0 = a[4 * ks], a1 = a[8 * LD + 4 * ks];
        double bv[NBW];
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) bv[nb] = (nb0 + nb < NB) ? b[nb * 8 * LD + 4 * ks] : 0.0;
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            if (nb0 + nb < NB) {
                dmma(acc[0][nb][0], acc[0][nb][1], a0, bv[nb]);
                dmma(acc[1][nb][0], acc[1][nb][1], a1, bv[nb]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_prep_mma: A~[r][l] = sw_r U[pa_r][l] U[pq_r][l] (all rows), B~ = A~ Lambda (this shard's panels)
// ------------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NTHR, 1) k_prep_mma(const double* __restrict__ U, const double* __restrict__ LamP,
                                                      const int* __restrict__ pa, const int* __restrict__ pq, const double* __restrict__ sw,
                                                      int n, int prow_begin, int prow_end, double* __restrict__ At, double* __restrict__ Bt) {
    using G = Geo<NB>;
    constexpr int LD = G::LD, NBW = G::NBW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sL = sA + G::panel_elems;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sL + G::lam_elems);
    const int panel = blockIdx.x;
    const bool need_b = panel >= prow_begin && panel < prow_end;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(mbar);
        if (need_b) {
            mbar_expect(mbar, (uint32_t)(G::lam_elems * 8));
            bulk(sL, LamP, (uint32_t)(G::lam_elems * 8), mbar);
        }
    }
    // gather the A~ panel: warp w builds rows 8w .. 8w+7, lanes sweep the columns (coalesced reads of U rows)
#pragma unroll 2
    for (int rr = 0; rr < 8; ++rr) {
        const int r = panel * ROWS + 8 * warp + rr;
        const double w = sw[r];
        const double* ua = U + (size_t)pa[r] * n;
        const double* uq = U + (size_t)pq[r] * n;
#pragma unroll
        for (int l = lane; l < LD; l += 32) {
            const double v = (l < n && w != 0.0) ? w * ua[l] * uq[l] : 0.0;
            sA[(8 * warp + rr) * LD + l] = v;
            At[(size_t)r * LD + l] = v;
        }
    }
    if (!need_b) return;
    __syncthreads();
    mbar_wait(mbar, 0);
    double acc[2][NBW][2];
    gemm_kk<NB>(sA, sL, acc);
    const int t0 = lane & 3, t1 = lane >> 2, wr = warp >> 1, wc = warp & 1;
    double* out = Bt + (size_t)((panel - prow_begin) * ROWS + 16 * wr + t1) * LD;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            const int col = 8 * (wc * NBW + nb) + 2 * t0;
            if (wc * NBW + nb < NB) {
                double2 v2 = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                *reinterpret_cast<double2*>(out + (size_t)mb * 8 * LD + col) = v2;
            }
        }
}

// ------------------------------------------------------------------------------------------------------------
// k_post_mma: per local panel: E~ (piece sum) ; F = E~ Lambda ; Spart = A~_panel^T E~
// ------------------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(NTHR, 1) k_post_mma(const double* __restrict__ Epiece, const int* __restrict__ piece_first,
                                                      const double* __restrict__ LamP, const double* __restrict__ At, int n, int row_begin,
                                                      double* __restrict__ F, double* __restrict__ Spart, int do_S) {
    using G = Geo<NB>;
    constexpr int Np = G::Np, LD = G::LD, NBW = G::NBW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sE = reinterpret_cast<double*>(smem_raw);
    double* sAp = sE + G::panel_elems;
    double* sL = sAp + G::panel_elems;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sL + G::lam_elems);
    const int panel = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2;
    if (threadIdx.x == 0) {
        mbar_init(mbar);
        mbar_expect(mbar, (uint32_t)((G::lam_elems + G::panel_elems) * 8));
        bulk(sL, LamP, (uint32_t)(G::lam_elems * 8), mbar);
        bulk(sAp, At + (size_t)(row_begin + panel * ROWS) * LD, (uint32_t)(G::panel_elems * 8), mbar);
    }
    // E~ panel = sum of pieces in a fixed order (two warp-column partials per visit); pad columns zero
    {
        const int v0 = piece_first[panel], v1 = piece_first[panel + 1];
        for (int base = threadIdx.x; base < ROWS * Np; base += 4 * NTHR) {
            double acc4[4] = {0.0, 0.0, 0.0, 0.0};
            for (int v = v0; v < v1; ++v) {
                const double* p0 = Epiece + (size_t)(v * 2 + 0) * ROWS * Np;
                const double* p1 = Epiece + (size_t)(v * 2 + 1) * ROWS * Np;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = base + u * NTHR;
                    if (idx < ROWS * Np) acc4[u] += p0[idx] + p1[idx];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * NTHR;
                if (idx < ROWS * Np) sE[(idx / Np) * LD + (idx % Np)] = acc4[u];
            }
        }
        for (int idx = threadIdx.x; idx < ROWS * 4; idx += NTHR) sE[(idx >> 2) * LD + Np + (idx & 3)] = 0.0;
    }
    __syncthreads();
    mbar_wait(mbar, 0);
    // ---- F = E~ Lambda --------------------------------------------------------------------------------------
    {
        double acc[2][NBW][2];
        gemm_kk<NB>(sE, sL, acc);
        const int wr = warp >> 1, wc = warp & 1;
        double* out = F + (size_t)(panel * ROWS + 16 * wr + t1) * Np;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
                const int col = 8 * (wc * NBW + nb) + 2 * t0;
                if (wc * NBW + nb < NB) *reinterpret_cast<double2*>(out + (size_t)mb * 8 * Np + col) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            }
    }
    if (!do_S) return;
    // ---- Spart[i][j] = sum_r A~[r][i] E~[r][j] : warps = 4 (i groups) x 2 (j groups) -----------------------------------
    {
        constexpr int MBW = (NB + 3) / 4;                   // m-blocks per warp row group
        const int wr = warp >> 1, wc = warp & 1;
        const int mb0 = wr * MBW, nb0 = wc * NBW;
        double acc[MBW][NBW][2];
#pragma unroll
        for (int mb = 0; mb < MBW; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        const double* a = sAp + t0 * LD + 8 * mb0 + t1;
        const double* b = sE + t0 * LD + 8 * nb0 + t1;
#pragma unroll 2
        for (int ks = 0; ks < ROWS / 4; ++ks) {
            double av[MBW], bv[NBW];
#pragma unroll
            for (int mb = 0; mb < MBW; ++mb) av[mb] = (mb0 + mb < NB) ? a[4 * ks * LD + 8 * mb] : 0.0;
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) bv[nb] = (nb0 + nb < NB) ? b[4 * ks * LD + 8 * nb] : 0.0;
#pragma unroll
            for (int mb = 0; mb < MBW; ++mb)
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb)
                    if (mb0 + mb < NB && nb0 + nb < NB) dmma(acc[mb][nb][0], acc[mb][nb][1], av[mb], bv[nb]);
        }
        double* out = Spart + (size_t)panel * n * n;
#pragma unroll
        for (int mb = 0; mb < MBW; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
                const int i = 8 * (mb0 + mb) + t1, j = 8 * (nb0 + nb) + 2 * t0;
                if (mb0 + mb < NB && nb0 + nb < NB && i < n) {
                    if (j < n) out[(size_t)i * n + j] = acc[mb][nb][0];
                    if (j + 1 < n) out[(size_t)i * n + j + 1] = acc[mb][nb][1];
                }
            }
    }
}

}  // namespace panel
