// panel_kernels.cuh -- the O(N^4) operand build and gradient epilogues as FP64 tensor-core panel GEMMs (sm_100a).
//
//   k_gram     : G = A~^T A~ = (U^T U) o (U^T U) and the staged product G Lambda (single-GEMM formulation, see k_ty)
//   k_prep_mma : A~ panel (64 rows, gathered from U) -> global; B~_I = A~_I Lambda            (fermionic.jl:74-83 operands)
//                and, for the single-GEMM formulation, C~_I = A~_I (Lambda G) with the partial <B~_I, C~_I> = |W~_I|^2
//   k_post_mma : E~_I = fixed-order sum of the fused kernel's pieces; F_I = E~_I Lambda (theta gradient, gradient.jl:57-74);
//                Spart_I = A~_I^T E~_I (lambda gradient, gradient.jl:13-41)
//
// One CTA per 64-row panel, 8 warps, DMMA.8x8x4.  Lambda is symmetric, so it is staged as [column m][k] with the same
// pitch LD = Np + 4 as the panels: every fragment load of the k-contiguous products follows the conflict-free
// (t1 * LD + t0) pattern.  Lambda (padded, written by chain::k_chain_u) and the A~ panel arrive by single TMA bulk copies.
// The template parameter CS splits the output columns (and with them the rows of the staged Lambda) over CS CTAs per
// panel so that the working set fits in shared memory for every supported N (CS = 1 up to N = 104, CS = 4 above).
// k_esum adds the fused kernel's E pieces in a fixed order once, so the GEMM CTAs fetch their E~ panel with one TMA copy.
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

template <int NB, int CS = 1>
struct Geo {
    static constexpr int Np = 8 * NB, LD = Np + 4;
    static constexpr int NBC = (NB + CS - 1) / CS;    // n-blocks per CTA (column split)
    static constexpr int NBW = (NBC + 1) / 2;         // n-blocks per warp column (2 warp columns)
    static constexpr size_t panel_elems = (size_t)ROWS * LD;
    static constexpr size_t lam_elems = (size_t)8 * NBC * LD;   // the Lambda rows (= output columns) of one CTA
};
template <int NB, int CS> inline size_t prep_smem() { return (Geo<NB, CS>::panel_elems + Geo<NB, CS>::lam_elems) * 8 + 64; }
template <int NB, int CS> inline size_t post_smem() { return (2 * Geo<NB, CS>::panel_elems + Geo<NB, CS>::lam_elems) * 8 + 64; }
// k_post_mma restricted to one of its two products: E~ panel + Lambda slice (F) or E~ panel + A~ panel (S)
template <int NB, int CS> inline size_t post_smem_part(bool do_F, bool do_S) {
    return (Geo<NB, CS>::panel_elems + (do_S ? Geo<NB, CS>::panel_elems : 0) + (do_F ? Geo<NB, CS>::lam_elems : 0)) * 8 + 64;
}

// C[16 rows of this warp][its n-blocks] += A[rows][k] . B[cols][k]^T ; both operands k-contiguous with pitch LD
template <int NB, int CS>
__device__ __forceinline__ void gemm_kk(const double* __restrict__ sA, const double* __restrict__ sB, int nblk,
                                        double (&acc)[2][Geo<NB, CS>::NBW][2]) {
    constexpr int LD = Geo<NB, CS>::LD, NBW = Geo<NB, CS>::NBW, KS = Geo<NB, CS>::Np / 4;
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
        for (int nb = 0; nb < NBW; ++nb) bv[nb] = (nb0 + nb < nblk) ? b[nb * 8 * LD + 4 * ks] : 0.0;
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            if (nb0 + nb < nblk) {
                dmma(acc[0][nb][0], acc[0][nb][1], a0, bv[nb]);
                dmma(acc[1][nb][0], acc[1][nb][1], a1, bv[nb]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_prep_mma: A~[r][l] = sw_r U[pa_r][l] U[pq_r][l] (all rows), B~ = A~ Lambda (this shard's panels)
//   grid (panels, CS): CTA (p, c) produces output n-blocks [c*NBC, min((c+1)*NBC, NB)).
// ------------------------------------------------------------------------------------------------------------
template <int NB, int CS>
__global__ void __launch_bounds__(NTHR, 1) k_prep_mma(const double* __restrict__ U, const double* __restrict__ LamP,
                                                      const int* __restrict__ pa, const int* __restrict__ pq, const double* __restrict__ sw,
                                                      int n, int prow_begin, int prow_end, double* __restrict__ At, double* __restrict__ Bt,
                                                      const double* __restrict__ GLP, double* __restrict__ Ct, double* __restrict__ wpart,
                                                      const int* __restrict__ run_if) {
    using G = Geo<NB, CS>;
    constexpr int LD = G::LD, NBW = G::NBW, NBC = G::NBC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_wred[NTHR / 32];
    if (run_if && *run_if == 0) return;      // exact-cost fallback of the lean path: B~ is only built when the flag is up
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sL = sA + G::panel_elems;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sL + G::lam_elems);
    const int panel = blockIdx.x, cs = blockIdx.y;
    const int nb_lo = cs * NBC, nblk = min(NBC, NB - nb_lo);
    const bool need_b = panel >= prow_begin && panel < prow_end && nblk > 0;
    if (!need_b && cs != 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(mbar);
        if (need_b) {
            const uint32_t bytes = (uint32_t)((size_t)8 * nblk * LD * 8);
            mbar_expect(mbar, bytes);
            bulk(sL, LamP + (size_t)8 * nb_lo * LD, bytes, mbar);
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
            if (cs == 0) At[(size_t)r * LD + l] = v;
        }
    }
    if (!need_b) return;
    __syncthreads();
    mbar_wait(mbar, 0);
    double acc[2][NBW][2];
    gemm_kk<NB, CS>(sA, sL, nblk, acc);
    const int t0 = lane & 3, t1 = lane >> 2, wr = warp >> 1, wc = warp & 1;
    double* out = Bt + (size_t)((panel - prow_begin) * ROWS + 16 * wr + t1) * LD;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            const int blk = wc * NBW + nb;
            if (blk < nblk)
                *reinterpret_cast<double2*>(out + (size_t)mb * 8 * LD + 8 * (nb_lo + blk) + 2 * t0) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
        }
    if (!GLP) return;
    // ---- single-GEMM formulation: C~ = A~ (Lambda G) (the staged matrix is G Lambda = (Lambda G)^T), <B~, C~> ----------
    __syncthreads();                                        // every warp is done reading Lambda from sL
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = (uint32_t)((size_t)8 * nblk * LD * 8);
        mbar_expect(mbar, bytes);
        bulk(sL, GLP + (size_t)8 * nb_lo * LD, bytes, mbar);
    }
    mbar_wait(mbar, 1);
    double acc2[2][NBW][2];
    gemm_kk<NB, CS>(sA, sL, nblk, acc2);
    double* outc = Ct + (size_t)((panel - prow_begin) * ROWS + 16 * wr + t1) * LD;
    double wsum = 0.0;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) {
            const int blk = wc * NBW + nb;
            if (blk < nblk) {
                *reinterpret_cast<double2*>(outc + (size_t)mb * 8 * LD + 8 * (nb_lo + blk) + 2 * t0) = make_double2(acc2[mb][nb][0], acc2[mb][nb][1]);
                wsum = fma(acc[mb][nb][0], acc2[mb][nb][0], fma(acc[mb][nb][1], acc2[mb][nb][1], wsum));
            }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    if (lane == 0) s_wred[warp] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NTHR / 32; ++w) t += s_wred[w];
        wpart[(size_t)(panel - prow_begin) * CS + cs] = t;
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_build_a (lean single-GEMM path): A~[r][l] = sw_r U[pa_r][l] U[pq_r][l] for ALL rows (the only operand fused::k_ty needs),
// zero in the padding columns / rows; the LAST CTA instead computes the O(N^2) closed-form pieces that replace the panel
// GEMMs B~ = A~ Lambda and C~ = A~ (Lambda G) of the first formulation:
//     o_l = sum_p U[p][l]^2      (diagonal of U^T U;  G = A~^T A~ = (U^T U) o (U^T U) has off-diagonal entries
//     d_l = o_l^2                 (U^T U)_lm^2 <= (2^s eps)^2 ~ 1e-24: below rounding, so G == diag(d) to working precision)
//     t_l = sum_k Lambda[l][k]^2 d_k        (diagonal of Lambda G Lambda)
//     wn  = sum_l d_l t_l = |W~|^2 = tr(Lambda G Lambda G)
// aux = [o (n) | d (n) | t (n) | wn].  With them   S = d_a Lambda_ab d_b - A~^T Y,   G_ab = 4 U_ab o_b t_b - gather(Y Lambda),
// cost = |T~|^2 + wn - 2 <Lambda, A~^T Y>   (k_post_reduce), where Y = T~ A~ is the hot kernel's product.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_build_a(const double* __restrict__ U, const double* __restrict__ Lam, const int* __restrict__ pa,
                                                  const int* __restrict__ pq, const double* __restrict__ sw, int n, int LD, int nrows,
                                                  double* __restrict__ At, double* aux) {
    if (blockIdx.x == gridDim.x - 1) {
        // 1024 threads = 8 row groups x 128 columns: every thread's <= ceil(n/8) loads are independent and in flight together, so
        // each of the two phases costs about one memory round trip; group partials are combined in a fixed order.
        __shared__ double red[8][128];
        __shared__ double s_d[256], s_w[256];
        const int tx = threadIdx.x & 127, grp = threadIdx.x >> 7;
        for (int l0 = 0; l0 < n; l0 += 128) {
            const int l = l0 + tx;
            double a = 0.0;
            if (l < n) {
#pragma unroll 8
                for (int p = grp; p < n; p += 8) {
                    const double u = U[(size_t)p * n + l];
                    a = fma(u, u, a);
                }
            }
            red[grp][tx] = a;
            __syncthreads();
            if (grp == 0 && l < n) {
                double o = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) o += red[k][tx];
                aux[l] = o;
                aux[n + l] = o * o;
                s_d[l] = o * o;
            }
            __syncthreads();
        }
        for (int l0 = 0; l0 < n; l0 += 128) {                 // t: Lambda is symmetric, so column l is swept (coalesced)
            const int l = l0 + tx;
            double a = 0.0;
            if (l < n) {
#pragma unroll 8
                for (int k = grp; k < n; k += 8) {
                    const double v = Lam[(size_t)k * n + l];
                    a = fma(v * v, s_d[k], a);
                }
            }
            red[grp][tx] = a;
            __syncthreads();
            if (grp == 0 && l < n) {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += red[k][tx];
                aux[2 * n + l] = t;
                s_w[l] = s_d[l] * t;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {                               // fixed-order sum (n <= 168 terms)
            double s = 0.0;
            for (int l = 0; l < n; ++l) s += s_w[l];
            aux[3 * n] = s;
        }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = gridDim.x - 1;
    for (int r = blockIdx.x * 32 + warp; r < nrows; r += nb * 32) {
        const double w = sw[r];
        const double* ua = U + (size_t)pa[r] * n;
        const double* uq = U + (size_t)pq[r] * n;
        for (int l = lane; l < LD; l += 32) At[(size_t)r * LD + l] = (l < n && w != 0.0) ? w * ua[l] * uq[l] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_gram: G[l][m] = (sum_p U[p][l] U[p][m])^2 (= A~^T A~ over all N^2 rows; the identity up to the rounding of exp(K)) and
//         GLP[m][k] = (G Lambda)[m][k], padded pitch LD: the [column][k] staging of Lambda G for k_prep_mma's second GEMM.
//   one CTA per row m; U and Lambda are dense n x n, row-major.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_gram(const double* __restrict__ U, const double* __restrict__ Lam, int n, int LD,
                                               double* __restrict__ GLP) {
    // 1024 threads = 8 row groups x 128 columns: every thread issues its <= ceil(n/8) independent loads back to back, so a
    // phase costs about one memory round trip instead of n of them; group partials are combined in a fixed order.
    __shared__ double um[256];        // column m of U
    __shared__ double g[256];         // row m of G
    __shared__ double red[8][128];
    const int m = blockIdx.x, tx = threadIdx.x & 127, grp = threadIdx.x >> 7;
    for (int p = threadIdx.x; p < n; p += 1024) um[p] = U[(size_t)p * n + m];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += 128) {
        const int j = j0 + tx;
        double d = 0.0;
        if (j < n) {
#pragma unroll 8
            for (int p = grp; p < n; p += 8) d = fma(um[p], U[(size_t)p * n + j], d);
        }
        red[grp][tx] = d;
        __syncthreads();
        if (grp == 0 && j < n) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k][tx];
            g[j] = t * t;
        }
        __syncthreads();
    }
    for (int j0 = 0; j0 < n; j0 += 128) {
        const int j = j0 + tx;
        double a = 0.0;
        if (j < n) {
#pragma unroll 8
            for (int i = grp; i < n; i += 8) a = fma(g[i], Lam[(size_t)i * n + j], a);
        }
        red[grp][tx] = a;
        __syncthreads();
        if (grp == 0 && j < n) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k][tx];
            GLP[(size_t)m * LD + j] = t;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_esum: E~[row][LD] = fixed-order sum of the fused kernel's pieces (two warp-column partials per visit)
// ------------------------------------------------------------------------------------------------------------
//         single-GEMM formulation (Ct != null): the pieces hold Y = T~ A~ and E~ = C~ - Y
__global__ void __launch_bounds__(256) k_esum(const double* __restrict__ Epiece, const int* __restrict__ piece_first, int Np, int LD,
                                              double* __restrict__ Et, const double* __restrict__ Ct) {
    const int panel = blockIdx.x;
    const int v0 = piece_first[panel], v1 = piece_first[panel + 1];
    const int total = ROWS * Np;
    for (int base = blockIdx.y * 1024 + threadIdx.x; base < total; base += gridDim.y * 1024) {
        double acc4[4] = {0.0, 0.0, 0.0, 0.0};
        for (int v = v0; v < v1; ++v) {
            const double* p0 = Epiece + (size_t)(v * 2 + 0) * ROWS * Np;
            const double* p1 = Epiece + (size_t)(v * 2 + 1) * ROWS * Np;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 256;
                if (idx < total) acc4[u] += __ldcs(p0 + idx) + __ldcs(p1 + idx);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * 256;
            if (idx < total) {
                const size_t o = (size_t)(panel * ROWS + idx / Np) * LD + (idx % Np);
                Et[o] = Ct ? Ct[o] - acc4[u] : acc4[u];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_post_mma: per local panel and column split: F[:, cols] = E~ Lambda[:, cols] ; Spart[:, cols] = A~_panel^T E~[:, cols]
// ------------------------------------------------------------------------------------------------------------
//   Without column split (CS = 1) the E~ panel itself is assembled in shared memory first: fixed-order sum of the hot
//   kernel's pieces (two warp-column partials per visit), subtracted from C~ in the single-GEMM formulation (E~ = C~ - Y),
//   so it never round-trips HBM.  With CS > 1 the CS CTAs of a panel would each repeat that sum: k_esum does it once.
template <int NB, int CS>
__global__ void __launch_bounds__(NTHR, 1) k_post_mma(const double* __restrict__ Epiece, const int* __restrict__ piece_first,
                                                      const double* __restrict__ Ct, const double* __restrict__ Et,
                                                      const double* __restrict__ LamP, const double* __restrict__ At, int n,
                                                      int row_begin, double* __restrict__ F, double* __restrict__ Spart, int do_S,
                                                      int do_F) {
    // do_F / do_S select the products (both by default); a launch that does only one of them needs only that product's
    // operands in shared memory (post_smem_part), so two such launches can run side by side on different streams.
    using G = Geo<NB, CS>;
    constexpr int Np = G::Np, LD = G::LD, NBW = G::NBW, NBC = G::NBC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sE = reinterpret_cast<double*>(smem_raw);
    double* sAp = sE + G::panel_elems;
    double* sL = do_S ? sAp + G::panel_elems : sAp;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sL + (do_F ? G::lam_elems : 0));
    const int panel = blockIdx.x, cs = blockIdx.y;
    const int nb_lo = cs * NBC, nblk = min(NBC, NB - nb_lo);
    if (nblk <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = lane & 3, t1 = lane >> 2;
    if (threadIdx.x == 0) {
        mbar_init(mbar);
        const uint32_t lbytes = (uint32_t)((size_t)8 * nblk * LD * 8), pbytes = (uint32_t)(G::panel_elems * 8);
        mbar_expect(mbar, (do_F ? lbytes : 0u) + (do_S ? pbytes : 0u) + (Et ? pbytes : 0u));
        if (Et) bulk(sE, Et + (size_t)panel * ROWS * LD, pbytes, mbar);   // column-split launches: E~ was summed once by k_esum
        if (do_F) bulk(sL, LamP + (size_t)8 * nb_lo * LD, lbytes, mbar);
        if (do_S) bulk(sAp, At + (size_t)(row_begin + panel * ROWS) * LD, pbytes, mbar);
    }
    if (!Et) {   // E~ panel -> shared memory while the TMA copies are in flight.  Four elements per thread and trip so that the
        // 4 x (2 visits x 2 + 1) independent loads are in flight together (one memory round trip per trip, not per load).
        const int v0 = piece_first[panel], v1 = piece_first[panel + 1];
        const double* cpanel = Ct ? Ct + (size_t)panel * ROWS * LD : nullptr;
        constexpr int HALF = LD / 2, TOTAL = ROWS * HALF, UN = 4;
        for (int base = threadIdx.x; base < TOTAL; base += UN * NTHR) {
            double2 a[UN];
            int row[UN], c2[UN];
            bool live[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int idx = base + u * NTHR;
                row[u] = idx / HALF;
                c2[u] = 2 * (idx % HALF);
                live[u] = idx < TOTAL && c2[u] < Np;
                a[u] = make_double2(0.0, 0.0);
            }
            for (int v = v0; v < v1; ++v) {
                double2 y0[UN], y1[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    y0[u] = y1[u] = make_double2(0.0, 0.0);
                    if (live[u]) {
                        y0[u] = __ldcs(reinterpret_cast<const double2*>(Epiece + ((size_t)(v * 2 + 0) * ROWS + row[u]) * Np + c2[u]));
                        y1[u] = __ldcs(reinterpret_cast<const double2*>(Epiece + ((size_t)(v * 2 + 1) * ROWS + row[u]) * Np + c2[u]));
                    }
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    a[u].x += y0[u].x + y1[u].x;
                    a[u].y += y0[u].y + y1[u].y;
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (base + u * NTHR >= TOTAL) continue;
                if (cpanel && live[u]) {
                    const double2 cc = *reinterpret_cast<const double2*>(cpanel + (size_t)row[u] * LD + c2[u]);
                    a[u].x = cc.x - a[u].x;
                    a[u].y = cc.y - a[u].y;
                }
                *reinterpret_cast<double2*>(sE + (size_t)row[u] * LD + c2[u]) = a[u];
            }
        }
    }
    __syncthreads();
    mbar_wait(mbar, 0);
    // ---- F = E~ Lambda --------------------------------------------------------------------------------------
    if (do_F) {
        double acc[2][NBW][2];
        gemm_kk<NB, CS>(sE, sL, nblk, acc);
        const int wr = warp >> 1, wc = warp & 1;
        double* out = F + (size_t)(panel * ROWS + 16 * wr + t1) * Np;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
                const int blk = wc * NBW + nb;
                if (blk < nblk)
                    *reinterpret_cast<double2*>(out + (size_t)mb * 8 * Np + 8 * (nb_lo + blk) + 2 * t0) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
            }
    }
    if (!do_S) return;
    // ---- Spart[i][j] = sum_r A~[r][i] E~[r][j], j in this CTA's columns: warps = 4 (i groups) x 2 (j groups) -----------
    {
        constexpr int MBW = (NB + 3) / 4;                   // m-blocks per warp row group
        const int wr = warp >> 1, wc = warp & 1;
        const int mb0 = wr * MBW, nbw0 = wc * NBW;
        double acc[MBW][NBW][2];
#pragma unroll
        for (int mb = 0; mb < MBW; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        const double* a = sAp + t0 * LD + 8 * mb0 + t1;
        const double* b = sE + t0 * LD + 8 * (nb_lo + nbw0) + t1;
#pragma unroll 2
        for (int ks = 0; ks < ROWS / 4; ++ks) {
            double av[MBW], bv[NBW];
#pragma unroll
            for (int mb = 0; mb < MBW; ++mb) av[mb] = (mb0 + mb < NB) ? a[4 * ks * LD + 8 * mb] : 0.0;
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) bv[nb] = (nbw0 + nb < nblk) ? b[4 * ks * LD + 8 * nb] : 0.0;
#pragma unroll
            for (int mb = 0; mb < MBW; ++mb)
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb)
                    if (mb0 + mb < NB && nbw0 + nb < nblk) dmma(acc[mb][nb][0], acc[mb][nb][1], av[mb], bv[nb]);
        }
        double* out = Spart + (size_t)panel * n * n;
#pragma unroll
        for (int mb = 0; mb < MBW; ++mb)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
                const int i = 8 * (mb0 + mb) + t1, j = 8 * (nb_lo + nbw0 + nb) + 2 * t0;
                if (mb0 + mb < NB && nbw0 + nb < nblk && i < n) {
                    if (j < n) out[(size_t)i * n + j] = acc[mb][nb][0];
                    if (j + 1 < n) out[(size_t)i * n + j + 1] = acc[mb][nb][1];
                }
            }
    }
}

}  // namespace panel
