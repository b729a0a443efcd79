// fused_kernel.cuh -- the hot kernels of the CSA engine (sm_100a): k_ty (one R x R x N product per evaluation, further
// down) and k_fused (two chained products; residual update, cost-only evaluation, exact-cost fallback).
//
// k_fused: for a residual matrix T~ (rows i, cols j; j contiguous) and operand matrices A~ (rows x Np) and
// B~ = A~ Lambda it computes, per 64x64 tile,
//      D   = B~_I A~_J^T - T~_IJ            (GEMM #1, K = Np, FP64 DMMA.8x8x4)
//      cost += sum D^2
//      E_I += D A~_J                        (GEMM #2, K = 64, D never leaves registers)
// which is the reference's reconstruction (unitaries.jl:254-260), residual (cost.jl:16-25) and the
// contraction shared by every gradient term (gradient.jl:13-28, 57-74) in one pass over T~.
//
// Layout / schedule
//   * persistent CTAs (one per SM), 8 warps (4 row groups x 2 column groups, warp tile 16 x 32); the TMA producer
//     role rotates over lane 0 of the eight warps (a ninth warp would cap every thread at 168 registers, a fixed
//     producer warp paces the whole CTA); the flattened (row panel, column tile) steps are split evenly over CTAs.
//   * B~ panel (64 x ld) resident in shared memory per row panel; A~ column sub-tiles (32 x ld) stream
//     through an S-slot ring filled by the TMA engine (cp.async.bulk, SASS UBLKCP) and handed over with
//     mbarrier full/empty pairs.
//   * ld = Np + 4 (Np = 8*NB) and the row/column slot permutation perm8 make every LDS.128 of both GEMMs
//     bank-conflict free; the accumulator fragment of GEMM #1 is, with that permutation, directly the
//     A fragment of GEMM #2 (no shuffle, no shared-memory round trip).
//   * T~ is read exactly once, straight into the (negated) accumulators, one step ahead.
//   * E partials are written per (CTA visit, column group) and reduced in a fixed order later:
//     results are bit-reproducible run to run.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fused {

constexpr int ROWS = 64;        // rows per panel (CTA)
constexpr int COLS = 64;        // columns per step (CTA)
constexpr int SUB = 32;         // columns per warp column group = rows of one A~ ring slot
constexpr int NCONS = 8;        // consumer warps
constexpr int NTHREADS = NCONS * 32;

struct Params {
    const double* T;        // [P*64][ldT] local rows of T~
    double* Tst;            // STORE: destination for T~ - W~ (may alias T)
    const double* Brow;     // B~ rows of the local panels  [P*64][ld]
    const double* Acol;     // A~ rows of all columns       [Q*64][ld]
    double* Epiece;         // [visits][2][64][Np]
    double* cost_part;      // [gridDim.x]
    const int* visit_start; // [gridDim.x]
    long long ldT;
    int P, Q, ld, Np, S;
    const int* run_if;      // optional device flag: the kernel returns at once when *run_if == 0 (exact-cost fallback)
    int dephase;            // cycles by which column group 1 starts late (0 = off)
    int lookahead;          // sub-tiles requested ahead of the current step (0 = default)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
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
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// TMA engine, non-tensor form: contiguous global -> shared, completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// slot -> position inside an 8-group.  Even slots 2j -> j, odd slots 2j+1 -> 4 + (j+2)%4.
__host__ __device__ __forceinline__ int perm8(int m) { return (m & 1) ? (4 + (((m >> 1) + 2) & 3)) : (m >> 1); }

inline size_t smem_bytes(int ld, int S) { return (size_t)(ROWS + S * SUB) * ld * 8 + (2 * S + 2) * 8 + NCONS * 8 + 64; }

template <int NB, bool DO_E, bool STORE, bool TPREF>
__global__ void __launch_bounds__(NTHREADS, 1) k_fused(Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ld = 8 * NB + 4;   // == p.ld: compile-time pitch turns the fragment addressing into immediates
    if (p.run_if && *p.run_if == 0) return;
    const int S = p.S, Q = p.Q;
    double* sB = reinterpret_cast<double*>(smem_raw);
    double* sA = sB + ROWS * ld;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * SUB * ld);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* bfull = bars + 2 * S;
    uint64_t* bempty = bfull + 1;
    double* sred = reinterpret_cast<double*>(bempty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)p.P * Q;
    const long long tbeg = total * blockIdx.x / gridDim.x;
    const long long tend = total * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NCONS / 2);
        }
        mbar_init(bfull, 1);
        mbar_init(bempty, NCONS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---------------------------------- consumers -------------------------------------------------------
    const int t0 = lane & 3, t1 = lane >> 2;
    // Warps w and w+4 share a scheduler (SMSP w%4): give each scheduler one warp of either column group, and start
    // column group 1 half a step late, so that the non-DMMA stretches of one warp (operand/T loads, barrier waits,
    // phase changes) fall into the DMMA stretches of the other instead of lining up.
    const int wc = (warp >> 2) & 1, wr = warp & 3;
    if (p.dephase && wc == 1) {
        const long long t_start = clock64();
        while (clock64() - t_start < (long long)p.dephase) { }
    }
    const int prow = perm8(t1);
    const int pc0 = t0, pc1 = 4 + ((t0 + 2) & 3);  // perm8(2*t0), perm8(2*t0+1)

    double e[2][NB][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int b = 0; b < NB; ++b) e[mb][b][0] = e[mb][b][1] = 0.0;
    double cost = 0.0;
    double acc[2][4][2];
    double tn[2][4][2];

    // element (mb, nb, v) of this lane: row 16*wr + 8*mb + prow, col 32*wc + 8*nb + (v ? pc1 : pc0)
    const long long lane_off = (long long)(16 * wr + prow) * p.ldT + 32 * wc;
    auto loadT = [&](double(&dst)[2][4][2], int panel_, int ct_) {
        const double* base = p.T + (long long)panel_ * ROWS * p.ldT + (long long)ct_ * COLS + lane_off;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                dst[mb][nb][0] = __ldcs(base + (long long)mb * 8 * p.ldT + nb * 8 + pc0);
                dst[mb][nb][1] = __ldcs(base + (long long)mb * 8 * p.ldT + nb * 8 + pc1);
            }
    };

    int panel = (int)(tbeg / Q), ct = (int)(tbeg % Q);
    int gi = 0, pv = 0;
    if (TPREF && tbeg < tend) loadT(tn, panel, ct);

    // ---- operand ring: sub-tile g = 2*step + column group lives in slot g % S; the producer role rotates over the warps
    // of the column group that consumes the sub-tile (see k_ty below); no single warp carries the ~650 cycles per request
    // and paces the CTA.  The B~ panel (once per row panel) stays with warp 0.
    const int g_total = 2 * (int)(tend - tbeg);
    int lookahead = p.lookahead > 0 ? min(p.lookahead, S - 2) : ((S >= 6) ? 2 : (S - 2));   // sub-tiles requested ahead of use
    lookahead &= ~1;
    const int ct0 = ct;
    const uint32_t panel_bytes = (uint32_t)(ROWS * ld * 8), sub_bytes = (uint32_t)(SUB * ld * 8);
    auto request = [&](int g) {
        const int slot = g % S, use = g / S, c = g & 1;
        const int ctg = (ct0 + (g >> 1)) % Q;
        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
        mbar_expect_tx(&full[slot], sub_bytes);
        bulk_g2s(sA + (size_t)slot * SUB * ld, p.Acol + ((size_t)ctg * COLS + c * SUB) * ld, sub_bytes, &full[slot]);
    };
    auto requester = [&](int g) { return 4 * (g & 1) + ((g >> 1) & 3); };   // a warp of column group g & 1 (wc = warp >> 2)
    if (lane == 0)
        for (int g = 0; g < lookahead && g < g_total; ++g)
            if (requester(g) == warp) request(g);
    __syncwarp();

    const double* sBrow = sB + (16 * wr + prow) * ld + 2 * t0;

    for (long long t = tbeg; t < tend; ++t, ++gi) {
        const bool newpanel = (t == tbeg || ct == 0);
        {
            const int g1 = 2 * gi + lookahead;         // the sub-tiles of step gi + lookahead/2
            if (lane == 0) {
                if (warp == 0 && newpanel) {
                    if (pv > 0) mbar_wait(bempty, (pv - 1) & 1);   // every warp is done with the old B~ panel
                    mbar_expect_tx(bfull, panel_bytes);
                    bulk_g2s(sB, p.Brow + (size_t)panel * ROWS * ld, panel_bytes, bfull);
                }
                if (g1 < g_total && requester(g1) == warp) request(g1);
                if (g1 + 1 < g_total && requester(g1 + 1) == warp) request(g1 + 1);
            }
            __syncwarp();
        }
        if (newpanel) {
            mbar_wait(bfull, pv & 1);
            ++pv;
        }
        if (TPREF) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    acc[mb][nb][0] = -tn[mb][nb][0];
                    acc[mb][nb][1] = -tn[mb][nb][1];
                }
        } else {
            loadT(acc, panel, ct);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    acc[mb][nb][0] = -acc[mb][nb][0];
                    acc[mb][nb][1] = -acc[mb][nb][1];
                }
        }
        const int g = 2 * gi + wc;
        const int slot = g % S;
        mbar_wait(&full[slot], (g / S) & 1);
        const double* sAs = sA + (size_t)slot * SUB * ld;

        // ---------------- GEMM #1: acc += B~_I (16 x Np) . A~_J^T (Np x 32) ----------------------------
        {
            const double* sArow = sAs + prow * ld + 2 * t0;
#pragma unroll
            for (int kc = 0; kc < NB * 8; kc += 8) {
                double2 a[2], b[4];
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) a[mb] = *reinterpret_cast<const double2*>(sBrow + mb * 8 * ld + kc);
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = *reinterpret_cast<const double2*>(sArow + nb * 8 * ld + kc);
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        dmma(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
                        dmma(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
                    }
            }
        }
        const bool lastp = (ct == Q - 1) || (t == tend - 1);
        if (lastp) {  // the B~ panel is no longer needed by this warp: let the producer refill it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before the async-proxy refill
            __syncwarp();
            if (lane == 0) mbar_arrive(bempty);
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) cost += acc[mb][nb][0] * acc[mb][nb][0] + acc[mb][nb][1] * acc[mb][nb][1];

        if (STORE) {
            double* base = p.Tst + (long long)panel * ROWS * p.ldT + (long long)ct * COLS + lane_off;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    base[(long long)mb * 8 * p.ldT + nb * 8 + pc0] = -acc[mb][nb][0];
                    base[(long long)mb * 8 * p.ldT + nb * 8 + pc1] = -acc[mb][nb][1];
                }
        }
        // next step's coordinates
        int npanel = panel, nct = ct + 1;
        if (nct == Q) { nct = 0; ++npanel; }
        if (TPREF && t + 1 < tend) loadT(tn, npanel, nct);

        // ---------------- GEMM #2: e += D (16 x 32, registers) . A~_J (32 x Np) ------------------------
        if (DO_E) {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
                const double* r0 = sAs + (8 * kb + pc0) * ld + 2 * t1;
                const double* r1 = sAs + (8 * kb + pc1) * ld + 2 * t1;
                // two column pairs (4 n-blocks x 2 m-blocks = 8 independent accumulators) per round: the two DMMAs
                // that hit the same accumulator (v = 0, 1) are 8 issues apart
#pragma unroll
                for (int pr = 0; pr + 1 < NB / 2; pr += 2) {
                    const double2 b00 = *reinterpret_cast<const double2*>(r0 + 16 * pr);
                    const double2 b01 = *reinterpret_cast<const double2*>(r0 + 16 * (pr + 1));
                    const double2 b10 = *reinterpret_cast<const double2*>(r1 + 16 * pr);
                    const double2 b11 = *reinterpret_cast<const double2*>(r1 + 16 * (pr + 1));
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][0], b00.x);
                        dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][0], b00.y);
                        dmma(e[mb][2 * pr + 2][0], e[mb][2 * pr + 2][1], acc[mb][kb][0], b01.x);
                        dmma(e[mb][2 * pr + 3][0], e[mb][2 * pr + 3][1], acc[mb][kb][0], b01.y);
                    }
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][1], b10.x);
                        dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][1], b10.y);
                        dmma(e[mb][2 * pr + 2][0], e[mb][2 * pr + 2][1], acc[mb][kb][1], b11.x);
                        dmma(e[mb][2 * pr + 3][0], e[mb][2 * pr + 3][1], acc[mb][kb][1], b11.y);
                    }
                }
                if ((NB / 2) & 1) {   // one column pair left (+ the odd tail block): 4 or 6 independent accumulators
                    constexpr int pr = (NB / 2) - 1;
                    const double2 b0 = *reinterpret_cast<const double2*>(r0 + 16 * pr);
                    const double2 b1 = *reinterpret_cast<const double2*>(r1 + 16 * pr);
                    double t0v = 0.0, t1v = 0.0;
                    if (NB & 1) {
                        t0v = (r0 - 2 * t1)[8 * (NB - 1) + t1];
                        t1v = (r1 - 2 * t1)[8 * (NB - 1) + t1];
                    }
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][0], b0.x);
                        dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][0], b0.y);
                        if (NB & 1) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][0], t0v);
                    }
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][1], b1.x);
                        dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][1], b1.y);
                        if (NB & 1) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][1], t1v);
                    }
                } else if (NB & 1) {  // only the odd tail block is left
                    const double t0v = (r0 - 2 * t1)[8 * (NB - 1) + t1];
                    const double t1v = (r1 - 2 * t1)[8 * (NB - 1) + t1];
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][0], t0v);
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][1], t1v);
                }
            }
        }
        // The slot is refilled by the TMA engine (async proxy) as soon as the empty barrier completes; this warp's operand loads
        // (generic proxy) have been ISSUED at this point, not necessarily performed -- nvcc schedules the arrive right behind
        // the last LDS and the DMMAs that consume it afterwards.  The proxy fence orders them before the arrive.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);

        if (DO_E && lastp) {
            // flush this visit's partial E rows: Epiece[visit][wc][row][n]
            const int visit = p.visit_start[blockIdx.x] + (pv - 1);
            double* dst = p.Epiece + ((size_t)(visit * 2 + wc) * ROWS + 16 * wr + prow) * p.Np;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    int n0, n1;
                    if (b < 2 * (NB / 2)) {
                        n0 = 16 * (b >> 1) + 2 * (2 * t0) + (b & 1);
                        n1 = n0 + 2;
                    } else {
                        n0 = 8 * (NB - 1) + 2 * t0;
                        n1 = n0 + 1;
                    }
                    dst[(size_t)mb * 8 * p.Np + n0] = e[mb][b][0];
                    dst[(size_t)mb * 8 * p.Np + n1] = e[mb][b][1];
                    e[mb][b][0] = e[mb][b][1] = 0.0;
                }
            }
        }
        panel = npanel;
        ct = nct;
    }

    // deterministic cost reduction: lanes -> warp -> CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
    if (lane == 0) sred[warp] = cost;
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS * 32) : "memory");
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NCONS; ++w) s += sred[w];
        p.cost_part[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_ty -- the hot kernel of the single-GEMM formulation:  Y_I = sum_J T~_IJ A~_J   (pieces per visit, as above).
//
// With G = A~^T A~ (= (U^T U) o (U^T U), the identity up to rounding) the contraction shared by every gradient term is
//      E = D A~ = (B~ A~^T - T~) A~ = B~ G - T~ A~ = C~ - Y,          C~ = A~ (Lambda G)   (O(R N^2), panel kernels)
// and the cost follows from N x N quantities (cost = |T~|^2 - <B~, C~> + 2 <Lambda, A~^T E>), so one evaluation needs ONE
// R x R x N product instead of two: the T~ tile is loaded straight into DMMA A fragments (same perm8 layout as the
// accumulators of k_fused) and multiplied with the A~ sub-tile streamed through the TMA ring.  Half the DMMA work of
// k_fused for the same single pass over T~.  k_fused remains the kernel for the residual update (STORE), the cost-only
// evaluation and the exact-cost fallback when the residual is tiny against the tensor (cancellation in the expansion).
// ---------------------------------------------------------------------------------------------------------------------
struct YParams {
    const double* T;        // [P*64][ldT] local rows of T~
    const double* Acol;     // A~ rows of all columns [Q*64][ld]
    double* Ypiece;         // [visits][2][64][Np]
    const int* visit_start; // [gridDim.x]
    long long ldT;
    int P, Q, ld, Np, S;
    int lookahead;
    int dephase;            // cycles by which column group 1 starts late (0 = off)
    long long* dbg;         // development aid (MAMBO_TY_DEBUG=1): clock64 stamps of CTA 0, [warp][step][4]
};

inline size_t smem_bytes_y(int ld, int S) { return (size_t)(S * SUB) * ld * 8 + (2 * S) * 8 + 64; }

// MB = m-blocks (8 rows) per warp: 2 -> 8 warps (4 row groups x 2 column groups, 2 warps per scheduler), 1 -> 16 warps
// (8 x 2, 4 warps per scheduler, <= 128 registers): more warps to cover ring waits and operand loads at every step boundary.
template <int NB, bool TPREF, int MB>
__global__ void __launch_bounds__(64 * (8 / MB), 1) k_ty(YParams p) {
    constexpr int NWR = 8 / MB;            // row groups (warps per column group)
    if (p.dbg && threadIdx.x == 0) p.dbg[16 * 64 * 4 + blockIdx.x * 4 + 0] = gtime();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ld = 8 * NB + 4;
    const int S = p.S, Q = p.Q;
    double* sA = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(sA + (size_t)S * SUB * ld);
    uint64_t* empty = full + S;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long total = (long long)p.P * Q;
    const long long tbeg = total * blockIdx.x / gridDim.x;
    const long long tend = total * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NWR);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int t0 = lane & 3, t1 = lane >> 2;
    const int wc = (warp / NWR) & 1, wr = warp % NWR;
    const int prow = perm8(t1);
    const int pc0 = t0, pc1 = 4 + ((t0 + 2) & 3);
    // Warps w and w+4 share a scheduler (one of each column group).  Starting column group 1 half a step late keeps the
    // two out of phase, so the DMMA-free stretch at every step boundary of one warp (ring wait, T~ fragment moves, first
    // operand loads) falls into the DMMA stream of the other instead of leaving the tensor pipe idle.
    if (p.dephase && wc == 1) {
        const long long t_start = clock64();
        while (clock64() - t_start < (long long)p.dephase) { }
    }

    double e[MB][NB][2];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
        for (int b = 0; b < NB; ++b) e[mb][b][0] = e[mb][b][1] = 0.0;
    double acc[MB][4][2];
    double tn[MB][4][2];

    const long long lane_off = (long long)(8 * MB * wr + prow) * p.ldT + 32 * wc;
    auto loadT = [&](double(&dst)[MB][4][2], int panel_, int ct_) {
        const double* base = p.T + (long long)panel_ * ROWS * p.ldT + (long long)ct_ * COLS + lane_off;
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                dst[mb][nb][0] = __ldcs(base + (long long)mb * 8 * p.ldT + nb * 8 + pc0);
                dst[mb][nb][1] = __ldcs(base + (long long)mb * 8 * p.ldT + nb * 8 + pc1);
            }
    };

    int panel = (int)(tbeg / Q), ct = (int)(tbeg % Q);
    int gi = 0, pv = 0;
    if (TPREF && tbeg < tend) loadT(tn, panel, ct);

    // ---- operand ring: sub-tile g = 2*step + column group lives in slot g % S.  The producer role ROTATES over the warps
    // of the column group that consumes the sub-tile: sub-tile g is requested by lane 0 of warp NWR (g & 1) + (g / 2) % NWR,
    // `lookahead` sub-tiles ahead of its use (a requester of the consumer group has itself consumed the slot's previous
    // occupant, so the parity it waits for on empty[slot] cannot be confused with an older phase).
    // Cross-proxy hazard (found in round 2): the consumers read a slot with LDS (generic proxy) and the TMA engine (async proxy)
    // refills it as soon as empty[slot] completes.  Without a proxy fence in front of the consumers' arrive the refill can
    // overtake operand loads that are issued but not yet performed: rare single-row operand errors (1e-5 relative in Y, 0.2 - 60 %
    // of launches) at every ring shape that refills a slot in the step right after its consumption (S - lookahead = 2:
    // N = 129..144 and, in k_fused, N = 121..144), never where a slot rests for a step (the benchmark sizes).  One
    // request (empty-slot wait + expect_tx + UBLKCP) costs ~650 cycles of the issuing warp; with a fixed producer warp
    // those 2 x 650 cycles per step made that warp 17 % slower than the others, and because every other warp can only
    // run as far ahead as the ring holds, the whole CTA ran at its pace (clock64 stamps, tools/ty_stamps.py).
    constexpr int NWARPS = 2 * NWR;
    const int g_total = 2 * (int)(tend - tbeg);
    int lookahead = p.lookahead > 0 ? min(p.lookahead, S - 2) : ((S >= 6) ? 4 : (S - 2));
    lookahead &= ~1;                                   // even: both sub-tiles of a step are requested in the same step
    const int ct0 = ct;
    const uint32_t sub_bytes = (uint32_t)(SUB * ld * 8);
    auto request = [&](int g) {                        // called by lane 0 of warp requester(g)
        const int slot = g % S, use = g / S, c = g & 1;
        const int ctg = (ct0 + (g >> 1)) % Q;
        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
        mbar_expect_tx(&full[slot], sub_bytes);
        bulk_g2s(sA + (size_t)slot * SUB * ld, p.Acol + ((size_t)ctg * COLS + c * SUB) * ld, sub_bytes, &full[slot]);
    };
    auto requester = [&](int g) { return NWR * (g & 1) + ((g >> 1) % NWR); };   // a warp of column group g & 1 (wc = warp / NWR)
    if (lane == 0)
        for (int g = 0; g < lookahead && g < g_total; ++g)
            if (requester(g) == warp) request(g);
    __syncwarp();

    const bool dbg = p.dbg && blockIdx.x == 0 && lane == 0;
    if (p.dbg && threadIdx.x == 0) p.dbg[16 * 64 * 4 + blockIdx.x * 4 + 1] = gtime();
    for (long long t = tbeg; t < tend; ++t, ++gi) {
        if (t == tbeg || ct == 0) ++pv;
        if (dbg && gi < 64) p.dbg[(warp * 64 + gi) * 4 + 0] = clock64();
        {
            const int g1 = 2 * gi + lookahead;
            if (lane == 0) {
                if (g1 < g_total && requester(g1) == warp) request(g1);
                if (g1 + 1 < g_total && requester(g1 + 1) == warp) request(g1 + 1);
            }
            __syncwarp();
        }
        int npanel = panel, nct = ct + 1;
        if (nct == Q) { nct = 0; ++npanel; }
        if (TPREF) {
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    acc[mb][nb][0] = tn[mb][nb][0];
                    acc[mb][nb][1] = tn[mb][nb][1];
                }
            if (t + 1 < tend) loadT(tn, npanel, nct);     // next tile's loads fly during this step's DMMAs
        } else {
            loadT(acc, panel, ct);                        // large N: no register room for the look-ahead tile
        }

        const int g = 2 * gi + wc;
        const int slot = g % S;
        if (dbg && gi < 64) p.dbg[(warp * 64 + gi) * 4 + 1] = clock64();
        mbar_wait(&full[slot], (g / S) & 1);
        if (dbg && gi < 64) p.dbg[(warp * 64 + gi) * 4 + 2] = clock64();
        const double* sAs = sA + (size_t)slot * SUB * ld;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            const double* r0 = sAs + (8 * kb + pc0) * ld + 2 * t1;
            const double* r1 = sAs + (8 * kb + pc1) * ld + 2 * t1;
#pragma unroll
            for (int pr = 0; pr + 1 < NB / 2; pr += 2) {
                const double2 b00 = *reinterpret_cast<const double2*>(r0 + 16 * pr);
                const double2 b01 = *reinterpret_cast<const double2*>(r0 + 16 * (pr + 1));
                const double2 b10 = *reinterpret_cast<const double2*>(r1 + 16 * pr);
                const double2 b11 = *reinterpret_cast<const double2*>(r1 + 16 * (pr + 1));
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][0], b00.x);
                    dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][0], b00.y);
                    dmma(e[mb][2 * pr + 2][0], e[mb][2 * pr + 2][1], acc[mb][kb][0], b01.x);
                    dmma(e[mb][2 * pr + 3][0], e[mb][2 * pr + 3][1], acc[mb][kb][0], b01.y);
                }
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][1], b10.x);
                    dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][1], b10.y);
                    dmma(e[mb][2 * pr + 2][0], e[mb][2 * pr + 2][1], acc[mb][kb][1], b11.x);
                    dmma(e[mb][2 * pr + 3][0], e[mb][2 * pr + 3][1], acc[mb][kb][1], b11.y);
                }
            }
            if ((NB / 2) & 1) {
                constexpr int pr = (NB / 2) - 1;
                const double2 b0 = *reinterpret_cast<const double2*>(r0 + 16 * pr);
                const double2 b1 = *reinterpret_cast<const double2*>(r1 + 16 * pr);
                double t0v = 0.0, t1v = 0.0;
                if (NB & 1) {
                    t0v = (r0 - 2 * t1)[8 * (NB - 1) + t1];
                    t1v = (r1 - 2 * t1)[8 * (NB - 1) + t1];
                }
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][0], b0.x);
                    dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][0], b0.y);
                    if (NB & 1) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][0], t0v);
                }
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    dmma(e[mb][2 * pr][0], e[mb][2 * pr][1], acc[mb][kb][1], b1.x);
                    dmma(e[mb][2 * pr + 1][0], e[mb][2 * pr + 1][1], acc[mb][kb][1], b1.y);
                    if (NB & 1) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][1], t1v);
                }
            } else if (NB & 1) {
                const double t0v = (r0 - 2 * t1)[8 * (NB - 1) + t1];
                const double t1v = (r1 - 2 * t1)[8 * (NB - 1) + t1];
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][0], t0v);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) dmma(e[mb][NB - 1][0], e[mb][NB - 1][1], acc[mb][kb][1], t1v);
            }
        }
        // The slot is refilled by the TMA engine (async proxy) as soon as the empty barrier completes; this warp's operand loads
        // (generic proxy) have been ISSUED at this point, not necessarily performed -- nvcc schedules the arrive right behind
        // the last LDS and the DMMAs that consume it afterwards.  The proxy fence orders them before the arrive.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (dbg && gi < 64) p.dbg[(warp * 64 + gi) * 4 + 3] = clock64();

        const bool lastp = (ct == Q - 1) || (t == tend - 1);
        if (lastp) {
            const int visit = p.visit_start[blockIdx.x] + (pv - 1);
            double* dst = p.Ypiece + ((size_t)(visit * 2 + wc) * ROWS + 8 * MB * wr + prow) * p.Np;
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    int n0, n1;
                    if (b < 2 * (NB / 2)) {
                        n0 = 16 * (b >> 1) + 2 * (2 * t0) + (b & 1);
                        n1 = n0 + 2;
                    } else {
                        n0 = 8 * (NB - 1) + 2 * t0;
                        n1 = n0 + 1;
                    }
                    dst[(size_t)mb * 8 * p.Np + n0] = e[mb][b][0];
                    dst[(size_t)mb * 8 * p.Np + n1] = e[mb][b][1];
                    e[mb][b][0] = e[mb][b][1] = 0.0;
                }
            }
        }
        panel = npanel;
        ct = nct;
    }
    if (p.dbg && threadIdx.x == 0) {
        p.dbg[16 * 64 * 4 + blockIdx.x * 4 + 2] = gtime();
        p.dbg[16 * 64 * 4 + blockIdx.x * 4 + 3] = (long long)(tend - tbeg);
    }
}

}  // namespace fused
