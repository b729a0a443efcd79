// mambo_csa.cu -- host side of libmambo_b200.so: handle, buffers, launch sequence, C-ABI (include/mambo_csa.h).
// Replaces the closures of /root/reference/src/UTILS/decompose.jl:96-105, 221-231 and the residual update
// decompose.jl:56,182.  There is deliberately no CPU code path for any of the numerics in this file.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include <cuda.h>            // types of the driver entry point fetched at run time (no link dependency on libcuda)
#include <cuda_runtime.h>

#include "../../include/mambo_csa.h"
#include "internal.h"
#include "fused_kernel.cuh"
#include "small_kernels.cuh"
#include "chain_kernels.cuh"
#include "panel_kernels.cuh"
#include "lanczos_kernels.cuh"
#include "tiny_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? MAMBO_ENOMEM : MAMBO_ECUDA,                        \
                        std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define TRY(expr)               \
    do {                        \
        int rc__ = (expr);      \
        if (rc__) return rc__;  \
    } while (0)

// every multiple of 8 columns up to 168 has its own instantiation: the DMMA work is never padded by more than one n-block
const int kNBBuckets[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21};

// Opt-in dynamic shared memory of a kernel is a per-function (per-process) attribute: handles of different N share the
// non-templated kernels, so the limit is only ever RAISED -- a later, smaller problem must not lower it under an
// earlier handle that is still in use.
template <class K>
cudaError_t raise_smem(K kernel, int bytes) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, (const void*)kernel);
    if (e != cudaSuccess) return e;
    if (a.maxDynamicSharedSizeBytes >= bytes) return cudaSuccess;
    return cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

struct GraphSeg {
    cudaGraphExec_t exec = nullptr;
    int kernels = 0;
};

// Residual state shared by a handle and the evaluation lanes created from it (mambo_csa_create_lane): the tensor,
// its row tables and the one-body residual exist once per GPU; every lane has its own operands and scratch.
struct SharedResidual {
    std::atomic<int> refs{1};             // the last handle out frees T1/T2/hT/pa/pq/sw (every holder has the same pointers)
    std::atomic<long long> version{0};   // bumped by every residual update: invalidates the memo of every lane
    double* tnorm2 = nullptr;            // device scalar |T~|^2 of the local rows (two-body part), kept current by the updates
};

}  // namespace

struct mambo_csa {
    int N = 0, flavour = 0, mode = 0, device = 0;
    int L = 0, cL = 0, uL = 0, lam_off = 0;
    int nR = 0, nRpad = 0;          // logical / padded rows of the residual matrix
    int NB = 0, Np = 0, ld = 0;     // n-blocks, padded N, operand pitch
    long long ldT = 0;
    int shard = 0, nshards = 1;
    SharedResidual* sh = nullptr;       // owner of T1/T2/hT/pa/pq/sw (shared with the lanes of this handle)
    bool is_lane = false;
    long long memo_version = -1;
    int prow_begin = 0, prow_end = 0;   // local panels [begin, end)
    int row_begin = 0, row_end = 0;     // local padded rows
    int P = 0, Q = 0, G = 0, S = 0, visits = 0;
    int num_sms = 0;
    size_t smem = 0;
    bool packed = false;
    double asym_pair = 0, asym_pq = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // device buffers
    double *T1 = nullptr, *T2 = nullptr;       // residual (and its transpose in GENERAL mode)
    double *hT = nullptr;                      // one-body residual (column-major), CSA_SD
    double *At = nullptr, *Bt = nullptr;
    int *pa = nullptr, *pq = nullptr;
    double* sw = nullptr;
    double *Epiece = nullptr, *Et = nullptr, *F1 = nullptr, *F2 = nullptr, *Spart = nullptr, *cost_part = nullptr;
    int *visit_start = nullptr, *piece_first = nullptr;
    double* mats = nullptr;                    // Lam, Dh, DU, hfrag (dense n x n each)
    // persistent-chain kernels: padded-pitch matrices (chain::mat_stride each) and their transposed copies
    double* cmats = nullptr;
    chain::UParams cu;
    chain::DParams cd;
    double* U = nullptr;                       // dense n x n copy of exp(K) written by k_chain_u
    int* d_zero = nullptr;
    double *Lam = nullptr, *Dh = nullptr, *DU = nullptr, *hfrag = nullptr, *g1 = nullptr;
    int* d_s = nullptr;
    unsigned* d_bar = nullptr;                 // grid-barrier counters of the persistent chain kernels
    long long* d_dbg = nullptr;                // MAMBO_CHAIN_DEBUG=1: clock64 stamps of the derivative chain
    int chain_ct = 16;                         // tile edge of the chain kernels
    bool chain_cluster = false;                // N <= 64: the <= 4 x 4 tiles form ONE thread-block cluster (hardware barrier)
    int dephase = 0;                           // fused kernel: start offset (cycles) of column group 1
    int lookahead = 0;                         // fused kernel: operand sub-tiles requested ahead (0 = default)
    // single-GEMM formulation (fused::k_ty): E = C~ - T~ A~, cost from N x N quantities, exact-cost fallback behind a flag
    bool fast = false;
    // lean single-GEMM path (unsharded handles): no B~/C~ panel GEMMs and no Gram kernel -- the N x N closed forms of
    // panel::k_build_a replace them; B~ is only built for the two-GEMM kernel (residual update, cost-only, exact-cost pass)
    bool lean = false;
    // molecule-sized problems (N <= MAMBO_TINY_MAX, default 12; packed rows; unsharded): the whole evaluation is ONE single-CTA
    // kernel (tiny::k_eval) instead of the 11-launch pipeline, whose every launch is latency-bound at that size
    bool tiny = false;
    double* aux = nullptr;                     // [o | d | t | wn] of k_build_a (3 N + 1)
    double* obc = nullptr;                     // one-body cost of the last evaluation (CSA_SD), 0 for CSA
    cudaStream_t side_stream = nullptr;        // the flagged exact-cost pass runs beside the derivative chain
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // how the host closures wait for the device: spinning (lowest latency, one busy core per caller) or sleeping on an event
    // created with cudaEventBlockingSync (callers that outnumber the host cores: many lanes / many ranks per box)
    bool blocking_wait = false;
    cudaEvent_t ev_done = nullptr;
    double* Ct = nullptr;                      // C~ = A~ (Lambda G) rows of the local panels [P*64][ld]
    double* GLP = nullptr;                     // G Lambda with the panel pitch [Np][ld] (k_gram)
    double* wpart = nullptr;                   // per (panel, column split) <B~, C~>
    int* d_flag = nullptr;                     // 1: the expansion cancelled, take the exact cost
    double* ls_part = nullptr;                 // per output row <Lambda_a, S_a> (k_post_reduce)
    unsigned* d_done = nullptr;                // last-CTA-out counter of k_post_reduce
    double* tnorm2 = nullptr;                  // == sh->tnorm2
    int prep_cs = 1;                           // column split of k_prep_mma (NB_PANEL_DISPATCH)
    int Sy = 0;                                // ring depth of k_ty
    long long* ty_dbg = nullptr;               // MAMBO_TY_DEBUG=1: clock64 stamps of CTA 0 of k_ty
    int ty_mb = 2;                             // k_ty: m-blocks per warp (2: 8 warps, 1: 16 warps)
    int ty_dephase = 0;                        // start offset (cycles) of column group 1 in k_ty
    size_t smem_y = 0;
    double tau = 1e-4;                         // fallback threshold: relative error of the expanded cost ~ 1e-15 / tau
    double* LamP = nullptr;                    // Lambda with the panel pitch [Np][ld] (written by k_chain_u)
    double* EtP = nullptr;                     // E~ rows with the panel pitch [P*64][ld] (k_esum)
    double* d_scale = nullptr;
    int* d_status = nullptr;
    double *d_x = nullptr, *d_out = nullptr, *d_part = nullptr;
    double *h_x = nullptr, *h_out = nullptr;   // pinned staging
    int* h_status = nullptr;
    // memo of the last evaluation through the host API
    std::vector<double> memo_x;
    bool memo_valid = false;
    long long launches = 0;
    int launches_last_eval = 0;
    // peer exchange of row-panel shards (one process per GPU, CUDA IPC over NVLink)
    double* peer_recv = nullptr;               // [2][world][plen_pad] receive slots of THIS rank
    unsigned long long* peer_flags = nullptr;  // [2][MAXPEERS] epochs published by the peers
    unsigned* peer_done = nullptr;
    small::PeerTable peer_tab;
    void* peer_opened[small::MAXPEERS] = {nullptr};   // IPC mappings to close
    int peer_attached = 0;
    int plen_pad = 0;
    unsigned long long peer_epoch = 0;
    bool peer_stream_wait = false;             // mambo_csa_set_peer_wait: the stream, not a spinning kernel, waits for the peers' flags
    // Lanczos workspace of mambo_csa_leading_eig (allocated on first use, kept: cudaMalloc/cudaFree of tens of MB cost more
    // than the iteration itself)
    double* eig_ws = nullptr;
    size_t eig_ws_doubles = 0;
    // scratch of the built-in optimiser (vectors + dense inverse Hessian or L-BFGS ring), kept between runs for the same reason
    void* opt_ws = nullptr;
    size_t opt_ws_bytes = 0;
    void* opt_hs = nullptr;                    // pinned host scalars of the optimiser (128 bytes)
    // double factorisation of the resident residual (df.cu): opaque state + its destructor
    void* df_state = nullptr;
    void (*df_destroy)(void*) = nullptr;
    // CUDA-graph segments of the small-kernel sequences
    bool use_graphs = true;
    bool capturing = false;               // a graph segment is being captured: nested segments run inline
    std::vector<GraphSeg> g_pre, g_fin;   // index 0: device-decided s; 1+s: s squarings
    std::vector<GraphSeg> g_pre_b;        // lean handles: chain + full operand build (A~ and B~) for the two-GEMM kernel
    std::vector<GraphSeg> g_fin_seq;      // finish half without the side branch (two-call evaluations: partial / finish)
    GraphSeg g_post_seq;
    std::vector<GraphSeg> g_host;         // the WHOLE host-API evaluation (H2D of x ... D2H of cost+grad) as one graph per s
    GraphSeg g_mid, g_post, g_costonly;
    // optional live profiling of the fused kernel (CUDA events on the launching stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;   // pairs: [2k] before, [2k+1] after
    size_t ev_used = 0;
    std::vector<cudaEvent_t> seg_pool;  // 4 marks per evaluation: start, after pre, after fused+post, after finish
    size_t seg_used = 0;
};

namespace {

using namespace fused;

int seg_mark(mambo_csa* h) {
    if (!h->profiling) return MAMBO_OK;
    if (h->seg_used >= h->seg_pool.size()) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        h->seg_pool.push_back(e);
    }
    CU(cudaEventRecord(h->seg_pool[h->seg_used++], h->stream));
    return MAMBO_OK;
}

template <int NB, bool DO_E, bool STORE>
int launch_fused_nb(mambo_csa* h, const Params& p) {
    constexpr bool TPREF = (NB <= 13);
    auto kern = k_fused<NB, DO_E, STORE, TPREF>;
    if (p.P < 0) {   // configuration call from create
        CU(raise_smem(kern, (int)h->smem));
        return MAMBO_OK;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profiling && !p.run_if) {   // the flagged exact-cost pass normally returns at once: not a timed hot launch
        if (h->ev_used + 2 > h->ev_pool.size()) {
            cudaEvent_t a, b;
            CU(cudaEventCreate(&a));
            CU(cudaEventCreate(&b));
            h->ev_pool.push_back(a);
            h->ev_pool.push_back(b);
        }
        e0 = h->ev_pool[h->ev_used++];
        e1 = h->ev_pool[h->ev_used++];
        CU(cudaEventRecord(e0, h->stream));
    }
    kern<<<h->G, NTHREADS, h->smem, h->stream>>>(p);
    if (e1) CU(cudaEventRecord(e1, h->stream));
    CU(cudaGetLastError());
    h->launches++;
    return MAMBO_OK;
}

template <bool DO_E, bool STORE>
int launch_fused(mambo_csa* h, const Params& p) {
    switch (h->NB) {
        case 1: return launch_fused_nb<1, DO_E, STORE>(h, p);
        case 2: return launch_fused_nb<2, DO_E, STORE>(h, p);
        case 3: return launch_fused_nb<3, DO_E, STORE>(h, p);
        case 4: return launch_fused_nb<4, DO_E, STORE>(h, p);
        case 5: return launch_fused_nb<5, DO_E, STORE>(h, p);
        case 6: return launch_fused_nb<6, DO_E, STORE>(h, p);
        case 7: return launch_fused_nb<7, DO_E, STORE>(h, p);
        case 8: return launch_fused_nb<8, DO_E, STORE>(h, p);
        case 9: return launch_fused_nb<9, DO_E, STORE>(h, p);
        case 10: return launch_fused_nb<10, DO_E, STORE>(h, p);
        case 11: return launch_fused_nb<11, DO_E, STORE>(h, p);
        case 12: return launch_fused_nb<12, DO_E, STORE>(h, p);
        case 13: return launch_fused_nb<13, DO_E, STORE>(h, p);
        case 14: return launch_fused_nb<14, DO_E, STORE>(h, p);
        case 15: return launch_fused_nb<15, DO_E, STORE>(h, p);
        case 16: return launch_fused_nb<16, DO_E, STORE>(h, p);
        case 17: return launch_fused_nb<17, DO_E, STORE>(h, p);
        case 18: return launch_fused_nb<18, DO_E, STORE>(h, p);
        case 19: return launch_fused_nb<19, DO_E, STORE>(h, p);
        case 20: return launch_fused_nb<20, DO_E, STORE>(h, p);
        case 21: return launch_fused_nb<21, DO_E, STORE>(h, p);
    }
    return fail(MAMBO_EINVAL, "unsupported NB bucket");
}

template <int NB, int MB>
int launch_ty_mb(mambo_csa* h, const YParams& p) {
    constexpr bool TPREF = (NB <= 16);
    auto kern = k_ty<NB, TPREF, MB>;
    if (p.P < 0) {   // configuration call from create
        CU(raise_smem(kern, (int)h->smem_y));
        return MAMBO_OK;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profiling) {
        if (h->ev_used + 2 > h->ev_pool.size()) {
            cudaEvent_t a, b;
            CU(cudaEventCreate(&a));
            CU(cudaEventCreate(&b));
            h->ev_pool.push_back(a);
            h->ev_pool.push_back(b);
        }
        e0 = h->ev_pool[h->ev_used++];
        e1 = h->ev_pool[h->ev_used++];
        CU(cudaEventRecord(e0, h->stream));
    }
    kern<<<h->G, 64 * (8 / MB), h->smem_y, h->stream>>>(p);
    if (e1) CU(cudaEventRecord(e1, h->stream));
    CU(cudaGetLastError());
    h->launches++;
    return MAMBO_OK;
}

template <int NB>
int launch_ty_nb(mambo_csa* h, const YParams& p) {
    if (p.P < 0) {
        TRY((launch_ty_mb<NB, 1>(h, p)));
        return launch_ty_mb<NB, 2>(h, p);
    }
    return h->ty_mb == 1 ? launch_ty_mb<NB, 1>(h, p) : launch_ty_mb<NB, 2>(h, p);
}

int launch_ty(mambo_csa* h, const YParams& p) {
    switch (h->NB) {
        case 1: return launch_ty_nb<1>(h, p);
        case 2: return launch_ty_nb<2>(h, p);
        case 3: return launch_ty_nb<3>(h, p);
        case 4: return launch_ty_nb<4>(h, p);
        case 5: return launch_ty_nb<5>(h, p);
        case 6: return launch_ty_nb<6>(h, p);
        case 7: return launch_ty_nb<7>(h, p);
        case 8: return launch_ty_nb<8>(h, p);
        case 9: return launch_ty_nb<9>(h, p);
        case 10: return launch_ty_nb<10>(h, p);
        case 11: return launch_ty_nb<11>(h, p);
        case 12: return launch_ty_nb<12>(h, p);
        case 13: return launch_ty_nb<13>(h, p);
        case 14: return launch_ty_nb<14>(h, p);
        case 15: return launch_ty_nb<15>(h, p);
        case 16: return launch_ty_nb<16>(h, p);
        case 17: return launch_ty_nb<17>(h, p);
        case 18: return launch_ty_nb<18>(h, p);
        case 19: return launch_ty_nb<19>(h, p);
        case 20: return launch_ty_nb<20>(h, p);
        case 21: return launch_ty_nb<21>(h, p);
    }
    return fail(MAMBO_EINVAL, "unsupported NB bucket");
}

YParams make_yparams(mambo_csa* h, const double* T) {
    YParams p;
    p.T = T;
    p.Acol = h->At;
    p.Ypiece = h->Epiece;
    p.visit_start = h->visit_start;
    p.ldT = h->ldT;
    p.P = h->P;
    p.Q = h->Q;
    p.ld = h->ld;
    p.Np = h->Np;
    p.S = h->Sy;
    p.lookahead = h->lookahead;
    p.dephase = h->ty_dephase;
    p.dbg = h->ty_dbg;
    return p;
}

Params make_params(mambo_csa* h, double* T) {
    Params p;
    p.T = T;
    p.Tst = T;
    p.Brow = h->Bt;
    p.Acol = h->At;
    p.Epiece = h->Epiece;
    p.cost_part = h->cost_part;
    p.visit_start = h->visit_start;
    p.ldT = h->ldT;
    p.P = h->P;
    p.Q = h->Q;
    p.ld = h->ld;
    p.Np = h->Np;
    p.S = h->S;
    p.run_if = nullptr;
    p.dephase = h->dephase;
    p.lookahead = h->lookahead;
    return p;
}

int host_s_for(const mambo_csa* h, const double* x) {
    // ||K||_1 on the host, same definition as small::k_build; returned s is passed down so host and device agree.
    const int n = h->N;
    const double* th = x + h->lam_off + h->cL;
    double mx = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            if (i == j) continue;
            const int a = std::min(i, j), b = std::max(i, j);
            s += std::fabs(th[(long long)a * n - (long long)a * (a + 1) / 2 + (b - a - 1)]);
        }
        if (!(s == s) || std::isinf(s)) return -2;   // NaN / Inf in theta
        mx = std::max(mx, s);
    }
    int s = 0;
    if (mx > 1.0) s = (int)std::ceil(std::log2(mx));
    return s;
}

// U = exp(K): Taylor-16 (Paterson-Stockmeyer, 6 GEMMs) + s squarings; every intermediate is kept for the
// derivative chain.  s_host >= 0: number of squarings (known when x came from the host; passed down so that
// host and device agree); s_host < 0: the device decides and SMAX guarded launches are enqueued.
int launch_coop(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelExC(&cfg, func, args));
    return MAMBO_OK;
}

// One thread-block cluster holding the whole grid (the chain kernels for N <= 128): co-scheduling is guaranteed by the
// cluster launch itself, the steps synchronise on the hardware cluster barrier.
int launch_cluster(const void* func, dim3 grid, dim3 block, size_t smem, cudaStream_t st, void** args) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = grid.x;
    attr[0].val.clusterDim.y = grid.y;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelExC(&cfg, func, args));
    return MAMBO_OK;
}

int launch_chain(mambo_csa* h, bool deriv, void** args) {
    const int n = h->N, ct = h->chain_ct, tiles = (n + ct - 1) / ct;
    const size_t smem = chain::smem_bytes(n, ct);
    if (h->chain_cluster) {
        const void* f = deriv ? (const void*)chain::k_chain_d<16, chain::SYNC_CLUSTER> : (const void*)chain::k_chain_u<16, chain::SYNC_CLUSTER>;
        return launch_cluster(f, dim3(tiles, tiles), dim3(chain::NTHR), smem, h->stream, args);
    }
    const void* f = deriv ? (const void*)chain::k_chain_d<16, chain::SYNC_GRID> : (const void*)chain::k_chain_u<16, chain::SYNC_GRID>;
    return launch_coop(f, dim3(tiles, tiles), dim3(chain::NTHR), smem, h->stream, args);
}

// U = exp(K) in ONE persistent kernel (chain::k_chain_u): grid barriers instead of kernel boundaries.
int enqueue_unitary(mambo_csa* h, const double* d_x, int s_host) {
    const int n = h->N;
    chain::UParams p = h->cu;
    p.x = d_x; p.n = n; p.lam_off = h->lam_off; p.s_override = s_host;
    p.Lam = h->Lam; p.U = h->U; p.LamP = h->LamP; p.ldl = h->ld; p.s_out = h->d_s; p.scale_out = h->d_scale; p.status = h->d_status; p.bar = h->d_bar;
    void* args[] = {&p};
    TRY(launch_chain(h, false, args));
    h->launches++;
    return MAMBO_OK;
}

// NB bucket -> (NBT, CST, CSP): panel-kernel instantiation, column split of k_post_mma (CST = 4 above N = 112, 6 above N = 160,
// so that the panels plus the Lambda slice fit in shared memory) and of k_prep_mma (CSP; a split of 2 for 72 < N <= 104 -- two
// 100 KB CTAs per SM, twice as many CTAs as panels -- was measured 2 us slower than 1: the kernel is not GEMM-bound)
#define NB_PANEL_DISPATCH(nb, CALL)                 \
    switch (nb) {                                   \
        case 1: { constexpr int NBT = 1, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 2: { constexpr int NBT = 2, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 3: { constexpr int NBT = 3, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 4: { constexpr int NBT = 4, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 5: { constexpr int NBT = 5, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 6: { constexpr int NBT = 6, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 7: { constexpr int NBT = 7, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 8: { constexpr int NBT = 8, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 9: { constexpr int NBT = 9, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 10: { constexpr int NBT = 10, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 11: { constexpr int NBT = 11, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 12: { constexpr int NBT = 12, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 13: { constexpr int NBT = 13, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 14: { constexpr int NBT = 14, CST = 1, CSP = 1; (void)CSP; (void)CST; CALL; break; }   \
        case 15: { constexpr int NBT = 15, CST = 4, CSP = 4; (void)CSP; (void)CST; CALL; break; }   \
        case 16: { constexpr int NBT = 16, CST = 4, CSP = 4; (void)CSP; (void)CST; CALL; break; }   \
        case 17: { constexpr int NBT = 17, CST = 4, CSP = 4; (void)CSP; (void)CST; CALL; break; }   \
        case 18: { constexpr int NBT = 18, CST = 4, CSP = 4; (void)CSP; (void)CST; CALL; break; }   \
        case 19: { constexpr int NBT = 19, CST = 4, CSP = 4; (void)CSP; (void)CST; CALL; break; }   \
        case 20: { constexpr int NBT = 20, CST = 4, CSP = 4; (void)CSP; (void)CST; CALL; break; }   \
        default: { constexpr int NBT = 21, CST = 6, CSP = 6; (void)CSP; (void)CST; CALL; break; } \
    }
int panel_cs(int nb) { return nb >= 21 ? 6 : (nb >= 15 ? 4 : 1); }   // == CST / CSP of NB_PANEL_DISPATCH

// Operand build after the orbital rotation.
//   PREP_LEAN : A~ (all rows) and the N x N closed forms only (panel::k_build_a) -- the lean single-GEMM evaluation
//   PREP_AB   : A~ and B~ = A~ Lambda (panel::k_prep_mma) -- the two-GEMM kernel (residual update, cost-only, exact-cost pass);
//               run_if makes it conditional on the device flag
//   PREP_ABC  : Gram kernel + A~, B~, C~ = A~ (Lambda G) and <B~, C~> -- the first single-GEMM formulation (row-panel shards)
enum PrepKind { PREP_LEAN, PREP_AB, PREP_ABC };
int enqueue_prep(mambo_csa* h, PrepKind kind, const int* run_if = nullptr) {
    if (kind == PREP_LEAN) {
        const int blocks = std::min(h->nRpad / 32, h->num_sms - 1) + 1;
        panel::k_build_a<<<blocks, 1024, 0, h->stream>>>(h->U, h->Lam, h->pa, h->pq, h->sw, h->N, h->ld, h->nRpad, h->At, h->aux);
        h->launches++;
        CU(cudaGetLastError());
        return MAMBO_OK;
    }
    const bool with_c = (kind == PREP_ABC);
    if (with_c) {
        panel::k_gram<<<h->N, 1024, 0, h->stream>>>(h->U, h->Lam, h->N, h->ld, h->GLP);
        h->launches++;
    }
    NB_PANEL_DISPATCH(h->NB, (panel::k_prep_mma<NBT, CSP><<<dim3(h->nRpad / panel::ROWS, CSP), panel::NTHR, panel::prep_smem<NBT, CSP>(), h->stream>>>(
                                 h->U, h->LamP, h->pa, h->pq, h->sw, h->N, h->prow_begin, h->prow_end, h->At, h->Bt,
                                 with_c ? h->GLP : nullptr, h->Ct, h->wpart, run_if)));
    h->launches++;
    CU(cudaGetLastError());
    return MAMBO_OK;
}
PrepKind eval_prep_kind(const mambo_csa* h) { return h->lean ? PREP_LEAN : (h->fast ? PREP_ABC : PREP_AB); }

// Row epilogue of one pass of the hot kernel: F = E~ Lambda and Spart = A~^T E~ per panel, where E~ is assembled from the
// hot kernel's pieces: E (two-GEMM kernel), C~ - Y (first single-GEMM formulation) or Y itself (lean path: F = Y Lambda,
// Spart = A~^T Y; the closed-form terms are added by k_post_reduce).
//   do_F / do_S select the products; esum: column-split launches (NB >= 15) first sum the pieces once (k_esum) -- callers that
//   run the two products on two streams do that before they fork.
int enqueue_post_f(mambo_csa* h, double* F, bool do_S = false, bool do_F = true, bool esum = true) {
    const bool split = panel_cs(h->NB) > 1;   // column-split panel kernels (NB_PANEL_DISPATCH): sum the pieces once, up front
    const double* ct = (h->fast && !h->lean) ? h->Ct : nullptr;
    if (split && esum) {
        const int ysplit = std::max(1, std::min(8, (2 * h->num_sms) / std::max(h->P, 1)));
        panel::k_esum<<<dim3(h->P, ysplit), 256, 0, h->stream>>>(h->Epiece, h->piece_first, h->Np, h->ld, h->EtP, ct);
        h->launches++;
    }
    if (!do_F && !do_S) return MAMBO_OK;
    NB_PANEL_DISPATCH(h->NB, (panel::k_post_mma<NBT, CST><<<dim3(h->P, CST), panel::NTHR, panel::post_smem_part<NBT, CST>(do_F, do_S), h->stream>>>(
                                 h->Epiece, h->piece_first, ct, split ? h->EtP : nullptr, h->LamP, h->At, h->N,
                                 h->row_begin, F, h->Spart, do_S ? 1 : 0, do_F ? 1 : 0)));
    h->launches += 1;
    CU(cudaGetLastError());
    return MAMBO_OK;
}

// The flagged exact-cost pass of the single-GEMM formulations: returns at once unless k_post_reduce raised d_flag.
//   out == nullptr: the exact two-body cost replaces part[0] (sequential tail: shards, two-call evaluations)
//   out != nullptr: k_cost_select writes the final cost of the evaluation to out[0] (lean tail, beside the derivative chain)
int enqueue_exact_pass(mambo_csa* h, double* out) {
    if (h->lean) TRY(enqueue_prep(h, PREP_AB, h->d_flag));     // the lean evaluation has not built B~
    Params px = make_params(h, h->T1);
    px.run_if = h->d_flag;
    TRY((launch_fused<false, false>(h, px)));
    small::k_cost_select<<<1, 32, 0, h->stream>>>(h->d_flag, h->cost_part, h->G, h->d_part, out, h->obc);
    h->launches++;
    CU(cudaGetLastError());
    return MAMBO_OK;
}

// N x N reduction of the row epilogues into part = [cost, S, G]; seq_tail: the exact-cost pass follows at once (part[0] is
// final when this returns); otherwise it is left to enqueue_finish, which runs it beside the derivative chain.
bool split_tail(const mambo_csa* h) { return h->lean && h->flavour == MAMBO_CSA; }
int enqueue_post_sr(mambo_csa* h, bool seq_tail) {
    const int n = h->N;
    const bool general = (h->mode == (int)MAMBO_SYM_GENERAL);
    small::k_post_reduce<<<n, 1024, 0, h->stream>>>(h->cost_part, h->fast ? 0 : h->G, h->Spart, h->P, h->F1, general ? h->F2 : nullptr,
                                                   h->sw, h->U, h->d_zero, n, h->Np, h->packed ? 1 : 0, general ? 1.0 : 2.0,
                                                   h->row_begin, std::min(h->row_end, h->nR), h->d_part,
                                                   h->fast ? h->Lam : nullptr, h->ls_part, h->d_done, h->wpart,
                                                   h->P * h->prep_cs /* partial <B~,C~> per (panel, column split of k_prep_mma) */,
                                                   h->tnorm2, h->tau, h->d_flag, h->lean ? h->aux : nullptr);
    h->launches += 1;
    CU(cudaGetLastError());
    if (h->fast && seq_tail) TRY(enqueue_exact_pass(h, nullptr));
    return MAMBO_OK;
}

// stage B: one-body terms, adjoint Frechet derivative, gradient assembly.  part lives in h->d_part.
//   side_tail (single-GEMM formulations, one-call evaluations): the flagged exact-cost pass and the cost hand-over run on the
//   handle's side stream beside the derivative chain instead of in front of it.
int enqueue_finish(mambo_csa* h, const double* d_x, double* d_out, bool side_tail) {
    const int n = h->N;
    if (h->flavour == MAMBO_CSA_SD) {
        small::k_sd_onebody<<<1, 1024, 0, h->stream>>>(d_x, h->hT, h->U, h->d_zero, n, h->d_part, h->g1, h->Dh, h->DU, nullptr, h->obc);
        h->launches++;
    }
    side_tail = side_tail && h->fast;
    const bool split = side_tail && split_tail(h);
    if (side_tail) {
        const bool general = (h->mode == (int)MAMBO_SYM_GENERAL);
        double* Flast = general ? h->F2 : h->F1;
        if (split && panel_cs(h->NB) > 1) TRY(enqueue_post_f(h, Flast, false, false, true));   // k_esum only: before the fork
        CU(cudaEventRecord(h->ev_fork, h->stream));
        CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        cudaStream_t main_stream = h->stream;
        h->stream = h->side_stream;
        int rc = MAMBO_OK;
        if (split) {
            // lean CSA evaluation: of the row epilogue only F = Y Lambda -> G feeds the derivative chain.  Z = A~^T Y, its
            // reduction to S, the lambda gradient, the cost and the flagged exact-cost pass all run beside the chain.
            rc = enqueue_post_f(h, Flast, true, false, false);
            small::k_reduce_s<<<n, 1024, 0, h->stream>>>(h->Spart, h->P, h->Lam, h->aux, n, h->lam_off, h->d_part, d_out, h->ls_part, h->d_done,
                                                        h->tnorm2, h->tau, h->d_flag);
            h->launches++;
        }
        if (!rc) rc = enqueue_exact_pass(h, d_out);
        h->stream = main_stream;
        TRY(rc);
        CU(cudaEventRecord(h->ev_join, h->side_stream));
        if (split) {
            TRY(enqueue_post_f(h, Flast, false, true, false));
            small::k_gather_g<<<n, 1024, 0, h->stream>>>(h->F1, general ? h->F2 : nullptr, h->sw, h->U, n, h->Np, h->packed ? 1 : 0,
                                                        general ? 1.0 : 2.0, h->row_begin, std::min(h->row_end, h->nR), h->aux, h->d_part);
            h->launches++;
            CU(cudaGetLastError());
        }
    }
    chain::DParams p = h->cd;
    p.part = h->d_part; p.g1 = h->g1; p.n = n; p.sd = h->flavour == MAMBO_CSA_SD ? 1 : 0;
    p.s_in = h->d_s; p.scale_in = h->d_scale; p.out = d_out; p.bar = h->d_bar; p.status = h->d_status; p.dbg = h->d_dbg;
    p.write_cost = side_tail ? 0 : 1;
    p.write_lambda = split ? 0 : 1;
    void* args[] = {&p};
    TRY(launch_chain(h, true, args));
    h->launches++;
    if (side_tail) CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    return MAMBO_OK;
}

// ---- CUDA-graph segments ------------------------------------------------------------------------------------
// The ~45 small launches of one evaluation are captured once into graph segments (keyed by the squaring count
// where it matters) and replayed; the fused kernel itself is launched directly between them so that it can
// be bracketed by profiling events.  All pointers inside a segment are handle-internal and therefore fixed.
template <class F>
int run_segment(mambo_csa* h, GraphSeg& seg, F&& body) {
    if (!h->use_graphs || h->capturing) return body();      // inside an enclosing capture: the kernels join that graph
    if (!seg.exec) {
        h->capturing = true;
        cudaStream_t saved = h->stream;
        h->stream = h->own_stream;
        const long long l0 = h->launches;
        cudaError_t e = cudaStreamBeginCapture(h->own_stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            h->stream = saved;
            h->capturing = false;
            CU(e);
        }
        int rc = body();
        cudaGraph_t graph = nullptr;
        e = cudaStreamEndCapture(h->own_stream, &graph);
        h->stream = saved;
        h->capturing = false;
        seg.kernels = (int)(h->launches - l0);
        h->launches = l0;
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        CU(e);
        e = cudaGraphInstantiate(&seg.exec, graph, 0);
        cudaGraphDestroy(graph);
        CU(e);
    }
    CU(cudaGraphLaunch(seg.exec, h->stream));
    h->launches += seg.kernels;
    return MAMBO_OK;
}

GraphSeg& seg_for(std::vector<GraphSeg>& v, int s_host) {  // index 0: device-decided s, 1 + s otherwise
    const size_t idx = s_host < 0 ? 0 : (size_t)(1 + std::min(s_host, small::SMAX));
    if (v.size() <= idx) v.resize(idx + 1);
    return v[idx];
}

// stage A: everything up to the per-shard partial vector [cost, S, G] (x must already be in h->d_x).
//   one_call: the finish half follows in the same call (eval_device / host closures), so the flagged exact-cost pass may be
//   left to it (it then runs beside the derivative chain); otherwise part[0] is final when this segment ends.
int enqueue_tiny(mambo_csa* h, bool need_grad) {
    tiny::Params p;
    p.x = h->d_x; p.T = h->T1; p.ldT = h->ldT; p.hT = h->flavour == MAMBO_CSA_SD ? h->hT : nullptr;
    p.pa = h->pa; p.pq = h->pq; p.sw = h->sw;
    p.n = h->N; p.M = h->nR; p.sd = h->flavour == MAMBO_CSA_SD ? 1 : 0; p.lam_off = h->lam_off; p.cL = h->cL;
    p.need_grad = need_grad ? 1 : 0;
    p.out = h->d_out; p.cost_out = h->d_part; p.obc = h->obc; p.status = h->d_status;
    tiny::k_eval<<<1, tiny::NT, tiny::smem_bytes(h->N, h->nR), h->stream>>>(p);
    h->launches++;
    CU(cudaGetLastError());
    return MAMBO_OK;
}

int enqueue_partial(mambo_csa* h, int s_host, bool need_grad, bool one_call) {
    if (h->tiny && (one_call || !need_grad)) {            // molecule sizes: cost (+ gradient) in one single-CTA kernel
        if (need_grad) TRY(seg_mark(h));
        TRY(enqueue_tiny(h, need_grad));
        if (need_grad) { TRY(seg_mark(h)); }
        return MAMBO_OK;
    }
    if (need_grad) TRY(seg_mark(h));
    if (!need_grad) {
        // cost only: reconstruction + residual norm with the two-GEMM kernel's first half (needs B~)
        TRY(run_segment(h, seg_for(h->g_pre_b, s_host), [&]() -> int {
            TRY(enqueue_unitary(h, h->d_x, s_host));
            return enqueue_prep(h, PREP_AB);
        }));
        Params p = make_params(h, h->T1);
        TRY((launch_fused<false, false>(h, p)));
        return run_segment(h, h->g_costonly, [&]() -> int {
            if (h->flavour == MAMBO_CSA_SD) {
                small::k_sd_onebody<<<1, 1024, 0, h->stream>>>(h->d_x, h->hT, h->U, h->d_zero, h->N, nullptr, nullptr, h->Dh, h->DU, nullptr);
                small::k_sumsq_add<<<1, 256, 0, h->stream>>>(h->Dh, h->shard == 0 ? h->N * h->N : 0, h->cost_part, h->G, h->d_part);
                h->launches += 2;
            } else {
                small::k_sumsq_add<<<1, 256, 0, h->stream>>>(nullptr, 0, h->cost_part, h->G, h->d_part);
                h->launches++;
            }
            CU(cudaGetLastError());
            return MAMBO_OK;
        });
    }
    TRY(run_segment(h, seg_for(h->g_pre, s_host), [&]() -> int {
        TRY(enqueue_unitary(h, h->d_x, s_host));
        return enqueue_prep(h, eval_prep_kind(h));
    }));
    TRY(seg_mark(h));
    const bool general = (h->mode == (int)MAMBO_SYM_GENERAL);
    GraphSeg& post = one_call ? h->g_post : h->g_post_seq;
    if (h->fast) {
        // single-GEMM formulations: Y = T~ A~ (k_ty); E~ = C~ - Y or, lean, Y itself goes through the row epilogue
        TRY(launch_ty(h, make_yparams(h, h->T1)));
        if (general) {
            TRY(run_segment(h, h->g_mid, [&]() -> int { return enqueue_post_f(h, h->F1); }));
            TRY(launch_ty(h, make_yparams(h, h->T2)));
        }
        if (one_call && split_tail(h)) return MAMBO_OK;    // the last pass's row epilogue is part of the finish half (two streams)
        if (general) {
            return run_segment(h, post, [&]() -> int {
                TRY(enqueue_post_f(h, h->F2, true));
                return enqueue_post_sr(h, !one_call);
            });
        }
        return run_segment(h, post, [&]() -> int {
            TRY(enqueue_post_f(h, h->F1, true));
            return enqueue_post_sr(h, !one_call);
        });
    }
    Params p = make_params(h, h->T1);
    TRY((launch_fused<true, false>(h, p)));
    if (general) {
        // second pass over the transposed residual: E = D A (pass 1 gave E' = D^T A); S uses E only (gradient.jl:22-26)
        TRY(run_segment(h, h->g_mid, [&]() -> int { return enqueue_post_f(h, h->F1); }));
        Params p2 = make_params(h, h->T2);
        p2.cost_part = h->cost_part + h->G;  // discarded
        TRY((launch_fused<true, false>(h, p2)));
        return run_segment(h, post, [&]() -> int {
            TRY(enqueue_post_f(h, h->F2, true));
            return enqueue_post_sr(h, !one_call);
        });
    }
    return run_segment(h, post, [&]() -> int {
        TRY(enqueue_post_f(h, h->F1, true));
        return enqueue_post_sr(h, !one_call);
    });
}

int enqueue_finish_seg(mambo_csa* h, int s_host, bool one_call) {
    if (h->tiny && one_call) {                              // tiny::k_eval has written d_out already
        TRY(seg_mark(h));
        return seg_mark(h);
    }
    TRY(seg_mark(h));
    TRY(run_segment(h, seg_for(one_call ? h->g_fin : h->g_fin_seq, s_host),
                    [&]() -> int { return enqueue_finish(h, h->d_x, h->d_out, one_call); }));
    return seg_mark(h);
}

int check_status(mambo_csa* h) {
    if (*h->h_status == 2) {
        *h->h_status = 0;
        cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream);
        return fail(MAMBO_ESTATE, "peer exchange timed out: a shard never published its partial vector");
    }
    if (*h->h_status != 0) {
        *h->h_status = 0;
        cudaMemsetAsync(h->d_status, 0, sizeof(int), h->stream);
        return fail(MAMBO_ERANGE, "||K||_1 exceeds 2^14 (or x contains NaN): orbital-rotation builder out of range");
    }
    return MAMBO_OK;
}

// full evaluation from a host x; results land in h->h_out (pinned) after synchronisation
int host_wait(mambo_csa* h) {        // wait for everything enqueued on the handle's stream, spinning or sleeping (blocking_wait)
    if (h->blocking_wait) {
        CU(cudaEventRecord(h->ev_done, h->stream));
        CU(cudaEventSynchronize(h->ev_done));
    } else {
        CU(cudaStreamSynchronize(h->stream));
    }
    return MAMBO_OK;
}
int eval_host_wait(mambo_csa* h) {   // second half of eval_host: results are in h->h_out afterwards
    CU(cudaSetDevice(h->device));
    TRY(host_wait(h));
    return check_status(h);
}

// nowait: enqueue only (H2D, kernels, D2H into the pinned staging); the caller finishes with eval_host_wait
int eval_host(mambo_csa* h, const double* x, bool need_grad, bool nowait = false) {
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "host closures need an unsharded handle (use the *_device entry points)");
    CU(cudaSetDevice(h->device));
    const int s_host = host_s_for(h, x);
    if (s_host < -1 || s_host > small::SMAX) return fail(MAMBO_ERANGE, "||K||_1 exceeds 2^14 (or x contains NaN)");
    std::memcpy(h->h_x, x, sizeof(double) * h->L);
    const long long l0 = h->launches;
    if (need_grad && h->use_graphs && !h->profiling && s_host >= 0) {
        // one graph launch for the whole closure: H2D of x, every kernel, D2H of [cost, grad] and of the status word (all
        // addresses are handle-owned and fixed; the hot kernel needs no event bracket when profiling is off)
        TRY(run_segment(h, seg_for(h->g_host, s_host), [&]() -> int {
            CU(cudaMemcpyAsync(h->d_x, h->h_x, sizeof(double) * h->L, cudaMemcpyHostToDevice, h->stream));
            TRY(enqueue_partial(h, s_host, true, true));
            TRY(enqueue_finish_seg(h, s_host, true));
            CU(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * (1 + h->L), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            return MAMBO_OK;
        }));
        h->launches_last_eval = (int)(h->launches - l0);
        if (nowait) return MAMBO_OK;
        TRY(host_wait(h));
        return check_status(h);
    }
    CU(cudaMemcpyAsync(h->d_x, h->h_x, sizeof(double) * h->L, cudaMemcpyHostToDevice, h->stream));
    int rc = enqueue_partial(h, s_host, need_grad, true);
    if (rc) return rc;
    if (need_grad) {
        rc = enqueue_finish_seg(h, s_host, true);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * (1 + h->L), cudaMemcpyDeviceToHost, h->stream));
    } else {
        CU(cudaMemcpyAsync(h->h_out, h->d_part, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->launches_last_eval = (int)(h->launches - l0);
    if (nowait) return MAMBO_OK;
    TRY(host_wait(h));
    return check_status(h);
}

// frees a temporary device allocation on every exit of the enclosing scope (error returns included)
struct DevTemp {
    void* p = nullptr;
    ~DevTemp() { if (p) cudaFree(p); }
};

template <typename T>
int dalloc(T** p, size_t count) {
    CU(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    CU(cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(T)));
    return MAMBO_OK;
}

int build_residual(mambo_csa* h, const double* tbt, const double* obt, unsigned flags, unsigned* mode_out) {
    const int n = h->N;
    h->sh = new (std::nothrow) SharedResidual();
    if (!h->sh) return fail(MAMBO_ENOMEM, "host allocation failed");
    // ---- upload the tensor (or use it in place when it is already on the device) and analyse its symmetry ----
    const size_t n4 = (size_t)n * n * n * n;
    const bool on_dev = flags & MAMBO_TBT_ON_DEVICE;
    if (h->flavour == MAMBO_CSA_SD && !obt) return fail(MAMBO_EINVAL, "CSA_SD needs the one-body tensor");
    if (h->nshards > 1 && (flags & MAMBO_SYM_MASK) == MAMBO_SYM_GENERAL)
        return fail(MAMBO_EINVAL, "row-panel sharding needs a (pq)<->(rs) symmetric tensor (SURVEY 8e.2)");
    double* mem = nullptr;
    DevTemp upload, stats_buf;              // the 8 N^4-byte upload and the statistics words never outlive this function
    if (on_dev) {
        // the caller's producer may still be running on another stream (this handle's stream does not synchronise with the
        // legacy default stream): wait for the device once before reading the tensor
        CU(cudaDeviceSynchronize());
        mem = const_cast<double*>(tbt);
    } else {
        CU(cudaMalloc((void**)&mem, n4 * sizeof(double)));
        upload.p = mem;
        CU(cudaMemcpyAsync(mem, tbt, n4 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    double* d_stats = nullptr;
    TRY(dalloc(&d_stats, 4));
    stats_buf.p = d_stats;
    small::k_sym_stats<<<h->num_sms * 8, 256, 0, h->stream>>>(mem, n, d_stats);
    double stats[4];
    CU(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const double tmax = stats[0] > 0 ? stats[0] : 1.0;
    h->asym_pair = stats[1] / tmax;
    h->asym_pq = std::max(stats[2], stats[3]) / tmax;
    const double tol = 1e-13;
    unsigned mode = flags & MAMBO_SYM_MASK;
    if (mode == MAMBO_SYM_AUTO) {
        if (h->asym_pair <= tol && h->asym_pq <= tol) mode = MAMBO_SYM_FULL;
        else if (h->asym_pair <= tol) mode = MAMBO_SYM_PAIR;
        else mode = MAMBO_SYM_GENERAL;
    }
    h->mode = (int)mode;
    h->packed = (mode == MAMBO_SYM_FULL);
    if (h->nshards > 1 && mode == MAMBO_SYM_GENERAL)
        return fail(MAMBO_EINVAL, "row-panel sharding needs a (pq)<->(rs) symmetric tensor (SURVEY 8e.2)");

    // ---- row tables ----
    h->nR = h->packed ? n * (n + 1) / 2 : n * n;
    h->nRpad = ((h->nR + ROWS - 1) / ROWS) * ROWS;
    h->ldT = h->nRpad;
    const int panels = h->nRpad / ROWS;
    h->Q = panels;
    h->prow_begin = (int)((long long)panels * h->shard / h->nshards);
    h->prow_end = (int)((long long)panels * (h->shard + 1) / h->nshards);
    h->P = h->prow_end - h->prow_begin;
    if (h->P <= 0) return fail(MAMBO_EINVAL, "more shards than 64-row panels");
    h->row_begin = h->prow_begin * ROWS;
    h->row_end = h->prow_end * ROWS;
    std::vector<int> pa(h->nRpad, 0), pq(h->nRpad, 0);
    std::vector<double> sw(h->nRpad, 0.0);
    if (h->packed) {
        int r = 0;
        for (int q = 0; q < n; ++q)
            for (int a = 0; a <= q; ++a, ++r) {
                pa[r] = a;
                pq[r] = q;
                sw[r] = (a == q) ? 1.0 : std::sqrt(2.0);
            }
    } else {
        for (int r = 0; r < h->nR; ++r) {
            pa[r] = r % n;
            pq[r] = r / n;
            sw[r] = 1.0;
        }
    }
    TRY(dalloc(&h->pa, h->nRpad));
    TRY(dalloc(&h->pq, h->nRpad));
    TRY(dalloc(&h->sw, h->nRpad));
    CU(cudaMemcpy(h->pa, pa.data(), sizeof(int) * h->nRpad, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->pq, pq.data(), sizeof(int) * h->nRpad, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->sw, sw.data(), sizeof(double) * h->nRpad, cudaMemcpyHostToDevice));

    // ---- residual matrix T~ (local rows) ----
    const size_t tcount = (size_t)h->P * ROWS * h->ldT;
    CU(cudaMalloc((void**)&h->T1, tcount * sizeof(double)));
    small::k_pack<<<h->num_sms * 16, 256, 0, h->stream>>>(mem, n, h->pa, h->pq, h->sw, h->row_begin, h->P * ROWS, h->nR, h->ldT, 0, h->T1);
    if (mode == MAMBO_SYM_GENERAL) {
        CU(cudaMalloc((void**)&h->T2, tcount * sizeof(double)));
        small::k_pack<<<h->num_sms * 16, 256, 0, h->stream>>>(mem, n, h->pa, h->pq, h->sw, h->row_begin, h->P * ROWS, h->nR, h->ldT, 1, h->T2);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    if (upload.p) {                         // release the upload before the remaining allocations (high-water mark)
        cudaFree(upload.p);
        upload.p = nullptr;
    }

    if (h->flavour == MAMBO_CSA_SD) {
        TRY(dalloc(&h->hT, (size_t)n * n));
        CU(cudaMemcpy(h->hT, obt, sizeof(double) * n * n, on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    }

    TRY(dalloc(&h->sh->tnorm2, 1));
    h->tnorm2 = h->sh->tnorm2;
    *mode_out = mode;
    return MAMBO_OK;
}

int build_handle(mambo_csa* h, const double* tbt, const double* obt, unsigned flags, const mambo_csa* parent = nullptr) {
    const int n = h->N;
    CU(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) return fail(MAMBO_ECUDA, "libmambo_b200 needs an sm_100a (B200) device; found sm_" + std::to_string(prop.major * 10 + prop.minor));
    h->num_sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;

    h->cL = n * (n + 1) / 2;
    h->uL = n * (n - 1) / 2;
    h->lam_off = (h->flavour == MAMBO_CSA_SD) ? n : 0;
    h->L = h->lam_off + h->cL + h->uL;
    int nb = (n + 7) / 8;
    h->NB = 0;
    for (int b : kNBBuckets)
        if (b >= nb) {
            h->NB = b;
            break;
        }
    if (!h->NB) return fail(MAMBO_EINVAL, "N too large for the compiled kernel set (N <= 168)");
    h->Np = 8 * h->NB;
    h->ld = h->Np + 4;

    unsigned mode = 0;
    if (parent) {
        // evaluation lane: same problem, same GPU, same residual; only the per-evaluation state below is new
        mode = (unsigned)parent->mode;
        h->mode = parent->mode; h->packed = parent->packed; h->asym_pair = parent->asym_pair; h->asym_pq = parent->asym_pq;
        h->nR = parent->nR; h->nRpad = parent->nRpad; h->ldT = parent->ldT; h->Q = parent->Q; h->P = parent->P;
        h->prow_begin = parent->prow_begin; h->prow_end = parent->prow_end; h->row_begin = parent->row_begin; h->row_end = parent->row_end;
        h->sh = parent->sh;
        h->sh->refs.fetch_add(1);
        h->T1 = parent->T1; h->T2 = parent->T2; h->hT = parent->hT; h->pa = parent->pa; h->pq = parent->pq; h->sw = parent->sw;
        h->tnorm2 = parent->tnorm2;
    } else {
        TRY(build_residual(h, tbt, obt, flags, &mode));
    }

    // ---- operands and work buffers ----
    TRY(dalloc(&h->At, (size_t)h->nRpad * h->ld));
    TRY(dalloc(&h->Bt, (size_t)h->P * ROWS * h->ld));
    // schedule of the persistent fused kernel
    const long long total = (long long)h->P * h->Q;
    h->G = (int)std::min<long long>(h->num_sms, total);
    std::vector<int> vstart(h->G + 1, 0), pfirst(h->P + 1, 0);
    {
        std::vector<int> seq;
        for (int b = 0; b < h->G; ++b) {
            const long long tb = total * b / h->G, te = total * (b + 1) / h->G;
            vstart[b] = (int)seq.size();
            if (te > tb)
                for (int p = (int)(tb / h->Q); p <= (int)((te - 1) / h->Q); ++p) seq.push_back(p);
        }
        vstart[h->G] = (int)seq.size();
        h->visits = (int)seq.size();
        int v = 0;
        for (int p = 0; p <= h->P; ++p) {
            while (v < h->visits && seq[v] < p) ++v;
            pfirst[p] = v;
        }
    }
    TRY(dalloc(&h->visit_start, h->G + 1));
    TRY(dalloc(&h->piece_first, h->P + 1));
    CU(cudaMemcpy(h->visit_start, vstart.data(), sizeof(int) * (h->G + 1), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->piece_first, pfirst.data(), sizeof(int) * (h->P + 1), cudaMemcpyHostToDevice));
    TRY(dalloc(&h->Epiece, (size_t)h->visits * 2 * ROWS * h->Np));
    TRY(dalloc(&h->Et, (size_t)h->P * ROWS * h->Np));
    TRY(dalloc(&h->F1, (size_t)h->P * ROWS * h->Np));
    if (mode == MAMBO_SYM_GENERAL) TRY(dalloc(&h->F2, (size_t)h->P * ROWS * h->Np));
    TRY(dalloc(&h->Spart, (size_t)h->P * n * n));
    TRY(dalloc(&h->cost_part, (size_t)2 * h->G));
    // shared-memory ring depth
    const size_t limit = prop.sharedMemPerBlockOptin;
    int S = 8;
    while (S > 2 && smem_bytes(h->ld, S) > limit) --S;
    if (smem_bytes(h->ld, S) > limit) return fail(MAMBO_EINVAL, "operand panel does not fit in shared memory");
    h->S = S;
    h->smem = smem_bytes(h->ld, S);

    const size_t nn = (size_t)n * n;
    TRY(dalloc(&h->mats, 4 * nn));
    h->Lam = h->mats; h->Dh = h->mats + nn; h->DU = h->mats + 2 * nn; h->hfrag = h->mats + 3 * nn;
    TRY(dalloc(&h->U, nn));
    TRY(dalloc(&h->d_zero, 1));
    {
        // N <= 64: the <= 4 x 4 tiles of the chain run as ONE thread-block cluster and synchronise on the hardware cluster
        // barrier.  (A 32 x 32 tiling would bring N <= 128 under the 16-CTA limit too, but was measured slower: a step is
        // then bound by the DMMA rate of 16 SMs instead of the barrier.)
        h->chain_ct = 16;
        h->chain_cluster = false;
        if (n <= 64 && std::getenv("MAMBO_CHAIN_GRID") == nullptr) {
            // only if the device can actually place one such cluster (<= 16 CTAs with this much shared memory in one GPC)
            const int tiles = (n + 15) / 16, sm_chain = (int)chain::smem_bytes(n, 16);
            bool ok = true;
            for (const void* f : {(const void*)chain::k_chain_u<16, chain::SYNC_CLUSTER>, (const void*)chain::k_chain_d<16, chain::SYNC_CLUSTER>}) {
                ok = ok && raise_smem(f, sm_chain) == cudaSuccess;
                ok = ok && cudaFuncSetAttribute(f, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
                cudaLaunchConfig_t cfg;
                std::memset(&cfg, 0, sizeof(cfg));
                cfg.gridDim = dim3(tiles, tiles);
                cfg.blockDim = dim3(chain::NTHR);
                cfg.dynamicSmemBytes = sm_chain;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = tiles;
                attr[0].val.clusterDim.y = tiles;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                int nclusters = 0;
                ok = ok && cudaOccupancyMaxActiveClusters(&nclusters, f, &cfg) == cudaSuccess && nclusters >= 1;
            }
            cudaGetLastError();
            h->chain_cluster = ok;
        }
        const size_t ms = chain::mat_stride(n, h->chain_ct);
        const size_t ncm = 11 + 11 + 4 * (size_t)(chain::SMAXC + 1);
        TRY(dalloc(&h->cmats, ncm * ms));
        double* cm = h->cmats;
        auto ctake = [&](size_t k) { double* r = cm; cm += k * ms; return r; };
        chain::UParams& u = h->cu;
        u.X = ctake(1); u.Xt = ctake(1); u.X2 = ctake(1); u.X2t = ctake(1); u.X4 = ctake(1);
        u.P0 = ctake(1); u.P1 = ctake(1); u.P2 = ctake(1); u.Q3t = ctake(1); u.Q2t = ctake(1); u.Q1t = ctake(1);
        u.R = ctake(chain::SMAXC + 1); u.Rt = ctake(chain::SMAXC + 1);
        chain::DParams& d = h->cd;
        d.X = u.X; d.Xt = u.Xt; d.X2 = u.X2; d.X2t = u.X2t; d.X4 = u.X4; d.Q3t = u.Q3t; d.Q2t = u.Q2t; d.Q1t = u.Q1t; d.R = u.R; d.Rt = u.Rt;
        d.dX2 = ctake(1); d.dX2t = ctake(1); d.dX4 = ctake(1); d.dP0 = ctake(1); d.dP1 = ctake(1); d.dP2 = ctake(1);
        d.dQ3t = ctake(1); d.dQ2t = ctake(1); d.dQ1t = ctake(1); (void)ctake(2);
        d.dR = ctake(chain::SMAXC + 1); d.dRt = ctake(chain::SMAXC + 1);
    }
    TRY(dalloc(&h->g1, (size_t)n));
    TRY(dalloc(&h->obc, 1));
    CU(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming | cudaEventBlockingSync));
    h->blocking_wait = parent ? parent->blocking_wait : (std::getenv("MAMBO_BLOCKING_WAIT") != nullptr);
    TRY(dalloc(&h->d_s, 1));
    TRY(dalloc(&h->d_bar, 4));                  // [0] barrier counter, [1] exit counter, [2] scratch target of release operations
    if (std::getenv("MAMBO_CHAIN_DEBUG")) TRY(dalloc(&h->d_dbg, 6 * 16));
    TRY(dalloc(&h->d_scale, 1));
    TRY(dalloc(&h->d_status, 1));
    TRY(dalloc(&h->d_x, (size_t)h->L));
    TRY(dalloc(&h->d_out, (size_t)h->L + 1));
    TRY(dalloc(&h->d_part, (size_t)1 + 2 * nn));
    CU(cudaMallocHost((void**)&h->h_x, sizeof(double) * h->L));
    CU(cudaMallocHost((void**)&h->h_out, sizeof(double) * (h->L + 1)));
    CU(cudaMallocHost((void**)&h->h_status, sizeof(int)));
    *h->h_status = 0;
    h->memo_x.assign(h->L, 0.0);
    h->use_graphs = (std::getenv("MAMBO_NO_GRAPH") == nullptr);
    h->dephase = 0;                             // measured: no gain from offsetting the column groups (kept as a knob)
    if (const char* e = std::getenv("MAMBO_DEPHASE")) h->dephase = std::atoi(e);
    if (const char* e = std::getenv("MAMBO_LOOKAHEAD")) h->lookahead = std::atoi(e);
    {
        size_t need_post = 0, need_prep = 0;
        NB_PANEL_DISPATCH(h->NB, need_post = (panel::post_smem<NBT, CST>()));
        NB_PANEL_DISPATCH(h->NB, need_prep = (panel::prep_smem<NBT, CSP>()));
        if (std::max(need_post, need_prep) > prop.sharedMemPerBlockOptin || chain::smem_bytes(n, h->chain_ct) > prop.sharedMemPerBlockOptin)
            return fail(MAMBO_EINVAL, "N too large for the shared-memory budget of the panel / chain kernels");
    }
    {
        int tiny_max = 12;                                   // measured (tools/tiny_threshold.py): 40 vs 65 us at N = 12, break-even at N = 15
        if (const char* e = std::getenv("MAMBO_TINY_MAX")) tiny_max = std::max(0, std::min(tiny::NMAX, std::atoi(e)));
        if (std::getenv("MAMBO_NO_TINY")) tiny_max = 0;
        h->tiny = h->packed && h->nshards == 1 && n <= tiny_max && tiny::smem_bytes(n, h->nR) <= prop.sharedMemPerBlockOptin;
        if (h->tiny) CU(raise_smem(tiny::k_eval, (int)tiny::smem_bytes(n, h->nR)));
    }
    h->fast = std::getenv("MAMBO_TWO_GEMM") == nullptr;
    // lean path: unsharded handles only (a shard needs its own rows' |W~|^2, which has no N x N closed form); MAMBO_NO_LEAN=1
    // keeps the first single-GEMM formulation (Gram kernel + B~/C~ panel GEMMs) as the A/B reference of the parity tests
    h->lean = h->fast && h->nshards == 1 && std::getenv("MAMBO_NO_LEAN") == nullptr;
    if (const char* e = std::getenv("MAMBO_TAU")) h->tau = std::atof(e);
    if (std::getenv("MAMBO_TY_DEBUG") && h->fast) TRY(dalloc(&h->ty_dbg, (size_t)16 * 64 * 4 + 160 * 4));
    h->ty_dephase = 256 * h->NB;                // half a step (16 NB DMMAs of 16 cycles per warp and step): measured 1 % faster than 0
    if (const char* e = std::getenv("MAMBO_TY_DEPHASE")) h->ty_dephase = std::atoi(e);
    if (const char* e = std::getenv("MAMBO_TY_MB")) h->ty_mb = (std::atoi(e) == 1) ? 1 : 2;
    if (h->fast) {
        TRY(dalloc(&h->Ct, (size_t)h->P * ROWS * h->ld));
        TRY(dalloc(&h->GLP, (size_t)h->Np * h->ld));
        TRY(dalloc(&h->wpart, (size_t)h->P * 8));
        TRY(dalloc(&h->aux, (size_t)3 * n + 1));
        TRY(dalloc(&h->d_flag, 1));
        TRY(dalloc(&h->ls_part, (size_t)n));
        TRY(dalloc(&h->d_done, 1));
        int Sy = 8;
        if (const char* e = std::getenv("MAMBO_TY_RING")) Sy = std::max(2, std::min(8, std::atoi(e)));
        while (Sy > 2 && smem_bytes_y(h->ld, Sy) > prop.sharedMemPerBlockOptin) --Sy;
        if (smem_bytes_y(h->ld, Sy) > prop.sharedMemPerBlockOptin) return fail(MAMBO_EINVAL, "operand ring does not fit in shared memory");
        h->Sy = Sy;
        h->smem_y = smem_bytes_y(h->ld, Sy);
        YParams cfgy = make_yparams(h, h->T1);
        cfgy.P = -1;
        TRY(launch_ty(h, cfgy));
    }
    {
        TRY(dalloc(&h->LamP, (size_t)h->Np * h->ld));
        TRY(dalloc(&h->EtP, (size_t)h->P * ROWS * h->ld));
        NB_PANEL_DISPATCH(h->NB, CU(raise_smem(panel::k_prep_mma<NBT, CSP>, (int)(panel::prep_smem<NBT, CSP>()))));
        h->prep_cs = panel_cs(h->NB);
        NB_PANEL_DISPATCH(h->NB, CU(raise_smem(panel::k_post_mma<NBT, CST>, (int)(panel::post_smem<NBT, CST>()))));
    }
    {
        const int sm_chain = (int)chain::smem_bytes(n, h->chain_ct);
        for (const void* f : {(const void*)chain::k_chain_u<16, chain::SYNC_GRID>, (const void*)chain::k_chain_d<16, chain::SYNC_GRID>,
                              (const void*)chain::k_chain_u<16, chain::SYNC_CLUSTER>, (const void*)chain::k_chain_d<16, chain::SYNC_CLUSTER>})
            CU(raise_smem(f, sm_chain));
        if (h->chain_cluster)
            for (const void* f : {(const void*)chain::k_chain_u<16, chain::SYNC_CLUSTER>, (const void*)chain::k_chain_d<16, chain::SYNC_CLUSTER>})
                CU(cudaFuncSetAttribute(f, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }

    // opt-in shared-memory sizes, set once per handle
    {
        Params cfg = make_params(h, h->T1);
        cfg.P = -1;
        TRY((launch_fused<true, false>(h, cfg)));
        TRY((launch_fused<false, false>(h, cfg)));
        TRY((launch_fused<false, true>(h, cfg)));
    }
    if (!parent) {
        // |T~|^2 of the local rows (zero padding contributes nothing); kept current by subtract_impl
        small::k_sumsq_partial<<<h->G, 256, 0, h->stream>>>(h->T1, (long long)h->P * ROWS * h->ldT, h->cost_part);
        small::k_sumsq_add<<<1, 256, 0, h->stream>>>(nullptr, 0, h->cost_part, h->G, h->tnorm2);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    }
    return MAMBO_OK;
}

int subtract_impl(mambo_csa* h, const double* d_x, int s_host) {
    TRY(enqueue_unitary(h, d_x, s_host));
    TRY(enqueue_prep(h, PREP_AB));
    Params p = make_params(h, h->T1);
    TRY((launch_fused<false, true>(h, p)));
    if (h->mode == (int)MAMBO_SYM_GENERAL) {
        Params p2 = make_params(h, h->T2);
        p2.cost_part = h->cost_part + h->G;
        TRY((launch_fused<false, true>(h, p2)));
    }
    const int n = h->N, nn = n * n;
    small::k_sumsq_add<<<1, 256, 0, h->stream>>>(nullptr, 0, h->cost_part, h->G, h->tnorm2);   // |T~_new|^2 (two-body part)
    h->launches++;
    if (h->flavour == MAMBO_CSA_SD) {
        small::k_sd_onebody<<<1, 1024, 0, h->stream>>>(d_x, h->hT, h->U, h->d_zero, n, nullptr, nullptr, h->Dh, h->DU, h->hfrag);
        small::k_obt_sub<<<(nn + 127) / 128, 128, 0, h->stream>>>(h->hT, h->hfrag, n);
        // sharded handles return shard-local costs that the caller sums: the (replicated) one-body term counts once
        small::k_sumsq_add<<<1, 256, 0, h->stream>>>(h->Dh, h->shard == 0 ? nn : 0, h->cost_part, h->G, h->d_part);
        h->launches += 3;
    } else {
        small::k_sumsq_add<<<1, 256, 0, h->stream>>>(nullptr, 0, h->cost_part, h->G, h->d_part);
        h->launches++;
    }
    CU(cudaGetLastError());
    return MAMBO_OK;
}

}  // namespace

// =========================================================================================================
//                                                C ABI
// =========================================================================================================
extern "C" {

const char* mambo_last_error(void) { return g_err.c_str(); }

// Library-internal (used by optimizer.cu, not declared in the public header): a device scratch area owned by the handle,
// grown on demand and released with it.  cudaMalloc / cudaFree per optimisation run would synchronise the whole device
// and stall sibling lanes.
int mambo_csa_scratch(mambo_csa* h, size_t bytes, void** ptr) {
    if (!h || !ptr) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    if (h->opt_ws_bytes < bytes) {
        if (h->opt_ws) CU(cudaFree(h->opt_ws));
        h->opt_ws = nullptr;
        h->opt_ws_bytes = 0;
        CU(cudaMalloc(&h->opt_ws, bytes));
        h->opt_ws_bytes = bytes;
    }
    *ptr = h->opt_ws;
    return MAMBO_OK;
}
int mambo_csa_scratch_host(mambo_csa* h, void** ptr) {   // 1 KB of pinned host memory owned by the handle
    if (!h || !ptr) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    if (!h->opt_hs) CU(cudaMallocHost(&h->opt_hs, 1024));
    *ptr = h->opt_hs;
    return MAMBO_OK;
}
void mambo_set_error(const char* msg) { g_err = msg ? msg : ""; }
// Library-internal (optimizer.cu): the device status word travels with the scalars of every optimiser evaluation, so an
// out-of-range orbital rotation or a peer time-out ends the run with an error code instead of a garbage minimum.
int mambo_csa_status_fetch(mambo_csa* h) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    CU(cudaMemcpyAsync(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    return MAMBO_OK;
}
int mambo_csa_status_check(mambo_csa* h) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    return check_status(h);
}

// Library-internal (df.cu): the resident residual of an unsharded handle, read-only.
int mambo_csa_resident_view(mambo_csa* h, mambo_resident_view* v) {
    if (!h || !v) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "the double factorisation needs an unsharded handle");
    v->N = h->N; v->mode = h->mode; v->packed = h->packed ? 1 : 0; v->rows = h->nR; v->device = h->device;
    v->asym_pair = h->asym_pair; v->ldT = h->ldT; v->T1 = h->T1; v->T2 = h->T2; v->pa = h->pa; v->pq = h->pq; v->sw = h->sw; v->stream = h->stream;
    return MAMBO_OK;
}
void** mambo_csa_df_slot(mambo_csa* h, void (*destroy)(void*)) {
    if (!h) return nullptr;
    if (destroy) h->df_destroy = destroy;
    return &h->df_state;
}
void mambo_csa_count_launches(mambo_csa* h, long long n) { if (h) h->launches += n; }
int mambo_csa_io_buffers(mambo_csa* h, double** d_x, double** d_out) {
    if (!h || !d_x || !d_out) return fail(MAMBO_EINVAL, "null argument");
    *d_x = h->d_x;
    *d_out = h->d_out;
    return MAMBO_OK;
}

int mambo_csa_get_stream(mambo_csa* h, void** stream, int* device) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    if (stream) *stream = (void*)h->stream;
    if (device) *device = h->device;
    return MAMBO_OK;
}

int mambo_csa_create_sharded(mambo_csa** out, int N, int flavour, const double* tbt, const double* obt, int device, unsigned flags,
                             int shard, int nshards) {
    if (!out || !tbt || N < 1) return fail(MAMBO_EINVAL, "bad arguments to mambo_csa_create");
    if (flavour != MAMBO_CSA && flavour != MAMBO_CSA_SD) return fail(MAMBO_EINVAL, "flavour must be MAMBO_CSA or MAMBO_CSA_SD");
    if (nshards < 1 || shard < 0 || shard >= nshards) return fail(MAMBO_EINVAL, "bad shard index");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(MAMBO_ECUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(MAMBO_EINVAL, "device index out of range");
    mambo_csa* h = new (std::nothrow) mambo_csa();
    if (!h) return fail(MAMBO_ENOMEM, "host allocation failed");
    h->N = N;
    h->flavour = flavour;
    h->device = device;
    h->shard = shard;
    h->nshards = nshards;
    int rc = build_handle(h, tbt, obt, flags);
    if (rc) {
        std::string keep = g_err;
        mambo_csa_destroy(h);
        g_err = keep;
        return rc;
    }
    *out = h;
    return MAMBO_OK;
}

// A second evaluation context on the same GPU and the same residual: own stream, operands, scratch and CUDA graphs,
// shared T~ / one-body residual.  Lanes let independent evaluations (multistart, line-search probes of different runs)
// overlap: the latency-bound orbital-rotation chains and epilogues of one lane run beside those of another while
// the tensor-pipe-bound fused kernels queue back to back.
int mambo_csa_create_lane(mambo_csa* parent, mambo_csa** out) {
    if (!parent || !out) return fail(MAMBO_EINVAL, "null argument");
    mambo_csa* h = new (std::nothrow) mambo_csa();
    if (!h) return fail(MAMBO_ENOMEM, "host allocation failed");
    h->N = parent->N;
    h->flavour = parent->flavour;
    h->device = parent->device;
    h->shard = parent->shard;
    h->nshards = parent->nshards;
    h->is_lane = true;
    int rc = build_handle(h, nullptr, nullptr, 0, parent);
    if (rc) {
        std::string keep = g_err;
        mambo_csa_destroy(h);
        g_err = keep;
        return rc;
    }
    *out = h;
    return MAMBO_OK;
}

int mambo_csa_create(mambo_csa** out, int N, int flavour, const double* tbt, const double* obt, int device, unsigned flags) {
    return mambo_csa_create_sharded(out, N, flavour, tbt, obt, device, flags, 0, 1);
}

void mambo_csa_destroy(mambo_csa* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    if (h->side_stream) cudaStreamSynchronize(h->side_stream);
    if (h->df_state && h->df_destroy) h->df_destroy(h->df_state);
    for (auto* v : {&h->g_pre, &h->g_fin, &h->g_host, &h->g_pre_b, &h->g_fin_seq})
        for (GraphSeg& g : *v)
            if (g.exec) cudaGraphExecDestroy(g.exec);
    for (GraphSeg* g : {&h->g_mid, &h->g_post, &h->g_post_seq, &h->g_costonly})
        if (g->exec) cudaGraphExecDestroy(g->exec);
    if (!h->sh || h->sh->refs.fetch_sub(1) == 1) {
        void* shared[] = {h->T1, h->T2, h->hT, h->pa, h->pq, h->sw, h->tnorm2};
        for (void* b : shared)
            if (b) cudaFree(b);
        delete h->sh;
    }
    void* bufs[] = {h->aux, h->obc, h->ty_dbg, h->Ct, h->GLP, h->wpart, h->d_flag, h->ls_part, h->d_done, h->At, h->Bt, h->Epiece, h->Et, h->F1, h->F2, h->Spart, h->cost_part,
                    h->eig_ws, h->opt_ws, h->visit_start, h->piece_first, h->mats, h->cmats, h->LamP, h->EtP, h->U, h->d_zero, h->g1, h->d_s, h->d_bar, h->d_dbg, h->d_scale, h->d_status, h->d_x, h->d_out, h->d_part};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (h->opt_hs) cudaFreeHost(h->opt_hs);
    if (h->h_x) cudaFreeHost(h->h_x);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_status) cudaFreeHost(h->h_status);
    for (void* m : h->peer_opened)
        if (m) cudaIpcCloseMemHandle(m);
    if (h->peer_recv) cudaFree(h->peer_recv);
    if (h->peer_done) cudaFree(h->peer_done);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->seg_pool) cudaEventDestroy(e);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int mambo_csa_num_params(const mambo_csa* h) { return h ? h->L : MAMBO_EINVAL; }

int mambo_csa_get_info(const mambo_csa* h, mambo_csa_info_t* info) {
    if (!h || !info) return fail(MAMBO_EINVAL, "null argument");
    info->N = h->N;
    info->flavour = h->flavour;
    info->mode = h->mode;
    info->num_params = h->L;
    info->rows = h->nR;
    info->rows_padded = h->nRpad;
    info->shard = h->shard;
    info->nshards = h->nshards;
    info->local_row_begin = h->row_begin;
    info->local_row_end = std::min(h->row_end, h->nR);
    info->partial_len = 1 + 2 * h->N * h->N;
    info->num_sms = h->num_sms;
    info->launches_per_eval = h->launches_last_eval;
    const double passes = (h->mode == (int)MAMBO_SYM_GENERAL) ? 2.0 : 1.0;
    info->flops_per_eval = passes * (h->fast ? 2.0 : 4.0) * (double)h->P * ROWS * (double)h->nRpad * h->Np;
    info->tensor_bytes = passes * (double)h->P * ROWS * (double)h->ldT * 8.0;
    info->asym_pair = h->asym_pair;
    info->asym_pq = h->asym_pq;
    info->hot_gemms = h->fast ? 1 : 2;
    {
        const int tiles = (h->N + h->chain_ct - 1) / h->chain_ct;
        info->lanes_hint = std::max(1, std::min(8, (4 * h->num_sms) / (tiles * tiles)));   // measured at N=100: 2/4/8 lanes -> 3.3k/4.35k/4.48k evals/s
    }
    return MAMBO_OK;
}

int mambo_csa_cost_grad(mambo_csa* h, const double* x, double* cost, double* grad) {
    if (!h || !x) return fail(MAMBO_EINVAL, "null argument");
    const long long version = h->sh->version.load();   // residual updates through any lane invalidate every memo
    if (!(h->memo_valid && h->memo_version == version && std::memcmp(h->memo_x.data(), x, sizeof(double) * h->L) == 0)) {
        h->memo_valid = false;
        TRY(eval_host(h, x, true));
        std::memcpy(h->memo_x.data(), x, sizeof(double) * h->L);
        h->memo_valid = true;
        h->memo_version = version;
    }
    if (cost) *cost = h->h_out[0];
    if (grad) std::memcpy(grad, h->h_out + 1, sizeof(double) * h->L);
    return MAMBO_OK;
}

// Optim.jl calls cost(x) and grad!(g, x) at the same x back to back (also inside the HagerZhang line
// search): one fused evaluation serves both.
int mambo_csa_cost(mambo_csa* h, const double* x, double* cost) { return mambo_csa_cost_grad(h, x, cost, nullptr); }
int mambo_csa_grad(mambo_csa* h, const double* x, double* grad) { return mambo_csa_cost_grad(h, x, nullptr, grad); }

// K independent evaluations (multistart: one point per run) in ONE call: point k goes to lane k mod nlanes.  Every lane is kept
// busy: the call enqueues one evaluation per lane without waiting (H2D from the lane's pinned staging, kernels, D2H), then walks
// the points in order -- wait for point k's lane, copy its result out, and at once enqueue that lane's next point (k + nlanes)
// while the other lanes are still computing.  The lanes overlap on the device as they do under one caller thread each, without
// needing those threads.
int mambo_csa_cost_grad_batch(mambo_csa** lanes, int nlanes, int K, const double* xs, double* costs, double* grads) {
    if (!lanes || nlanes < 1 || K < 0 || !xs) return fail(MAMBO_EINVAL, "bad arguments");
    const int L = lanes[0] ? lanes[0]->L : 0;
    for (int i = 0; i < nlanes; ++i)
        if (!lanes[i] || lanes[i]->L != L) return fail(MAMBO_EINVAL, "lanes must be handles of the same problem");
    int rc_first = MAMBO_OK;
    std::string first_msg;
    std::vector<char> inflight(nlanes, 0);
    auto note = [&](int rc) {
        if (rc && !rc_first) {
            rc_first = rc;
            first_msg = g_err;
        }
    };
    auto enqueue = [&](int k) {
        mambo_csa* ln = lanes[k % nlanes];
        ln->memo_valid = false;
        const int rc = eval_host(ln, xs + (size_t)k * L, true, true);
        if (!rc) inflight[k % nlanes] = 1;
        note(rc);
    };
    for (int k = 0; k < std::min(K, nlanes) && !rc_first; ++k) enqueue(k);
    for (int k = 0; k < K && !rc_first; ++k) {
        mambo_csa* ln = lanes[k % nlanes];
        const int rc = eval_host_wait(ln);
        inflight[k % nlanes] = 0;
        note(rc);
        if (rc) break;
        if (costs) costs[k] = ln->h_out[0];
        if (grads) std::memcpy(grads + (size_t)k * L, ln->h_out + 1, sizeof(double) * L);
        if (k + nlanes < K) enqueue(k + nlanes);
    }
    if (rc_first) {
        for (int i = 0; i < nlanes; ++i)                  // every lane still in flight is drained before the error is returned
            if (inflight[i]) eval_host_wait(lanes[i]);
        g_err = first_msg;
        return rc_first;
    }
    return MAMBO_OK;
}

// 0: the host closures spin in cudaStreamSynchronize (default; lowest latency); 1: they sleep on a blocking-sync event, which
// frees the core while the device works -- for callers that outnumber the host cores (lanes x ranks per box).
int mambo_csa_set_wait_mode(mambo_csa* h, int blocking) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    h->blocking_wait = blocking != 0;
    return MAMBO_OK;
}

int mambo_csa_set_stream(mambo_csa* h, void* stream) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return MAMBO_OK;
}

int mambo_csa_set_profiling(mambo_csa* h, int enable) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    h->profiling = enable != 0;
    h->ev_used = 0;
    h->seg_used = 0;
    return MAMBO_OK;
}

// Sum of the device durations of the fused-kernel launches recorded since profiling was (re)enabled.
int mambo_csa_get_profile(mambo_csa* h, double* fused_ms, int* fused_launches, long long* total_launches) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    double ms = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, h->ev_pool[i], h->ev_pool[i + 1]));
        ms += t;
    }
    if (fused_ms) *fused_ms = ms;
    if (fused_launches) *fused_launches = (int)(h->ev_used / 2);
    if (total_launches) *total_launches = h->launches;
    h->ev_used = 0;
    return MAMBO_OK;
}

// Segment breakdown of the evaluations recorded since profiling was (re)enabled:
// seg_ms[0] = unitary chain + operand build, [1] = fused kernel(s) + row epilogues, [2] = Frechet chain + assembly.
int mambo_csa_get_segments(mambo_csa* h, double* seg_ms, int* evals) {
    if (!h || !seg_ms) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    seg_ms[0] = seg_ms[1] = seg_ms[2] = 0.0;
    int n = 0;
    for (size_t i = 0; i + 3 < h->seg_used; i += 4, ++n)
        for (int k = 0; k < 3; ++k) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, h->seg_pool[i + k], h->seg_pool[i + k + 1]));
            seg_ms[k] += t;
        }
    if (evals) *evals = n;
    h->seg_used = 0;
    return MAMBO_OK;
}

int mambo_csa_debug_ty_stamps(mambo_csa* h, long long* out4096) {
    if (!h || !h->ty_dbg) return fail(MAMBO_ESTATE, "set MAMBO_TY_DEBUG=1 before create");
    CU(cudaMemcpy(out4096, h->ty_dbg, sizeof(long long) * (16 * 64 * 4 + 160 * 4), cudaMemcpyDeviceToHost));
    return MAMBO_OK;
}

int mambo_csa_debug_stamps(mambo_csa* h, long long* out96) {
    if (!h || !h->d_dbg) return fail(MAMBO_ESTATE, "set MAMBO_CHAIN_DEBUG=1 before create");
    CU(cudaMemcpy(out96, h->d_dbg, sizeof(long long) * 96, cudaMemcpyDeviceToHost));
    return MAMBO_OK;
}

int mambo_csa_synchronize(mambo_csa* h) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(h->h_status, h->d_status, sizeof(int), cudaMemcpyDeviceToHost));
    return check_status(h);
}

int mambo_csa_eval_partial_device(mambo_csa* h, const double* d_x, double* d_partial) {
    if (!h || !d_x || !d_partial) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const long long l0 = h->launches;
    if (d_x != h->d_x) CU(cudaMemcpyAsync(h->d_x, d_x, sizeof(double) * h->L, cudaMemcpyDeviceToDevice, h->stream));
    TRY(enqueue_partial(h, -1, true, false));
    CU(cudaMemcpyAsync(d_partial, h->d_part, sizeof(double) * (1 + 2 * (size_t)h->N * h->N), cudaMemcpyDeviceToDevice, h->stream));
    h->launches_last_eval = (int)(h->launches - l0);
    h->memo_valid = false;
    return MAMBO_OK;
}

int mambo_csa_eval_finish_device(mambo_csa* h, const double* d_partial, double* d_out) {
    if (!h || !d_partial || !d_out) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const long long l0 = h->launches;
    CU(cudaMemcpyAsync(h->d_part, d_partial, sizeof(double) * (1 + 2 * (size_t)h->N * h->N), cudaMemcpyDeviceToDevice, h->stream));
    TRY(enqueue_finish_seg(h, -1, false));
    if (d_out != h->d_out) CU(cudaMemcpyAsync(d_out, h->d_out, sizeof(double) * (1 + h->L), cudaMemcpyDeviceToDevice, h->stream));
    h->launches_last_eval += (int)(h->launches - l0);
    return MAMBO_OK;
}

int mambo_csa_is_tiny(mambo_csa* h) { return h && h->tiny ? 1 : 0; }

int mambo_csa_eval_device(mambo_csa* h, const double* d_x, double* d_out) {
    if (!h || !d_x || !d_out) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "sharded handle: use eval_partial_device / eval_finish_device around an allreduce");
    CU(cudaSetDevice(h->device));
    return mambo_csa_enqueue_eval(h, d_x, d_out);
}

int mambo_csa_enqueue_eval(mambo_csa* h, const double* d_x, double* d_out) {
    const long long l0 = h->launches;
    if (d_x != h->d_x) CU(cudaMemcpyAsync(h->d_x, d_x, sizeof(double) * h->L, cudaMemcpyDeviceToDevice, h->stream));
    TRY(enqueue_partial(h, -1, true, true));
    TRY(enqueue_finish_seg(h, -1, true));
    if (d_out != h->d_out) CU(cudaMemcpyAsync(d_out, h->d_out, sizeof(double) * (1 + h->L), cudaMemcpyDeviceToDevice, h->stream));
    h->launches_last_eval = (int)(h->launches - l0);
    h->memo_valid = false;
    return MAMBO_OK;
}

int mambo_csa_cost_device(mambo_csa* h, const double* d_x, double* d_cost) {
    if (!h || !d_x || !d_cost) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "cost_device needs an unsharded handle");
    CU(cudaSetDevice(h->device));
    if (d_x != h->d_x) CU(cudaMemcpyAsync(h->d_x, d_x, sizeof(double) * h->L, cudaMemcpyDeviceToDevice, h->stream));
    TRY(enqueue_partial(h, -1, false, false));
    CU(cudaMemcpyAsync(d_cost, h->d_part, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    h->memo_valid = false;
    return MAMBO_OK;
}

// ---- peer exchange of row-panel shards -----------------------------------------------------------------------------
// Layout of the exported allocation: [2][nshards][plen_pad] doubles, then [2][MAXPEERS] u64 flags.
int mambo_csa_peer_export(mambo_csa* h, unsigned char* handle64) {
    if (!h || !handle64) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards > small::MAXPEERS) return fail(MAMBO_EINVAL, "too many shards for the peer exchange");
    CU(cudaSetDevice(h->device));
    if (!h->peer_recv) {
        const int plen = 1 + 2 * h->N * h->N;
        h->plen_pad = (plen + 15) & ~15;
        const size_t bytes = (size_t)2 * h->nshards * h->plen_pad * sizeof(double) + (size_t)2 * small::MAXPEERS * sizeof(unsigned long long);
        CU(cudaMalloc((void**)&h->peer_recv, bytes));
        CU(cudaMemset(h->peer_recv, 0, bytes));
        h->peer_flags = reinterpret_cast<unsigned long long*>(h->peer_recv + (size_t)2 * h->nshards * h->plen_pad);
        TRY(dalloc(&h->peer_done, 1));
        std::memset(&h->peer_tab, 0, sizeof(h->peer_tab));
        h->peer_tab.recv[h->shard] = h->peer_recv;
        h->peer_tab.flags[h->shard] = h->peer_flags;
        h->peer_attached = 1;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    CU(cudaIpcGetMemHandle(&hd, h->peer_recv));
    std::memcpy(handle64, &hd, 64);
    return MAMBO_OK;
}

static int peer_set(mambo_csa* h, int rank, double* base) {
    if (rank < 0 || rank >= h->nshards) return fail(MAMBO_EINVAL, "peer rank out of range");
    if (!h->peer_recv) return fail(MAMBO_ESTATE, "call mambo_csa_peer_export first");
    if (rank == h->shard) return MAMBO_OK;
    if (!h->peer_tab.recv[rank]) h->peer_attached++;
    h->peer_tab.recv[rank] = base;
    h->peer_tab.flags[rank] = reinterpret_cast<unsigned long long*>(base + (size_t)2 * h->nshards * h->plen_pad);
    return MAMBO_OK;
}

// Map the receive buffer another PROCESS exported (cudaIpcOpenMemHandle; NVLink peer access is enabled lazily).
int mambo_csa_peer_attach(mambo_csa* h, int rank, const unsigned char* handle64) {
    if (!h || !handle64) return fail(MAMBO_EINVAL, "null argument");
    if (rank == h->shard) return MAMBO_OK;
    CU(cudaSetDevice(h->device));
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle64, 64);
    void* base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
    if (rank >= 0 && rank < small::MAXPEERS) h->peer_opened[rank] = base;
    return peer_set(h, rank, static_cast<double*>(base));
}

// Same-process variant: the peer handle lives in this process (tests, single-process multi-GPU drivers).
int mambo_csa_peer_attach_local(mambo_csa* h, int rank, mambo_csa* peer) {
    if (!h || !peer) return fail(MAMBO_EINVAL, "null argument");
    if (!peer->peer_recv) return fail(MAMBO_ESTATE, "peer has not exported its receive buffer");
    if (peer->device != h->device) {
        CU(cudaSetDevice(h->device));
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        cudaGetLastError();
    }
    return peer_set(h, rank, peer->peer_recv);
}

int mambo_csa_eval_sharded_push_device(mambo_csa* h, const double* d_x) {
    if (!h || !d_x) return fail(MAMBO_EINVAL, "null argument");
    if (h->peer_attached != h->nshards) return fail(MAMBO_ESTATE, "peer exchange not set up: export + attach every shard first");
    CU(cudaSetDevice(h->device));
    const long long l0 = h->launches;
    if (d_x != h->d_x) CU(cudaMemcpyAsync(h->d_x, d_x, sizeof(double) * h->L, cudaMemcpyDeviceToDevice, h->stream));
    TRY(enqueue_partial(h, -1, true, false));
    h->peer_epoch++;
    const int plen = 1 + 2 * h->N * h->N;
    small::k_peer_push<<<std::min(64, (plen + 255) / 256), 256, 0, h->stream>>>(h->d_part, plen, h->plen_pad, h->peer_tab, h->nshards, h->shard,
                                                                              h->peer_epoch, h->peer_done);
    h->launches++;
    CU(cudaGetLastError());
    h->launches_last_eval = (int)(h->launches - l0);
    h->memo_valid = false;
    return MAMBO_OK;
}

// The peers' epoch flags are awaited by the STREAM (cuStreamWaitValue64, a front-end memory operation) before the sum kernel is
// released: a kernel that spins on a flag holds its SM, and fused::k_ty needs every SM empty, so with several sharded
// evaluations in flight per GPU (lanes on shard handles) spinning sums of different lanes could wait for each other's hot kernels
// across two GPUs.  The entry point comes from the driver at run time (cudaGetDriverEntryPoint): no link dependency on libcuda.
// k_peer_sum keeps its acquire-polling loop (satisfied at once) as the fall-back when the entry point is unavailable.
typedef CUresult (*stream_wait64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
static stream_wait64_fn stream_wait64() {
    static stream_wait64_fn fn = []() -> stream_wait64_fn {
        if (std::getenv("MAMBO_NO_STREAM_WAIT")) return nullptr;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<stream_wait64_fn>(p);
    }();
    return fn;
}

int mambo_csa_eval_sharded_finish_device(mambo_csa* h, double* d_out) {
    if (!h || !d_out) return fail(MAMBO_EINVAL, "null argument");
    if (h->peer_attached != h->nshards) return fail(MAMBO_ESTATE, "peer exchange not set up");
    CU(cudaSetDevice(h->device));
    const long long l0 = h->launches;
    const int plen = 1 + 2 * h->N * h->N;
    stream_wait64_fn wait = h->peer_stream_wait ? stream_wait64() : nullptr;
    if (wait) {
        const int parity = (int)(h->peer_epoch & 1ull);
        for (int r = 0; r < h->nshards; ++r) {
            const unsigned long long* f = h->peer_flags + (size_t)parity * small::MAXPEERS + r;
            if (wait((CUstream)h->stream, (CUdeviceptr)(uintptr_t)f, (cuuint64_t)h->peer_epoch, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
                return fail(MAMBO_ECUDA, "cuStreamWaitValue64 failed");
        }
    }
    small::k_peer_sum<<<std::min(64, (plen + 255) / 256), 256, 0, h->stream>>>(h->peer_recv, h->peer_flags, plen, h->plen_pad, h->nshards,
                                                                             h->peer_epoch, h->d_part, h->d_status);
    h->launches++;
    TRY(enqueue_finish_seg(h, -1, false));
    if (d_out != h->d_out) CU(cudaMemcpyAsync(d_out, h->d_out, sizeof(double) * (1 + h->L), cudaMemcpyDeviceToDevice, h->stream));
    h->launches_last_eval += (int)(h->launches - l0);
    return MAMBO_OK;
}

// How the finish half waits for the peers: 0 (default) the sum kernel polls the flags and gives up after ~4 s (a dead peer
// becomes MAMBO_ESTATE); 1 the stream waits (cuStreamWaitValue64) and the kernel starts only when every flag is there -- no SM
// is held while waiting, which several sharded evaluations in flight per GPU (lanes of a shard handle) need; a peer that
// never publishes then blocks the stream for good.  Returns MAMBO_ESTATE if the driver offers no stream memory operations.
int mambo_csa_set_peer_wait(mambo_csa* h, int stream_ordered) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    if (stream_ordered && !stream_wait64()) return fail(MAMBO_ESTATE, "cuStreamWaitValue64 is not available from this driver");
    h->peer_stream_wait = stream_ordered != 0;
    return MAMBO_OK;
}

int mambo_csa_eval_sharded_device(mambo_csa* h, const double* d_x, double* d_out) {
    TRY(mambo_csa_eval_sharded_push_device(h, d_x));
    return mambo_csa_eval_sharded_finish_device(h, d_out);
}

// ---- leading eigenpair of the residual (initial guess of a greedy step, guesses.jl:51-84) ---------------------------
namespace {

// number of eigenvalues of the symmetric tridiagonal (a, b) strictly below x (Sturm sequence)
int sturm_count(const std::vector<double>& a, const std::vector<double>& b, int m, double x) {
    int cnt = 0;
    double d = 1.0;
    for (int i = 0; i < m; ++i) {
        const double off = i ? b[i - 1] * b[i - 1] : 0.0;
        d = a[i] - x - (i ? off / d : 0.0);
        if (d == 0.0) d = -1e-300;
        if (d < 0.0) ++cnt;
    }
    return cnt;
}
// k-th smallest eigenvalue (0-based) of the m x m tridiagonal by bisection
double tridiag_eigval(const std::vector<double>& a, const std::vector<double>& b, int m, int k) {
    double lo = a[0], hi = a[0];
    for (int i = 0; i < m; ++i) {
        const double r = (i ? std::fabs(b[i - 1]) : 0.0) + (i + 1 < m ? std::fabs(b[i]) : 0.0);
        lo = std::min(lo, a[i] - r);
        hi = std::max(hi, a[i] + r);
    }
    for (int it = 0; it < 200 && hi - lo > 4e-16 * std::max(std::fabs(lo), std::fabs(hi)); ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (sturm_count(a, b, m, mid) > k) hi = mid;
        else lo = mid;
    }
    return 0.5 * (lo + hi);
}
// unit eigenvector of the tridiagonal for eigenvalue theta: inverse iteration; each solve is Gaussian elimination with
// partial pivoting on the shifted tridiagonal (second super-diagonal kept in dl)
void tridiag_eigvec(const std::vector<double>& a, const std::vector<double>& b, int m, double theta, std::vector<double>& z) {
    z.assign(m, 1.0 / std::sqrt((double)m));
    if (m == 1) { z[0] = 1.0; return; }
    double scale = 0.0;
    for (int i = 0; i < m; ++i) scale = std::max(scale, std::fabs(a[i]) + (i + 1 < m ? std::fabs(b[i]) : 0.0));
    const double shift = theta + 1e-15 * std::max(scale, 1e-300);   // keeps the elimination away from an exact zero pivot
    std::vector<double> d(m), du(m), dl(m);
    auto nz = [](double v) { return v == 0.0 ? 1e-300 : v; };
    for (int rep = 0; rep < 4; ++rep) {
        for (int i = 0; i < m; ++i) { d[i] = a[i] - shift; du[i] = i + 1 < m ? b[i] : 0.0; dl[i] = i + 1 < m ? b[i] : 0.0; }
        for (int i = 0; i + 1 < m; ++i) {
            if (std::fabs(d[i]) >= std::fabs(dl[i])) {
                const double mult = dl[i] / nz(d[i]);
                d[i + 1] -= mult * du[i];
                z[i + 1] -= mult * z[i];
                dl[i] = 0.0;
            } else {
                const double mult = d[i] / dl[i];
                d[i] = dl[i];
                const double tmp = d[i + 1];
                d[i + 1] = du[i] - mult * tmp;
                if (i + 2 < m) {
                    dl[i] = du[i + 1];
                    du[i + 1] = -mult * dl[i];
                } else {
                    dl[i] = 0.0;
                }
                du[i] = tmp;
                const double t = z[i];
                z[i] = z[i + 1];
                z[i + 1] = t - mult * z[i + 1];
            }
        }
        z[m - 1] /= nz(d[m - 1]);
        z[m - 2] = (z[m - 2] - du[m - 2] * z[m - 1]) / nz(d[m - 2]);
        for (int i = m - 3; i >= 0; --i) z[i] = (z[i] - du[i] * z[i + 1] - dl[i] * z[i + 2]) / nz(d[i]);
        double nrm = 0.0, big = 0.0;
        for (double v : z) big = std::max(big, std::fabs(v));
        if (!(big > 0.0) || std::isinf(big) || big != big) { z.assign(m, 0.0); z[0] = 1.0; big = 1.0; }
        for (double& v : z) { v /= big; nrm += v * v; }
        nrm = std::sqrt(nrm);
        for (double& v : z) v /= nrm;
    }
}

}  // namespace

// y = T~ v on the device, in the engine's row space (rows_padded entries each; rows beyond `rows` must be zero in v and
// are left untouched in y): the HBM-bound building block of the Lanczos driver, exported for callers that bring their
// own Krylov method and for bandwidth measurements.
int mambo_csa_symv_device(mambo_csa* h, const double* d_v, double* d_y) {
    if (!h || !d_v || !d_y) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "symv needs an unsharded handle");
    CU(cudaSetDevice(h->device));
    if (h->mode == (int)MAMBO_SYM_GENERAL) lanczos::k_symv<1><<<(h->nR + lanczos::SYMV_RPC - 1) / lanczos::SYMV_RPC, 256, 0, h->stream>>>(h->T2, h->T1, h->ldT, h->nR, d_v, d_y);
    else lanczos::k_symv<0><<<(h->nR + lanczos::SYMV_RPC - 1) / lanczos::SYMV_RPC, 256, 0, h->stream>>>(h->T1, nullptr, h->ldT, h->nR, d_v, d_y);
    h->launches++;
    CU(cudaGetLastError());
    return MAMBO_OK;
}

// Host half of the Lanczos driver on its own (no GPU involved): extremal Ritz value of largest magnitude of the m x m
// tridiagonal (alpha, beta) and its unit eigenvector.  Exposed so that the CPU test-suite can pin it against LAPACK.
int mambo_debug_tridiag_leading(const double* alpha, const double* beta, int m, double* theta, double* z) {
    if (!alpha || !beta || m < 1 || !theta || !z) return fail(MAMBO_EINVAL, "bad arguments");
    std::vector<double> a(alpha, alpha + m), b(beta, beta + m), v;
    const double lo = tridiag_eigval(a, b, m, 0), hi = tridiag_eigval(a, b, m, m - 1);
    *theta = std::fabs(lo) > std::fabs(hi) ? lo : hi;
    tridiag_eigvec(a, b, m, *theta, v);
    std::memcpy(z, v.data(), sizeof(double) * m);
    return MAMBO_OK;
}

// Largest-|eigenvalue| eigenpair of Symmetric(reshape(residual tbt, (N^2, N^2))): Lanczos with full (twice-applied
// classical Gram-Schmidt) re-orthogonalisation on the device-resident T~.  vec: N x N column-major, unit Frobenius norm.
int mambo_csa_leading_eig(mambo_csa* h, double tol, int max_iter, double* eigval, double* vec, int* iterations, double* residual) {
    if (!h || !eigval || !vec) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "leading_eig needs an unsharded handle");
    CU(cudaSetDevice(h->device));
    const int n = h->N, R = h->nR;
    const long long ldv = h->ldT;
    if (!(tol > 0.0)) tol = 1e-13;
    int mmax = max_iter > 0 ? max_iter : 600;
    mmax = std::min(mmax, R);
    cudaStream_t st = h->stream;
    const bool dbg = std::getenv("MAMBO_EIG_DEBUG") != nullptr;
    auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    const double t_begin = now();
    // workspace: V [(mmax+1)][ldv] | w [ldv] | c [mmax+1] | alpha, beta [2 (mmax+1)] | s [mmax+1] | vec [n n]
    // sized for the default Krylov dimension at least, so that a short first call does not force a re-allocation later
    const int mcap = std::max(mmax, std::min(600, R));
    const size_t need = (size_t)(mcap + 2) * ldv + 4 * (size_t)(mcap + 1) + (size_t)n * n;
    if (h->eig_ws_doubles < need) {
        if (h->eig_ws) CU(cudaFree(h->eig_ws));
        h->eig_ws = nullptr;
        h->eig_ws_doubles = 0;
        CU(cudaMalloc((void**)&h->eig_ws, sizeof(double) * need));
        h->eig_ws_doubles = need;
    }
    double* V = h->eig_ws;
    double* w = V + (size_t)(mcap + 1) * ldv;
    double* c = w + ldv;
    double* ab = c + (mcap + 1);
    double* d_s = ab + 2 * (mcap + 1);
    double* d_vec = d_s + (mcap + 1);
    auto cleanup = [&]() {};
#define LCU(call)                          \
    do {                                   \
        cudaError_t le__ = (call);         \
        if (le__ != cudaSuccess) {         \
            cleanup();                     \
            CU(le__);                      \
        }                                  \
    } while (0)
    LCU(cudaMemsetAsync(w, 0, sizeof(double) * ldv, st));
    LCU(cudaMemsetAsync(ab, 0, sizeof(double) * 2 * (mmax + 1), st));
    double* alpha = ab;
    double* beta = ab + (mmax + 1);
    const int vblocks = (R + 255) / 256;
    const int mv_grid = (R + lanczos::SYMV_RPC - 1) / lanczos::SYMV_RPC;   // one CTA per group of rows, scheduled dynamically
    const bool general = (h->mode == (int)MAMBO_SYM_GENERAL);
    lanczos::k_start<<<vblocks, 256, 0, st>>>(w, R, h->pa, h->pq, n, h->packed ? 1 : 0);
    lanczos::k_normalise<<<1, 1024, 0, st>>>(w, R, (int)ldv, V, beta, mmax);   // beta[mmax] is scratch
    h->launches += 2;
    std::vector<double> ha(mmax + 1), hb(mmax + 1), z;
    double theta = 0.0, res = 0.0;
    int m = 0;
    bool done = false;
    const double t_alloc = now() - t_begin;
    double t_enq = 0, t_sync = 0, t_host = 0, tq = now();
    while (!done) {
        tq = now();
        const int stop = std::min(mmax, m + (m < 16 ? 4 : 12));
        for (; m < stop; ++m) {
            const double* vj = V + (size_t)m * ldv;
            if (general) lanczos::k_symv<1><<<mv_grid, 256, 0, st>>>(h->T2, h->T1, h->ldT, R, vj, w);   // T1 holds M^T, T2 holds M (k_pack)
            else lanczos::k_symv<0><<<mv_grid, 256, 0, st>>>(h->T1, nullptr, h->ldT, R, vj, w);
            for (int pass = 0; pass < 2; ++pass) {
                lanczos::k_dots<<<m + 1, 256, 0, st>>>(V, ldv, R, w, c);
                lanczos::k_project<<<vblocks, 256, sizeof(double) * (m + 1), st>>>(V, ldv, R, m + 1, c, w, alpha, m, pass);
            }
            lanczos::k_normalise<<<1, 1024, 0, st>>>(w, R, (int)ldv, V + (size_t)(m + 1) * ldv, beta, m);
            h->launches += 6;
        }
        LCU(cudaGetLastError());
        t_enq += now() - tq; tq = now();
        LCU(cudaMemcpyAsync(ha.data(), alpha, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        LCU(cudaMemcpyAsync(hb.data(), beta, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
        LCU(cudaStreamSynchronize(st));
        t_sync += now() - tq; tq = now();
        // an invariant subspace was reached at step k (beta_k ~ 0): the Krylov space is exhausted there
        double bscale = 0.0;
        for (int i = 0; i < m; ++i) bscale = std::max(bscale, std::max(std::fabs(ha[i]), std::fabs(hb[i])));
        int meff = m;
        for (int i = 0; i < m; ++i)
            if (!(hb[i] > 1e-14 * std::max(bscale, 1e-300))) { meff = i + 1; break; }
        const double lo = tridiag_eigval(ha, hb, meff, 0), hi = tridiag_eigval(ha, hb, meff, meff - 1);
        theta = std::fabs(lo) > std::fabs(hi) ? lo : hi;       // sortperm(abs.(L))[end] (guesses.jl:61)
        tridiag_eigvec(ha, hb, meff, theta, z);
        res = std::fabs(hb[meff - 1] * z[meff - 1]);
        if (meff < m || res <= tol * std::max(std::fabs(theta), 1e-300) || m >= mmax || theta == 0.0) {
            if (meff < m) res = 0.0;
            m = meff;
            done = true;
        }
        t_host += now() - tq;
    }
    const double t_loop_end = now();
    LCU(cudaMemcpyAsync(d_s, z.data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
    LCU(cudaMemsetAsync(d_vec, 0, sizeof(double) * (size_t)n * n, st));
    lanczos::k_combine<<<vblocks, 256, sizeof(double) * m, st>>>(V, ldv, R, m, d_s, h->pa, h->pq, h->sw, n, h->packed ? 1 : 0, d_vec);
    h->launches++;
    LCU(cudaGetLastError());
    LCU(cudaMemcpyAsync(vec, d_vec, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, st));
    LCU(cudaStreamSynchronize(st));
#undef LCU
    if (dbg)
        fprintf(stderr, "[mambo eig] m=%d setup %.3f ms, enqueue %.3f ms, sync %.3f ms, host tridiagonal %.3f ms, Ritz vector + D2H %.3f ms\n", m,
                t_alloc * 1e3, t_enq * 1e3, t_sync * 1e3, t_host * 1e3, (now() - t_loop_end) * 1e3);
    // fix the sign so that the result does not depend on the start vector: largest-magnitude entry positive
    {
        size_t big = 0;
        for (size_t i = 1; i < (size_t)n * n; ++i)
            if (std::fabs(vec[i]) > std::fabs(vec[big]) * (1.0 + 1e-12)) big = i;
        if (vec[big] < 0.0)
            for (size_t i = 0; i < (size_t)n * n; ++i) vec[i] = -vec[i];
    }
    *eigval = theta;
    if (iterations) *iterations = m;
    if (residual) *residual = res;
    return MAMBO_OK;
}

int mambo_csa_subtract_fragment(mambo_csa* h, const double* x, double* new_cost) {
    if (!h || !x) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    const int s_host = host_s_for(h, x);
    if (s_host < -1 || s_host > small::SMAX) return fail(MAMBO_ERANGE, "||K||_1 exceeds 2^14 (or x contains NaN)");
    std::memcpy(h->h_x, x, sizeof(double) * h->L);
    CU(cudaMemcpyAsync(h->d_x, h->h_x, sizeof(double) * h->L, cudaMemcpyHostToDevice, h->stream));
    TRY(subtract_impl(h, h->d_x, s_host));
    CU(cudaMemcpyAsync(h->h_out, h->d_part, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->memo_valid = false;
    h->sh->version.fetch_add(1);
    if (new_cost) *new_cost = h->h_out[0];
    return MAMBO_OK;
}

int mambo_csa_residual_cost(mambo_csa* h, double* cost) {
    if (!h || !cost) return fail(MAMBO_EINVAL, "null argument");
    CU(cudaSetDevice(h->device));
    // L2_partial_cost(F_rem): evaluate the residual against the zero fragment (lambda = 0 => W = 0), theta = 0
    CU(cudaMemsetAsync(h->d_x, 0, sizeof(double) * h->L, h->stream));
    h->memo_valid = false;
    TRY(enqueue_partial(h, 0, false, false));
    CU(cudaMemcpyAsync(h->h_out, h->d_part, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *cost = h->h_out[0];
    return MAMBO_OK;
}

int mambo_csa_get_residual(mambo_csa* h, double* tbt_out, double* obt_out) {
    if (!h) return fail(MAMBO_EINVAL, "null handle");
    CU(cudaSetDevice(h->device));
    const int n = h->N;
    if (tbt_out) {
        const size_t n4 = (size_t)n * n * n * n;
        double* tmp = nullptr;
        CU(cudaMalloc((void**)&tmp, n4 * sizeof(double)));
        small::k_unpack<<<h->num_sms * 16, 256, 0, h->stream>>>(h->T1, n, h->packed ? 1 : 0, h->sw, h->row_begin, std::min(h->row_end, h->nR),
                                                              h->ldT, tmp);
        cudaError_t e = cudaMemcpyAsync(tbt_out, tmp, n4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(tmp);
        CU(e);
    }
    if (obt_out && h->flavour == MAMBO_CSA_SD) {
        CU(cudaMemcpyAsync(obt_out, h->hT, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return MAMBO_OK;
}

// ---- final-norm pieces (lcu.jl:139-208, 262-267; fermionic.jl:457-479) ----------------------------------------------------
// The one-body matrix whose eigenvalues give one_body_L1 (lcu.jl:262-267): H.mbts[2] + ob_correction(H) of the RESIDENT residual,
// ob_correction = (spin_orb ? 1 : 2) sum_r tbt[:,:,r,r].  CSA handles hold no one-body tensor: obt_in (N x N column-major, may be
// NULL) is added instead; CSA_SD handles add their own resident one-body residual.  out: N x N column-major.  The N x N `eigen`
// of to_OBF (fermionic.jl:527-538) stays with the caller, like the other N x N eigenproblems of the path.
int mambo_csa_one_body_matrix(mambo_csa* h, int spin_orb, const double* obt_in, double* out) {
    if (!h || !out) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "one_body_matrix needs an unsharded handle");
    CU(cudaSetDevice(h->device));
    const int n = h->N, nn = n * n;
    const double* Tt = h->T1;                 // T1[pair(r,s)][pair(p,q)] = tbt[p,q,r,s] in every mode (see k_unpack)
    const double* add = nullptr;
    if (h->flavour == MAMBO_CSA_SD) add = h->hT;
    else if (obt_in) {
        CU(cudaMemcpyAsync(h->Dh, obt_in, sizeof(double) * nn, cudaMemcpyHostToDevice, h->stream));
        add = h->Dh;
    }
    CU(cudaMemsetAsync(h->DU, 0, sizeof(double) * nn, h->stream));
    small::k_ob_correction<<<(h->nR + 255) / 256, 256, 0, h->stream>>>(Tt, n, h->packed ? 1 : 0, h->sw, h->pa, h->pq, h->nR, h->ldT,
                                                                       spin_orb ? 1.0 : 2.0, add, h->DU);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, h->DU, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->memo_valid = false;       // Dh / DU are scratch of the evaluation as well
    return MAMBO_OK;
}

// CSA_L1 of one fragment (lcu.jl:139-182): a weighted abs-sum over the fragment's two-body Cartan coefficients -- O(N^2) host
// arithmetic on x, no tensor work.  flavour picks where lambda starts in x (CSA: 0, CSA_SD: N).
int mambo_csa_frag_l1(int N, int flavour, int spin_orb, const double* x, double* l1) {
    if (N < 1 || !x || !l1) return fail(MAMBO_EINVAL, "bad arguments");
    const double* lam = x + (flavour == MAMBO_CSA_SD ? N : 0);
    double s = 0.0;
    int idx = 0;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j <= i; ++j, ++idx) {
            if (spin_orb) {
                if (i != j) s += 0.5 * std::fabs(lam[idx]);
            } else {
                s += (i != j ? 2.0 : 0.5) * std::fabs(lam[idx]);
            }
        }
    *l1 = s;
    return MAMBO_OK;
}

// ---- checkpoint block (decompose.jl:58-64): the dataset "<group>/x" = tot_L x alpha Float64, column-major, as a raw file ----
// Layout: 8 bytes magic "MAMBOX01", 16 bytes group name (zero padded), int64 tot_L, int64 alpha, then tot_L*alpha doubles.
// There is no libhdf5 in this image; INTEGRATION.md shows the ten Julia lines that turn the file into the HDF5 dataset.
int mambo_csa_write_x_block(const char* path, const char* group, long long tot_L, long long alpha, const double* X) {
    if (!path || !group || tot_L < 1 || alpha < 0 || (!X && alpha > 0)) return fail(MAMBO_EINVAL, "bad arguments");
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) return fail(MAMBO_EINVAL, std::string("cannot open ") + tmp);
    char name[16] = {0};
    std::strncpy(name, group, 15);
    bool ok = std::fwrite("MAMBOX01", 1, 8, f) == 8 && std::fwrite(name, 1, 16, f) == 16 && std::fwrite(&tot_L, 8, 1, f) == 1 &&
              std::fwrite(&alpha, 8, 1, f) == 1;
    const size_t cnt = (size_t)tot_L * (size_t)alpha;
    ok = ok && (cnt == 0 || std::fwrite(X, sizeof(double), cnt, f) == cnt);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok || std::rename(tmp.c_str(), path) != 0) {    // written beside the target and renamed: a crash never leaves half a checkpoint
        std::remove(tmp.c_str());
        return fail(MAMBO_EINVAL, std::string("cannot write ") + path);
    }
    return MAMBO_OK;
}
// X == NULL: only the header is read (tot_L, alpha, group); otherwise X must hold tot_L*alpha doubles (cap = its capacity).
int mambo_csa_read_x_block(const char* path, char* group16, long long* tot_L, long long* alpha, double* X, long long cap) {
    if (!path || !tot_L || !alpha) return fail(MAMBO_EINVAL, "bad arguments");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(MAMBO_EINVAL, std::string("cannot open ") + path);
    char magic[8], name[16];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "MAMBOX01", 8) == 0 && std::fread(name, 1, 16, f) == 16 &&
              std::fread(tot_L, 8, 1, f) == 1 && std::fread(alpha, 8, 1, f) == 1 && *tot_L >= 1 && *alpha >= 0;
    if (ok && group16) std::memcpy(group16, name, 16);
    if (ok && X) {
        const long long cnt = *tot_L * *alpha;
        ok = cnt <= cap && (cnt == 0 || std::fread(X, sizeof(double), (size_t)cnt, f) == (size_t)cnt);
    }
    std::fclose(f);
    if (!ok) return fail(MAMBO_EINVAL, std::string("not a valid x block: ") + path);
    return MAMBO_OK;
}

int mambo_csa_reconstruct(mambo_csa* h, const double* x, double* tbt_out, double* obt_out) {
    if (!h || !x) return fail(MAMBO_EINVAL, "null argument");
    if (h->nshards != 1) return fail(MAMBO_ESTATE, "reconstruct needs an unsharded handle");
    CU(cudaSetDevice(h->device));
    const int n = h->N, nn = n * n;
    const int s_host = host_s_for(h, x);
    if (s_host < -1 || s_host > small::SMAX) return fail(MAMBO_ERANGE, "||K||_1 exceeds 2^14 (or x contains NaN)");
    std::memcpy(h->h_x, x, sizeof(double) * h->L);
    CU(cudaMemcpyAsync(h->d_x, h->h_x, sizeof(double) * h->L, cudaMemcpyHostToDevice, h->stream));
    TRY(enqueue_unitary(h, h->d_x, s_host));
    TRY(enqueue_prep(h, PREP_AB));
    h->memo_valid = false;
    if (tbt_out) {
        const size_t n4 = (size_t)nn * nn;
        double* tmp = nullptr;
        CU(cudaMalloc((void**)&tmp, n4 * sizeof(double)));
        small::k_reconstruct<<<h->num_sms * 16, 256, 0, h->stream>>>(h->At, h->Bt, n, h->ld, h->packed ? 1 : 0, h->sw, tmp);
        cudaError_t e = cudaMemcpyAsync(tbt_out, tmp, n4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(tmp);
        CU(e);
    }
    if (obt_out && h->flavour == MAMBO_CSA_SD) {
        // h = U diag(l1) U^T: reuse the one-body kernel against the stored target, keep h in hfrag
        small::k_sd_onebody<<<1, 1024, 0, h->stream>>>(h->d_x, h->hT, h->U, h->d_zero, n, nullptr, nullptr, h->Dh, h->DU, h->hfrag);
        small::k_obt_colmajor<<<(nn + 127) / 128, 128, 0, h->stream>>>(h->hfrag, n, h->Dh);
        CU(cudaMemcpyAsync(obt_out, h->Dh, sizeof(double) * nn, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return MAMBO_OK;
}

// Development aid (not part of the public header): copy one of the handle's internal device buffers to the host after a
// synchronise, so that repeated evaluations can be compared buffer by buffer.  which: 0 U, 1 A~, 2 Y pieces, 3 E~ panels (column-
// split launches), 4 F, 5 Spart, 6 part, 7 out, 8 aux.  Returns the number of doubles of the buffer (copies min(count, that)).
long long mambo_debug_fetch(mambo_csa* h, int which, double* out, long long count) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->side_stream) cudaStreamSynchronize(h->side_stream);
    const long long n = h->N;
    const double* src = nullptr;
    long long len = 0;
    switch (which) {
        case 0: src = h->U; len = n * n; break;
        case 1: src = h->At; len = (long long)h->nRpad * h->ld; break;
        case 2: src = h->Epiece; len = (long long)h->visits * 2 * ROWS * h->Np; break;
        case 3: src = h->EtP; len = h->EtP ? (long long)h->P * ROWS * h->ld : 0; break;
        case 4: src = h->F1; len = (long long)h->P * ROWS * h->Np; break;
        case 5: src = h->Spart; len = (long long)h->P * n * n; break;
        case 6: src = h->d_part; len = 1 + 2 * n * n; break;
        case 7: src = h->d_out; len = 1 + h->L; break;
        case 8: src = h->aux; len = h->aux ? 3 * n + 1 : 0; break;
        default: return -1;
    }
    if (out && src && len > 0) cudaMemcpy(out, src, sizeof(double) * (size_t)std::min(count, len), cudaMemcpyDeviceToHost);
    return len;
}

}  // extern "C"
