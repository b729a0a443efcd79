// df.cu -- double factorisation of the device-resident residual: host driver of df_kernels.cuh (C-ABI: include/mambo_csa.h).
// Replaces DF_decomposition (/root/reference/src/UTILS/decompose.jl:297-398) and DF_based_greedy (:400-436).
// There is no CPU code path for the numerics: the host only sorts eigenvalues and steers the block structure.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "internal.h"
#include "lanczos_kernels.cuh"
#include "df_kernels.cuh"

namespace {

int dfail(int code, const std::string& msg) {
    mambo_set_error(msg.c_str());
    return code;
}
#define DCU(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess)                                                                                     \
            return dfail(e__ == cudaErrorMemoryAllocation ? MAMBO_ENOMEM : MAMBO_ECUDA,                             \
                         std::string(#call) + ": " + cudaGetErrorString(e__) + " (df.cu:" + std::to_string(__LINE__) + ")"); \
    } while (0)

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct DfState {
    int device = 0, N = 0, nfrag = 0;
    bool valid = false;               // a decomposition has completed on this handle (it may hold zero fragments)
    double *d_coeff = nullptr, *d_omega = nullptr, *d_U = nullptr, *d_L = nullptr;
    std::vector<double> coeff;
    std::vector<int> kind;
    void release() {
        cudaSetDevice(device);
        for (double** p : {&d_coeff, &d_omega, &d_U, &d_L})
            if (*p) { cudaFree(*p); *p = nullptr; }
        nfrag = 0;
        valid = false;
        coeff.clear();
        kind.clear();
    }
};
void df_destroy(void* p) {
    DfState* s = static_cast<DfState*>(p);
    if (!s) return;
    s->release();
    delete s;
}

// every temporary device buffer of one decomposition; freed on every exit path
struct Scratch {
    std::vector<void*> bufs;
    ~Scratch() {
        for (void* b : bufs)
            if (b) cudaFree(b);
    }
    cudaError_t get(void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
        if (e == cudaSuccess) bufs.push_back(*p);
        return e;
    }
};

// Opt-in dynamic shared memory is only ever RAISED (per device): handles of different N on one GPU share these kernels.
template <class K>
cudaError_t raise_smem(K kernel, size_t bytes) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, (const void*)kernel);
    if (e != cudaSuccess) return e;
    if ((size_t)a.maxDynamicSharedSizeBytes >= bytes) return cudaSuccess;
    return cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// opt-in dynamic shared memory of the GEMM instantiations: a per-device attribute, so it is (re)set for the current device on
// every entry (a few microseconds) rather than remembered in a process-wide flag
cudaError_t gemm_setup() {
    cudaError_t e;
    const int bytes = (int)df::GEMM_SMEM;
    if ((e = cudaFuncSetAttribute(df::k_dgemm<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(df::k_dgemm<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(df::k_dgemm<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
template <bool TA, int EPI>
void gemm(cudaStream_t st, int M, int N, int K, const double* A, long long lda, const double* B, long long ldb, double* C, long long ldc) {
    dim3 grid((N + df::GBN - 1) / df::GBN, (M + df::GBM - 1) / df::GBM);
    df::k_dgemm<TA, EPI><<<grid, 256, df::GEMM_SMEM, st>>>(M, N, K, A, lda, B, ldb, C, ldc);
}

__global__ void __launch_bounds__(256) k_eig_residual(const double* __restrict__ Ty, const double* __restrict__ y, double theta, int len,
                                                      double* __restrict__ out) {
    __shared__ double red[8];
    double a = 0.0;
    for (int i = threadIdx.x; i < len; i += 256) {
        const double r = Ty[i] - theta * y[i];
        a = fma(r, r, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += red[q];
        out[0] = sqrt(s);
    }
}

// ---- eigenvectors of numerically degenerate clusters of the tridiagonal (host) -------------------------------------------------
// k_twisted takes one step of inverse iteration from the twist index: two eigenvalues of one unreduced block that agree to
// rounding (a degenerate eigenvalue of the tensor found twice in one block once the recurrence runs on rounding noise --
// molecular integrals: the pi shells, rank deficiency) get the SAME vector.  Such clusters are rare and short, so they are
// redone here the way LAPACK's dstein does it: inverse iteration on (T - sigma) with partial pivoting, shifts of a cluster
// spread by a few ulps of |T|, pseudo-random start, modified Gram-Schmidt against the cluster's earlier vectors in every
// iteration.  The vectors span the cluster's invariant subspace to working precision; which basis of it is immaterial (the
// Newton-Schulz step mixes inside close clusters anyway, and neither residuals nor the sum of the fragments change).
struct TriLU {          // (T - sigma I) = P L U of a symmetric tridiagonal, row interchanges as in dgttrf
    std::vector<double> d, du, du2, l;
    std::vector<char> swp;
};
void tri_factor(const double* a, const double* b, int n, double sigma, double tiny, TriLU& f) {
    f.d.assign(n, 0.0); f.du.assign(n, 0.0); f.du2.assign(n, 0.0); f.l.assign(n, 0.0); f.swp.assign(n, 0);
    for (int i = 0; i < n; ++i) f.d[i] = a[i] - sigma;
    for (int i = 0; i + 1 < n; ++i) f.du[i] = b[i];
    for (int i = 0; i + 1 < n; ++i) {
        const double sub = b[i];
        if (std::fabs(f.d[i]) >= std::fabs(sub)) {
            if (std::fabs(f.d[i]) < tiny) f.d[i] = f.d[i] < 0.0 ? -tiny : tiny;
            const double m = sub / f.d[i];
            f.l[i] = m;
            f.d[i + 1] -= m * f.du[i];
        } else {                                        // interchange rows i and i + 1
            const double m = f.d[i] / sub;
            f.swp[i] = 1;
            f.l[i] = m;
            f.d[i] = sub;
            const double t = f.du[i];
            f.du[i] = f.d[i + 1];
            f.d[i + 1] = t - m * f.d[i + 1];
            if (i + 2 < n) {
                f.du2[i] = f.du[i + 1];
                f.du[i + 1] = -m * f.du[i + 1];
            }
        }
    }
    if (std::fabs(f.d[n - 1]) < tiny) f.d[n - 1] = f.d[n - 1] < 0.0 ? -tiny : tiny;
}
void tri_solve(const TriLU& f, int n, std::vector<double>& x) {   // x <- (T - sigma)^-1 x
    for (int i = 0; i + 1 < n; ++i) {
        if (f.swp[i]) std::swap(x[i], x[i + 1]);
        x[i + 1] -= f.l[i] * x[i];
    }
    x[n - 1] /= f.d[n - 1];
    if (n > 1) x[n - 2] = (x[n - 2] - f.du[n - 2] * x[n - 1]) / f.d[n - 2];
    for (int i = n - 3; i >= 0; --i) x[i] = (x[i] - f.du[i] * x[i + 1] - f.du2[i] * x[i + 2]) / f.d[i];
}
// a[0..n), b[0..n-1): one unreduced block; theta[0..k) ascending: a cluster of its eigenvalues; Z: k vectors of length n.
void tridiag_cluster_vectors(const double* a, const double* b, int n, const double* theta, int k, double tnorm, std::vector<double>& Z) {
    const double eps = 2.220446049250313e-16;
    const double tiny = eps * std::max(tnorm, 1e-300);
    Z.assign((size_t)k * n, 0.0);
    std::vector<double> x(n);
    TriLU f;
    double sig_prev = 0.0;
    unsigned long long rng = 0x9E3779B97F4A7C15ull;
    for (int t = 0; t < k; ++t) {
        double sigma = theta[t];
        if (t > 0 && sigma - sig_prev < 10.0 * tiny) sigma = sig_prev + 10.0 * tiny;   // dstein's spread of coincident shifts
        sig_prev = sigma;
        tri_factor(a, b, n, sigma, tiny, f);
        for (int i = 0; i < n; ++i) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            x[i] = ((double)(rng >> 11) / 9007199254740992.0) - 0.5;
        }
        for (int it = 0; it < 6; ++it) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s = std::max(s, std::fabs(x[i]));
            s = (s > 0.0 ? 1.0 / s : 1.0);
            for (int i = 0; i < n; ++i) x[i] *= s;
            tri_solve(f, n, x);
            for (int pass = 0; pass < 2; ++pass)
                for (int u = 0; u < t; ++u) {
                    const double* zu = Z.data() + (size_t)u * n;
                    double dot = 0.0;
                    for (int i = 0; i < n; ++i) dot = std::fma(zu[i], x[i], dot);
                    for (int i = 0; i < n; ++i) x[i] -= dot * zu[i];
                }
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) nrm = std::fma(x[i], x[i], nrm);
            nrm = std::sqrt(nrm);
            if (!(nrm > 0.0)) {                          // annihilated by the projection: another start
                for (int i = 0; i < n; ++i) {
                    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                    x[i] = ((double)(rng >> 11) / 9007199254740992.0) - 0.5;
                }
                continue;
            }
            for (int i = 0; i < n; ++i) x[i] /= nrm;
        }
        int imax = 0;
        for (int i = 1; i < n; ++i)
            if (std::fabs(x[i]) > std::fabs(x[imax])) imax = i;
        const double sgn = x[imax] < 0.0 ? -1.0 : 1.0;
        for (int i = 0; i < n; ++i) Z[(size_t)t * n + i] = sgn * x[i];
    }
}
// Groups of kept eigenvalues of one block whose neighbours are closer than ctol: indices into sel (positions 0..ne).
void find_tight_clusters(const std::vector<double>& sel_theta, const std::vector<int>& sel_blk, int nblk, double ctol,
                         std::vector<std::vector<int>>& clusters) {
    clusters.clear();
    std::vector<std::vector<int>> per(nblk);
    for (int i = 0; i < (int)sel_theta.size(); ++i) per[sel_blk[i]].push_back(i);
    for (auto& v : per) {
        std::sort(v.begin(), v.end(), [&](int x, int y) { return sel_theta[x] != sel_theta[y] ? sel_theta[x] < sel_theta[y] : x < y; });
        size_t i = 0;
        while (i < v.size()) {
            size_t j = i;
            while (j + 1 < v.size() && sel_theta[v[j + 1]] - sel_theta[v[j]] <= ctol) ++j;
            if (j > i) clusters.emplace_back(v.begin() + i, v.begin() + j + 1);
            i = j + 1;
        }
    }
}

}  // namespace

extern "C" {

// Host half of the double factorisation on its own (no GPU): unit eigenvectors of a cluster of k (nearly) coincident eigenvalues
// theta[0..k) (ascending) of the unreduced m x m tridiagonal (alpha, beta[0..m-1)); Z receives k rows of m doubles.  Exposed so
// that the CPU test-suite can pin it against LAPACK.
int mambo_debug_tridiag_cluster(const double* alpha, const double* beta, int m, const double* theta, int k, double* Z) {
    if (!alpha || !beta || !theta || !Z || m < 1 || k < 1 || k > m) return dfail(MAMBO_EINVAL, "bad arguments");
    double tn = 0.0;
    for (int i = 0; i < m; ++i) tn = std::max(tn, std::fabs(alpha[i]) + (i > 0 ? std::fabs(beta[i - 1]) : 0.0) + (i + 1 < m ? std::fabs(beta[i]) : 0.0));
    std::vector<double> z;
    tridiag_cluster_vectors(alpha, beta, m, theta, k, tn, z);
    std::memcpy(Z, z.data(), sizeof(double) * (size_t)k * m);
    return MAMBO_OK;
}

int mambo_dgemm_device(int device, void* stream, int trans_a, int M, int N, int K, const double* d_A, long long lda, const double* d_B,
                       long long ldb, double* d_C, long long ldc) {
    if (M < 1 || N < 1 || K < 1 || !d_A || !d_B || !d_C) return dfail(MAMBO_EINVAL, "bad arguments");
    if ((lda & 1) || (ldb & 1) || (ldc & 1) || ((uintptr_t)d_A & 15) || ((uintptr_t)d_B & 15) || ((uintptr_t)d_C & 15))
        return dfail(MAMBO_EINVAL, "leading dimensions must be even and the pointers 16-byte aligned");
    DCU(cudaSetDevice(device));
    DCU(gemm_setup());
    cudaStream_t st = (cudaStream_t)stream;
    if (trans_a) gemm<true, 0>(st, M, N, K, d_A, lda, d_B, ldb, d_C, ldc);
    else gemm<false, 0>(st, M, N, K, d_A, lda, d_B, ldb, d_C, ldc);
    DCU(cudaGetLastError());
    return MAMBO_OK;
}

int mambo_csa_df_decompose(mambo_csa* h, double tol, int max_frags, int verify, int* num_frags, mambo_df_stats_t* stats) {
    if (!h) return dfail(MAMBO_EINVAL, "null handle");
    mambo_resident_view v;
    int rc = mambo_csa_resident_view(h, &v);
    if (rc) return rc;
    // decompose.jl:317-322: a matrix view that is not symmetric goes through a general (complex) `eigen` in the reference; that
    // branch is not reproduced here (every input of the path is symmetric: ferm_utils.py:471-507, bliss.jl:17-26, residuals)
    if (v.mode == (int)MAMBO_SYM_GENERAL && v.asym_pair > 1e-12)
        return dfail(MAMBO_EINVAL, "double factorisation needs a (pq)<->(rs)-symmetric two-body tensor");
    DCU(cudaSetDevice(v.device));
    DCU(gemm_setup());
    if (!(tol > 0.0)) tol = 1e-8;                       // SVD_tol, config.jl:22
    const int n = v.N, R = v.rows;
    const long long ldv = v.ldT;
    cudaStream_t st = v.stream;
    const bool general = (v.mode == (int)MAMBO_SYM_GENERAL);
    void** slot = mambo_csa_df_slot(h, df_destroy);
    if (!*slot) *slot = new DfState();
    DfState* S = static_cast<DfState*>(*slot);
    S->device = v.device;
    S->N = n;
    S->release();
    mambo_df_stats_t sts;
    std::memset(&sts, 0, sizeof(sts));
    long long launches = 0;

    // ---- 1. Lanczos tridiagonalisation, full re-orthogonalisation, restart on breakdown ----------------------------------
    Scratch scr;
    const int KSMAX = 32;
    double *V = nullptr, *w = nullptr, *part = nullptr, *cvec = nullptr, *d_alpha = nullptr, *d_beta = nullptr, *d_rnorm = nullptr;
    DCU(scr.get((void**)&V, sizeof(double) * (size_t)(R + 1) * ldv));
    DCU(scr.get((void**)&w, sizeof(double) * ldv));
    DCU(scr.get((void**)&part, sizeof(double) * (size_t)KSMAX * ldv));
    DCU(scr.get((void**)&cvec, sizeof(double) * (R + 2)));
    DCU(scr.get((void**)&d_alpha, sizeof(double) * (R + 2)));
    DCU(scr.get((void**)&d_beta, sizeof(double) * (R + 2)));
    DCU(scr.get((void**)&d_rnorm, sizeof(double) * 2));
    double* npart = nullptr;
    DCU(scr.get((void**)&npart, sizeof(double) * ((size_t)(R + 255) / 256 + 1)));
    DCU(cudaMemsetAsync(d_alpha, 0, sizeof(double) * (R + 2), st));
    DCU(cudaMemsetAsync(d_beta, 0, sizeof(double) * (R + 2), st));
    const int vblocks = (R + 255) / 256;
    const int mv_grid = (R + lanczos::SYMV_RPC - 1) / lanczos::SYMV_RPC;
    const double t0 = now_s();
    // orthogonalise `vec` (ldv doubles, updated in place) against V[0..nv) and normalise it into `dst`
    auto reorth = [&](double* vec, int nv, double* dst, int j, int mode) {
        const int KS = std::max(1, std::min(KSMAX, nv / 16));
        if (nv > 0) {
            lanczos::k_dots<<<nv, 256, 0, st>>>(V, ldv, R, vec, cvec);
            df::k_project_split<<<dim3(vblocks, KS), 256, 0, st>>>(V, ldv, R, nv, cvec, part, (int)ldv);
            launches += 2;
        }
        df::k_apply_part<<<vblocks, 256, 0, st>>>(vec, R, (int)ldv, part, nv > 0 ? KS : 0, npart);
        df::k_scale_norm<<<(int)((ldv + 255) / 256), 256, 0, st>>>(vec, R, (int)ldv, npart, vblocks, dst, cvec, d_alpha, d_beta, j, mode, d_rnorm);
        launches += 2;
    };
    lanczos::k_start<<<vblocks, 256, 0, st>>>(w, R, v.pa, v.pq, n, v.packed);
    reorth(w, 0, V, 0, 2);
    launches += 1;
    std::vector<double> ha(R + 1, 0.0), hb(R + 1, 0.0);
    std::vector<int> block_start(1, 0);
    int m = 0;
    double scale = 0.0;
    bool finished = false;
    unsigned long long seed = 1;
    while (!finished && m < R) {
        const int stop = std::min(R, m + 32);
        for (int j = m; j < stop; ++j) {
            const double* vj = V + (size_t)j * ldv;
            double* vn = V + (size_t)(j + 1) * ldv;
            if (general) lanczos::k_symv<1><<<mv_grid, 256, 0, st>>>(v.T2, v.T1, v.ldT, R, vj, w);
            else lanczos::k_symv<0><<<mv_grid, 256, 0, st>>>(v.T1, nullptr, v.ldT, R, vj, w);
            launches += 1;
            reorth(w, j + 1, vn, j, 0);      // alpha_j, beta_j, v_{j+1} = w / beta_j
            reorth(vn, j + 1, vn, j, 1);     // second pass on the unit vector: orthogonal to working precision whatever beta_j is
        }
        DCU(cudaGetLastError());
        DCU(cudaMemcpyAsync(ha.data() + m, d_alpha + m, sizeof(double) * (stop - m), cudaMemcpyDeviceToHost, st));
        DCU(cudaMemcpyAsync(hb.data() + m, d_beta + m, sizeof(double) * (stop - m), cudaMemcpyDeviceToHost, st));
        DCU(cudaStreamSynchronize(st));
        int brk = -1;
        for (int j = m; j < stop; ++j) {
            if (!(ha[j] == ha[j]) || !(hb[j] == hb[j])) return dfail(MAMBO_ERANGE, "NaN in the Lanczos recurrence (NaN in the residual?)");
            scale = std::max(scale, std::fabs(ha[j]));
            if (j == R - 1 || !(hb[j] > 1e-13 * scale)) { brk = j; break; }
            scale = std::max(scale, hb[j]);
        }
        if (brk < 0) { m = stop; continue; }
        // the Krylov space is exhausted at row brk: beta_brk := 0 closes the block
        hb[brk] = 0.0;
        m = brk + 1;
        if (m >= R) break;
        const int blen = m - block_start.back();
        if (blen == 1 && std::fabs(ha[brk]) <= 1e-12 * scale && block_start.size() > 1) break;   // a random vector of the complement is annihilated
        if (scale == 0.0) break;                                                                 // zero tensor
        // restart: a pseudo-random vector, orthogonalised twice against everything found so far
        double* vn = V + (size_t)m * ldv;
        df::k_random_vec<<<(int)((ldv + 255) / 256), 256, 0, st>>>(w, R, (int)ldv, seed++);
        launches += 1;
        reorth(w, m, vn, 0, 2);
        reorth(vn, m, vn, 0, 2);
        double rn[2];
        DCU(cudaMemcpyAsync(rn, d_rnorm, sizeof(double), cudaMemcpyDeviceToHost, st));
        DCU(cudaStreamSynchronize(st));
        if (!(rn[0] > 0.5)) return dfail(MAMBO_ESTATE, "Lanczos restart vector lost its norm");
        block_start.push_back(m);
        sts.restarts++;
    }
    DCU(cudaStreamSynchronize(st));
    sts.krylov_dim = m;
    sts.seconds_lanczos = now_s() - t0;
    if (m < 1) return dfail(MAMBO_ESTATE, "empty Krylov space");

    // ---- 2. the tridiagonal: eigenvalues by bisection, eigenvectors by twisted factorisation --------------------------------
    const double t1 = now_s();
    const int nblk = (int)block_start.size();
    std::vector<int> blk(3 * nblk);
    std::vector<double> glo(nblk), ghi(nblk);
    double maxb2 = 0.0;
    for (int b = 0; b < nblk; ++b) {
        const int r0 = block_start[b], r1 = b + 1 < nblk ? block_start[b + 1] : m;
        blk[3 * b] = r0; blk[3 * b + 1] = r1 - r0; blk[3 * b + 2] = r0;
        double lo = ha[r0], hi = ha[r0];
        for (int i = r0; i < r1; ++i) {
            const double rad = (i > r0 ? std::fabs(hb[i - 1]) : 0.0) + (i + 1 < r1 ? std::fabs(hb[i]) : 0.0);
            lo = std::min(lo, ha[i] - rad);
            hi = std::max(hi, ha[i] + rad);
            if (i + 1 < r1) maxb2 = std::max(maxb2, hb[i] * hb[i]);
        }
        const double pad = 1e-14 * std::max(std::fabs(lo), std::fabs(hi)) + 1e-300;
        glo[b] = lo - pad; ghi[b] = hi + pad;
    }
    hb[m - 1] = 0.0;
    const double pivmin = 1e-290 * std::max(1.0, maxb2);
    int* d_blk = nullptr;
    double *d_glo = nullptr, *d_ghi = nullptr, *d_theta = nullptr;
    DCU(scr.get((void**)&d_blk, sizeof(int) * 3 * nblk));
    DCU(scr.get((void**)&d_glo, sizeof(double) * nblk));
    DCU(scr.get((void**)&d_ghi, sizeof(double) * nblk));
    DCU(scr.get((void**)&d_theta, sizeof(double) * (m + 1)));
    DCU(cudaMemcpyAsync(d_blk, blk.data(), sizeof(int) * 3 * nblk, cudaMemcpyHostToDevice, st));
    DCU(cudaMemcpyAsync(d_glo, glo.data(), sizeof(double) * nblk, cudaMemcpyHostToDevice, st));
    DCU(cudaMemcpyAsync(d_ghi, ghi.data(), sizeof(double) * nblk, cudaMemcpyHostToDevice, st));
    DCU(cudaMemcpyAsync(d_alpha, ha.data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));   // with the closed blocks (beta = 0)
    DCU(cudaMemcpyAsync(d_beta, hb.data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
    df::k_bisect<<<(m + 63) / 64, 64, 0, st>>>(d_alpha, d_beta, m, d_blk, nblk, d_glo, d_ghi, pivmin, d_theta);
    launches += 1;
    std::vector<double> theta(m);
    DCU(cudaMemcpyAsync(theta.data(), d_theta, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    DCU(cudaStreamSynchronize(st));
    // the reference's order and truncation (decompose.jl:332-350): decreasing |eigenvalue|, cut at the first one below tol
    std::vector<int> order;
    order.reserve(m);
    for (int e = 0; e < m; ++e)
        if (std::fabs(theta[e]) >= tol) order.push_back(e);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        const double fa = std::fabs(theta[a]), fb = std::fabs(theta[b]);
        return fa != fb ? fa > fb : theta[a] > theta[b];
    });
    if (max_frags > 0 && (int)order.size() > max_frags) order.resize(max_frags);
    const int ne = (int)order.size();
    sts.num_frags = ne;
    if (ne == 0) {
        S->valid = true;
        if (num_frags) *num_frags = 0;
        sts.seconds_tridiag = now_s() - t1;
        if (stats) *stats = sts;
        mambo_csa_count_launches(h, launches);
        return MAMBO_OK;
    }
    std::vector<double> sel_theta(ne);
    std::vector<int> sel_blk(ne);
    for (int i = 0; i < ne; ++i) {
        sel_theta[i] = theta[order[i]];
        int b = 0;
        while (b + 1 < nblk && order[i] >= blk[3 * (b + 1) + 2]) ++b;
        sel_blk[i] = b;
    }
    const long long ldz = ((long long)ne + 63) / 64 * 64;
    double *d_sel = nullptr, *Z = nullptr, *B1 = nullptr, *B2 = nullptr, *G = nullptr;
    int* d_selblk = nullptr;
    unsigned long long* d_dev = nullptr;
    const size_t big = std::max((size_t)m * ldz, (size_t)ne * ldv);
    DCU(scr.get((void**)&d_sel, sizeof(double) * ne));
    DCU(scr.get((void**)&d_selblk, sizeof(int) * ne));
    DCU(scr.get((void**)&Z, sizeof(double) * (size_t)m * ldz));
    DCU(scr.get((void**)&B1, sizeof(double) * big));
    DCU(scr.get((void**)&B2, sizeof(double) * big));
    DCU(scr.get((void**)&G, sizeof(double) * (size_t)ne * ldz));
    DCU(scr.get((void**)&d_dev, sizeof(unsigned long long)));
    DCU(cudaMemcpyAsync(d_sel, sel_theta.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, st));
    DCU(cudaMemcpyAsync(d_selblk, sel_blk.data(), sizeof(int) * ne, cudaMemcpyHostToDevice, st));
    df::k_twisted<<<(ne + 63) / 64, 64, 0, st>>>(d_alpha, d_beta, m, d_blk, nblk, d_sel, d_selblk, ne, pivmin, B1, B2, ldz, Z, ldz);
    launches += 1;
    DCU(cudaGetLastError());
    {   // numerically degenerate clusters inside one block: redo their vectors on the host (tridiag_cluster_vectors)
        double tnorm = 0.0;
        for (int i = 0; i < m; ++i) tnorm = std::max(tnorm, std::fabs(ha[i]) + (i > 0 ? std::fabs(hb[i - 1]) : 0.0) + std::fabs(hb[i]));
        std::vector<std::vector<int>> clusters;
        find_tight_clusters(sel_theta, sel_blk, nblk, 1e-11 * tnorm, clusters);
        std::vector<double> zc, col;
        for (const auto& cl : clusters) {
            const int b = sel_blk[cl[0]], r0 = blk[3 * b], nr = blk[3 * b + 1], k = (int)cl.size();
            std::vector<double> th(k);
            for (int t = 0; t < k; ++t) th[t] = sel_theta[cl[t]];
            tridiag_cluster_vectors(ha.data() + r0, hb.data() + r0, nr, th.data(), k, tnorm, zc);
            for (int t = 0; t < k; ++t) {
                col.assign(m, 0.0);
                std::copy(zc.begin() + (size_t)t * nr, zc.begin() + (size_t)(t + 1) * nr, col.begin() + r0);
                DCU(cudaMemcpy2DAsync(Z + cl[t], sizeof(double) * ldz, col.data(), sizeof(double), sizeof(double), m, cudaMemcpyHostToDevice, st));
                DCU(cudaStreamSynchronize(st));          // col is reused
            }
            sts.cluster_vectors += k;
        }
    }
    // Newton-Schulz re-orthonormalisation Z <- Z (3 I - Z^T Z) / 2: eigenvectors of close eigenvalues come out of the twisted
    // factorisation orthogonal only to eps / gap; a mix inside such a cluster changes neither the residuals nor the sum of
    // the fragments.  C = 1.5 I - 0.5 Z^T Z comes straight out of the Gram GEMM; max |Z^T Z - I| = 2 max |C - I|.
    auto ortho_dev = [&](double* out) -> int {
        gemm<true, 1>(st, ne, ne, m, Z, ldz, Z, ldz, G, ldz);
        DCU(cudaMemsetAsync(d_dev, 0, sizeof(unsigned long long), st));
        df::k_max_dev_identity<<<256, 256, 0, st>>>(G, ne, ldz, d_dev);
        launches += 2;
        unsigned long long bits = 0;
        DCU(cudaMemcpyAsync(&bits, d_dev, sizeof(bits), cudaMemcpyDeviceToHost, st));
        DCU(cudaStreamSynchronize(st));
        double d;
        std::memcpy(&d, &bits, sizeof(d));
        *out = 2.0 * d;
        return MAMBO_OK;
    };
    double dev = 0.0;
    rc = ortho_dev(&dev);
    if (rc) return rc;
    sts.ortho_dev_before = sts.ortho_dev_after = dev;
    if (!(dev < 0.5))
        return dfail(MAMBO_ESTATE, "tridiagonal eigenvectors of a numerically degenerate cluster coincide (max |Z^T Z - I| = " + std::to_string(dev) + ")");
    for (int it = 0; it < 5 && dev > 2e-14; ++it) {
        gemm<false, 0>(st, m, ne, ne, Z, ldz, G, ldz, B1, ldz);
        launches += 1;
        std::swap(Z, B1);
        rc = ortho_dev(&dev);
        if (rc) return rc;
        sts.ortho_dev_after = dev;
    }
    sts.seconds_tridiag = now_s() - t1;

    // ---- 3. back to the row space: Y = Z^T V  (ne x R) -------------------------------------------------------------------------
    const double t2 = now_s();
    double* Y = B2;
    DCU(cudaMemsetAsync(Y, 0, sizeof(double) * (size_t)ne * ldv, st));
    gemm<true, 0>(st, ne, R, m, Z, ldz, V, ldv, Y, ldv);
    launches += 1;
    DCU(cudaGetLastError());
    if (verify) {
        double* d_res = nullptr;
        DCU(scr.get((void**)&d_res, sizeof(double) * ne));
        for (int e = 0; e < ne; ++e) {
            const double* ye = Y + (size_t)e * ldv;
            if (general) lanczos::k_symv<1><<<mv_grid, 256, 0, st>>>(v.T2, v.T1, v.ldT, R, ye, w);
            else lanczos::k_symv<0><<<mv_grid, 256, 0, st>>>(v.T1, nullptr, v.ldT, R, ye, w);
            k_eig_residual<<<1, 256, 0, st>>>(w, ye, sel_theta[e], R, d_res + e);
        }
        launches += 2LL * ne;
        std::vector<double> res(ne);
        DCU(cudaMemcpyAsync(res.data(), d_res, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
        DCU(cudaStreamSynchronize(st));
        double mx = 0.0;
        for (double r : res) mx = std::max(mx, r);
        sts.max_rel_residual = mx / std::max(std::fabs(sel_theta[0]), 1e-300);
    }
    DCU(cudaStreamSynchronize(st));
    sts.seconds_backtransform = now_s() - t2;

    // ---- 4. fragments: L_e = reshape(y_e), symmetry class, eigen-decomposition of every L_e ----------------------------------
    const double t3 = now_s();
    const size_t nn = (size_t)n * n;
    DCU(cudaMalloc((void**)&S->d_L, sizeof(double) * nn * ne));
    DCU(cudaMalloc((void**)&S->d_U, sizeof(double) * nn * ne));
    DCU(cudaMalloc((void**)&S->d_omega, sizeof(double) * (size_t)n * ne));
    DCU(cudaMalloc((void**)&S->d_coeff, sizeof(double) * ne));
    if (!v.packed) DCU(cudaMemsetAsync(S->d_L, 0, sizeof(double) * nn * ne, st));
    df::k_frag_matrix<<<dim3(vblocks, ne), 256, 0, st>>>(Y, ldv, R, v.pa, v.pq, v.sw, n, v.packed, S->d_L);
    df::k_frag_fix_sign<<<ne, 256, 0, st>>>(S->d_L, n);
    launches += 1;
    double *d_sym = nullptr, *d_asym = nullptr;
    int* d_status = nullptr;
    DCU(scr.get((void**)&d_sym, sizeof(double) * ne));
    DCU(scr.get((void**)&d_asym, sizeof(double) * ne));
    DCU(scr.get((void**)&d_status, sizeof(int) * ne));
    df::k_frag_symmetry<<<ne, 256, 0, st>>>(S->d_L, n, d_sym, d_asym);
    launches += 2;
    std::vector<double> hsym(ne), hasym(ne);
    DCU(cudaMemcpyAsync(hsym.data(), d_sym, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
    DCU(cudaMemcpyAsync(hasym.data(), d_asym, sizeof(double) * ne, cudaMemcpyDeviceToHost, st));
    const size_t sm_full = df::jacobi_smem(n, true), sm_half = df::jacobi_smem(n, false);
    const bool usm = sm_full <= 232448;
    if (!usm && sm_half > 232448) return dfail(MAMBO_EINVAL, "N too large for the shared-memory Jacobi eigensolver");
    const int jthreads = n >= 64 ? 1024 : (n >= 24 ? 256 : 64);
    if (usm) {
        DCU(raise_smem(df::k_jacobi_batch<true>, sm_full));
        df::k_jacobi_batch<true><<<ne, jthreads, sm_full, st>>>(S->d_L, n, S->d_omega, S->d_U, nullptr, d_status, 60);
    } else {
        double* Ut = nullptr;                              // ne x n x n: the transposed eigenvector accumulators live in global memory / L2
        DCU(scr.get((void**)&Ut, sizeof(double) * nn * ne));
        DCU(raise_smem(df::k_jacobi_batch<false>, sm_half));
        df::k_jacobi_batch<false><<<ne, jthreads, sm_half, st>>>(S->d_L, n, S->d_omega, S->d_U, Ut, d_status, 60);
    }
    launches += 1;
    DCU(cudaGetLastError());
    std::vector<int> hstat(ne);
    DCU(cudaMemcpyAsync(hstat.data(), d_status, sizeof(int) * ne, cudaMemcpyDeviceToHost, st));
    DCU(cudaStreamSynchronize(st));
    S->coeff.assign(sel_theta.begin(), sel_theta.end());
    S->kind.assign(ne, 0);
    const double tiny = 1e-12;                             // SVD_tiny, config.jl:21
    int sweeps_max = 0;
    bool unconverged = false;
    for (int e = 0; e < ne; ++e) {
        if (hsym[e] > tiny) {                              // decompose.jl:355-361
            if (hasym[e] > tiny) {
                S->release();
                return dfail(MAMBO_ESTATE, "DF fragment " + std::to_string(e + 1) + " is neither Hermitian or anti-Hermitian!");
            }
            S->kind[e] = 1;
            S->coeff[e] = -S->coeff[e];
            sts.anti_frags++;
            continue;
        }
        if (hstat[e] < 0) {
            if (!unconverged && std::getenv("MAMBO_DF_DEBUG")) {
                std::vector<double> Lh(nn);
                cudaMemcpy(Lh.data(), S->d_L + (size_t)e * nn, sizeof(double) * nn, cudaMemcpyDeviceToHost);
                double mx = 0.0, mn = 1e300, fro = 0.0;
                for (double x : Lh) { mx = std::max(mx, std::fabs(x)); if (x != 0.0) mn = std::min(mn, std::fabs(x)); fro += x * x; }
                fprintf(stderr, "[mambo df] fragment %d (theta %.3e) Jacobi status %d: max|L| %.3e min nonzero %.3e ||L||^2 %.6e\n", e, sel_theta[e],
                        hstat[e], mx, mn, fro);
            }
            unconverged = true;
        }
        sweeps_max = std::max(sweeps_max, std::abs(hstat[e]));
    }
    sts.jacobi_sweeps_max = unconverged ? -sweeps_max : sweeps_max;
    DCU(cudaMemcpyAsync(S->d_coeff, S->coeff.data(), sizeof(double) * ne, cudaMemcpyHostToDevice, st));
    DCU(cudaStreamSynchronize(st));
    S->nfrag = ne;
    S->valid = !unconverged;
    sts.seconds_fragments = now_s() - t3;
    mambo_csa_count_launches(h, launches);
    if (num_frags) *num_frags = ne;
    if (stats) *stats = sts;
    if (unconverged) return dfail(MAMBO_ESTATE, "Jacobi eigensolver did not converge for some fragment");
    return MAMBO_OK;
}

int mambo_csa_df_get(mambo_csa* h, int first, int count, double* coeff, double* omega, double* U, double* L, int* kind) {
    if (!h) return dfail(MAMBO_EINVAL, "null handle");
    void** slot = mambo_csa_df_slot(h, nullptr);
    DfState* S = slot ? static_cast<DfState*>(*slot) : nullptr;
    if (!S || !S->valid) return dfail(MAMBO_ESTATE, "no double factorisation on this handle (call mambo_csa_df_decompose first)");
    if (first < 0 || count < 0 || first + count > S->nfrag) return dfail(MAMBO_EINVAL, "fragment range out of bounds");
    if (count == 0) return MAMBO_OK;
    DCU(cudaSetDevice(S->device));
    const size_t n = S->N, nn = n * n;
    if (coeff) std::memcpy(coeff, S->coeff.data() + first, sizeof(double) * count);
    if (kind) std::memcpy(kind, S->kind.data() + first, sizeof(int) * count);
    if (omega) DCU(cudaMemcpy(omega, S->d_omega + (size_t)first * n, sizeof(double) * n * count, cudaMemcpyDeviceToHost));
    if (U) DCU(cudaMemcpy(U, S->d_U + (size_t)first * nn, sizeof(double) * nn * count, cudaMemcpyDeviceToHost));
    if (L) DCU(cudaMemcpy(L, S->d_L + (size_t)first * nn, sizeof(double) * nn * count, cudaMemcpyDeviceToHost));
    for (int i = 0; i < count; ++i)
        if (S->kind[first + i] == 1) {      // anti-symmetric fragment: the complex eigen-decomposition stays with the caller
            if (omega) std::memset(omega + (size_t)i * n, 0, sizeof(double) * n);
            if (U) std::memset(U + (size_t)i * nn, 0, sizeof(double) * nn);
        }
    return MAMBO_OK;
}

int mambo_csa_df_based_greedy(mambo_csa* h, double* lambda, double* U1, double* Lmat) {
    if (!h || !lambda || !U1) return dfail(MAMBO_EINVAL, "null argument");
    void** slot = mambo_csa_df_slot(h, nullptr);
    DfState* S = slot ? static_cast<DfState*>(*slot) : nullptr;
    if (!S || !S->valid) return dfail(MAMBO_ESTATE, "no double factorisation on this handle (call mambo_csa_df_decompose first)");
    if (S->nfrag == 0) return dfail(MAMBO_ESTATE, "the double factorisation holds no fragment (every |eigenvalue| is below the tolerance)");
    for (int k : S->kind)
        if (k) return dfail(MAMBO_ESTATE, "DF_based_greedy with anti-symmetric (complex) fragments is not available on the device");
    mambo_resident_view v;
    int rc = mambo_csa_resident_view(h, &v);
    if (rc) return rc;
    DCU(cudaSetDevice(S->device));
    cudaStream_t st = v.stream;
    const int n = S->N, nf = S->nfrag;
    Scratch scr;
    double *vv = nullptr, *d_Lmat = nullptr;
    DCU(scr.get((void**)&vv, sizeof(double) * (size_t)n * nf));
    DCU(scr.get((void**)&d_Lmat, sizeof(double) * (size_t)n * n));
    const size_t smem = sizeof(double) * (size_t)n * n;
    DCU(raise_smem(df::k_df_frame, smem));
    df::k_df_frame<<<nf, 256, smem, st>>>(S->d_U, S->d_omega, n, nf, vv);
    df::k_df_lmat<<<(n * n + 255) / 256, 256, 0, st>>>(vv, S->d_coeff, n, nf, d_Lmat);
    DCU(cudaGetLastError());
    mambo_csa_count_launches(h, 2);
    std::vector<double> Lm((size_t)n * n);
    DCU(cudaMemcpyAsync(Lm.data(), d_Lmat, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
    DCU(cudaMemcpyAsync(U1, S->d_U, sizeof(double) * n * n, cudaMemcpyDeviceToHost, st));
    DCU(cudaStreamSynchronize(st));
    int idx = 0;
    for (int i = 0; i < n; ++i)                       // cartan_mat_to_2b, fermionic.jl:34-48
        for (int j = 0; j <= i; ++j) lambda[idx++] = 0.5 * (Lm[i + (size_t)n * j] + Lm[j + (size_t)n * i]);
    if (Lmat) std::memcpy(Lmat, Lm.data(), sizeof(double) * n * n);
    return MAMBO_OK;
}

}  // extern "C"
