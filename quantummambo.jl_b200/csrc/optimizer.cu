// optimizer.cu -- device-resident BFGS and the multistart scheduler of libmambo_b200.so.
//
// Stands in for `optimize(cost, grad!, x0, BFGS())` of /root/reference/src/UTILS/decompose.jl:107-119, 233-243
// (Optim.jl: dense inverse-Hessian BFGS, g_tol = 1e-8 on the inf-norm of the gradient, 1000 iterations) when the
// whole greedy step should stay on the GPU: x, the gradient and the dense L x L inverse Hessian live in HBM,
// one fused pass per iteration applies the pending rank-2 update and forms H*g (2 * 8 L^2 bytes of traffic),
// only scalars cross PCIe.  The line search is a strong-Wolfe bracketing/zoom search (Nocedal & Wright
// alg. 3.5/3.6; c1 = 1e-4, c2 = 0.9) -- Optim's HagerZhang is a different member of the same family, so
// trajectories are compared at convergence (final cost / 1-norm), not step by step.
//
// mambo_multistart_run: K independent starts over a set of handles (one per GPU), one host thread per
// handle pulling start indices from a shared counter; precedent orbitals.jl:39-67 (OO_reps).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/mambo_csa.h"

extern "C" void mambo_set_error(const char* msg);   // defined in mambo_csa.cu (thread-local message)
extern "C" int mambo_csa_scratch(mambo_csa* h, size_t bytes, void** ptr);   // handle-owned device scratch (mambo_csa.cu)
extern "C" int mambo_csa_scratch_host(mambo_csa* h, void** ptr);            // handle-owned pinned scalars (128 bytes)
extern "C" int mambo_csa_status_fetch(mambo_csa* h);    // enqueue the D2H of the device status word on the handle's stream
extern "C" int mambo_csa_status_check(mambo_csa* h);    // after a synchronise: MAMBO_ERANGE / MAMBO_ESTATE if it was raised

namespace {

#define OCU(call)                                           \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) {                           \
            snprintf(o->err, sizeof(o->err), "%s: %s", #call, cudaGetErrorString(e__)); \
            return MAMBO_ECUDA;                             \
        }                                                   \
    } while (0)

// ---- vector kernels (deterministic: fixed-order tree reductions, single CTA for the final sum) ------------------
__global__ void k_axpy(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ p, double a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fma(a, p[i], x[i]);
}
__global__ void k_sub(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] - b[i];
}
// res[0] = a.b, res[1] = max|a|   (single CTA, 1024 threads)
__global__ void __launch_bounds__(1024) k_dot_absmax(const double* __restrict__ a, const double* __restrict__ b, int n,
                                                     double* __restrict__ res) {
    __shared__ double sd[32], sm[32];
    double d = 0.0, m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        d = fma(a[i], b[i], d);
        m = fmax(m, fabs(a[i]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        d += __shfl_xor_sync(0xffffffffu, d, o);
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) {
        sd[threadIdx.x >> 5] = d;
        sm[threadIdx.x >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double D = 0.0, M = 0.0;
        for (int w = 0; w < 32; ++w) {
            D += sd[w];
            M = fmax(M, sm[w]);
        }
        res[0] = D;
        res[1] = M;
    }
}
__global__ void k_set_identity(double* __restrict__ H, long long n, double diag) {
    const long long total = n * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        H[idx] = (idx / n == idx % n) ? diag : 0.0;
}
__global__ void k_scale(double* __restrict__ out, const double* __restrict__ a, double f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f * a[i];
}
// One pass over the dense inverse Hessian: (optionally) apply the pending (self-scaled) BFGS update
//     H <- tau H - rho (s Hy^T + Hy s^T) + c s s^T          (rho already carries the factor tau)
// and form u = H g with the updated matrix.  One warp per row, fixed-order shuffle reduction.
__global__ void __launch_bounds__(256) k_update_gemv(double* __restrict__ H, int n, const double* __restrict__ s,
                                                     const double* __restrict__ Hy, double tau, double rho, double c, int apply,
                                                     const double* __restrict__ g, double* __restrict__ u) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= n) return;
    double* hr = H + (size_t)row * n;
    const double si = apply ? s[row] : 0.0, hyi = apply ? Hy[row] : 0.0;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) {
        double h = hr[j];
        if (apply) {
            h = tau * h - rho * (si * Hy[j] + hyi * s[j]) + c * si * s[j];
            hr[j] = h;
        }
        acc = fma(h, g[j], acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) u[row] = acc;
}
__global__ void k_scale_neg(double* __restrict__ out, const double* __restrict__ a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = -a[i];
}
// p = -(tau u - rho (s a + Hy b) + c s b)   with a = Hy.g, b = s.g   (the search direction of the *updated* H)
__global__ void k_direction(double* __restrict__ p, const double* __restrict__ u, const double* __restrict__ s,
                            const double* __restrict__ Hy, double tau, double rho, double c, double a, double b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = -(tau * u[i] - rho * (s[i] * a + Hy[i] * b) + c * s[i] * b);
}

// Up to 6 dot products a_k.b_k and max|m| in ONE single-CTA pass (one host round trip instead of one per scalar):
// res[k] = a_k.b_k, res[6] = max|m|.  Fixed-order reductions (bit-reproducible).
struct DotSet {
    const double* a[6];
    const double* b[6];
    const double* m;
    int nd;
};
__global__ void __launch_bounds__(1024) k_multi_dot(DotSet ds, int n, double* __restrict__ res) {
    __shared__ double sd[7][32];
    double acc[6] = {0, 0, 0, 0, 0, 0}, mx = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (k < ds.nd) acc[k] = fma(ds.a[k][i], ds.b[k][i], acc[k]);
        if (ds.m) mx = fmax(mx, fabs(ds.m[i]));
    }
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) sd[k][threadIdx.x >> 5] = acc[k];
        sd[6][threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double r = 0.0;
        if (threadIdx.x < 6) {
            for (int w = 0; w < 32; ++w) r += sd[threadIdx.x][w];
        } else {
            for (int w = 0; w < 32; ++w) r = fmax(r, sd[6][w]);
        }
        res[threadIdx.x] = r;
    }
}

// L-BFGS two-loop recursion in one single-CTA kernel (Nocedal & Wright alg. 7.4): p = -H_k g from the `used` newest
// pairs of the ring (S, Y: [mem][n]; rho[mem]); `head` is the slot of the newest pair; H_0 = gamma I.
// Also returns res[0] = g.p and res[1] = max|g| so that the host needs one round trip per iteration.
__global__ void __launch_bounds__(1024) k_lbfgs_direction(const double* __restrict__ g, const double* __restrict__ S, const double* __restrict__ Y,
                                                          const double* __restrict__ rho, int mem, int used, int head, double gamma, int n,
                                                          double* __restrict__ p, double* __restrict__ res) {
    __shared__ double sred[32];
    __shared__ double s_val;
    __shared__ double s_alpha[64];
    auto block_sum = [&](double v) -> double {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 32; ++w) t += sred[w];
            s_val = t;
        }
        __syncthreads();
        const double out = s_val;
        __syncthreads();
        return out;
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = g[i];          // q = g
    __syncthreads();
    for (int k = 0; k < used; ++k) {                                        // newest -> oldest
        const int slot = (head - k + mem) % mem;
        const double* s = S + (size_t)slot * n;
        const double* y = Y + (size_t)slot * n;
        double d = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) d = fma(s[i], p[i], d);
        const double a = rho[slot] * block_sum(d);
        if (threadIdx.x == 0) s_alpha[k] = a;
        for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = fma(-a, y[i], p[i]);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] *= gamma;        // r = H_0 q
    __syncthreads();
    for (int k = used - 1; k >= 0; --k) {                                   // oldest -> newest
        const int slot = (head - k + mem) % mem;
        const double* s = S + (size_t)slot * n;
        const double* y = Y + (size_t)slot * n;
        double d = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) d = fma(y[i], p[i], d);
        const double b = rho[slot] * block_sum(d);
        const double a = s_alpha[k];
        for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = fma(a - b, s[i], p[i]);
        __syncthreads();
    }
    double gp = 0.0, gm = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = -p[i];
        p[i] = v;
        gp = fma(g[i], v, gp);
        gm = fmax(gm, fabs(g[i]));
    }
    const double tot = block_sum(gp);
    for (int o = 16; o > 0; o >>= 1) gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = gm;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < 32; ++w) m = fmax(m, sred[w]);
        res[0] = tot;
        res[1] = m;
    }
}
__global__ void k_store_pair(double* __restrict__ Sslot, double* __restrict__ Yslot, const double* __restrict__ s, const double* __restrict__ y,
                             double* __restrict__ rho_slot, double rho, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        Sslot[i] = s[i];
        Yslot[i] = y[i];
    }
    if (i == 0) *rho_slot = rho;
}

struct Opt {
    mambo_csa* h = nullptr;
    int L = 0, device = 0;
    cudaStream_t stream = nullptr;
    double *x = nullptr, *g = nullptr, *p = nullptr, *xt = nullptr, *out = nullptr, *s = nullptr, *y = nullptr, *Hy = nullptr,
           *s2 = nullptr, *Hy2 = nullptr, *u = nullptr, *H = nullptr, *scal = nullptr;
    double *LS = nullptr, *LY = nullptr, *Lrho = nullptr;   // L-BFGS ring (lbfgs_memory > 0)
    double* h_scal = nullptr;  // pinned
    int evals = 0;
    bool sharded = false;
    char err[256] = {0};
};

int blocks(int n) { return (n + 255) / 256; }

int opt_free(Opt* o) {   // device buffers and pinned scalars live in the handle's scratch areas: nothing to release
    o->h_scal = nullptr;
    return 0;
}

// doubles of device scratch one optimisation run needs (layout: optimize_impl)
size_t opt_scratch_doubles(int L, int mem) {
    const size_t Lp = ((size_t)L + 1 + 15) & ~(size_t)15;
    const size_t big = mem > 0 ? (2 * (size_t)mem * L + 64) : (size_t)L * L;
    return 11 * Lp + 16 + big;
}

// several dot products and max|m| in one kernel + one host round trip; out[k] = a_k.b_k, *amax = max|m|
int dots_host(Opt* o, std::initializer_list<std::pair<const double*, const double*>> pairs, const double* m, double* out, double* amax) {
    DotSet ds;
    ds.nd = 0;
    for (auto& pr : pairs) {
        ds.a[ds.nd] = pr.first;
        ds.b[ds.nd] = pr.second;
        ++ds.nd;
    }
    for (int k = ds.nd; k < 6; ++k) ds.a[k] = ds.b[k] = nullptr;
    ds.m = m;
    k_multi_dot<<<1, 1024, 0, o->stream>>>(ds, o->L, o->scal);
    OCU(cudaMemcpyAsync(o->h_scal, o->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    OCU(cudaStreamSynchronize(o->stream));
    for (int k = 0; k < ds.nd; ++k) out[k] = o->h_scal[k];
    if (amax) *amax = o->h_scal[6];
    return MAMBO_OK;
}

// phi(alpha) = cost(x + alpha p), dphi = grad(x + alpha p).p ; leaves [cost, grad] in o->out and the point in o->xt
int phi(Opt* o, double alpha, double* f, double* df) {
    k_axpy<<<blocks(o->L), 256, 0, o->stream>>>(o->xt, o->x, o->p, alpha, o->L);
    // Row-panel shard: every rank runs this same optimiser on its own shard handle; the evaluation exchanges the partial
    // vectors over NVLink peer memory inside the library and returns identical bits on every rank, so the replicated
    // BFGS states never diverge and no host collective is needed.
    int rc = o->sharded ? mambo_csa_eval_sharded_device(o->h, o->xt, o->out) : mambo_csa_eval_device(o->h, o->xt, o->out);
    if (rc) {
        snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
        return rc;
    }
    o->evals++;
    k_dot_absmax<<<1, 1024, 0, o->stream>>>(o->out + 1, o->p, o->L, o->scal);
    OCU(cudaMemcpyAsync(o->h_scal, o->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    OCU(cudaMemcpyAsync(o->h_scal + 8, o->out, sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    rc = mambo_csa_status_fetch(o->h);            // device status word (||K||_1 out of range, NaN in x, peer time-out) ...
    if (rc) {
        snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
        return rc;
    }
    OCU(cudaStreamSynchronize(o->stream));
    rc = mambo_csa_status_check(o->h);            // ... checked with the scalars of the same evaluation
    if (rc) {
        snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
        return rc;
    }
    *df = o->h_scal[0];
    *f = o->h_scal[8];
    return MAMBO_OK;
}

double cubic_min(double a, double fa, double fpa, double b, double fb, double c, double fc) {
    // minimiser of the cubic through (a,fa,fpa), (b,fb), (c,fc); NaN if not usable
    const double db = b - a, dc = c - a;
    const double denom = (db * dc) * (db * dc) * (db - dc);
    if (denom == 0.0) return NAN;
    const double d1 = fb - fa - fpa * db, d2 = fc - fa - fpa * dc;
    const double A = (dc * dc * d1 - db * db * d2) / denom;
    const double B = (-dc * dc * dc * d1 + db * db * db * d2) / denom;
    const double rad = B * B - 3 * A * fpa;
    if (rad < 0 || A == 0.0) return NAN;
    return a + (-B + std::sqrt(rad)) / (3 * A);
}
double quad_min(double a, double fa, double fpa, double b, double fb) {
    const double db = b - a;
    const double B = (fb - fa - fpa * db) / (db * db);
    if (B <= 0) return NAN;
    return a - fpa / (2 * B);
}

// Strong-Wolfe line search.  On success the accepted point is the last one evaluated (o->xt / o->out).
int line_search(Opt* o, double f0, double df0, double alpha1, double* alpha_out, double* f_out, bool* ok) {
    const double c1 = 1e-4, c2 = 0.9, amax = 50.0;
    double a_prev = 0.0, f_prev = f0, df_prev = df0;
    double a = alpha1, f = 0, df = 0;
    *ok = false;
    auto zoom = [&](double lo, double flo, double dlo, double hi, double fhi, double dhi) -> int {
        double a_rec = 0, f_rec = f0;
        (void)dhi;
        for (int it = 0; it < 30; ++it) {
            const double dalpha = hi - lo;
            double lo_b = std::min(lo, hi), hi_b = std::max(lo, hi);
            double aj = NAN;
            if (it > 0) aj = cubic_min(lo, flo, dlo, hi, fhi, a_rec, f_rec);
            const double cchk = 0.2 * std::fabs(dalpha), qchk = 0.1 * std::fabs(dalpha);
            if (!(aj == aj) || aj > hi_b - cchk || aj < lo_b + cchk) {
                aj = quad_min(lo, flo, dlo, hi, fhi);
                if (!(aj == aj) || aj > hi_b - qchk || aj < lo_b + qchk) aj = lo + 0.5 * dalpha;
            }
            double fj, dfj;
            int rc = phi(o, aj, &fj, &dfj);
            if (rc) return rc;
            if (fj > f0 + c1 * aj * df0 || fj >= flo) {
                a_rec = hi; f_rec = fhi;
                hi = aj; fhi = fj;
            } else {
                if (std::fabs(dfj) <= -c2 * df0) {
                    *alpha_out = aj; *f_out = fj; *ok = true;
                    return MAMBO_OK;
                }
                if (dfj * (hi - lo) >= 0) {
                    a_rec = hi; f_rec = fhi;
                    hi = lo; fhi = flo;
                } else {
                    a_rec = lo; f_rec = flo;
                }
                lo = aj; flo = fj; dlo = dfj;
            }
            if (std::fabs(hi - lo) < 1e-16 * std::max(1.0, std::fabs(lo))) break;
        }
        // not converged: fall back to the best sufficient-decrease point found (lo), re-evaluated so that out/xt match
        if (lo > 0 && flo < f0) {
            double fj, dfj;
            int rc = phi(o, lo, &fj, &dfj);
            if (rc) return rc;
            *alpha_out = lo; *f_out = fj; *ok = true;
        }
        return MAMBO_OK;
    };
    bool last_valid = false;      // the last point evaluated (o->xt / o->out) is a_prev with f_prev, a sufficient-decrease point
    for (int i = 0; i < 25; ++i) {
        int rc = phi(o, a, &f, &df);
        if (rc) return rc;
        if (!(f == f)) {  // NaN: shrink
            a = 0.5 * (a_prev + a);
            last_valid = false;
            continue;
        }
        if (f > f0 + c1 * a * df0 || (i > 0 && f >= f_prev)) return zoom(a_prev, f_prev, df_prev, a, f, df);
        if (std::fabs(df) <= -c2 * df0) {
            *alpha_out = a; *f_out = f; *ok = true;
            return MAMBO_OK;
        }
        if (df >= 0) return zoom(a, f, df, a_prev, f_prev, df_prev);
        a_prev = a; f_prev = f; df_prev = df;
        last_valid = true;
        a = std::min(2.0 * a, amax);
        if (a_prev >= amax) break;
    }
    // The bracket phase ran out of doublings (or reached amax) with the cost still falling along p: a long step that satisfies
    // the sufficient-decrease condition is a usable step, not a failure (Optim's HagerZhang would carry on from it as well).
    // o->xt / o->out hold exactly this point, so alpha and f come from the same evaluation.
    if (last_valid && f_prev < f0) {
        *alpha_out = a_prev; *f_out = f_prev; *ok = true;
    }
    return MAMBO_OK;
}

int optimize_impl(mambo_csa* h, const double* x0, double* x_min, const mambo_opt_options_t* optin, mambo_opt_result_t* res,
                  std::string* errmsg) {
    mambo_opt_options_t opt = {1000, 1e-8, 0, 0};
    if (optin) {
        if (optin->max_iter > 0) opt.max_iter = optin->max_iter;
        if (optin->g_tol > 0) opt.g_tol = optin->g_tol;
        opt.lbfgs_memory = optin->lbfgs_memory;
        opt.verbose = optin->verbose;
    }
    mambo_csa_info_t info;
    int rc = mambo_csa_get_info(h, &info);
    if (rc) return rc;
    Opt ob, *o = &ob;
    o->h = h;
    o->L = info.num_params;
    o->sharded = info.nshards > 1;
    void* st = nullptr;
    rc = mambo_csa_get_stream(h, &st, &o->device);
    if (rc) return rc;
    o->stream = (cudaStream_t)st;
    const int L = o->L;
    auto finish = [&](int code) {
        if (code && errmsg) *errmsg = o->err;
        opt_free(o);
        return code;
    };
    if (cudaSetDevice(o->device) != cudaSuccess) return finish(MAMBO_ECUDA);
    const int mem = std::min(std::max(opt.lbfgs_memory, 0), 64);
    const bool lbfgs = mem > 0;
    {
        // one carve-up of the handle's scratch: 10 vectors, out (L+1), 8 scalars, then H (L x L) or the L-BFGS ring
        const size_t Lp = ((size_t)L + 1 + 15) & ~(size_t)15;     // keeps every piece 128-byte aligned
        const size_t total = opt_scratch_doubles(L, mem);
        void* base = nullptr;
        rc = mambo_csa_scratch(h, total * sizeof(double), &base);
        if (rc) {
            snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
            return finish(rc);
        }
        double* q = static_cast<double*>(base);
        double** vecs[] = {&o->x, &o->g, &o->p, &o->xt, &o->s, &o->y, &o->Hy, &o->s2, &o->Hy2, &o->u, &o->out};
        for (double** v : vecs) {
            *v = q;
            q += Lp;
        }
        o->scal = q;
        q += 16;
        if (lbfgs) {
            o->LS = q;
            o->LY = q + (size_t)mem * L;
            o->Lrho = q + 2 * (size_t)mem * L;
        } else {
            o->H = q;
        }
        void* hs = nullptr;
        rc = mambo_csa_scratch_host(h, &hs);
        if (rc) {
            snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
            return finish(rc);
        }
        o->h_scal = static_cast<double*>(hs);
    }
    auto CK = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess) {
            snprintf(o->err, sizeof(o->err), "%s: %s", what, cudaGetErrorString(e));
            return true;
        }
        return false;
    };
    if (CK(cudaMemcpyAsync(o->x, x0, sizeof(double) * L, cudaMemcpyHostToDevice, o->stream), "H2D x0")) return finish(MAMBO_ECUDA);
    if (!lbfgs) k_set_identity<<<1184, 256, 0, o->stream>>>(o->H, L, 1.0);

    // f, g at x0
    double f = 0, dfdummy = 0;
    if (CK(cudaMemsetAsync(o->p, 0, sizeof(double) * L, o->stream), "memset")) return finish(MAMBO_ECUDA);
    rc = phi(o, 0.0, &f, &dfdummy);
    if (rc) return finish(rc);
    cudaMemcpyAsync(o->g, o->out + 1, sizeof(double) * L, cudaMemcpyDeviceToDevice, o->stream);
    double gnorm = 0, gg = 0;
    rc = dots_host(o, {{o->g, o->g}}, o->g, &gg, &gnorm);
    if (rc) return finish(rc);
    k_scale_neg<<<blocks(L), 256, 0, o->stream>>>(o->p, o->g, L);   // H = I
    double df0 = -gg;                                                // g.p of the current direction, kept up to date below

    int it = 0, converged = 0;
    // Dense mode: the rank-2 update of iteration k is applied to H in memory during iteration k+1, in the same pass that
    // forms u = H g_new: (s_prev, Hy_prev, rho, cc) describe that pending update.  Host round trips per iteration: one
    // per line-search evaluation plus two fused multi-dot reads (was seven single dots).
    bool pending = false;
    bool fresh = true;             // H is still the (re)start identity: rescale it by s.y / y.y before the first update
    double rho = 0, cc = 0, tau = 1.0;
    // Optional self-scaling (Oren-Luenberger with Al-Baali's cap, MAMBO_BFGS_SELFSCALE=1): H is multiplied by
    // tau = min(1, s.y / y.Hy) before each update.  Off by default: on the CSA landscape the factors compound and the
    // metric collapses (N=100 planted fragment, 1000 iterations: cost 8.5e3 with, 3.6e2 without).  The limited-memory mode
    // (lbfgs_memory > 0), which re-scales gamma = s.y / y.y every iteration, reaches 1.4e2 in the same budget.
    const bool selfscale = std::getenv("MAMBO_BFGS_SELFSCALE") != nullptr;
    double *s_cur = o->s, *Hy_cur = o->Hy, *s_prev = o->s2, *Hy_prev = o->Hy2;
    double alpha_prev = 1.0;       // step length accepted by the previous line search
    double df0_prev = -1.0;        // directional derivative at the start of the previous line search
    int l_used = 0, l_head = -1;   // L-BFGS ring state
    double l_gamma = 1.0;
    for (; it < opt.max_iter; ++it) {
        if (gnorm <= opt.g_tol) {
            converged = 1;
            break;
        }
        if (!(df0 < 0)) {  // not a descent direction: restart from steepest descent
            if (!lbfgs) k_set_identity<<<1184, 256, 0, o->stream>>>(o->H, L, 1.0);
            pending = false;
            fresh = true;
            l_used = 0;
            l_head = -1;
            l_gamma = 1.0;
            k_scale_neg<<<blocks(L), 256, 0, o->stream>>>(o->p, o->g, L);
            rc = dots_host(o, {{o->g, o->g}}, nullptr, &gg, nullptr);
            if (rc) return finish(rc);
            df0 = -gg;
            if (!(df0 < 0)) break;
        }
        double alpha = 0, fnew = f;
        bool ok = false;
        // First trial step: the unit step once the metric is learnt, otherwise twice the step accepted last time.  With
        // L ~ N^2 parameters of very different curvature (lambda vs theta) and H = gamma I in every direction not yet
        // explored, the unit step is ~30x too long for hundreds of iterations; starting every search there costs 3-4
        // evaluations per iteration in the zoom phase.
        // (Nocedal & Wright (3.60): alpha_prev * phi'_prev(0) / phi'(0), the step that would give the same first-order change.)
        double a1 = (it == 0) ? std::min(1.0, 1.0 / std::max(gnorm, 1e-300)) : alpha_prev * (df0_prev / df0);
        if (it > 0) a1 = std::min(1.0, std::max(0.5 * alpha_prev, std::min(a1, 2.0 * alpha_prev)));
        df0_prev = df0;
        rc = line_search(o, f, df0, a1, &alpha, &fnew, &ok);
        if (rc) return finish(rc);
        if (!ok || !(fnew <= f)) {
            if (opt.verbose) fprintf(stderr, "[mambo bfgs] it %d: line search failed (f=%.6e)\n", it, f);
            break;
        }
        // accepted point = last evaluated: xt, out.  s = xt - x, y = g_new - g
        k_sub<<<blocks(L), 256, 0, o->stream>>>(s_cur, o->xt, o->x, L);
        k_sub<<<blocks(L), 256, 0, o->stream>>>(o->y, o->out + 1, o->g, L);
        double d2[2] = {0, 0};
        rc = dots_host(o, {{s_cur, o->y}, {o->y, o->y}}, nullptr, d2, nullptr);
        if (rc) return finish(rc);
        const double sy = d2[0], yy = d2[1];
        f = fnew;
        alpha_prev = alpha;
        if (lbfgs) {
            if (sy > 1e-300) {
                l_head = (l_head + 1) % mem;
                l_used = std::min(l_used + 1, mem);
                k_store_pair<<<blocks(L), 256, 0, o->stream>>>(o->LS + (size_t)l_head * L, o->LY + (size_t)l_head * L, s_cur, o->y,
                                                               o->Lrho + l_head, 1.0 / sy, L);
                l_gamma = sy / yy;
            }
            cudaMemcpyAsync(o->x, o->xt, sizeof(double) * L, cudaMemcpyDeviceToDevice, o->stream);
            cudaMemcpyAsync(o->g, o->out + 1, sizeof(double) * L, cudaMemcpyDeviceToDevice, o->stream);
            k_lbfgs_direction<<<1, 1024, 0, o->stream>>>(o->g, o->LS, o->LY, o->Lrho, mem, l_used, std::max(l_head, 0), l_gamma, L, o->p, o->scal);
            if (CK(cudaMemcpyAsync(o->h_scal, o->scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, o->stream), "D2H scalars") ||
                CK(cudaStreamSynchronize(o->stream), "sync"))
                return finish(MAMBO_ECUDA);
            df0 = o->h_scal[0];
            gnorm = o->h_scal[1];
        } else {
            if (fresh && sy > 1e-300) {
                // Nocedal & Wright (6.20): H0 <- (s.y / y.y) I before the first update, so the first curvature pair sets the scale
                const double gamma = sy / yy;
                k_set_identity<<<1184, 256, 0, o->stream>>>(o->H, L, gamma);
                k_scale<<<blocks(L), 256, 0, o->stream>>>(o->u, o->out + 1, gamma, L);     // u  = H g_new
                k_scale<<<blocks(L), 256, 0, o->stream>>>(Hy_cur, o->y, gamma, L);          // Hy = H y
                fresh = false;
            } else {
                // one pass over H: fold in the pending update of the previous iteration and form u = H g_new
                k_update_gemv<<<(L + 7) / 8, 256, 0, o->stream>>>(o->H, L, s_prev, Hy_prev, tau, rho, cc, pending ? 1 : 0, o->out + 1, o->u);
                // Hy = H y = H g_new - H g = u + p   (p = -H g)
                k_axpy<<<blocks(L), 256, 0, o->stream>>>(Hy_cur, o->u, o->p, 1.0, L);
            }
            pending = false;
            cudaMemcpyAsync(o->x, o->xt, sizeof(double) * L, cudaMemcpyDeviceToDevice, o->stream);
            cudaMemcpyAsync(o->g, o->out + 1, sizeof(double) * L, cudaMemcpyDeviceToDevice, o->stream);
            double d4[4] = {0, 0, 0, 0};   // y.Hy, Hy.g, s.g, u.g
            rc = dots_host(o, {{o->y, Hy_cur}, {Hy_cur, o->g}, {s_cur, o->g}, {o->u, o->g}}, o->g, d4, &gnorm);
            if (rc) return finish(rc);
            const double yHy = d4[0], hyg = d4[1], sg = d4[2], ug = d4[3];
            if (sy > 1e-300) {
                const double r1 = 1.0 / sy;
                tau = (selfscale && yHy > 0.0) ? std::min(1.0, sy / yHy) : 1.0;
                rho = tau * r1;                          // coefficient of the cross terms
                cc = r1 * (1.0 + tau * r1 * yHy);        // coefficient of s s^T
                // direction of the updated matrix from vectors only; the matrix itself is updated next iteration
                k_direction<<<blocks(L), 256, 0, o->stream>>>(o->p, o->u, s_cur, Hy_cur, tau, rho, cc, hyg, sg, L);
                df0 = -(tau * ug - rho * (sg * hyg + hyg * sg) + cc * sg * sg);
                pending = true;
                std::swap(s_cur, s_prev);
                std::swap(Hy_cur, Hy_prev);
            } else {
                k_scale_neg<<<blocks(L), 256, 0, o->stream>>>(o->p, o->u, L);
                df0 = -ug;
            }
        }
        if (opt.verbose && (it % opt.verbose == 0))
            fprintf(stderr, "[mambo bfgs] it %d f=%.10e |g|inf=%.3e alpha=%.3e evals=%d\n", it, f, gnorm, alpha, o->evals);
    }
    if (gnorm <= opt.g_tol) converged = 1;
    if (CK(cudaMemcpyAsync(x_min, o->x, sizeof(double) * L, cudaMemcpyDeviceToHost, o->stream), "D2H x") ||
        CK(cudaStreamSynchronize(o->stream), "sync"))
        return finish(MAMBO_ECUDA);
    if (o->sharded) {   // a peer that never published its partial vector shows up as a status flag, not as a CUDA error
        rc = mambo_csa_synchronize(h);
        if (rc) {
            snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
            return finish(rc);
        }
    }
    if (res) {
        res->minimum = f;
        res->iterations = it;
        res->evaluations = o->evals;
        res->converged = converged;
        res->device = o->device;
    }
    return finish(MAMBO_OK);
}

thread_local std::string g_opt_err;

}  // namespace

extern "C" {

// Allocate everything mambo_csa_optimize will need on this handle (device scratch for the given mode, pinned scalars) so
// that the run itself allocates nothing.  cudaMalloc / cudaMallocHost synchronise the whole device: with several handles
// on one GPU (lanes, shards in one process) an allocation in one thread would otherwise wait for kernels of another.
int mambo_csa_optimize_reserve(mambo_csa* h, int lbfgs_memory) {
    if (!h) return MAMBO_EINVAL;
    const int L = mambo_csa_num_params(h);
    if (L <= 0) return MAMBO_EINVAL;
    const int mem = std::min(std::max(lbfgs_memory, 0), 64);
    void* p = nullptr;
    int rc = mambo_csa_scratch(h, opt_scratch_doubles(L, mem) * sizeof(double), &p);
    if (rc) return rc;
    return mambo_csa_scratch_host(h, &p);
}

int mambo_csa_optimize(mambo_csa* h, const double* x0, double* x_min, const mambo_opt_options_t* opt, mambo_opt_result_t* res) {
    if (!h || !x0 || !x_min) return MAMBO_EINVAL;
    std::string err;
    int rc = optimize_impl(h, x0, x_min, opt, res, &err);
    if (rc && !err.empty()) mambo_set_error(err.c_str());
    return rc;
}

int mambo_multistart_run(mambo_csa** handles, int nhandles, int K, const double* x0, double* x_min, const mambo_opt_options_t* opt,
                         mambo_opt_result_t* res, int* best) {
    if (!handles || nhandles < 1 || K < 1 || !x0 || !x_min || !res) return MAMBO_EINVAL;
    const int L = mambo_csa_num_params(handles[0]);
    if (L <= 0) return MAMBO_EINVAL;
    for (int i = 1; i < nhandles; ++i)
        if (mambo_csa_num_params(handles[i]) != L) return MAMBO_EINVAL;
    for (int i = 0; i < nhandles; ++i) {   // allocate before any worker starts: an allocation stalls every handle of that GPU
        int rc = mambo_csa_optimize_reserve(handles[i], opt ? opt->lbfgs_memory : 0);
        if (rc) return rc;
    }
    std::atomic<int> next(0), failed(0);
    std::vector<std::string> errs(nhandles);
    std::vector<std::thread> workers;
    for (int w = 0; w < nhandles; ++w)
        workers.emplace_back([&, w]() {
            for (;;) {
                const int k = next.fetch_add(1);
                if (k >= K || failed.load()) break;
                int rc = optimize_impl(handles[w], x0 + (size_t)k * L, x_min + (size_t)k * L, opt, &res[k], &errs[w]);
                if (rc) {
                    failed.store(rc);
                    break;
                }
            }
        });
    for (auto& t : workers) t.join();
    if (failed.load()) {
        for (auto& e : errs)
            if (!e.empty()) {
                mambo_set_error(e.c_str());
                break;
            }
        return failed.load();
    }
    if (best) {
        int b = 0;
        for (int k = 1; k < K; ++k)
            if (res[k].minimum < res[b].minimum) b = k;
        *best = b;
    }
    return MAMBO_OK;
}

}  // extern "C"
