// optimizer.cu -- device-resident BFGS and the multistart scheduler of libmambo_b200.so.
//
// Stands in for `optimize(cost, grad!, x0, BFGS())` of /root/reference/src/UTILS/decompose.jl:107-119, 233-243
// (Optim.jl: dense inverse-Hessian BFGS, g_tol = 1e-8 on the inf-norm of the gradient, 1000 iterations) when the
// whole greedy step should stay on the GPU: x, the gradient and the dense L x L inverse Hessian live in HBM,
// one fused pass per iteration over the lower-triangular tiles of the (symmetric) matrix applies the pending rank-2 update and
// forms H*g (8 L^2 bytes of traffic: every stored element is read and written once, and used for both u_i and u_j),
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

#include "internal.h"
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
// The dense inverse Hessian is symmetric: only the 128 x 128 tiles on and below the diagonal are ever read or written (the
// allocation stays square, pitch ldh = L rounded up to 4 doubles so that rows start on 32-byte sectors).
constexpr int STB = 128;
__global__ void k_set_identity(double* __restrict__ H, long long ldh, long long n, double diag) {
    const long long total = n * ldh;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        H[idx] = (idx / ldh == idx % ldh) ? diag : 0.0;
}
__global__ void k_scale(double* __restrict__ out, const double* __restrict__ a, double f, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f * a[i];
}

struct DevState;
__device__ __forceinline__ bool dev_halted(const DevState* st);
__device__ __forceinline__ void dev_pending(const DevState* st, const double* s0, const double* s1, const double* hy0, const double* hy1,
                                            const double** s, const double** Hy, double* rho, double* c, int* apply);

// One pass over the lower-triangular tiles of the inverse Hessian: (optionally) apply the pending (self-scaled) BFGS update
//     H <- tau H - rho (s Hy^T + Hy s^T) + c s s^T          (rho already carries the factor tau)
// and accumulate u = H g with the updated matrix: tile (I, J), J <= I, contributes its row sums to block I of u and -- below the
// diagonal -- its column sums (the mirrored tile) to block J.  Both go to P[block][slot][128], slot = the OTHER tile index, so every
// (block, slot) pair is written exactly once and the fixed-order sum over slots (sym_gemv_sum) is bit-reproducible.
// 8 L^2 bytes of HBM traffic per iteration instead of 16 L^2.  st != NULL: the parameters come from the device state.
__global__ void __launch_bounds__(256) k_sym_update_gemv(double* __restrict__ H, long long ldh, int n, const DevState* __restrict__ st,
                                                         const double* __restrict__ s0, const double* __restrict__ s1,
                                                         const double* __restrict__ hy0, const double* __restrict__ hy1, double tau, double rho,
                                                         double c, int apply, const double* __restrict__ g, double* __restrict__ P, int ntb) {
    const int I = blockIdx.y, J = blockIdx.x;
    if (J > I) return;
    const double *s = s0, *Hy = hy0;
    if (st) {
        if (dev_halted(st)) return;
        dev_pending(st, s0, s1, hy0, hy1, &s, &Hy, &rho, &c, &apply);
        tau = 1.0;
    }
    __shared__ double sI[STB], hI[STB], gI[STB], sJ[STB], hJ[STB], gJ[STB], rowsum[STB];
    __shared__ double colsum[8][STB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < STB) {
        const int i = I * STB + tid, j = J * STB + tid;
        sI[tid] = (apply && i < n) ? s[i] : 0.0;
        hI[tid] = (apply && i < n) ? Hy[i] : 0.0;
        gI[tid] = i < n ? g[i] : 0.0;
        sJ[tid] = (apply && j < n) ? s[j] : 0.0;
        hJ[tid] = (apply && j < n) ? Hy[j] : 0.0;
        gJ[tid] = j < n ? g[j] : 0.0;
    }
    __syncthreads();
    const bool diag = (I == J);
    const int c0 = 2 * lane, c1 = 64 + 2 * lane;                       // two double2 per row and lane
    const bool v0 = J * STB + c0 < n, v1 = J * STB + c1 < n;           // (an odd n leaves one pad column inside the last double2: g = 0 there)
    double ca0 = 0.0, ca1 = 0.0, ca2 = 0.0, ca3 = 0.0;
    for (int r = warp; r < STB; r += 8) {
        const int i = I * STB + r;
        if (i >= n) {
            if (lane == 0) rowsum[r] = 0.0;
            continue;
        }
        double* hr = H + (long long)i * ldh + (long long)J * STB;
        double2 a = v0 ? *reinterpret_cast<double2*>(hr + c0) : make_double2(0.0, 0.0);
        double2 b = v1 ? *reinterpret_cast<double2*>(hr + c1) : make_double2(0.0, 0.0);
        if (apply) {
            const double si = sI[r], hi = hI[r];
            a.x = tau * a.x - rho * (si * hJ[c0] + hi * sJ[c0]) + c * (si * sJ[c0]);
            a.y = tau * a.y - rho * (si * hJ[c0 + 1] + hi * sJ[c0 + 1]) + c * (si * sJ[c0 + 1]);
            b.x = tau * b.x - rho * (si * hJ[c1] + hi * sJ[c1]) + c * (si * sJ[c1]);
            b.y = tau * b.y - rho * (si * hJ[c1 + 1] + hi * sJ[c1 + 1]) + c * (si * sJ[c1 + 1]);
            if (v0) *reinterpret_cast<double2*>(hr + c0) = a;
            if (v1) *reinterpret_cast<double2*>(hr + c1) = b;
        }
        double acc = fma(a.x, gJ[c0], fma(a.y, gJ[c0 + 1], fma(b.x, gJ[c1], b.y * gJ[c1 + 1])));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) rowsum[r] = acc;
        if (!diag) {
            const double gi = gI[r];
            ca0 = fma(a.x, gi, ca0);
            ca1 = fma(a.y, gi, ca1);
            ca2 = fma(b.x, gi, ca2);
            ca3 = fma(b.y, gi, ca3);
        }
    }
    if (!diag) {
        colsum[warp][c0] = ca0; colsum[warp][c0 + 1] = ca1; colsum[warp][c1] = ca2; colsum[warp][c1 + 1] = ca3;
    }
    __syncthreads();
    if (tid < STB) {
        P[((long long)I * ntb + J) * STB + tid] = rowsum[tid];
        if (!diag) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += colsum[w][tid];
            P[((long long)J * ntb + I) * STB + tid] = t;
        }
    }
}
// u[i] = sum over slots of P[block(i)][slot][i mod 128], fixed order
__device__ __forceinline__ double sym_gemv_sum(const double* __restrict__ P, int ntb, int i) {
    const double* q = P + (long long)(i / STB) * ntb * STB + (i % STB);
    double t = 0.0;
    for (int sl = 0; sl < ntb; ++sl) t += q[(long long)sl * STB];
    return t;
}
__global__ void k_sym_reduce(const double* __restrict__ P, int ntb, int n, double* __restrict__ u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) u[i] = sym_gemv_sum(P, ntb, i);
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


// ---- device-driven iterations (the common case: the first trial step satisfies the strong-Wolfe conditions) -------------
// The host loop above costs three host round trips per iteration (line-search scalars, s.y / y.y, the direction's dots): at
// N = 50 they are longer than the evaluation itself.  In the fast path every scalar of the iteration lives in DevState on
// the device and four kernels around the evaluation take all decisions there; the host only enqueues iterations (two in
// flight) and reads a copy of the state that trails each one.  Whenever the device cannot decide alone -- the trial step
// fails a Wolfe condition (zoom / extension needed), the direction is not a descent direction, a NaN appears -- it raises
// `halt`, every later kernel of the fast path returns at once (the evaluations queued behind it are wasted, nothing else
// is touched), and the host runs that one iteration through the line search above before it resumes the fast path.
struct DevState {
    double f, df0, alpha_prev, df0_prev, gnorm, rho, cc, a1, gtol, sy, yy, f_new, gmax_new, l_gamma;
    int pending, parity, halt, it;     // halt: 0 running, 1 converged, 2 line search needs the host, 3 not a descent direction
    int l_used, l_head;                // L-BFGS ring (limited-memory mode)
};
__device__ __forceinline__ bool dev_halted(const DevState* st) { return st->halt != 0; }
// the pending pair is the one stored by the PREVIOUS iteration: the buffer the current parity does not point at
__device__ __forceinline__ void dev_pending(const DevState* st, const double* s0, const double* s1, const double* hy0, const double* hy1,
                                            const double** s, const double** Hy, double* rho, double* c, int* apply) {
    *s = st->parity ? s0 : s1;
    *Hy = st->parity ? hy0 : hy1;
    *rho = st->rho;
    *c = st->cc;
    *apply = st->pending;
}

// fixed-order sum (or max) of per-CTA partials part[cta][4] by one warp; every CTA that calls it gets the same bits
__device__ __forceinline__ void sum_partials(const double* __restrict__ part, int nb, double* sh /* 4 doubles of shared memory */, bool last_is_max) {
    if (threadIdx.x < 32) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int b = threadIdx.x; b < nb; b += 32) {
            a0 += part[4 * b];
            a1 += part[4 * b + 1];
            a2 += part[4 * b + 2];
            a3 = last_is_max ? fmax(a3, part[4 * b + 3]) : a3 + part[4 * b + 3];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            const double t = __shfl_xor_sync(0xffffffffu, a3, o);
            a3 = last_is_max ? fmax(a3, t) : a3 + t;
        }
        if (threadIdx.x == 0) { sh[0] = a0; sh[1] = a1; sh[2] = a2; sh[3] = a3; }
    }
    __syncthreads();
}
// per-CTA reduction of four running values into part[blockIdx.x][4]
__device__ __forceinline__ void store_partials(double v0, double v1, double v2, double v3, bool last_is_max, double* __restrict__ part) {
    __shared__ double sd[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        const double t = __shfl_xor_sync(0xffffffffu, v3, o);
        v3 = last_is_max ? fmax(v3, t) : v3 + t;
    }
    if ((threadIdx.x & 31) == 0) {
        sd[0][threadIdx.x >> 5] = v0; sd[1][threadIdx.x >> 5] = v1; sd[2][threadIdx.x >> 5] = v2; sd[3][threadIdx.x >> 5] = v3;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t = (last_is_max && threadIdx.x == 3) ? fmax(t, sd[3][w]) : t + sd[threadIdx.x][w];
        part[4 * blockIdx.x + threadIdx.x] = t;
    }
}

// All kernels below: 256 threads, blocks(L) CTAs unless stated otherwise.
__global__ void __launch_bounds__(256) k_fast_trial(DevState* __restrict__ st, const double* __restrict__ x, const double* __restrict__ p,
                                                    double* __restrict__ xt, int n) {
    if (st->halt) return;
    const double df0 = st->df0;
    if (!(df0 < 0.0)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) st->halt = 3;
        return;
    }
    const double ap = st->alpha_prev;
    double a1 = ap * (st->df0_prev / df0);                        // Nocedal & Wright (3.60), same clamps as the host rule
    a1 = fmin(1.0, fmax(0.5 * ap, fmin(a1, 2.0 * ap)));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) xt[i] = fma(a1, p[i], x[i]);
    if (i == 0) st->a1 = a1;
}

// after the evaluation at xt (out = [cost, grad]): partial sums of g_new.p, s.y, y.y and max|g_new|  (s = xt - x, y = g_new - g)
__global__ void __launch_bounds__(256) k_fast_wolfe_part(const DevState* __restrict__ st, const double* __restrict__ x, const double* __restrict__ xt,
                                                         const double* __restrict__ g, const double* __restrict__ p, const double* __restrict__ out,
                                                         double* __restrict__ part, int n) {
    if (st->halt) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double df = 0.0, sy = 0.0, yy = 0.0, gm = 0.0;
    if (i < n) {
        const double gi = out[1 + i], si = xt[i] - x[i], yi = gi - g[i];
        df = gi * p[i];
        sy = si * yi;
        yy = yi * yi;
        gm = fabs(gi);
    }
    store_partials(df, sy, yy, gm, true, part);
}
// strong-Wolfe test of the trial step (every CTA takes the same decision from the same partials); on success s and y are
// written into the current pair buffers.  CTA 0 records the outcome in the state (none of those fields is read by this kernel).
__global__ void __launch_bounds__(256) k_fast_wolfe_apply(DevState* __restrict__ st, const double* __restrict__ x, const double* __restrict__ xt,
                                                          const double* __restrict__ g, const double* __restrict__ out, const double* __restrict__ part,
                                                          int nb, double* __restrict__ s0, double* __restrict__ s1, double* __restrict__ y, int n) {
    __shared__ double sh[4];
    if (st->halt) return;
    sum_partials(part, nb, sh, true);
    const double c1 = 1e-4, c2 = 0.9;
    const double f = out[0], f0 = st->f, df0 = st->df0, a = st->a1;
    const bool ok = (f <= f0 + c1 * a * df0) && (fabs(sh[0]) <= -c2 * df0) && (f <= f0);      // NaN fails every comparison
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (ok) {
            st->sy = sh[1]; st->yy = sh[2]; st->f_new = f; st->gmax_new = sh[3];
            st->df0_prev = df0; st->alpha_prev = a;
        } else {
            st->halt = 2;
        }
    }
    if (!ok) return;
    double* s = st->parity ? s1 : s0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        s[i] = xt[i] - x[i];
        y[i] = out[1 + i] - g[i];
    }
}

// u = sum of the tile partials of k_sym_update_gemv, Hy = u + p, partial sums of y.Hy, Hy.g_new, s.g_new, u.g_new
__global__ void __launch_bounds__(256) k_fast_dots(const DevState* __restrict__ st, const double* __restrict__ P, int ntb, const double* __restrict__ p,
                                                   const double* __restrict__ out, const double* __restrict__ s0, const double* __restrict__ s1,
                                                   double* __restrict__ hy0, double* __restrict__ hy1, const double* __restrict__ y,
                                                   double* __restrict__ u, double* __restrict__ part, int n) {
    if (st->halt) return;
    const double* s = st->parity ? s1 : s0;
    double* Hy = st->parity ? hy1 : hy0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double yHy = 0.0, hyg = 0.0, sg = 0.0, ug = 0.0;
    if (i < n) {
        const double ui = sym_gemv_sum(P, ntb, i);
        const double hy = ui + p[i];              // H y = H g_new - H g = u + p   (p = -H g)
        u[i] = ui;
        Hy[i] = hy;
        const double gi = out[1 + i];
        yHy = y[i] * hy;
        hyg = hy * gi;
        sg = s[i] * gi;
        ug = ui * gi;
    }
    store_partials(yHy, hyg, sg, ug, false, part);
}
// the direction of the updated matrix (the matrix itself is updated by the next pass over H), x <- xt, g <- g_new
__global__ void __launch_bounds__(256) k_fast_direction(const DevState* __restrict__ st, const double* __restrict__ part, int nb, double* __restrict__ x,
                                                        const double* __restrict__ xt, double* __restrict__ g, double* __restrict__ p,
                                                        const double* __restrict__ out, const double* __restrict__ u, const double* __restrict__ s0,
                                                        const double* __restrict__ s1, const double* __restrict__ hy0, const double* __restrict__ hy1, int n) {
    __shared__ double sh[4];
    if (st->halt) return;
    sum_partials(part, nb, sh, false);
    const double yHy = sh[0], hyg = sh[1], sg = sh[2];
    const double sy = st->sy;
    const bool upd = sy > 1e-300;
    const double r1 = upd ? 1.0 / sy : 0.0;
    const double rho = r1, cc = r1 * (1.0 + r1 * yHy);
    const double* s = st->parity ? s1 : s0;
    const double* Hy = st->parity ? hy1 : hy0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        p[i] = upd ? -(u[i] - rho * (s[i] * hyg + Hy[i] * sg) + cc * s[i] * sg) : -u[i];
        x[i] = xt[i];
        g[i] = out[1 + i];
    }
}
// the scalars of the finished iteration (one warp): the only kernel that advances parity / pending / f / it
__global__ void __launch_bounds__(32) k_fast_state(DevState* __restrict__ st, const double* __restrict__ part, int nb) {
    __shared__ double sh[4];
    if (st->halt) return;
    sum_partials(part, nb, sh, false);
    if (threadIdx.x == 0) {
        const double yHy = sh[0], hyg = sh[1], sg = sh[2], ug = sh[3];
        const double sy = st->sy;
        const bool upd = sy > 1e-300;
        const double r1 = upd ? 1.0 / sy : 0.0;
        const double rho = r1, cc = r1 * (1.0 + r1 * yHy);
        st->df0 = upd ? -(ug - rho * (sg * hyg + hyg * sg) + cc * sg * sg) : -ug;
        st->rho = rho; st->cc = cc;
        st->pending = upd ? 1 : 0;
        if (upd) st->parity ^= 1;
        st->f = st->f_new;
        st->gnorm = st->gmax_new;
        st->it += 1;
        if (st->gnorm <= st->gtol) st->halt = 1;
    }
}

// ---- single-CTA problems (L <= 256: molecule sizes): the same iteration in fewer launches ------------------------------------
// With one CTA every "per-CTA partial -> fixed-order sum" pair collapses into one block reduction, so k_fast_wolfe_part +
// k_fast_wolfe_apply become one kernel and k_fast_dots + k_fast_direction + k_fast_state + the NEXT iteration's k_fast_trial
// another: an iteration is evaluation, A, k_sym_update_gemv, B instead of eight launches.  The arithmetic and its order are
// those of the kernels above (block_sum4 adds in store_partials' order, and sum_partials over one CTA adds zeros), so the
// trajectory is bit-identical to the unfused path (MAMBO_OPT_NO_FUSE=1; tested).
__device__ __forceinline__ void block_sum4(double v0, double v1, double v2, double v3, bool last_is_max, double* sh /* 4 */) {
    __shared__ double sd4[4][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        const double t = __shfl_xor_sync(0xffffffffu, v3, o);
        v3 = last_is_max ? fmax(v3, t) : v3 + t;
    }
    __syncthreads();                                   // sd4 / sh may still be read from an earlier call
    if ((threadIdx.x & 31) == 0) {
        sd4[0][threadIdx.x >> 5] = v0; sd4[1][threadIdx.x >> 5] = v1; sd4[2][threadIdx.x >> 5] = v2; sd4[3][threadIdx.x >> 5] = v3;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t = (last_is_max && threadIdx.x == 3) ? fmax(t, sd4[3][w]) : t + sd4[threadIdx.x][w];
        // sum_partials over a single CTA: 0 + t (0 max t) in lane 0, zeros elsewhere
        sh[threadIdx.x] = (last_is_max && threadIdx.x == 3) ? fmax(0.0, t) : 0.0 + t;
    }
    __syncthreads();
}
// A: Wolfe sums + strong-Wolfe test + (on success) s, y            (1 CTA of 256 threads, n <= 256)
__global__ void __launch_bounds__(256) k_fast_small_a(DevState* st, const double* x, const double* xt, const double* g, const double* p,
                                                      const double* out, double* s0, double* s1, double* y, int n) {
    __shared__ double sh[4];
    if (st->halt) return;
    const int i = threadIdx.x;
    double df = 0.0, sy = 0.0, yy = 0.0, gm = 0.0;
    if (i < n) {
        const double gi = out[1 + i], si = xt[i] - x[i], yi = gi - g[i];
        df = gi * p[i];
        sy = si * yi;
        yy = yi * yi;
        gm = fabs(gi);
    }
    block_sum4(df, sy, yy, gm, true, sh);
    const double c1 = 1e-4, c2 = 0.9;
    const double f = out[0], f0 = st->f, df0 = st->df0, a = st->a1;
    const int parity = st->parity;
    const bool ok = (f <= f0 + c1 * a * df0) && (fabs(sh[0]) <= -c2 * df0) && (f <= f0);      // NaN fails every comparison
    __syncthreads();                                   // every thread has read the state before thread 0 changes it
    if (threadIdx.x == 0) {
        if (ok) {
            st->sy = sh[1]; st->yy = sh[2]; st->f_new = f; st->gmax_new = sh[3];
            st->df0_prev = df0; st->alpha_prev = a;
        } else {
            st->halt = 2;
        }
    }
    if (!ok) return;
    double* s = parity ? s1 : s0;
    if (i < n) {
        s[i] = xt[i] - x[i];
        y[i] = out[1 + i] - g[i];
    }
}
// B: u, Hy and their dots; direction, x <- xt, g <- g_new; the scalars of the finished iteration; the next trial point
__global__ void __launch_bounds__(256) k_fast_small_b(DevState* st, const double* P, int ntb, double* x, double* xt, double* g, double* p,
                                                      const double* out, double* u, const double* s0, const double* s1, double* hy0, double* hy1,
                                                      const double* y, int n) {
    __shared__ double sh[4];
    __shared__ double nxt[4];                          // halt, df0, alpha_prev, df0_prev of the NEW state
    if (st->halt) return;
    const int parity = st->parity;
    const double* s = parity ? s1 : s0;
    double* Hy = parity ? hy1 : hy0;
    const int i = threadIdx.x;
    double yHy = 0.0, hyg = 0.0, sg = 0.0, ug = 0.0, ui = 0.0, hy = 0.0, si = 0.0;
    if (i < n) {
        ui = sym_gemv_sum(P, ntb, i);
        hy = ui + p[i];                                // H y = H g_new - H g = u + p   (p = -H g)
        u[i] = ui;
        Hy[i] = hy;
        const double gi = out[1 + i];
        si = s[i];
        yHy = y[i] * hy;
        hyg = hy * gi;
        sg = si * gi;
        ug = ui * gi;
    }
    block_sum4(yHy, hyg, sg, ug, false, sh);
    const double syv = st->sy;
    const bool upd = syv > 1e-300;
    const double r1 = upd ? 1.0 / syv : 0.0;
    const double rho = r1, cc = r1 * (1.0 + r1 * sh[0]);
    if (i < n) {
        p[i] = upd ? -(ui - rho * (si * sh[1] + hy * sh[2]) + cc * si * sh[2]) : -ui;
        x[i] = xt[i];
        g[i] = out[1 + i];
    }
    __syncthreads();                                   // every thread has read the state; p and x are complete
    if (threadIdx.x == 0) {
        const double hygv = sh[1], sgv = sh[2], ugv = sh[3];
        const double df0n = upd ? -(ugv - rho * (sgv * hygv + hygv * sgv) + cc * sgv * sgv) : -ugv;
        st->df0 = df0n;
        st->rho = rho; st->cc = cc;
        st->pending = upd ? 1 : 0;
        if (upd) st->parity = parity ^ 1;
        st->f = st->f_new;
        const double gn = st->gmax_new;
        st->gnorm = gn;
        st->it += 1;
        int halt = 0;
        if (gn <= st->gtol) halt = 1;
        else if (!(df0n < 0.0)) halt = 3;              // what the next iteration's k_fast_trial would find
        if (halt) st->halt = halt;
        nxt[0] = (double)halt; nxt[1] = df0n; nxt[2] = st->alpha_prev; nxt[3] = st->df0_prev;
    }
    __syncthreads();
    if (nxt[0] != 0.0) return;
    // k_fast_trial of the next iteration
    const double ap = nxt[2];
    double a1 = ap * (nxt[3] / nxt[1]);                // Nocedal & Wright (3.60), same clamps as the host rule
    a1 = fmin(1.0, fmax(0.5 * ap, fmin(a1, 2.0 * ap)));
    if (i < n) xt[i] = fma(a1, p[i], x[i]);
    if (i == 0) st->a1 = a1;
}

// ---- limited-memory mode of the device-driven iterations ------------------------------------------------------------------
// the ring slot the pair of this iteration goes to, the number of pairs in use and H_0's scale, from the state BEFORE the update
__device__ __forceinline__ void lbfgs_next(const DevState* st, int mem, int* head, int* used, double* gamma) {
    const bool upd = st->sy > 1e-300;
    *head = upd ? (st->l_head + 1) % mem : (st->l_head < 0 ? 0 : st->l_head);
    *used = upd ? min(st->l_used + 1, mem) : st->l_used;
    *gamma = upd ? st->sy / st->yy : st->l_gamma;
}
// x <- xt, g <- g_new, and the accepted pair (s, y, 1 / s.y) into its ring slot
__global__ void __launch_bounds__(256) k_fast_lbfgs_store(const DevState* __restrict__ st, const double* __restrict__ s, const double* __restrict__ y,
                                                          double* __restrict__ LS, double* __restrict__ LY, double* __restrict__ Lrho, int mem,
                                                          double* __restrict__ x, const double* __restrict__ xt, double* __restrict__ g,
                                                          const double* __restrict__ out, int n) {
    if (st->halt) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = xt[i];
    g[i] = out[1 + i];
    if (st->sy > 1e-300) {
        int head, used;
        double gamma;
        lbfgs_next(st, mem, &head, &used, &gamma);
        LS[(size_t)head * n + i] = s[i];
        LY[(size_t)head * n + i] = y[i];
        if (i == 0) Lrho[head] = 1.0 / st->sy;
    }
}
// The two-loop recursion is 2 m dependent reductions over L: a single-CTA kernel that costs as much as a third of the evaluation
// at N = 152.  The device-driven iterations use the compact representation instead (Byrd, Nocedal & Schnabel 1994, eq. 3.1):
//     H = gamma I + [S  gamma Y] [ R^-T (D + gamma Y^T Y) R^-1   -R^-T ] [ S^T       ]
//                                [ -R^-1                          0    ] [ gamma Y^T ]
// with R the upper triangle of S^T Y (pairs ordered oldest to newest) and D its diagonal: ONE multi-CTA pass forms the Gram
// blocks S^T Y, Y^T Y and S^T g, Y^T g, g.g (per-CTA partials, fixed-order sum), one warp solves the two m x m triangular
// systems, and one more multi-CTA pass assembles p = -H g.  The directional derivative g.p comes out of the same scalars.
constexpr int LB_FAST_MAX = 24;      // largest memory the shared-memory Gram kernel takes (2 x 24 x 257 doubles)
constexpr int LB_W = 256, LB_WP = 257;
__host__ __device__ inline int lb_ldp(int mem) { return 2 * mem * mem + 2 * mem + 1; }
__device__ __forceinline__ int lb_slot(int head, int used, int mem, int k) { return (head - (used - 1) + k + 2 * mem) % mem; }   // k = 0: oldest

__global__ void __launch_bounds__(256) k_fast_lbfgs_gram(const DevState* __restrict__ st, const double* __restrict__ S, const double* __restrict__ Y,
                                                         const double* __restrict__ g, int mem, int n, double* __restrict__ part) {
    extern __shared__ double lsm[];
    if (st->halt) return;
    int head, used;
    double gamma;
    lbfgs_next(st, mem, &head, &used, &gamma);
    double* Ss = lsm;                              // [used][LB_WP]
    double* Ys = lsm + (size_t)mem * LB_WP;
    double* gs = lsm + (size_t)2 * mem * LB_WP;    // [LB_W]
    const int i = blockIdx.x * LB_W + threadIdx.x;
    for (int k = 0; k < used; ++k) {
        const int slot = lb_slot(head, used, mem, k);
        Ss[k * LB_WP + threadIdx.x] = i < n ? S[(size_t)slot * n + i] : 0.0;
        Ys[k * LB_WP + threadIdx.x] = i < n ? Y[(size_t)slot * n + i] : 0.0;
    }
    gs[threadIdx.x] = i < n ? g[i] : 0.0;
    __syncthreads();
    double* out = part + (size_t)blockIdx.x * lb_ldp(mem);
    const int nsy = used * used, total = 2 * nsy + 2 * used + 1;
    for (int e = threadIdx.x; e < total; e += 256) {
        const double *a, *b;
        if (e < nsy) { a = Ss + (e / used) * LB_WP; b = Ys + (e % used) * LB_WP; }                       // S^T Y [k][l]
        else if (e < 2 * nsy) { a = Ys + ((e - nsy) / used) * LB_WP; b = Ys + ((e - nsy) % used) * LB_WP; }   // Y^T Y [k][l]
        else if (e < 2 * nsy + used) { a = Ss + (e - 2 * nsy) * LB_WP; b = gs; }                         // S^T g
        else if (e < 2 * nsy + 2 * used) { a = Ys + (e - 2 * nsy - used) * LB_WP; b = gs; }              // Y^T g
        else { a = gs; b = gs; }                                                                         // g.g
        double acc = 0.0;
        for (int c = 0; c < LB_W; ++c) acc = fma(a[c], b[c], acc);
        out[e] = acc;
    }
}
// one CTA: fixed-order sum of the partials, the two triangular solves, the coefficients of p over the ring slots, state update
__global__ void __launch_bounds__(256) k_fast_lbfgs_solve(DevState* __restrict__ st, const double* __restrict__ part, int nb, int mem,
                                                          double* __restrict__ coef) {
    extern __shared__ double lsm[];
    if (st->halt) return;
    int head, used;
    double gamma;
    lbfgs_next(st, mem, &head, &used, &gamma);
    const int nsy = used * used, total = 2 * nsy + 2 * used + 1, ldp = lb_ldp(mem);
    double* G = lsm;                                 // the summed Gram data, layout of k_fast_lbfgs_gram
    double* a = lsm + ldp;                           // [mem]
    double* t = a + mem;                             // [mem]
    for (int e = threadIdx.x; e < total; e += 256) {
        double acc = 0.0;
        for (int b = 0; b < nb; ++b) acc += part[(size_t)b * ldp + e];
        G[e] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *SY = G, *YY = G + nsy, *u = G + 2 * nsy, *v = u + used;
        const double gg = G[2 * nsy + 2 * used];
        for (int k = used - 1; k >= 0; --k) {        // a = R^-1 u,  R = upper triangle of S^T Y
            double r = u[k];
            for (int l = k + 1; l < used; ++l) r -= SY[k * used + l] * a[l];
            a[k] = r / SY[k * used + k];
        }
        for (int k = 0; k < used; ++k) {             // t = (D + gamma Y^T Y) a - gamma v
            double r = SY[k * used + k] * a[k] - gamma * v[k];
            for (int l = 0; l < used; ++l) r = fma(gamma * YY[k * used + l], a[l], r);
            t[k] = r;
        }
        for (int k = 0; k < used; ++k) {             // top = R^-T t  (forward substitution), kept in t
            double r = t[k];
            for (int l = 0; l < k; ++l) r -= SY[l * used + k] * t[l];
            t[k] = r / SY[k * used + k];
        }
        double hgg = gamma * gg;                     // g.H g
        for (int k = 0; k < used; ++k) hgg += t[k] * u[k] - gamma * a[k] * v[k];
        for (int sl = 0; sl < 2 * mem; ++sl) coef[sl] = 0.0;
        for (int k = 0; k < used; ++k) {
            const int slot = lb_slot(head, used, mem, k);
            coef[slot] = t[k];                       // coefficient of S[slot] in H g
            coef[mem + slot] = -gamma * a[k];        // coefficient of Y[slot]
        }
        coef[2 * mem] = gamma;
        const bool upd = st->sy > 1e-300;
        st->df0 = -hgg;
        st->l_head = upd ? head : st->l_head;
        st->l_used = used;
        st->l_gamma = gamma;
        st->f = st->f_new;
        st->gnorm = st->gmax_new;
        st->it += 1;
        if (st->gnorm <= st->gtol) st->halt = 1;
    }
}
// p = -(gamma g + sum_slots cS S + cY Y); runs after the state has advanced (l_used / l_head are the new ring state)
__global__ void __launch_bounds__(256) k_fast_lbfgs_apply(const DevState* __restrict__ st, const double* __restrict__ coef, const double* __restrict__ S,
                                                          const double* __restrict__ Y, const double* __restrict__ g, int mem, int n,
                                                          double* __restrict__ p) {
    if (st->halt > 1) return;                        // (halt == 1: converged in this very iteration; p is not needed any more but harmless)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int used = st->l_used, head = st->l_head < 0 ? 0 : st->l_head;
    double acc = coef[2 * mem] * g[i];
    for (int k = 0; k < used; ++k) {
        const int slot = lb_slot(head, used, mem, k);
        acc = fma(coef[slot], S[(size_t)slot * n + i], acc);
        acc = fma(coef[mem + slot], Y[(size_t)slot * n + i], acc);
    }
    p[i] = -acc;
}

struct Opt {
    mambo_csa* h = nullptr;
    int L = 0, device = 0;
    cudaStream_t stream = nullptr;
    double *x = nullptr, *g = nullptr, *p = nullptr, *xt = nullptr, *out = nullptr, *s = nullptr, *y = nullptr, *Hy = nullptr,
           *s2 = nullptr, *Hy2 = nullptr, *u = nullptr, *H = nullptr, *scal = nullptr;
    double *LS = nullptr, *LY = nullptr, *Lrho = nullptr;   // L-BFGS ring (lbfgs_memory > 0)
    double* h_scal = nullptr;  // pinned (128 bytes of scalars, then the status ring of the device-driven iterations)
    DevState* dstate = nullptr;
    double *P = nullptr, *part = nullptr;   // tile partials of the symmetric H pass; per-CTA partial sums of the fast path
    double *lpart = nullptr, *lcoef = nullptr;   // limited-memory mode: per-CTA Gram partials, coefficients of p over the ring slots
    long long ldh = 0;
    int ntb = 0;
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
    const size_t ldh = ((size_t)L + 3) & ~(size_t)3, ntb = ((size_t)L + STB - 1) / STB;
    const size_t nb = ((size_t)L + 255) / 256;
    const size_t big = mem > 0 ? (2 * (size_t)mem * L + 64 + nb * lb_ldp(mem) + 2 * mem + 16) : ldh * L + ntb * ntb * STB;
    return 11 * Lp + 16 + 32 + 4 * (((size_t)L + 255) / 256) + 16 + big;      // + DevState and the per-CTA partials of the device-driven iterations
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
        // the trial point and the evaluation's output live in the handle's own buffers: no device-to-device copy per evaluation
        rc = mambo_csa_io_buffers(h, &o->xt, &o->out);
        if (rc) {
            snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
            return finish(rc);
        }
        o->scal = q;
        q += 16;
        o->dstate = reinterpret_cast<DevState*>(q);
        q += 32;                     // (256 bytes)
        o->part = q;
        q += 4 * (((size_t)L + 255) / 256) + 16;
        o->ldh = (long long)(((size_t)L + 3) & ~(size_t)3);
        o->ntb = (L + STB - 1) / STB;
        if (lbfgs) {
            o->LS = q;
            o->LY = q + (size_t)mem * L;
            o->Lrho = q + 2 * (size_t)mem * L;
            o->lpart = o->Lrho + 64;
            o->lcoef = o->lpart + (((size_t)L + 255) / 256) * lb_ldp(mem);
        } else {
            o->H = q;
            o->P = q + (size_t)o->ldh * L;
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
    if (!lbfgs) k_set_identity<<<1184, 256, 0, o->stream>>>(o->H, o->ldh, L, 1.0);

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
    // ---- device-driven iterations (see DevState): eligible in the reference's mode (dense BFGS, no self-scaling) on unsharded handles
    constexpr int RING = 4, DEPTH = 2;
    static_assert(sizeof(DevState) <= 192, "status ring slots are 192 bytes");
    const bool fast_ok = !selfscale && std::getenv("MAMBO_OPT_HOSTLOOP") == nullptr && (!lbfgs || mem <= LB_FAST_MAX);
    const size_t lb_smem = sizeof(double) * ((size_t)2 * mem * LB_WP + LB_W);
    if (lbfgs && fast_ok && lb_smem > 48 * 1024) {      // raised only: concurrent runs on lanes of this GPU may use other memories
        cudaFuncAttributes fa;
        if (CK(cudaFuncGetAttributes(&fa, (const void*)k_fast_lbfgs_gram), "kernel attributes")) return finish(MAMBO_ECUDA);
        if ((size_t)fa.maxDynamicSharedSizeBytes < lb_smem &&
            CK(cudaFuncSetAttribute(k_fast_lbfgs_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb_smem), "smem attribute"))
            return finish(MAMBO_ECUDA);
    }
    cudaEvent_t ring_ev[RING] = {nullptr, nullptr, nullptr, nullptr};
    auto ring_slot = [&](int k) { return reinterpret_cast<DevState*>(reinterpret_cast<char*>(o->h_scal) + 128 + 192 * (k % RING)); };
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() {
            for (int k = 0; k < RING; ++k)
                if (e[k]) cudaEventDestroy(e[k]);
        }
    } ev_guard{ring_ev};
    int fast_iterations = 0;
    // Molecule-sized handles (single-kernel evaluation): the eight launches of one device-driven iteration cost more host time
    // than the device needs to run them, so the iteration is captured once into a CUDA graph and replayed (every pointer in it
    // is owned by the handle or by this optimiser's workspace).
    const bool graph_iter = !o->sharded && mambo_csa_is_tiny(o->h) && std::getenv("MAMBO_OPT_NO_GRAPH") == nullptr;
    struct GraphGuard {
        cudaGraphExec_t exec = nullptr;
        ~GraphGuard() { if (exec) cudaGraphExecDestroy(exec); }
    } iter_graph;
    // one device-driven iteration: trial point, evaluation, Wolfe test, update + direction, state
    // single-CTA problems: fused kernels (k_fast_small_a / _b); the trial point of an iteration is then computed by the previous
    // iteration's B kernel, so a run of iterations starts with one standalone k_fast_trial
    const bool fuse_small = !lbfgs && !o->sharded && blocks(L) == 1 && std::getenv("MAMBO_OPT_NO_FUSE") == nullptr;
    auto enqueue_iteration = [&](bool capturing) -> int {
        const int nb = blocks(L);
        if (fuse_small) {
            int rc2 = capturing ? mambo_csa_enqueue_eval(o->h, o->xt, o->out) : mambo_csa_eval_device(o->h, o->xt, o->out);
            if (rc2) {
                snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
                return rc2;
            }
            k_fast_small_a<<<1, 256, 0, o->stream>>>(o->dstate, o->x, o->xt, o->g, o->p, o->out, o->s, o->s2, o->y, L);
            k_sym_update_gemv<<<dim3(o->ntb, o->ntb), 256, 0, o->stream>>>(o->H, o->ldh, L, o->dstate, o->s, o->s2, o->Hy, o->Hy2, 1.0, 0.0, 0.0, 0,
                                                                             o->out + 1, o->P, o->ntb);
            k_fast_small_b<<<1, 256, 0, o->stream>>>(o->dstate, o->P, o->ntb, o->x, o->xt, o->g, o->p, o->out, o->u, o->s, o->s2, o->Hy, o->Hy2,
                                                     o->y, L);
            return MAMBO_OK;
        }
        k_fast_trial<<<nb, 256, 0, o->stream>>>(o->dstate, o->x, o->p, o->xt, L);
        // (row-panel shards: the exchange inside the evaluation returns identical bits on every rank, every rank takes the
        // same decisions from them and enqueues the same number of evaluations, so the peer epochs stay aligned)
        int rc2 = o->sharded ? mambo_csa_eval_sharded_device(o->h, o->xt, o->out)
                             : (capturing ? mambo_csa_enqueue_eval(o->h, o->xt, o->out) : mambo_csa_eval_device(o->h, o->xt, o->out));
        if (rc2) {
            snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
            return rc2;
        }
        k_fast_wolfe_part<<<nb, 256, 0, o->stream>>>(o->dstate, o->x, o->xt, o->g, o->p, o->out, o->part, L);
        k_fast_wolfe_apply<<<nb, 256, 0, o->stream>>>(o->dstate, o->x, o->xt, o->g, o->out, o->part, nb, o->s, o->s2, o->y, L);
        if (lbfgs) {
            k_fast_lbfgs_store<<<nb, 256, 0, o->stream>>>(o->dstate, o->s, o->y, o->LS, o->LY, o->Lrho, mem, o->x, o->xt, o->g, o->out, L);
            k_fast_lbfgs_gram<<<nb, 256, lb_smem, o->stream>>>(o->dstate, o->LS, o->LY, o->g, mem, L, o->lpart);
            k_fast_lbfgs_solve<<<1, 256, sizeof(double) * (lb_ldp(mem) + 2 * mem), o->stream>>>(o->dstate, o->lpart, nb, mem, o->lcoef);
            k_fast_lbfgs_apply<<<nb, 256, 0, o->stream>>>(o->dstate, o->lcoef, o->LS, o->LY, o->g, mem, L, o->p);
        } else {
            k_sym_update_gemv<<<dim3(o->ntb, o->ntb), 256, 0, o->stream>>>(o->H, o->ldh, L, o->dstate, o->s, o->s2, o->Hy, o->Hy2, 1.0, 0.0, 0.0, 0,
                                                                             o->out + 1, o->P, o->ntb);
            k_fast_dots<<<nb, 256, 0, o->stream>>>(o->dstate, o->P, o->ntb, o->p, o->out, o->s, o->s2, o->Hy, o->Hy2, o->y, o->u, o->part, L);
            k_fast_direction<<<nb, 256, 0, o->stream>>>(o->dstate, o->part, nb, o->x, o->xt, o->g, o->p, o->out, o->u, o->s, o->s2, o->Hy, o->Hy2, L);
            k_fast_state<<<1, 32, 0, o->stream>>>(o->dstate, o->part, nb);
        }
        return MAMBO_OK;
    };
    // runs device-driven iterations until the device halts or the iteration budget is used; returns the halt code in *halt_out
    auto fast_run = [&](int* halt_out) -> int {
        for (int k = 0; k < RING; ++k)
            if (!ring_ev[k] && CK(cudaEventCreateWithFlags(&ring_ev[k], cudaEventDisableTiming), "event")) return MAMBO_ECUDA;
        DevState hs;
        std::memset(&hs, 0, sizeof(hs));
        hs.f = f; hs.df0 = df0; hs.alpha_prev = alpha_prev; hs.df0_prev = df0_prev; hs.gnorm = gnorm; hs.rho = rho; hs.cc = cc;
        hs.gtol = opt.g_tol; hs.pending = pending ? 1 : 0; hs.parity = (s_cur == o->s) ? 0 : 1; hs.halt = 0; hs.it = it;
        hs.l_used = l_used; hs.l_head = l_head; hs.l_gamma = l_gamma;
        DevState* up = ring_slot(RING - 1);                 // pinned staging for the upload (not in use: nothing is in flight)
        *up = hs;
        if (CK(cudaMemcpyAsync(o->dstate, up, sizeof(DevState), cudaMemcpyHostToDevice, o->stream), "H2D state")) return MAMBO_ECUDA;
        if (CK(cudaStreamSynchronize(o->stream), "sync")) return MAMBO_ECUDA;       // `up` is reused as a ring slot below
        int budget = opt.max_iter - it, enq = 0, seen = 0, halt = 0;
        DevState last = hs;
        if (fuse_small) k_fast_trial<<<blocks(L), 256, 0, o->stream>>>(o->dstate, o->x, o->p, o->xt, L);
        while (true) {
            if (enq < budget && enq - seen < DEPTH) {
                if (graph_iter) {
                    if (!iter_graph.exec) {
                        if (CK(cudaStreamBeginCapture(o->stream, cudaStreamCaptureModeThreadLocal), "begin capture")) return MAMBO_ECUDA;
                        const int rcc = enqueue_iteration(true);
                        cudaGraph_t g = nullptr;
                        const cudaError_t ce = cudaStreamEndCapture(o->stream, &g);
                        if (rcc) {
                            if (g) cudaGraphDestroy(g);
                            return rcc;
                        }
                        if (CK(ce, "end capture")) return MAMBO_ECUDA;
                        const cudaError_t ci = cudaGraphInstantiate(&iter_graph.exec, g, 0);
                        cudaGraphDestroy(g);
                        if (CK(ci, "graph instantiate")) return MAMBO_ECUDA;
                    }
                    if (CK(cudaGraphLaunch(iter_graph.exec, o->stream), "graph launch")) return MAMBO_ECUDA;
                    mambo_csa_count_launches(o->h, 1);
                } else {
                    const int rci = enqueue_iteration(false);
                    if (rci) return rci;
                }
                o->evals++;
                if (CK(cudaMemcpyAsync(ring_slot(enq), o->dstate, sizeof(DevState), cudaMemcpyDeviceToHost, o->stream), "D2H state") ||
                    CK(cudaEventRecord(ring_ev[enq % RING], o->stream), "event record"))
                    return MAMBO_ECUDA;
                ++enq;
                continue;
            }
            if (seen == enq) break;
            if (CK(cudaEventSynchronize(ring_ev[seen % RING]), "event sync")) return MAMBO_ECUDA;
            last = *ring_slot(seen);
            ++seen;
            if (last.halt && !halt) {
                halt = last.halt;
                budget = enq;                               // nothing more is enqueued; the iterations in flight are no-ops
            }
            if (opt.verbose && !last.halt && (last.it % opt.verbose == 0))
                fprintf(stderr, "[mambo bfgs] it %d f=%.10e |g|inf=%.3e alpha=%.3e evals=%d (device-driven)\n", last.it, last.f, last.gnorm,
                        last.alpha_prev, o->evals);
        }
        int rc2 = mambo_csa_status_fetch(o->h);             // ||K||_1 out of range / NaN in x during any of the evaluations
        if (!rc2 && CK(cudaStreamSynchronize(o->stream), "sync")) return MAMBO_ECUDA;
        if (!rc2) rc2 = mambo_csa_status_check(o->h);
        if (rc2) {
            snprintf(o->err, sizeof(o->err), "%s", mambo_last_error());
            return rc2;
        }
        fast_iterations += last.it - it;
        f = last.f; df0 = last.df0; alpha_prev = last.alpha_prev; df0_prev = last.df0_prev; gnorm = last.gnorm; rho = last.rho; cc = last.cc;
        tau = 1.0;
        pending = last.pending != 0;
        if (last.parity == 0) { s_cur = o->s; Hy_cur = o->Hy; s_prev = o->s2; Hy_prev = o->Hy2; }
        else { s_cur = o->s2; Hy_cur = o->Hy2; s_prev = o->s; Hy_prev = o->Hy; }
        l_used = last.l_used; l_head = last.l_head; l_gamma = last.l_gamma;
        it = last.it;
        *halt_out = halt;
        return MAMBO_OK;
    };
    for (; it < opt.max_iter; ++it) {
        if (gnorm <= opt.g_tol) {
            converged = 1;
            break;
        }
        if (fast_ok && (lbfgs || !fresh) && it > 0 && df0 < 0) {
            int halt = 0;
            rc = fast_run(&halt);
            if (rc) return finish(rc);
            if (halt == 1 || gnorm <= opt.g_tol) {
                converged = 1;
                break;
            }
            if (it >= opt.max_iter) break;
            // halt 2 / 3: this iteration needs the host's line search (or a restart); it runs below, then the fast path resumes
        }
        if (!(df0 < 0)) {  // not a descent direction: restart from steepest descent
            if (!lbfgs) k_set_identity<<<1184, 256, 0, o->stream>>>(o->H, o->ldh, L, 1.0);
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
                k_set_identity<<<1184, 256, 0, o->stream>>>(o->H, o->ldh, L, gamma);
                k_scale<<<blocks(L), 256, 0, o->stream>>>(o->u, o->out + 1, gamma, L);     // u  = H g_new
                k_scale<<<blocks(L), 256, 0, o->stream>>>(Hy_cur, o->y, gamma, L);          // Hy = H y
                fresh = false;
            } else {
                // one pass over (the lower triangle of) H: fold in the pending update of the previous iteration and form u = H g_new
                k_sym_update_gemv<<<dim3(o->ntb, o->ntb), 256, 0, o->stream>>>(o->H, o->ldh, L, nullptr, s_prev, nullptr, Hy_prev, nullptr, tau, rho, cc,
                                                                                 pending ? 1 : 0, o->out + 1, o->P, o->ntb);
                k_sym_reduce<<<blocks(L), 256, 0, o->stream>>>(o->P, o->ntb, L, o->u);
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
// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, and that load waits for kernels that are
// running in the context -- a shard spinning for a peer whose thread is stuck in such a load (several shards driven from one
// process on one GPU) would time out.  Touching every optimiser kernel here loads them before any run starts.
static void preload_optimizer_kernels() {
    cudaFuncAttributes a;
    const void* ks[] = {(const void*)k_axpy, (const void*)k_sub, (const void*)k_dot_absmax, (const void*)k_set_identity, (const void*)k_scale,
                        (const void*)k_sym_update_gemv, (const void*)k_sym_reduce, (const void*)k_scale_neg, (const void*)k_direction,
                        (const void*)k_multi_dot, (const void*)k_lbfgs_direction, (const void*)k_store_pair, (const void*)k_fast_trial,
                        (const void*)k_fast_wolfe_part, (const void*)k_fast_wolfe_apply, (const void*)k_fast_dots,
                        (const void*)k_fast_direction, (const void*)k_fast_state, (const void*)k_fast_lbfgs_store,
                        (const void*)k_fast_lbfgs_gram, (const void*)k_fast_lbfgs_solve, (const void*)k_fast_lbfgs_apply};
    for (const void* k : ks) cudaFuncGetAttributes(&a, k);
    cudaGetLastError();
}

int mambo_csa_optimize_reserve(mambo_csa* h, int lbfgs_memory) {
    if (!h) return MAMBO_EINVAL;
    {
        void* st = nullptr;
        int dev = 0;
        if (mambo_csa_get_stream(h, &st, &dev) == MAMBO_OK && cudaSetDevice(dev) == cudaSuccess) preload_optimizer_kernels();
    }
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

// Measurement aid: `reps` passes of the inverse-Hessian kernel of the dense BFGS (pending rank-2 update + H g over the lower-
// triangular tiles, then the fixed-order reduction) on an L x L identity, timed with CUDA events on a stream of its own.
// ms_per_pass: average device time; bytes_per_pass: the algorithmic HBM traffic 16 B x L (L + 1) / 2 (every stored element read
// and written once).
int mambo_debug_bfgs_pass(int device, int L, int reps, double* ms_per_pass, double* bytes_per_pass) {
    if (L < 1 || reps < 1 || !ms_per_pass) return MAMBO_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return MAMBO_ECUDA;
    const long long ldh = ((long long)L + 3) & ~3LL;
    const int ntb = (L + STB - 1) / STB;
    double *H = nullptr, *v = nullptr, *P = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = MAMBO_ECUDA;
    do {
        if (cudaMalloc((void**)&H, sizeof(double) * ldh * L) != cudaSuccess) { rc = MAMBO_ENOMEM; break; }
        if (cudaMalloc((void**)&v, sizeof(double) * 4 * (size_t)L) != cudaSuccess) { rc = MAMBO_ENOMEM; break; }
        if (cudaMalloc((void**)&P, sizeof(double) * (size_t)ntb * ntb * STB) != cudaSuccess) { rc = MAMBO_ENOMEM; break; }
        if (cudaStreamCreate(&st) != cudaSuccess || cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) break;
        k_set_identity<<<1184, 256, 0, st>>>(H, ldh, L, 1.0);
        cudaMemsetAsync(v, 0, sizeof(double) * 4 * (size_t)L, st);          // s, Hy, g, u := 0
        double *s = v, *Hy = v + L, *g = v + 2 * (size_t)L, *u = v + 3 * (size_t)L;
        for (int r = 0; r < 2; ++r) {
            k_sym_update_gemv<<<dim3(ntb, ntb), 256, 0, st>>>(H, ldh, L, nullptr, s, nullptr, Hy, nullptr, 1.0, 1e-3, 1e-3, 1, g, P, ntb);
            k_sym_reduce<<<blocks(L), 256, 0, st>>>(P, ntb, L, u);
        }
        cudaEventRecord(e0, st);
        for (int r = 0; r < reps; ++r) {
            k_sym_update_gemv<<<dim3(ntb, ntb), 256, 0, st>>>(H, ldh, L, nullptr, s, nullptr, Hy, nullptr, 1.0, 1e-3, 1e-3, 1, g, P, ntb);
            k_sym_reduce<<<blocks(L), 256, 0, st>>>(P, ntb, L, u);
        }
        cudaEventRecord(e1, st);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_pass = ms / reps;
        if (bytes_per_pass) *bytes_per_pass = 16.0 * (double)L * ((double)L + 1.0) / 2.0;
        rc = MAMBO_OK;
    } while (false);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    cudaFree(H); cudaFree(v); cudaFree(P);
    if (rc) mambo_set_error("mambo_debug_bfgs_pass failed");
    return rc;
}

}  // extern "C"
