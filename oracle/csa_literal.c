/* CPU ORACLE, C flavour (test infrastructure, not product code).
 *
 * A literal, single-threaded restatement of the reference's loop nests for the CSA cost and its gradient:
 *   cost      : decompose.jl:96-99 -> fermionic.jl:20-32,74-83 -> unitaries.jl:254-260 (N^6 einsum) -> cost.jl:16-25
 *   grad_lr   : gradient.jl:5-41      (cL x N^4)
 *   del_w_u   : gradient.jl:57-74     (dense N^6 array, 4 product-rule terms)
 *   grad_theta: gradient.jl:131-146   (uL x N^6 contraction with dU/dtheta_k)
 * dU/dtheta_k (gradient.jl:76-129, eigen-based) and U = exp(K) (unitaries.jl:31-46, LAPACK) are passed in by the
 * caller (numpy side) -- LAPACK is not restated in C.  Used (a) as a second pin of oracle/csa_oracle.py and (b) as
 * the "reference-literal" CPU timing of BASELINE.md section 2 (N = 6..18; N = 100 would need 8 TB of scratch).
 * PARITY UNPINNED: the reference has no golden vectors for this path (see oracle/csa_oracle.py header).
 *
 * Arrays are C-ordered with the Julia index order kept: tbt[p][q][r][s], U[a][l], wu[p][q][r][s][a][b].
 */
#include <stdlib.h>
#include <string.h>

#define IDX4(n, p, q, r, s) ((((size_t)(p) * (n) + (q)) * (n) + (r)) * (n) + (s))

/* unitaries.jl:254-260 with tbt[l,l,m,m] = Lam[l][m] (fermionic.jl:20-32) */
void csa_lit_reconstruct(int n, const double* U, const double* Lam, double* W) {
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
            for (int c = 0; c < n; ++c)
                for (int d = 0; d < n; ++d) {
                    double s = 0.0;
                    for (int l = 0; l < n; ++l)
                        for (int m = 0; m < n; ++m)
                            s += U[a * n + l] * U[b * n + l] * U[c * n + m] * U[d * n + m] * Lam[l * n + m];
                    W[IDX4(n, a, b, c, d)] = s;
                }
}

/* cost.jl:16-25 */
double csa_lit_cost(int n, const double* T, const double* W) {
    const size_t n4 = (size_t)n * n * n * n;
    double c = 0.0;
    for (size_t i = 0; i < n4; ++i) {
        const double d = T[i] - W[i];
        c += d * d;
    }
    return c;
}

/* gradient.jl:5-41: cartan_grad[idx(i>=j)] = sum(2 diff .* del_wr_lr(i,j)) * (i != j ? 2 : 1) */
void csa_lit_grad_lambda(int n, const double* U, const double* diff, double* g) {
    int idx = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double cl = 0.0;
            for (int p = 0; p < n; ++p)
                for (int q = 0; q < n; ++q)
                    for (int r = 0; r < n; ++r)
                        for (int s = 0; s < n; ++s)
                            cl += 2.0 * diff[IDX4(n, p, q, r, s)] * U[p * n + i] * U[q * n + i] * U[r * n + j] * U[s * n + j];
            if (i != j) cl *= 2.0;
            g[idx++] = cl;
        }
}

/* gradient.jl:57-74: wu[p,q,r,s,a,b]; returns a malloc'ed n^6 array (caller frees with csa_lit_free) */
double* csa_lit_del_w_u(int n, const double* U, const double* Lam) {
    const size_t n2 = (size_t)n * n, n6 = n2 * n2 * n2;
    double* wu = (double*)calloc(n6, sizeof(double));
    if (!wu) return NULL;
#define WU(p, q, r, s, a, b) wu[(IDX4(n, p, q, r, s) * n + (a)) * n + (b)]
    for (int p = 0; p < n; ++p)
        for (int q = 0; q < n; ++q)
            for (int r = 0; r < n; ++r)
                for (int s = 0; s < n; ++s)
                    for (int b = 0; b < n; ++b) {
                        double t_rs = 0.0, t_pq = 0.0;
                        for (int m = 0; m < n; ++m) {
                            t_rs += U[r * n + m] * U[s * n + m] * Lam[b * n + m];
                            t_pq += U[p * n + m] * U[q * n + m] * Lam[m * n + b];
                        }
                        WU(p, q, r, s, p, b) += U[q * n + b] * t_rs;
                        WU(p, q, r, s, q, b) += U[p * n + b] * t_rs;
                        WU(p, q, r, s, r, b) += t_pq * U[s * n + b];
                        WU(p, q, r, s, s, b) += t_pq * U[r * n + b];
                    }
    return wu;
}

/* gradient.jl:131-146: u_grad[k] = 2 sum(diff .* (wu . u_theta_k)); dU holds the uL matrices dU/dtheta_k (n x n each) */
void csa_lit_grad_theta(int n, const double* wu, const double* dU, int uL, const double* diff, double* g) {
    const size_t n4 = (size_t)n * n * n * n, n2 = (size_t)n * n;
    for (int k = 0; k < uL; ++k) {
        const double* ut = dU + (size_t)k * n2;
        double tot = 0.0;
        for (size_t i = 0; i < n4; ++i) {
            const double* w = wu + i * n2;
            double wt = 0.0;
            for (size_t ab = 0; ab < n2; ++ab) wt += w[ab] * ut[ab];
            tot += diff[i] * wt;
        }
        g[k] = 2.0 * tot;
    }
}

/* ---- CSA_SD (decompose.jl:221-231 -> fermionic.jl:58-72,85-92 -> unitaries.jl:266-272,314-320,366-384; gradient.jl:157-186,218-284) ---- */

/* unitaries.jl:266-272: the GENERAL 4-index rotation the CSA_SD representer applies to the sparse cartan tensor (N^8 loop) */
void csa_lit_tbt_rotation(int n, const double* U, const double* tbt, double* W) {
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
            for (int c = 0; c < n; ++c)
                for (int d = 0; d < n; ++d) {
                    double s = 0.0;
                    for (int l = 0; l < n; ++l)
                        for (int k = 0; k < n; ++k)
                            for (int m = 0; m < n; ++m)
                                for (int p = 0; p < n; ++p)
                                    s += U[a * n + l] * U[b * n + k] * U[c * n + m] * U[d * n + p] * tbt[IDX4(n, l, k, m, p)];
                    W[IDX4(n, a, b, c, d)] = s;
                }
}

/* unitaries.jl:314-320 */
void csa_lit_obt_rotation(int n, const double* U, const double* obt, double* h) {
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            double s = 0.0;
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) s += U[a * n + i] * U[b * n + j] * obt[i * n + j];
            h[a * n + b] = s;
        }
}

/* cost.jl:21-25: one-body part of L2_partial_cost */
double csa_lit_cost_ob(int n, const double* hT, const double* h) {
    double c = 0.0;
    for (int i = 0; i < n * n; ++i) {
        const double d = hT[i] - h[i];
        c += d * d;
    }
    return c;
}

/* gradient.jl:157-186: cartan_ob_grad[k] = sum(2 diff_ob .* (U[:,k] U[:,k]')) */
void csa_lit_grad_l_ob(int n, const double* U, const double* diff_ob, double* g) {
    for (int k = 0; k < n; ++k) {
        double cl = 0.0;
        for (int r = 0; r < n; ++r)
            for (int s = 0; s < n; ++s) cl += 2.0 * diff_ob[r * n + s] * U[r * n + k] * U[s * n + k];
        g[k] = cl;
    }
}

/* gradient.jl:218-231: hu[r,s,a,j] (n^4), a = r and a = s terms */
void csa_lit_del_h_u(int n, const double* U, const double* lam1, double* hu) {
    memset(hu, 0, sizeof(double) * (size_t)n * n * n * n);
    for (int r = 0; r < n; ++r)
        for (int s = 0; s < n; ++s)
            for (int j = 0; j < n; ++j) {
                hu[IDX4(n, r, s, r, j)] += U[s * n + j] * lam1[j];
                hu[IDX4(n, r, s, s, j)] += U[r * n + j] * lam1[j];
            }
}

/* gradient.jl:266-284: u_grad[k] = 2 (sum(diff_tb .* (wu . u_theta_k)) + sum(diff_ob .* (hu . u_theta_k))) */
void csa_lit_grad_theta_sd(int n, const double* wu, const double* hu, const double* dU, int uL, const double* diff_tb,
                           const double* diff_ob, double* g) {
    const size_t n4 = (size_t)n * n * n * n, n2 = (size_t)n * n;
    for (int k = 0; k < uL; ++k) {
        const double* ut = dU + (size_t)k * n2;
        double tot = 0.0;
        for (size_t i = 0; i < n4; ++i) {
            const double* w = wu + i * n2;
            double wt = 0.0;
            for (size_t ab = 0; ab < n2; ++ab) wt += w[ab] * ut[ab];
            tot += diff_tb[i] * wt;
        }
        for (size_t pq = 0; pq < n2; ++pq) {
            const double* hh = hu + pq * n2;
            double ht = 0.0;
            for (size_t ab = 0; ab < n2; ++ab) ht += hh[ab] * ut[ab];
            tot += diff_ob[pq] * ht;
        }
        g[k] = 2.0 * tot;
    }
}

void csa_lit_free(void* p) { free(p); }
