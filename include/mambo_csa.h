/*
 * mambo_csa.h -- C-ABI of libmambo_b200.so: the B200-native engine for QuantumMAMBO's CSA / CSA_SD
 * fragment cost + analytic gradient (the per-iteration hot path of the greedy decomposition).
 *
 * The reference (pure Julia, /root/reference) has no FFI seam for this path; the boundary is placed at
 * the narrowest point it offers -- the two closures handed to Optim.jl and the residual update:
 *
 *   mambo_csa_cost / mambo_csa_grad / mambo_csa_cost_grad
 *        replace   cost(x), grad!(storage,x)      src/UTILS/decompose.jl:96-105   (CSA)
 *                  cost(x), SD_grad!(storage,x)   src/UTILS/decompose.jl:221-231  (CSA_SD)
 *        i.e. to_OP(CSA_x_to_F_FRAG(x)) + L2_partial_cost + gradient / gradient_csa_sd
 *        (fermionic.jl:74-92,413-427, unitaries.jl:31-46,254-276, cost.jl:16-25, gradient.jl:1-290)
 *   mambo_csa_subtract_fragment
 *        replaces  F_rem = F_rem - to_OP(frag)    decompose.jl:56, 182  and the resume replay :28-32, :154-158
 *   mambo_csa_residual_cost
 *        replaces  L2_partial_cost(F_rem)         decompose.jl:40, 166  (cost.jl:44-55)
 *   mambo_csa_reconstruct
 *        replaces  to_OP(frag)                    fermionic.jl:1-8
 *   mambo_csa_get_residual
 *        hands F_rem back to Julia (F_OP.mbts[2], F_OP.mbts[3])
 *
 * Conventions
 *   - All arrays are FP64.  `tbt` is the Julia Array{Float64,4} as it lies in memory: column-major,
 *     element (p,q,r,s) (0-based) at p + N*q + N^2*r + N^3*s.  `obt` is N x N column-major.
 *   - CSA     x = [lambda (N(N+1)/2, lower-tri row-major) ; theta (N(N-1)/2, upper-tri row-major)]
 *     CSA_SD  x = [lambda1 (N) ; lambda2 (N(N+1)/2) ; theta (N(N-1)/2)]          (fermionic.jl:413-427)
 *   - "cost" is the squared Frobenius norm (cost.jl:16-25); the gradient is the reference's expression,
 *     including its one-sided lambda gradient for tensors that are not (pq)<->(rs) symmetric.
 *   - Every function returns 0 on success, a negative MAMBO_E* code otherwise; mambo_last_error() gives the
 *     thread-local message.  Nothing throws or aborts across this boundary.
 *   - A handle is bound to one GPU and is not re-entrant; different handles may be driven concurrently
 *     from different host threads.  The caller owns every host buffer; buffers are borrowed per call.
 *   - There is no CPU fallback: if no CUDA device is usable, create fails with MAMBO_ECUDA.
 */
#ifndef MAMBO_CSA_H
#define MAMBO_CSA_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mambo_csa mambo_csa;

/* flavour */
#define MAMBO_CSA     0
#define MAMBO_CSA_SD  1

/* flags (mambo_csa_create): bits 0-1 choose how the symmetry of tbt is exploited */
#define MAMBO_SYM_AUTO     0u  /* detect on device (relative tolerance 1e-13) and pick the cheapest valid mode */
#define MAMBO_SYM_GENERAL  1u  /* no symmetry assumed: D and D^T passes (matches the reference for any tensor) */
#define MAMBO_SYM_PAIR     2u  /* assume tbt[p,q,r,s] == tbt[r,s,p,q]: one pass over the N^2 x N^2 matrix     */
#define MAMBO_SYM_FULL     3u  /* assume the 8-fold symmetry of real orbital integrals: packed M x M, M=N(N+1)/2 */
#define MAMBO_SYM_MASK     3u
#define MAMBO_TBT_ON_DEVICE 4u /* tbt/obt pointers given to create are device pointers on `device`            */

/* error codes */
#define MAMBO_OK        0
#define MAMBO_EINVAL   -1
#define MAMBO_ECUDA    -2
#define MAMBO_ENOMEM   -3
#define MAMBO_ESTATE   -4
#define MAMBO_ERANGE   -5   /* ||K||_1 needs more squarings than the engine provides */

typedef struct mambo_csa_info_t {
    int    N, flavour, mode;          /* mode: MAMBO_SYM_GENERAL / _PAIR / _FULL actually in use            */
    int    num_params;                /* length of x                                                        */
    int    rows;                      /* logical rows of the residual matrix (N^2 or N(N+1)/2)              */
    int    rows_padded;               /* rows padded to the 64-row panel                                    */
    int    shard, nshards;            /* row-panel shard owned by this handle                               */
    int    local_row_begin, local_row_end;
    int    partial_len;               /* doubles exchanged per evaluation when sharded: 1 + 2 N^2           */
    int    num_sms;
    int    launches_per_eval;         /* kernels launched by one fused cost+grad evaluation                 */
    double flops_per_eval;            /* DMMA flops executed by the fused kernel(s) per evaluation          */
    double tensor_bytes;              /* bytes of residual tensor resident on this GPU                      */
    double asym_pair, asym_pq;        /* measured relative asymmetries of the input tensor                  */
    int    hot_gemms;                 /* R x R x N products per pass of the hot kernel: 1 (single-GEMM formulation,
                                         fused::k_ty) or 2 (fused::k_fused, MAMBO_TWO_GEMM=1 / legacy paths)   */
    int    lanes_hint;                /* useful number of evaluation lanes: min(8, floor(4 SMs / chain CTAs)), >= 1 */
} mambo_csa_info_t;

/* --- lifetime ------------------------------------------------------------------------------------------ */
int  mambo_csa_create(mambo_csa** h, int N, int flavour, const double* tbt, const double* obt,
                      int device, unsigned flags);
/* Row-panel shard `shard` of `nshards`: this handle keeps only its rows of the residual. */
int  mambo_csa_create_sharded(mambo_csa** h, int N, int flavour, const double* tbt, const double* obt,
                              int device, unsigned flags, int shard, int nshards);
/* Evaluation lane: another handle on the same GPU that SHARES the device-resident residual of `parent` (tensor, one-body
 * residual, row tables) but owns its stream, operands, scratch and CUDA graphs.  Independent evaluations issued through
 * different lanes (multistart runs, one host thread per lane) overlap on the device: the latency-bound orbital-rotation
 * chains / epilogues of one lane fill the SMs while another lane's are waiting at a grid barrier.  A residual update
 * (mambo_csa_subtract_fragment) through any handle of the group is seen by all of them; the caller must not run it
 * concurrently with evaluations on sibling lanes.  The residual is freed when the last handle of the group is
 * destroyed.  Useful lane counts (mambo_csa_info_t.lanes_hint): 8 at N <= 100, 4 at N = 152.                       */
int  mambo_csa_create_lane(mambo_csa* parent, mambo_csa** lane);
void mambo_csa_destroy(mambo_csa* h);
int  mambo_csa_get_info(const mambo_csa* h, mambo_csa_info_t* info);
int  mambo_csa_num_params(const mambo_csa* h);
const char* mambo_last_error(void);

/* --- the closures (host buffers; H2D of x and D2H of cost/grad happen inside the call) ------------------ */
int  mambo_csa_cost(mambo_csa* h, const double* x, double* cost);
int  mambo_csa_grad(mambo_csa* h, const double* x, double* grad);
int  mambo_csa_cost_grad(mambo_csa* h, const double* x, double* cost, double* grad);
/* The same closure at K points x[k*L .. (k+1)*L) in one call (K independent runs of a multistart): point k is evaluated on
 * lanes[k mod nlanes] (a handle and lanes created from it, see mambo_csa_create_lane).  Every lane is kept busy: as soon as
 * point k has been collected, point k + nlanes is enqueued on the same lane while the other lanes are still computing.
 * costs: K doubles, grads: K*L doubles (either may be NULL).                                                          */
int  mambo_csa_cost_grad_batch(mambo_csa** lanes, int nlanes, int K, const double* x, double* costs, double* grads);

/* How the host closures wait for the device: 0 (default) spins in cudaStreamSynchronize -- lowest latency, one busy core per
 * caller; 1 sleeps on a cudaEventBlockingSync event -- for callers that outnumber the host cores (lanes x ranks on one box).
 * Lanes inherit the mode of their parent at creation; MAMBO_BLOCKING_WAIT=1 in the environment makes 1 the default.       */
int  mambo_csa_set_wait_mode(mambo_csa* h, int blocking);

/* --- greedy driver state (device-resident residual) ------------------------------------------------------ */
/* On a row-panel shard both return the shard-local part of the cost (the caller sums over shards; the replicated
 * one-body term of CSA_SD is counted by shard 0 only).                                                          */
int  mambo_csa_subtract_fragment(mambo_csa* h, const double* x, double* new_cost);
int  mambo_csa_residual_cost(mambo_csa* h, double* cost);
int  mambo_csa_get_residual(mambo_csa* h, double* tbt_out, double* obt_out);
int  mambo_csa_reconstruct(mambo_csa* h, const double* x, double* tbt_out, double* obt_out);

/* --- final norms and the checkpoint block --------------------------------------------------------------- */
/* H.mbts[2] + ob_correction(H) for the RESIDENT residual (lcu.jl:262-267 with fermionic.jl:457-479): out = obt +
 * (spin_orb ? 1 : 2) sum_r tbt[:,:,r,r], N x N column-major.  obt is the handle's one-body residual (CSA_SD) or obt_in (CSA
 * handles; may be NULL).  One pass over N of the residual's rows.  The N x N `eigen` of to_OBF stays with the caller:
 * one_body_L1 = sum |eigenvalues| (x 1/2 for spin orbitals).                                                             */
int  mambo_csa_one_body_matrix(mambo_csa* h, int spin_orb, const double* obt_in, double* out);
/* CSA_L1(frag) (lcu.jl:139-182) from the fragment's parameter vector x (flavour: MAMBO_CSA / MAMBO_CSA_SD).            */
int  mambo_csa_frag_l1(int N, int flavour, int spin_orb, const double* x, double* l1);
/* The checkpoint dataset "<group>/x" (decompose.jl:58-64: Float64 tot_L x alpha, column-major) as a raw file: "MAMBOX01",
 * 16-byte group name, int64 tot_L, int64 alpha, the doubles.  Written beside the target and renamed.  INTEGRATION.md has
 * the Julia lines that h5write it (no libhdf5 in this image).  read: X == NULL reads the header only.                   */
int  mambo_csa_write_x_block(const char* path, const char* group, long long tot_L, long long alpha, const double* X);
int  mambo_csa_read_x_block(const char* path, char* group16, long long* tot_L, long long* alpha, double* X, long long cap);

/* --- initial guess (guesses.jl:51-84) ------------------------------------------------------------------- */
/* The eigenpair of largest |eigenvalue| of Symmetric(reshape(tbt_residual, (N^2, N^2))) -- what tbt_svd_1st keeps of
 * its O(N^6) `eigen` (guesses.jl:55-63) -- by Lanczos with full re-orthogonalisation on the device-resident residual
 * (HBM-bound matrix-vector products).  vec receives the eigenvector as the N x N column-major operator
 * `reshape(U[:,1], (n,n))`, unit Frobenius norm, largest-magnitude entry positive.  tol <= 0: 1e-13 (relative Ritz
 * residual); max_iter <= 0: 600 Krylov vectors.  The O(N^3) remainder of tbt_svd_1st (eigen of the N x N operator,
 * SVD_to_CSA, log of the rotation) stays with the caller.                                                       */
int  mambo_csa_leading_eig(mambo_csa* h, double tol, int max_iter, double* eigval, double* vec,
                           int* iterations, double* residual);

/* y = T~ v in the engine's row space (device pointers, rows_padded doubles each, see mambo_csa_get_info; in the packed
 * mode row (p<=q) carries sqrt(2 - delta_pq) v[p,q]).  One pass over the resident residual: 8 rows^2 bytes, HBM-bound. */
int  mambo_csa_symv_device(mambo_csa* h, const double* d_v, double* d_y);
/* Host half of the Lanczos driver (no GPU): largest-|eigenvalue| Ritz pair of the m x m symmetric tridiagonal with
 * diagonal alpha[0..m) and off-diagonal beta[0..m-1) (bisection + inverse iteration); z receives the unit vector.   */
int  mambo_debug_tridiag_leading(const double* alpha, const double* beta, int m, double* theta, double* z);
/* Host half of the double factorisation (no GPU): unit eigenvectors of a cluster of k numerically coincident eigenvalues
 * theta[0..k) (ascending) of the unreduced m x m tridiagonal; Z receives k rows of m doubles (orthonormal).            */
int  mambo_debug_tridiag_cluster(const double* alpha, const double* beta, int m, const double* theta, int k, double* Z);

/* --- device-resident entry points (no host synchronisation; used by the multi-GPU host and bench) -------- */
/* Run all engine work on `stream` (a cudaStream_t); NULL restores the handle's own stream (so the legacy
 * default stream, whose handle is 0, cannot be selected: pass a created stream). */
int  mambo_csa_set_stream(mambo_csa* h, void* stream);
int  mambo_csa_get_stream(mambo_csa* h, void** stream, int* device);
/* d_x: num_params doubles on the device; d_out: 1 + num_params doubles [cost, grad...] on the device. */
int  mambo_csa_eval_device(mambo_csa* h, const double* d_x, double* d_out);
/* Sharded evaluation in two halves around the caller's allreduce(sum) of d_partial (partial_len doubles):
 *   partial = [cost_2body, S (N x N), G (N x N)] summed over this shard's rows.                        */
int  mambo_csa_eval_partial_device(mambo_csa* h, const double* d_x, double* d_partial);
int  mambo_csa_eval_finish_device(mambo_csa* h, const double* d_partial, double* d_out);
/* The same exchange without a host-launched collective (one process per GPU): every shard stores its partial vector
 * into a slot of every peer's receive buffer over NVLink peer memory and publishes an epoch flag; the finish half waits
 * for all flags and adds the slots in rank order (deterministic, identical on every rank).
 *   setup : peer_export on every rank -> exchange the 64-byte handles (e.g. all_gather) -> peer_attach(rank, handle) for
 *           every other rank (peer_attach_local when the peer handle lives in the same process).
 *   eval  : eval_sharded_device == eval_sharded_push_device + eval_sharded_finish_device.                            */
int  mambo_csa_peer_export(mambo_csa* h, unsigned char* handle64);
int  mambo_csa_peer_attach(mambo_csa* h, int rank, const unsigned char* handle64);
int  mambo_csa_peer_attach_local(mambo_csa* h, int rank, mambo_csa* peer);
/* How eval_sharded_finish waits for the peers' flags: 0 (default) the sum kernel polls them and gives up after ~4 s (a dead peer
 * becomes MAMBO_ESTATE); 1 the STREAM waits (cuStreamWaitValue64, fetched from the driver at run time) and no SM is held while
 * waiting -- required when several sharded evaluations are in flight per GPU (lanes created from a shard handle, each with its
 * own peer_export / peer_attach): a spinning kernel holds its SM and the hot kernel needs every SM, so spinning sums of different
 * lanes could wait for each other across two GPUs.                                                                            */
int  mambo_csa_set_peer_wait(mambo_csa* h, int stream_ordered);
int  mambo_csa_eval_sharded_push_device(mambo_csa* h, const double* d_x);
int  mambo_csa_eval_sharded_finish_device(mambo_csa* h, double* d_out);
int  mambo_csa_eval_sharded_device(mambo_csa* h, const double* d_x, double* d_out);
/* Cost-only evaluation (half the flops: reconstruction + residual norm, no gradient). */
int  mambo_csa_cost_device(mambo_csa* h, const double* d_x, double* d_cost);
int  mambo_csa_synchronize(mambo_csa* h);
/* Live profiling: CUDA events bracket every fused-kernel launch on the launching stream.  get_profile
 * synchronises, returns the summed device time (ms) and count of those launches since the last call, and
 * the handle's running total of kernel launches.                                                         */
int  mambo_csa_set_profiling(mambo_csa* h, int enable);
int  mambo_csa_get_profile(mambo_csa* h, double* fused_ms, int* fused_launches, long long* total_launches);
/* Per-segment device time (ms) summed over the evaluations profiled since the last call:
 * seg_ms[0] orbital-rotation chain + operand build, [1] fused kernel(s) + row epilogues, [2] adjoint Frechet
 * chain + gradient assembly.  Call before mambo_csa_get_profile (which resets the event pools).            */
int  mambo_csa_get_segments(mambo_csa* h, double* seg_ms, int* evals);

/* --- multistart scheduler: K independent BFGS runs, one handle per GPU, host thread each ----- */
typedef struct mambo_opt_options_t {
    int    max_iter;      /* default 1000 (Optim.jl default)                                   */
    double g_tol;         /* default 1e-8  (Optim.jl default, inf-norm of the gradient)        */
    int    lbfgs_memory;  /* 0 = dense BFGS (reference behaviour), >0 = L-BFGS with that memory */
    int    verbose;
} mambo_opt_options_t;

typedef struct mambo_opt_result_t {
    double minimum;
    int    iterations, evaluations, converged, device;
} mambo_opt_result_t;

/* Minimise cost(x) from x0 on one handle with the built-in BFGS (strong-Wolfe line search).  On a row-panel shard
 * (after the peer exchange has been set up) every rank calls this with the same x0 on its own shard handle: the
 * evaluations exchange their partial vectors inside the library and return identical bits on every rank, so the
 * replicated optimiser states stay in lock-step without any host collective.                                   */
int  mambo_csa_optimize(mambo_csa* h, const double* x0, double* x_min, const mambo_opt_options_t* opt,
                        mambo_opt_result_t* res);
/* Allocate up front what mambo_csa_optimize needs on this handle for the given mode (device scratch, pinned scalars).
 * Allocation synchronises the device, so with several handles on one GPU (lanes, or shards driven from one process) do
 * this before the runs start: a shard spinning for its peers must never be waited for by a peer's cudaMalloc.       */
int  mambo_csa_optimize_reserve(mambo_csa* h, int lbfgs_memory);
/* K starts x0[k*L .. (k+1)*L) distributed over the given handles (one worker thread per handle);
 * x_min receives the K minimisers, res the K results; *best receives argmin_k res[k].minimum. */
int  mambo_multistart_run(mambo_csa** handles, int nhandles, int K, const double* x0, double* x_min,
                          const mambo_opt_options_t* opt, mambo_opt_result_t* res, int* best);

/* Measurement aid: average device time (CUDA events) of one pass of the dense BFGS's inverse-Hessian kernel at L parameters
 * (pending rank-2 update + H g over the lower-triangular 128 x 128 tiles) and its algorithmic HBM bytes, 16 B x L (L + 1) / 2. */
int  mambo_debug_bfgs_pass(int device, int L, int reps, double* ms_per_pass, double* bytes_per_pass);

/* --- double factorisation of the RESIDENT residual ------------------------------------------------------
 * Replaces DF_decomposition (decompose.jl:297-398, do_Givens = false) and DF_based_greedy (decompose.jl:400-436).
 * The reference diagonalises the whole N^2 x N^2 view with LAPACK (:326-330), keeps the eigenpairs with |eigenvalue| >= tol
 * ordered by decreasing magnitude (:332-350, SVD_tol = 1e-8) and diagonalises every reshaped eigenvector (:352-354).  Here:
 * Lanczos tridiagonalisation with full re-orthogonalisation on the device-resident matrix (restarted on breakdown, so
 * degenerate eigenvalues and low-rank tensors are handled), bisection + twisted factorisation for the tridiagonal, a
 * tensor-core GEMM back to the row space, and a batched Jacobi eigensolver for the N x N fragment matrices.  The result
 * stays on the device, owned by the handle, until the next call or destroy.                                              */
typedef struct mambo_df_stats_t {
    int    krylov_dim;        /* Lanczos steps (= dimension of the tridiagonal)                                       */
    int    restarts;          /* breakdowns of the recurrence (each opens a new unreduced block)                      */
    int    num_frags;         /* eigenpairs kept: |eigenvalue| >= tol                                                 */
    int    anti_frags;        /* fragments whose matrix is anti-symmetric (the Hermitian(1im * L) branch, :356-361);
                                 their coefficient is already negated, omega / U are left zero: the complex N x N eigen
                                 stays with the caller                                                                */
    int    jacobi_sweeps_max; /* most sweeps any fragment needed (negative: some fragment did not converge)           */
    int    cluster_vectors;   /* tridiagonal eigenvectors recomputed on the host because their eigenvalues coincide to
                                 rounding inside one block (inverse iteration + Gram-Schmidt, like LAPACK's dstein)      */
    double ortho_dev_before;  /* max |Z^T Z - I| of the tridiagonal eigenvectors before / after the Newton-Schulz     */
    double ortho_dev_after;   /*   re-orthonormalisation                                                              */
    double max_rel_residual;  /* max ||T y - theta y|| / max|theta| over the kept eigenpairs (only with verify != 0)  */
    double seconds_lanczos, seconds_tridiag, seconds_backtransform, seconds_fragments;
} mambo_df_stats_t;
/* tol <= 0: the reference's SVD_tol = 1e-8.  max_frags <= 0: no limit, otherwise only the max_frags leading eigenpairs are
 * turned into fragments.  verify != 0 additionally measures the eigen-residuals with one matvec per fragment.           */
int  mambo_csa_df_decompose(mambo_csa* h, double tol, int max_frags, int verify, int* num_frags, mambo_df_stats_t* stats);
/* Fragments [first, first + count) of the last decomposition, in the reference's order (decreasing |eigenvalue|); any
 * output may be NULL.  coeff[count]: F_FRAG.coeff; omega[count][N]: cartan_1b lambda (ascending, like LAPACK);
 * U[count][N x N]: f_matrix_rotation.mat, column-major; L[count][N x N]: the fragment's matrix, column-major
 * (to_OP(frag).mbts[3] = coeff * L (x) L, fermionic.jl:94-116); kind[count]: 0 symmetric, 1 anti-symmetric.             */
int  mambo_csa_df_get(mambo_csa* h, int first, int count, double* coeff, double* omega, double* U, double* L, int* kind);
/* DF_based_greedy (decompose.jl:400-436) from the last decomposition: lambda (N(N+1)/2, lower-triangular row-major, the
 * cartan_mat_to_2b order), U1 (N x N column-major: the frame of the leading fragment), Lmat (N x N, may be NULL).        */
int  mambo_csa_df_based_greedy(mambo_csa* h, double* lambda, double* U1, double* Lmat);
/* C[M x N] = op(A) B on the device with the library's FP64 tensor-core GEMM (row-major, leading dimensions in doubles and
 * even, 16-byte aligned pointers; trans_a != 0: A is stored K x M).  Exported for tests and bandwidth/flop measurements. */
int  mambo_dgemm_device(int device, void* stream, int trans_a, int M, int N, int K, const double* d_A, long long lda,
                        const double* d_B, long long ldb, double* d_C, long long ldc);

#ifdef __cplusplus
}
#endif
#endif /* MAMBO_CSA_H */
