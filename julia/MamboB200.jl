# MamboB200.jl -- drop-in binding of libmambo_b200.so for QuantumMAMBO.jl (ccall, no extra Julia packages).
#
# Include this file after `using QuantumMAMBO` (or from inside src/QuantumMAMBO.jl after UTILS/decompose.jl).
# It redefines the two greedy steps so that the closures handed to Optim.jl evaluate on the B200:
#     CSA_greedy_step      src/UTILS/decompose.jl:82-120
#     CSA_SD_greedy_step   src/UTILS/decompose.jl:207-246
# and adds device-resident variants of the outer loops (decompose.jl:1-80, 122-205) that keep F_rem on the GPU.
# Everything else -- F_OP / F_FRAG, CSA_x_to_F_FRAG, Optim.jl BFGS, HDF5 checkpoints, L1 post-processing -- is
# untouched.  NOTE: this file could not be executed in the build image (no julia binary); tests/ drive the same
# C-ABI through Python ctypes with identical memory layout (column-major arrays).
module MamboB200

using QuantumMAMBO
import QuantumMAMBO: F_OP, F_FRAG, CSA_x_to_F_FRAG, CSA_SD_x_to_F_FRAG, cartan_2b_num_params,
                     real_orbital_rotation_num_params, tbt_svd_1st, SVD_to_CSA, SVD_for_CSA, SVD_for_CSA_SD, SVD_tiny, SVD_tol,
                     DECOMPOSITION_PRINT, ϵ, CSA, DF, cartan_1b, cartan_2b, f_matrix_rotation, c_matrix_rotation,
                     DF_decomposition, DF_based_greedy
using Optim
using LinearAlgebra

const LIB = get(ENV, "MAMBO_B200_LIB", joinpath(@__DIR__, "..", "quantummambo.jl_b200", "csrc", "libmambo_b200.so"))

const MAMBO_CSA = Cint(0)
const MAMBO_CSA_SD = Cint(1)

lasterr() = unsafe_string(ccall((:mambo_last_error, LIB), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error("libmambo_b200 error $rc: $(lasterr())")

mutable struct Engine
    h::Ptr{Cvoid}
    L::Int
end

# flags of mambo_csa_create (include/mambo_csa.h): how the symmetry of F.mbts[3] is exploited
const MAMBO_SYM_AUTO    = Cuint(0)   # detect on the device (relative tolerance 1e-13), pick the cheapest valid mode
const MAMBO_SYM_GENERAL = Cuint(1)   # assume nothing: matches the reference for ANY tensor (incl. its one-sided lambda gradient)
const MAMBO_SYM_PAIR    = Cuint(2)   # tbt[p,q,r,s] == tbt[r,s,p,q]
const MAMBO_SYM_FULL    = Cuint(3)   # 8-fold symmetry of real-orbital integrals (what get_tbt / BLISS produce)

"""
    Engine(F, flavour; device = 0, flags = MAMBO_SYM_AUTO)

Upload F.mbts[3] (and F.mbts[2] for CSA_SD) once; the residual then lives on the device until `destroy!(e)` (or, failing
that, the finalizer).  The default `flags` lets the library detect the symmetry of the tensor; pass `MAMBO_SYM_GENERAL` to
force the symmetry-agnostic mode for inputs that are only approximately symmetric (asymmetry above 1e-13 relative).
GPU memory is invisible to Julia's GC (it sees a 16-byte object), so long-lived loops should call `destroy!` themselves.
"""
function Engine(F::F_OP, flavour::Cint; device::Integer = 0, flags::Integer = MAMBO_SYM_AUTO)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    tbt = F.mbts[3]::Array{Float64,4}                       # column-major, passed as it lies in memory
    obt = flavour == MAMBO_CSA_SD ? Matrix{Float64}(F.mbts[2]) : Float64[]
    GC.@preserve tbt obt begin
        check(ccall((:mambo_csa_create, LIB), Cint,
                    (Ref{Ptr{Cvoid}}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Cint, Cuint),
                    href, F.N, flavour, tbt, flavour == MAMBO_CSA_SD ? pointer(obt) : C_NULL, device, flags))
    end
    e = Engine(href[], Int(ccall((:mambo_csa_num_params, LIB), Cint, (Ptr{Cvoid},), href[])))
    finalizer(x -> (x.h != C_NULL && ccall((:mambo_csa_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), e)
    return e
end

"Release the device memory of `e` now (idempotent; the finalizer becomes a no-op)."
function destroy!(e::Engine)
    if e.h != C_NULL
        ccall((:mambo_csa_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.h)
        e.h = C_NULL
    end
    return nothing
end

function cost(e::Engine, x::Vector{Float64})
    c = Ref{Float64}(0.0)
    GC.@preserve x check(ccall((:mambo_csa_cost, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), e.h, x, c))
    return c[]
end

function grad!(e::Engine, storage::Vector{Float64}, x::Vector{Float64})
    GC.@preserve storage x check(ccall((:mambo_csa_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), e.h, x, storage))
    return storage
end

"fg! form for Optim.only_fg!: one fused evaluation per point."
function fg!(e::Engine, F, G, x::Vector{Float64})
    c = Ref{Float64}(0.0)
    g = G === nothing ? Ptr{Float64}(C_NULL) : pointer(G)
    GC.@preserve x G check(ccall((:mambo_csa_cost_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}), e.h, x, c, g))
    return F === nothing ? nothing : c[]
end

function subtract_fragment!(e::Engine, x::Vector{Float64})
    c = Ref{Float64}(0.0)
    GC.@preserve x check(ccall((:mambo_csa_subtract_fragment, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), e.h, x, c))
    return c[]
end

function residual_cost(e::Engine)
    c = Ref{Float64}(0.0)
    check(ccall((:mambo_csa_residual_cost, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), e.h, c))
    return c[]
end

"Another evaluation context on the same GPU sharing the device-resident residual of `e` (mambo_csa_create_lane)."
function lane(e::Engine)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mambo_csa_create_lane, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), e.h, href))
    l = Engine(href[], e.L)
    finalizer(x -> (x.h != C_NULL && ccall((:mambo_csa_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), l)
    return l
end

"""
    tbt_svd_1st(e::Engine, n)

guesses.jl:51-84 on the device-resident residual: the O(N^6) `eigen(Symmetric(reshape(tbt,(N^2,N^2))))` of :55-63 is
replaced by `mambo_csa_leading_eig` (device Lanczos, only the eigenpair of largest |eigenvalue| is ever used); the
N x N `eigen` and `SVD_to_CSA` (:3-49) stay as they are.
"""
function QuantumMAMBO.tbt_svd_1st(e::Engine, n::Integer; tiny = SVD_tiny)
    Λ1 = Ref{Float64}(0.0)
    full_l = Matrix{Float64}(undef, n, n)
    iters = Ref{Cint}(0)
    resid = Ref{Float64}(0.0)
    GC.@preserve full_l check(ccall((:mambo_csa_leading_eig, LIB), Cint,
        (Ptr{Cvoid}, Cdouble, Cint, Ref{Float64}, Ptr{Float64}, Ref{Cint}, Ref{Float64}),
        e.h, 0.0, 0, Λ1, full_l, iters, resid))
    cur_l = Symmetric(full_l)
    lead = Λ1[]
    if sum(abs.(cur_l - full_l)) > tiny                                   # guesses.jl:67-75
        if sum(abs.(full_l + full_l')) > tiny
            error("SVD operator is neither Hermitian or anti-Hermitian, cannot do double factorization into Hermitian fragment!")
        end
        cur_l = Hermitian(1im * full_l)
        lead *= -1
    end
    ωl, Ul = eigen(cur_l)
    return SVD_to_CSA(lead, ωl, Ul)
end

"cost and gradient at the K columns of X in one call, spread over `engines` (an Engine and lanes of it)."
function cost_grad_batch(engines::Vector{Engine}, X::Matrix{Float64})
    L, K = size(X)
    costs = Vector{Float64}(undef, K)
    grads = Matrix{Float64}(undef, L, K)
    hs = [e.h for e in engines]
    GC.@preserve X costs grads hs check(ccall((:mambo_csa_cost_grad_batch, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), hs, length(hs), K, X, costs, grads))
    return costs, grads
end

# ---- optional: the library's own optimiser (dense BFGS / L-BFGS with all vectors on the device) ---------------------------
struct OptOptions            # mambo_opt_options_t
    max_iter::Cint
    g_tol::Cdouble
    lbfgs_memory::Cint
    verbose::Cint
end
struct OptResult             # mambo_opt_result_t
    minimum::Cdouble
    iterations::Cint
    evaluations::Cint
    converged::Cint
    device::Cint
end

"Minimise cost(x) from x0 on the GPU (stands in for `optimize(cost, grad!, x0, BFGS())`, decompose.jl:107-119)."
function optimize_native(e::Engine, x0::Vector{Float64}; max_iter = 1000, g_tol = 1e-8, lbfgs_memory = 0)
    xmin = similar(x0)
    opt = Ref(OptOptions(max_iter, g_tol, lbfgs_memory, 0))
    res = Ref(OptResult(0.0, 0, 0, 0, 0))
    GC.@preserve x0 xmin check(ccall((:mambo_csa_optimize, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{OptOptions}, Ref{OptResult}), e.h, x0, xmin, opt, res))
    return xmin, res[]
end

"K independent starts (columns of X0) scheduled over `engines` (lanes of one GPU and/or engines on several GPUs)."
function multistart(engines::Vector{Engine}, X0::Matrix{Float64}; max_iter = 1000, g_tol = 1e-8, lbfgs_memory = 0)
    L, K = size(X0)
    Xmin = similar(X0)
    hs = [e.h for e in engines]
    opt = Ref(OptOptions(max_iter, g_tol, lbfgs_memory, 0))
    res = Vector{OptResult}(undef, K)
    best = Ref{Cint}(-1)
    GC.@preserve X0 Xmin hs res check(ccall((:mambo_multistart_run, LIB), Cint,
        (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ref{OptOptions}, Ptr{OptResult}, Ref{Cint}),
        hs, length(hs), K, X0, Xmin, opt, res, best))
    return Xmin, res, Int(best[]) + 1
end

# ---- drop-in greedy steps (same signatures as decompose.jl:82, :207) ---------------------------------------------
function _optimize(e::Engine, x0, print, do_grad)
    c(x) = cost(e, x)                                       # decompose.jl:96-99   / :221-224
    g!(storage::Vector{Float64}, x::Vector{Float64}) = grad!(e, storage, x)   # decompose.jl:101-105 / :226-231
    if do_grad                                              # decompose.jl:107-119 / :233-245, branch for branch
        if print == false
            return optimize(c, g!, x0, BFGS())
        else
            return optimize(c, g!, x0, BFGS(), Optim.Options(show_every = print, show_trace = true, extended_trace = true))
        end
    else                                                    # finite-difference gradients of the (GPU) cost, as in the reference
        if print == false
            return optimize(c, x0, BFGS())
        else
            return optimize(c, x0, BFGS(), Optim.Options(show_every = print, show_trace = true, extended_trace = true))
        end
    end
end

# A step that had to create its own Engine (the path the UNMODIFIED CSA_greedy_decomposition takes: one call per fragment)
# releases it before returning: GPU memory must not wait for Julia's GC, which does not see it.
function _with_engine(f, F::F_OP, flavour::Cint, engine)
    own = engine === nothing
    e = own ? Engine(F, flavour) : engine
    try
        return f(e)
    finally
        own && destroy!(e)
    end
end

function QuantumMAMBO.CSA_greedy_step(F::F_OP, do_svd = SVD_for_CSA, print = DECOMPOSITION_PRINT, do_grad = true; engine = nothing)
    cartan_L = cartan_2b_num_params(F.N)
    unitary_L = real_orbital_rotation_num_params(F.N)
    x0 = zeros(cartan_L + unitary_L)
    return _with_engine(F, MAMBO_CSA, engine) do e               # with `engine`, F_rem is the engine's residual
        if do_svd
            frag = tbt_svd_1st(e, F.N)                             # largest SVD fragment of the *residual*
            x0[1:cartan_L] = frag.C.λ
            x0[cartan_L+1:end] = frag.U[1].θs
        else
            x0[cartan_L+1:end] = 2π * rand(unitary_L)
        end
        _optimize(e, x0, print, do_grad)
    end
end

function QuantumMAMBO.CSA_SD_greedy_step(F::F_OP, do_svd = SVD_for_CSA_SD, print = DECOMPOSITION_PRINT, do_grad = true; engine = nothing)
    cartan_L = cartan_2b_num_params(F.N)
    unitary_L = real_orbital_rotation_num_params(F.N)
    x0 = zeros(cartan_L + unitary_L + F.N)
    return _with_engine(F, MAMBO_CSA_SD, engine) do e
        if do_svd
            frag = tbt_svd_1st(e, F.N)
            x0[F.N+1:F.N+cartan_L] = frag.C.λ
            x0[F.N+cartan_L+1:end] = frag.U[1].θs
        else
            x0[F.N+cartan_L+1:end] = 2π * rand(unitary_L)
        end
        _optimize(e, x0, print, do_grad)
    end
end

"""
    CSA_greedy_decomposition_b200(H, α_max; decomp_tol = ϵ, verbose = false)

decompose.jl:1-80 with F_rem resident on the GPU: one upload of H.mbts[3], `subtract_fragment!` instead of
`F_rem = F_rem - to_OP(frag)` (decompose.jl:56).  The HDF5 checkpoint code of the original can be kept verbatim
around this loop (X[:, α] = sol.minimizer); on resume replay the saved columns with `subtract_fragment!`.
"""
function CSA_greedy_decomposition_b200(H::F_OP, α_max; decomp_tol = ϵ, verbose = false, flavour = MAMBO_CSA)
    e = Engine(H, flavour)
    cartan_L = cartan_2b_num_params(H.N)
    Farr = F_FRAG[]
    try
        curr_cost = residual_cost(e)                                           # decompose.jl:40
        α = 0
        while curr_cost > decomp_tol && α < α_max
            α += 1
            sol = flavour == MAMBO_CSA ? QuantumMAMBO.CSA_greedy_step(H; engine = e) : QuantumMAMBO.CSA_SD_greedy_step(H; engine = e)
            frag = flavour == MAMBO_CSA ? CSA_x_to_F_FRAG(sol.minimizer, H.N, H.spin_orb, cartan_L) :
                                          CSA_SD_x_to_F_FRAG(sol.minimizer, H.N, H.spin_orb, cartan_L)
            push!(Farr, frag)
            subtract_fragment!(e, sol.minimizer)                                # decompose.jl:56
            curr_cost = sol.minimum
            verbose && println("Current L2 cost after $α fragments is $(sol.minimum)")
        end
    finally
        destroy!(e)
    end
    return Farr
end

# ---- double factorisation on the resident residual (decompose.jl:297-436) -------------------------------------------------
struct DfStats               # mambo_df_stats_t
    krylov_dim::Cint; restarts::Cint; num_frags::Cint; anti_frags::Cint; jacobi_sweeps_max::Cint; cluster_vectors::Cint
    ortho_dev_before::Cdouble; ortho_dev_after::Cdouble; max_rel_residual::Cdouble
    seconds_lanczos::Cdouble; seconds_tridiag::Cdouble; seconds_backtransform::Cdouble; seconds_fragments::Cdouble
end

"""
    DF_decomposition(e::Engine, n; tol = SVD_tol) -> Vector{F_FRAG}

decompose.jl:297-398 (do_Givens = false) of the tensor resident in `e`: the N^2 x N^2 `eigen` (:326-330), the ordering and
truncation (:332-350) and the per-fragment `eigen(cur_l)` (:354) all run on the GPU; this wrapper only builds the F_FRAGs
(:372-378).  Anti-Hermitian fragments (:356-361) come back with their coefficient negated and the matrix L; their complex
`eigen(Hermitian(1im * L))` is done here.
"""
function QuantumMAMBO.DF_decomposition(e::Engine, n::Integer; tol = SVD_tol, spin_orb = false)
    nf = Ref{Cint}(0)
    st = Ref{DfStats}()
    check(ccall((:mambo_csa_df_decompose, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cint, Cint, Ref{Cint}, Ref{DfStats}), e.h, tol, 0, 0, nf, st))
    k = Int(nf[])
    coeff = Vector{Float64}(undef, k); omega = Matrix{Float64}(undef, n, k); kind = Vector{Cint}(undef, k)
    U = Array{Float64,3}(undef, n, n, k); L = Array{Float64,3}(undef, n, n, k)
    k > 0 && check(ccall((:mambo_csa_df_get, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
                         e.h, 0, k, coeff, omega, U, L, kind))
    FRAGS = F_FRAG[]
    for i in 1:k
        if kind[i] == 0
            Rl = f_matrix_rotation(n, U[:, :, i]); ωl = omega[:, i]
        else
            ωl, Ul = eigen(Hermitian(1im * L[:, :, i]))                           # decompose.jl:359
            Rl = c_matrix_rotation(n, Ul)
        end
        push!(FRAGS, F_FRAG(1, tuple(Rl), DF(), cartan_1b(spin_orb, ωl, n), n, spin_orb, coeff[i], true))
    end
    return FRAGS
end

"""
    DF_based_greedy(e::Engine, n) -> F_FRAG      (decompose.jl:400-436, after DF_decomposition(e, n))
"""
function QuantumMAMBO.DF_based_greedy(e::Engine, n::Integer; spin_orb = false)
    lam = Vector{Float64}(undef, cartan_2b_num_params(n)); U1 = Matrix{Float64}(undef, n, n)
    check(ccall((:mambo_csa_df_based_greedy, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), e.h, lam, U1, C_NULL))
    return F_FRAG(1, tuple(f_matrix_rotation(n, U1)), CSA(), cartan_2b(spin_orb, lam, n), n, spin_orb)
end

end # module
