"""GPU parity tests (run by the driver with `-m gpu` on a B200): the CUDA path, called through the C-ABI, against
the oracle on the same seeded inputs; against the committed golden vectors; and, at BASELINE.json's full sizes,
through size-independent properties.  Tolerance (north_star): 1e-10 relative for cost and (norm-wise) gradient in
FP64; final 1-norms 1e-8."""
import ctypes as C
import os

import numpy as np
import pytest

import csa_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10
KINDS = {"dense_sym": o.dense_sym_tbt, "pairsym": o.pairsym_tbt, "general": o.general_tbt}


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def _obt(n, seed):
    return np.random.default_rng(100 + seed).standard_normal((n, n)) / n


@pytest.fixture(scope="module")
def lib(mb):
    if not os.path.exists(os.path.join(os.path.dirname(mb.__file__), "quantummambo.jl_b200", "csrc", "libmambo_b200.so")):
        mb.build()
    return mb.load()


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 6, 7, 10, 17, 24, 33])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
def test_cost_and_gradient_match_oracle(mb, lib, n, flavour, kind):
    T = KINDS[kind](n, n)
    obt = _obt(n, n) if flavour == "CSA_SD" else None
    x = o.random_x(n, 3, flavour)
    c0, g0 = o.cost_grad_factored(x, T, obt, flavour)
    with mb.CSAEngine(T, obt, flavour) as e:
        c, g = e.cost_grad(x)
        assert e.cost(x) == c and np.array_equal(e.grad(x), g)          # memoised closures (decompose.jl:96-105)
        mode = e.info()["mode_name"]
    assert mode == {"dense_sym": "fullsym-packed", "pairsym": "pairsym", "general": "general"}[kind] or n == 1
    assert abs(c - c0) <= RTOL * abs(c0)
    assert rel(g, g0) <= RTOL


@pytest.mark.parametrize("sym", ["general", "pair", "full"])
def test_forced_modes_agree_on_a_symmetric_tensor(mb, lib, sym):
    n = 12
    T = o.dense_sym_tbt(n, 0)
    x = o.random_x(n, 1)
    c0, g0 = o.cost_grad_factored(x, T)
    flag = {"general": mb.SYM_GENERAL, "pair": mb.SYM_PAIR, "full": mb.SYM_FULL}[sym]
    with mb.CSAEngine(T, sym=flag) as e:
        c, g = e.cost_grad(x)
    assert abs(c - c0) <= RTOL * abs(c0) and rel(g, g0) <= RTOL


def test_golden_vectors(mb, lib, golden):
    for name, c in golden.items():
        n, sd = int(c["meta"][0]), int(c["meta"][1])
        fl = "CSA_SD" if sd else "CSA"
        with mb.CSAEngine(c["tbt"], c.get("obt"), fl) as e:
            cost, grad = e.cost_grad(c["x"])
            W, h = e.reconstruct(c["x"])
        assert abs(cost - float(c["cost"])) <= RTOL * abs(float(c["cost"])), name
        assert rel(grad, c["grad"]) <= RTOL, name
        # extended-precision ground truth (long-double factored form with its own expm, oracle/csa_oracle_ld.py)
        assert abs(cost - float(c["cost_ld"])) <= 1e-12 * abs(float(c["cost_ld"])), name
        assert rel(grad, c["grad_ld"]) <= 1e-12, name
        assert rel(W, c["W"]) <= RTOL, name
        if sd:
            assert rel(h, c["h"]) <= RTOL, name


def test_theta_zero_and_zero_lambda_edge_cases(mb, lib):
    n = 6
    T = o.dense_sym_tbt(n, 2)
    cL, uL = o.cartan_2b_num_params(n), o.real_orbital_rotation_num_params(n)
    lam = np.random.default_rng(0).standard_normal(cL)
    with mb.CSAEngine(T) as e:
        for x in (np.concatenate([lam, np.zeros(uL)]),                     # U = I
                  np.zeros(cL + uL),                                       # zero fragment: cost = |T|^2
                  np.concatenate([np.zeros(cL), 2 * np.pi * np.random.default_rng(1).random(uL)])):   # reference x0
            c0, g0 = o.cost_grad_factored(x, T)
            c, g = e.cost_grad(x)
            assert abs(c - c0) <= RTOL * abs(c0) and np.max(np.abs(g - g0)) <= RTOL * max(np.max(np.abs(g0)), 1.0)
        assert e.residual_cost() == pytest.approx(float(np.sum(T * T)), rel=1e-13)


def test_zero_tensor_and_planted_tensor(mb, lib):
    n = 8
    with mb.CSAEngine(np.zeros((n, n, n, n))) as e:
        x = o.random_x(n, 0)
        c0, g0 = o.cost_grad_factored(x, np.zeros((n, n, n, n)))
        c, g = e.cost_grad(x)
        assert abs(c - c0) <= RTOL * abs(c0) and rel(g, g0) <= RTOL
    T, xs = o.planted_tbt(n, 1, 0)
    with mb.CSAEngine(T) as e:
        c, g = e.cost_grad(xs[0])
        assert c <= 1e-20 * float(np.sum(T * T))
        assert np.max(np.abs(g)) <= 1e-9 * np.sqrt(float(np.sum(T * T)))


def test_error_paths(mb, lib):
    n = 4
    T = o.dense_sym_tbt(n, 0)
    with mb.CSAEngine(T) as e:
        x = o.random_x(n, 0)
        xb = x.copy()
        xb[-1] = np.nan
        with pytest.raises(mb.MamboError) as ei:
            e.cost_grad(xb)
        assert ei.value.code == -5                                   # MAMBO_ERANGE
        xb[-1] = 1e9                                                 # ||K||_1 > 2^14
        with pytest.raises(mb.MamboError):
            e.cost_grad(xb)
        c, g = e.cost_grad(x)                                        # the handle stays usable
        c0, g0 = o.cost_grad_factored(x, T)
        assert abs(c - c0) <= RTOL * abs(c0)
        with pytest.raises(ValueError):
            e.cost_grad(x[:-1])
    with pytest.raises(ValueError):
        mb.CSAEngine(T, flavour="CSA_SD")                            # CSA_SD without obt
    with pytest.raises(mb.MamboError):
        mb.CSAEngine(o.general_tbt(6, 0), shard=0, nshards=2)        # sharding needs a pair-symmetric tensor


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
def test_greedy_state_round_trip(mb, lib, flavour, kind):
    """subtract_fragment == F_rem - to_OP(frag) (decompose.jl:56,182); reconstruct == to_OP; get_residual hands F_rem back."""
    n = 9
    T = KINDS[kind](n, 4)
    obt = _obt(n, 4) if flavour == "CSA_SD" else None
    x = o.random_x(n, 5, flavour)
    h0, W0 = o.reconstruct_factored(x, n, flavour)
    c0 = o.cost_grad_factored(x, T, obt, flavour, need_grad=False)[0]
    with mb.CSAEngine(T, obt, flavour) as e:
        W, h = e.reconstruct(x)
        assert rel(W, W0) <= RTOL
        new_cost = e.subtract_fragment(x)
        assert abs(new_cost - c0) <= RTOL * abs(c0)                    # |T - W|^2 is the cost at x
        assert e.residual_cost() == pytest.approx(c0, rel=1e-12)
        Tr, orr = e.get_residual()
        assert rel(Tr, T - W0) <= RTOL
        if obt is not None:
            assert rel(h, h0) <= RTOL and rel(orr, obt - h0) <= RTOL
        # a second fragment is fitted against the updated residual
        x2 = o.random_x(n, 6, flavour)
        c2, g2 = e.cost_grad(x2)
        c20, g20 = o.cost_grad_factored(x2, T - W0, None if obt is None else obt - h0, flavour)
        assert abs(c2 - c20) <= RTOL * abs(c20) and rel(g2, g20) <= RTOL


def test_bitwise_reproducible(mb, lib):
    n = 20
    T = o.dense_sym_tbt(n, 1)
    x = o.random_x(n, 2)
    with mb.CSAEngine(T) as e:
        c1, g1 = e.cost_grad(x)
        e.cost_grad(o.random_x(n, 3))
        c2, g2 = e.cost_grad(x)
    assert c1 == c2 and np.array_equal(g1, g2)


def test_sharded_partials_sum_to_unsharded(mb, lib):
    """Row-panel sharding emulated on one GPU: the shards' partial vectors add up to the unsharded
    one, and finishing from the sum gives the unsharded cost+gradient."""
    import torch
    n = 14                                       # 105 packed rows -> 2 panels; 196 full rows -> 4 panels
    for kind, sym, world in (("dense_sym", None, 2), ("pairsym", None, 4), ("dense_sym", "pair", 3)):
        T = KINDS[kind](n, 0)
        x = o.random_x(n, 1)
        c0, g0 = o.cost_grad_factored(x, T)
        flag = 0 if sym is None else 2
        dev = torch.device("cuda", 0)
        d_x = torch.from_numpy(x).to(dev)
        total = None
        engines = [mb.CSAEngine(T, sym=flag, shard=r, nshards=world) for r in range(world)]
        try:
            plen = engines[0].info()["partial_len"]
            for e in engines:
                d_p = torch.zeros(plen, dtype=torch.float64, device=dev)
                e.eval_partial_device(d_x.data_ptr(), d_p.data_ptr())
                e.synchronize()
                total = d_p if total is None else total + d_p
            for e in engines:
                d_o = torch.zeros(len(x) + 1, dtype=torch.float64, device=dev)
                e.eval_finish_device(total.data_ptr(), d_o.data_ptr())
                e.synchronize()
                out = d_o.cpu().numpy()
                assert abs(out[0] - c0) <= RTOL * abs(c0) and rel(out[1:], g0) <= RTOL
        finally:
            for e in engines:
                e.close()


def test_peer_exchange_of_shards_on_one_gpu(mb, lib):
    """The NVLink peer-memory exchange (push -> flag -> rank-ordered sum) with all shards living on one GPU of one
    process: every shard pushes first, then every shard finishes (no kernel ever waits for a later launch)."""
    import torch
    n = 14
    dev = torch.device("cuda", 0)
    for kind, world in (("dense_sym", 2), ("pairsym", 3)):
        T = KINDS[kind](n, 2)
        engines = [mb.CSAEngine(T, shard=r, nshards=world) for r in range(world)]
        try:
            for e in engines:
                e.peer_export()
            for r, e in enumerate(engines):
                for q, peer in enumerate(engines):
                    if q != r:
                        e.peer_attach_local(q, peer)
            for rep in range(3):                                   # three epochs: exercises the parity double buffer
                x = o.random_x(n, 10 + rep)
                c0, g0 = o.cost_grad_factored(x, T)
                d_x = torch.from_numpy(x).to(dev)
                outs = [torch.zeros(len(x) + 1, dtype=torch.float64, device=dev) for _ in engines]
                for e in engines:
                    e.eval_sharded_push_device(d_x.data_ptr())
                for e in engines:
                    e.synchronize()
                for e, d_o in zip(engines, outs):
                    e.eval_sharded_finish_device(d_o.data_ptr())
                    e.synchronize()
                ref = outs[0].cpu().numpy()
                for d_o in outs:
                    out = d_o.cpu().numpy()
                    assert abs(out[0] - c0) <= RTOL * abs(c0) and rel(out[1:], g0) <= RTOL
                    assert np.array_equal(out, ref)                # identical bits on every shard
        finally:
            for e in engines:
                e.close()


def test_lanes_on_shards_with_stream_ordered_peer_wait(mb, lib):
    """Several sharded evaluations in flight per GPU: lanes created from shard handles (each with its own peer buffers), the
    peers' flags awaited by the stream (mambo_csa_set_peer_wait) instead of a spinning kernel.  All shards and lanes live on
    one GPU of one process here and every call is only enqueued -- with stream-ordered waits nothing blocks an SM, so the
    order of the launches does not matter (the spinning variant would need every push before any finish)."""
    import torch
    n, world, nl = 14, 2, 3
    dev = torch.device("cuda", 0)
    T = o.dense_sym_tbt(n, 2)
    shards = [mb.CSAEngine(T, shard=r, nshards=world) for r in range(world)]
    lanes = [[e] + [e.lane() for _ in range(nl - 1)] for e in shards]          # lanes[r][k]
    streams = [[torch.cuda.Stream(dev) for _ in range(nl)] for _ in range(world)]
    try:
        for k in range(nl):
            for r in range(world):
                lanes[r][k].peer_export()
            for r in range(world):
                for q in range(world):
                    if q != r:
                        lanes[r][k].peer_attach_local(q, lanes[q][k])
        for r in range(world):
            for k in range(nl):
                lanes[r][k].set_peer_wait(True)
                lanes[r][k].set_stream(streams[r][k].cuda_stream)
        xs = [o.random_x(n, 30 + i) for i in range(2 * nl)]                   # two points per lane: two epochs each
        d_x = torch.from_numpy(np.stack(xs)).to(dev)
        d_out = torch.zeros(world, len(xs), len(xs[0]) + 1, dtype=torch.float64, device=dev)
        for i in range(len(xs)):
            for r in range(world):                                            # shard 0 enqueues push AND finish before shard 1 pushes
                lanes[r][i % nl].eval_sharded_device(d_x[i].data_ptr(), d_out[r, i].data_ptr())
        torch.cuda.synchronize()
        for r in range(world):
            for k in range(nl):
                lanes[r][k].synchronize()
        res = d_out.cpu().numpy()
        for i, x in enumerate(xs):
            c0, g0 = o.cost_grad_factored(x, T)
            for r in range(world):
                assert abs(res[r, i, 0] - c0) <= RTOL * abs(c0) and rel(res[r, i, 1:], g0) <= RTOL
            assert np.array_equal(res[0, i], res[1, i])                       # identical bits on every shard
    finally:
        for r in range(world):
            for ln in lanes[r]:
                ln.set_stream(None)
            for ln in lanes[r][1:]:
                ln.close()
            shards[r].close()


def test_single_gemm_formulation_matches_two_gemm_kernel_and_falls_back(mb, lib, monkeypatch):
    """The default evaluation (fused::k_ty: E = C~ - T~ A~, cost from N x N quantities) against the two-GEMM kernel
    (MAMBO_TWO_GEMM=1: D = B~ A~^T - T~ explicitly) on the same handle inputs, including the regime where the expanded
    cost cancels (x at / near a planted fragment) and the exact-cost pass must take over."""
    import torch
    n = 13
    T, xs = o.planted_tbt(n, 1, 7)
    x_true = xs[0]
    rng = np.random.default_rng(2)
    probes = [o.random_x(n, 1), x_true, x_true + 1e-7 * rng.standard_normal(x_true.shape),
              x_true + 1e-3 * rng.standard_normal(x_true.shape)]
    fast = mb.CSAEngine(T)                                  # lean path: N x N closed forms instead of the B~/C~ panel GEMMs
    monkeypatch.setenv("MAMBO_NO_LEAN", "1")
    mid = mb.CSAEngine(T)                                   # first single-GEMM formulation (Gram kernel + B~, C~)
    monkeypatch.delenv("MAMBO_NO_LEAN")
    monkeypatch.setenv("MAMBO_TWO_GEMM", "1")
    slow = mb.CSAEngine(T)
    monkeypatch.delenv("MAMBO_TWO_GEMM")
    tt = float(np.sum(T * T))
    try:
        for x in probes:
            cs, gs = slow.cost_grad(x)
            c0, g0 = o.cost_grad_factored(x, T)
            # cost: 1e-10 relative, down to what FP64 resolves: every element of D = W - T carries eps |T|, so
            # |D|^2 carries ~ eps |D| |T| (and eps^2 |T|^2 at the exact solution)
            ctol = 1e-10 * abs(c0) + 1e-14 * np.sqrt(abs(c0) * tt) + 1e-26 * tt
            # gradient: 1e-10 of its size, down to the FP64 floor of D itself (each element of D = W - T carries eps |T|)
            gtol = 1e-10 * float(np.max(np.abs(g0))) + 1e-13 * np.sqrt(tt)
            for eng in (fast, mid):
                cf, gf = eng.cost_grad(x)
                assert cf >= 0.0
                assert abs(cf - cs) <= ctol
                assert abs(cf - c0) <= ctol
                assert float(np.max(np.abs(gf - gs))) <= gtol
                assert float(np.max(np.abs(gf - g0))) <= gtol
                # the device-resident entry point takes the same tail (exact-cost pass beside the derivative chain)
                d_x = torch.tensor(x, dtype=torch.float64, device="cuda")
                d_o = torch.zeros(1 + len(x), dtype=torch.float64, device="cuda")
                eng.eval_device(d_x.data_ptr(), d_o.data_ptr())
                eng.synchronize()
                out = d_o.cpu().numpy()
                assert out[0] == cf and np.array_equal(out[1:], gf)
    finally:
        fast.close()
        mid.close()
        slow.close()


def test_batch_closure_over_lanes_matches_single_closures(mb, lib):
    """mambo_csa_cost_grad_batch: K points in one call, spread over lanes in waves -- the same bits as K single closures."""
    for n, flavour in ((11, "CSA"), (8, "CSA_SD")):
        T = o.dense_sym_tbt(n, 6)
        obt = _obt(n, 6) if flavour == "CSA_SD" else None
        with mb.CSAEngine(T, obt, flavour) as e:
            lanes = [e, e.lane(), e.lane()]
            xs = np.stack([o.random_x(n, 30 + k, flavour) for k in range(7)])       # 7 points over 3 lanes: ragged last wave
            costs, grads = mb.cost_grad_batch(lanes, xs)
            for k in range(len(xs)):
                c, g = e.cost_grad(xs[k])
                assert costs[k] == c and np.array_equal(grads[k], g)
                c0, g0 = o.cost_grad_factored(xs[k], T, obt, flavour)
                assert abs(c - c0) <= RTOL * abs(c0) and rel(g, g0) <= RTOL
            c1, g1 = mb.cost_grad_batch(lanes[:1], xs[:2])                           # a single lane is allowed
            assert c1[1] == costs[1] and np.array_equal(g1[0], grads[0])
            for ln in lanes[1:]:
                ln.close()


def test_handles_of_different_sizes_coexist(mb, lib):
    """Opt-in shared-memory limits are per kernel function, not per handle: creating a smaller problem afterwards must not
    take them away from a larger handle that is still in use."""
    big_n, small_n = 40, 5
    Tb, Ts = o.dense_sym_tbt(big_n, 1), o.dense_sym_tbt(small_n, 2)
    xb, xs = o.random_x(big_n, 3), o.random_x(small_n, 4)
    with mb.CSAEngine(Tb) as big:
        c1, g1 = big.cost_grad(xb)
        with mb.CSAEngine(Ts) as small:
            cs, gs = small.cost_grad(xs)
            c2, g2 = big.cost_grad(xb + 0.0)
            xm, res = big.optimize(xb, max_iter=3)
            lam, vec, _ = big.leading_eig()
        c0, g0 = o.cost_grad_factored(xs, Ts)
        assert abs(cs - c0) <= RTOL * abs(c0) and rel(gs, g0) <= RTOL
        assert c1 == c2 and np.array_equal(g1, g2) and res["minimum"] <= c1


def test_lanes_share_the_residual_and_run_concurrently(mb, lib):
    """mambo_csa_create_lane: lanes evaluate independently (one host thread each, overlapping on the device), give the
    parent's bits, see residual updates made through a sibling, and keep the residual alive after the parent is closed."""
    import threading
    for n, flavour in ((13, "CSA"), (9, "CSA_SD"), (40, "CSA")):
        T = o.dense_sym_tbt(n, 4)
        obt = _obt(n, 4) if flavour == "CSA_SD" else None
        parent = mb.CSAEngine(T, obt, flavour)
        lanes = [parent.lane() for _ in range(3)]
        group = [parent] + lanes
        xs = [o.random_x(n, 20 + k, flavour) for k in range(len(group))]
        want = [o.cost_grad_factored(x, T, obt, flavour) for x in xs]
        c_par, g_par = parent.cost_grad(xs[1])
        c_lane, g_lane = lanes[0].cost_grad(xs[1])
        assert c_par == c_lane and np.array_equal(g_par, g_lane)              # same bits through either context
        got = [None] * len(group)

        def work(k):
            for _ in range(6):                                                # repeated so the lanes really overlap
                got[k] = group[k].cost_grad(xs[k] + 0.0)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(len(group))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for (c, g), (c0, g0) in zip(got, want):
            assert abs(c - c0) <= RTOL * abs(c0) and rel(g, g0) <= RTOL
        # a residual update through one handle is seen by every lane, memoised x included
        before = lanes[1].cost_grad(xs[0])[0]
        new_cost = parent.subtract_fragment(xs[0])
        W, h1 = parent.reconstruct(xs[0])
        T2 = T - W
        obt2 = None if obt is None else obt - h1
        c_after, g_after = lanes[1].cost_grad(xs[0])
        c0, g0 = o.cost_grad_factored(xs[0], T2, obt2, flavour)
        assert c_after != before
        assert abs(c_after - c0) <= 1e-9 * max(abs(c0), abs(before)) and rel(g_after, g0) <= 1e-9
        assert abs(lanes[2].residual_cost() - new_cost) <= RTOL * new_cost
        parent.close()                                                         # the lanes keep the shared residual alive
        c_again, _ = lanes[0].cost_grad(xs[0])
        assert c_again == c_after
        for e in lanes:
            e.close()


def test_multistart_scheduler_over_lanes(mb, lib):
    """mambo_multistart_run over lanes of one GPU: same minima as running every start alone, and the reported best
    start is the arg-min (multistart precedent: orbitals.jl:39-67)."""
    n = 6
    x_true = o.random_x(n, 5)
    T = o.reconstruct_factored(x_true, n)[1]
    cL, uL = o.cartan_2b_num_params(n), o.real_orbital_rotation_num_params(n)
    rng = np.random.default_rng(7)
    K = 5
    x0s = np.stack([np.concatenate([np.zeros(cL), 2 * np.pi * rng.random(uL)]) for _ in range(K)])
    with mb.CSAEngine(T) as e:
        lanes = [e, e.lane(), e.lane()]
        xm, res, best = mb.multistart(lanes, x0s, max_iter=200)
        alone = [e.optimize(x0, max_iter=200) for x0 in x0s]
        for k in range(K):
            assert abs(res[k]["minimum"] - alone[k][1]["minimum"]) <= 1e-9 * max(1.0, abs(alone[k][1]["minimum"]))
            c, _ = e.cost_grad(xm[k])
            assert abs(c - res[k]["minimum"]) <= 1e-9 * max(1.0, abs(c))
        assert best == int(np.argmin([r["minimum"] for r in res]))
        for ln in lanes[1:]:
            ln.close()


def _sym_matrix_view(T):
    """Symmetric(reshape(tbt, (N^2, N^2))) as Julia builds it (guesses.jl:57): upper triangle of the column-major view."""
    n = T.shape[0]
    M = np.reshape(np.asfortranarray(T), (n * n, n * n), order="F")
    return np.triu(M) + np.triu(M, 1).T


@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
@pytest.mark.parametrize("n", [1, 2, 3, 6, 10, 17])
def test_leading_eigenpair_matches_dense_eigen(mb, lib, n, kind):
    """mambo_csa_leading_eig (device Lanczos) against the full `eigen` the reference runs (guesses.jl:57-63)."""
    T = KINDS[kind](n, 30 + n)
    M = _sym_matrix_view(T)
    w, v = np.linalg.eigh(M)
    k = int(np.argmax(np.abs(w)))
    with mb.CSAEngine(T) as e:
        lam, vec, info = e.leading_eig()
    assert abs(lam - w[k]) <= 1e-11 * abs(w[k])
    got = vec.reshape(-1, order="F")
    assert abs(np.linalg.norm(got) - 1.0) <= 1e-12
    ref = v[:, k]
    gap = np.min(np.abs(np.delete(w, k) - w[k])) if len(w) > 1 else 1.0
    # rank-one projector: independent of the sign convention; accuracy limited by residual / gap
    assert rel(np.outer(got, got), np.outer(ref, ref)) <= 1e-9 * max(1.0, abs(w[k]) / gap)
    assert np.linalg.norm(M @ got - lam * got) <= 1e-10 * abs(lam)


@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
def test_symv_matches_the_matrix_view(mb, lib, kind):
    """mambo_csa_symv_device: y = Symmetric(reshape(tbt)) v for an operator v (N x N), through the engine's row space."""
    import torch
    n = 9
    T = KINDS[kind](n, 3)
    M = _sym_matrix_view(T)
    rng = np.random.default_rng(0)
    v = rng.standard_normal((n, n))
    if kind == "dense_sym":
        v = v + v.T                                                   # the packed mode carries symmetric operators
    want = (M @ v.reshape(-1, order="F")).reshape(n, n, order="F")
    with mb.CSAEngine(T) as e:
        info = e.info()
        rows, rp = info["rows"], info["rows_padded"]
        if info["mode_name"] == "fullsym-packed":
            idx = [(a, q) for q in range(n) for a in range(q + 1)]
            sw = np.array([1.0 if a == q else np.sqrt(2.0) for a, q in idx])
        else:
            idx = [(r % n, r // n) for r in range(n * n)]
            sw = np.ones(n * n)
        hv = np.zeros(rp)
        hv[:rows] = sw * np.array([v[a, q] for a, q in idx])
        d_v = torch.from_numpy(hv).cuda()
        d_y = torch.zeros(rp, dtype=torch.float64, device="cuda")
        e.symv_device(d_v.data_ptr(), d_y.data_ptr())
        e.synchronize()
        y = d_y.cpu().numpy()[:rows] / sw
    got = np.array([want[a, q] for a, q in idx])
    assert rel(y, got) <= 1e-12


def test_svd_initial_guess_matches_oracle(mb, lib):
    """tbt_svd_1st on the device-resident residual (guesses.jl:51-84): same fragment as the oracle's dense-eigen port.
    The sign of the leading eigenvector is a LAPACK convention that the fragment depends on, so the oracle
    is asked for the library's documented convention (largest-magnitude entry positive); the n x n `eigen` may still flip
    the columns of the rotation, hence the comparison on the operator W(x0)."""
    for n in (4, 7, 12):
        T = o.dense_sym_tbt(n, 40 + n)
        lam_o, th_o = o.tbt_svd_1st(T, fix_sign=True)
        with mb.CSAEngine(T) as e:
            lam_g, th_g = mb.tbt_svd_1st(e)
        W_o = o.reconstruct_factored(np.concatenate([lam_o, th_o]), n)[1]
        W_g = o.reconstruct_factored(np.concatenate([lam_g, th_g]), n)[1]
        assert rel(W_g, W_o) <= 1e-8
    with mb.CSAEngine(o.pairsym_tbt(5, 1)) as e:        # no p<->q symmetry: the reference refuses (guesses.jl:69-71)
        with pytest.raises(ValueError):
            mb.tbt_svd_1st(e)


def test_leading_eigenpair_in_the_antisymmetric_subspace(mb, lib):
    """A (pq)<->(rs)-symmetric tensor whose dominant eigenvector is ANTI-symmetric under p <-> q (the anti-Hermitian branch of
    tbt_svd_1st, guesses.jl:67-75).  The tensor commutes with the p <-> q swap, so a Krylov space started from a swap-symmetric
    vector would never leave the symmetric subspace: the start vector of the unpacked modes must not be symmetric."""
    n = 6
    rng = np.random.default_rng(12)
    a = rng.standard_normal((n, n))
    a = a - a.T
    sy = rng.standard_normal((n, n))
    sy = sy + sy.T
    T = 3.0 * np.einsum("pq,rs->pqrs", a, a) + 0.1 * np.einsum("pq,rs->pqrs", sy, sy)
    with mb.CSAEngine(T) as e:
        assert e.info()["mode_name"] == "pairsym"
        lam, vec, st = e.leading_eig()
        coeffs_g, _ = mb.tbt_svd_1st(e)
    assert abs(lam - 3.0 * float(np.sum(a * a))) <= 1e-10 * abs(lam)
    assert np.max(np.abs(vec + vec.T)) <= 1e-10                                   # antisymmetric operator
    assert min(rel(vec, a / np.linalg.norm(a)), rel(-vec, a / np.linalg.norm(a))) <= 1e-9
    coeffs_o, _ = o.tbt_svd_1st(T, fix_sign=True)
    assert rel(coeffs_g, coeffs_o) <= 1e-9
    # pure anti-symmetric rank one: T w0 vanished for the old (symmetric) start vector and the call returned 0
    with mb.CSAEngine(np.einsum("pq,rs->pqrs", a, a)) as e:
        lam2, _, _ = e.leading_eig()
    assert abs(lam2 - float(np.sum(a * a))) <= 1e-10 * abs(lam2)


def test_leading_eigenpair_full_size(mb, lib):
    """N = 50 (R = 1275 packed rows): eigen-residual and eigenvalue against LAPACK on the 2500 x 2500 view."""
    n = 50
    T = o.dense_sym_tbt(n, 0)
    M = _sym_matrix_view(T)
    with mb.CSAEngine(T) as e:
        lam, vec, info = e.leading_eig()
    got = vec.reshape(-1, order="F")
    w = np.linalg.eigvalsh(M)
    top = w[np.argmax(np.abs(w))]
    assert abs(lam - top) <= 1e-11 * abs(top)
    assert np.linalg.norm(M @ got - lam * got) <= 1e-9 * abs(lam)
    assert np.abs(vec - vec.T).max() <= 1e-14


def test_csa_sd_greedy_uses_the_svd_guess_by_default(mb, lib):
    """CSA_SD_greedy_step starts from tbt_svd_1st unless told otherwise (SVD_for_CSA_SD = true, config.jl:10): on a
    tensor that IS one CSA_SD fragment the guess is already the two-body solution and the step converges to ~0."""
    n = 6
    x_true = o.random_x(n, 9, "CSA_SD")
    h_t, W_t = o.reconstruct_factored(x_true, n, "CSA_SD")
    H = mb.F_OP(([0.0], h_t, W_t))
    frags = mb.CSA_SD_greedy_decomposition(H, 1, decomp_tol=1e-10)
    assert len(frags) == 1
    with mb.CSAEngine(W_t, h_t, "CSA_SD") as e:
        c = e.subtract_fragment(frags[0].x())
    assert c <= 1e-8 * (np.sum(W_t ** 2) + np.sum(h_t ** 2))


def test_sharded_greedy_step_runs_in_lock_step(mb, lib):
    """The library's BFGS on row-panel shards (BASELINE configs[4]): every shard runs the same optimiser on its own handle,
    the evaluations exchange partial vectors over peer memory; all shards must return identical bits, agree with the
    unsharded run over a short trajectory, and the residual update must leave shard-local costs that sum to the total."""
    import threading
    n, world = 20, 3                      # 210 packed rows = 4 panels of 64 over 3 shards
    T, xs = o.planted_tbt(n, 1, 11)
    rng = np.random.default_rng(4)
    x0 = xs[0] + 0.05 * rng.standard_normal(xs[0].shape)
    shards = [mb.CSAEngine(T, shard=r, nshards=world) for r in range(world)]
    try:
        for e in shards:
            e.peer_export()
        for r, e in enumerate(shards):
            for q, peer in enumerate(shards):
                if q != r:
                    e.peer_attach_local(q, peer)
        for e in shards:                   # one process drives all shards here: no allocation once a shard can spin for its peers
            e.optimize_reserve()
        out = [None] * world

        def work(r):
            out[r] = shards[r].optimize(x0, g_tol=1e-9, max_iter=12)

        th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        with mb.CSAEngine(T) as whole:
            xw, rw = whole.optimize(x0, g_tol=1e-9, max_iter=12)
            for r in range(world):
                assert np.array_equal(out[r][0], out[0][0]) and out[r][1]["minimum"] == out[0][1]["minimum"]
            assert out[0][1]["iterations"] == rw["iterations"]
            assert abs(out[0][1]["minimum"] - rw["minimum"]) <= 1e-8 * abs(rw["minimum"]) + 1e-20
            assert rel(out[0][0], xw) <= 1e-8
            # residual update: shard-local costs add up to the unsharded one
            total = sum(e.subtract_fragment(out[0][0]) for e in shards)
            cw = whole.subtract_fragment(out[0][0])
            assert abs(total - cw) <= 1e-10 * abs(cw) + 1e-22 * float(np.sum(T * T))
            assert abs(sum(e.residual_cost() for e in shards) - cw) <= 1e-10 * abs(cw) + 1e-22 * float(np.sum(T * T))
    finally:
        for e in shards:
            e.close()


def test_device_entry_points_match_host_closures(mb, lib):
    import torch
    n = 11
    T = o.dense_sym_tbt(n, 5)
    obt = _obt(n, 5)
    x = o.random_x(n, 6, "CSA_SD")
    dev = torch.device("cuda", 0)
    with mb.CSAEngine(T, obt, "CSA_SD") as e:
        c, g = e.cost_grad(x)
        s = torch.cuda.Stream(dev)
        e.set_stream(s.cuda_stream)
        d_x = torch.from_numpy(x).to(dev)
        d_o = torch.zeros(len(x) + 1, dtype=torch.float64, device=dev)
        d_c = torch.zeros(1, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        e.eval_device(d_x.data_ptr(), d_o.data_ptr())
        e.cost_device(d_x.data_ptr(), d_c.data_ptr())
        e.synchronize()
        out = d_o.cpu().numpy()
        assert abs(out[0] - c) <= 1e-13 * abs(c) and rel(out[1:], g) <= 1e-12
        assert abs(float(d_c.item()) - c) <= 1e-13 * abs(c)


# ------------------------------------------------------------------------------------------------------------
def test_native_bfgs_finds_planted_fragment(mb, lib):
    n = 6
    T, xs = o.planted_tbt(n, 1, 1)
    rng = np.random.default_rng(0)
    x0 = xs[0] + 0.02 * rng.standard_normal(xs[0].shape)
    with mb.CSAEngine(T) as e:
        c_start = e.cost(x0)
        xm, res = e.optimize(x0, g_tol=1e-9, max_iter=500)
        assert res["minimum"] < 1e-12 * float(np.sum(T * T)) and res["minimum"] < c_start
        assert res["evaluations"] >= res["iterations"] > 0
        assert e.cost(xm) == pytest.approx(res["minimum"], abs=1e-14)


def test_native_lbfgs_mode_finds_planted_fragment(mb, lib):
    """lbfgs_memory > 0 (mambo_opt_options_t): limited-memory two-loop direction instead of the dense inverse Hessian;
    same minimum as the dense mode on a planted fragment, and against the oracle's cost at the returned point."""
    n = 7
    T, xs = o.planted_tbt(n, 1, 4)
    rng = np.random.default_rng(1)
    x0 = xs[0] + 0.02 * rng.standard_normal(xs[0].shape)
    with mb.CSAEngine(T) as e:
        xd, rd = e.optimize(x0, g_tol=1e-9, max_iter=500)
        xl, rl = e.optimize(x0, g_tol=1e-9, max_iter=500, lbfgs_memory=12)
        tt = float(np.sum(T * T))
        assert rd["minimum"] < 1e-12 * tt and rl["minimum"] < 1e-12 * tt
        assert rl["iterations"] > 0 and rl["evaluations"] >= rl["iterations"]
        c_l = o.cost_grad_factored(xl, T, need_grad=False)
        c_l = c_l[0] if isinstance(c_l, tuple) else c_l
        assert abs(c_l - rl["minimum"]) <= 1e-12 * tt


def test_device_driven_iterations_agree_with_the_host_loop(mb, lib, monkeypatch):
    """mambo_csa_optimize keeps the scalars of an iteration on the device while the first trial step satisfies the strong-Wolfe
    conditions and hands single iterations back to the host's line search otherwise; MAMBO_OPT_HOSTLOOP=1 forces the host
    loop for every iteration.  Same decisions, same arithmetic up to the order of a few sums: on a non-trivial landscape
    (dense random tensor, random start -- many line searches need the host) both must follow the same trajectory for a
    short run, and on a planted tensor both must converge to the planted fragment; dense and limited-memory mode."""
    n = 8
    T = o.dense_sym_tbt(n, 5)
    x0 = o.random_x(n, 2)
    for mem in (0, 6, 40):                       # 40 > LB_FAST_MAX: the limited-memory mode falls back to the host loop by itself
        with mb.CSAEngine(T) as e:
            monkeypatch.delenv("MAMBO_OPT_HOSTLOOP", raising=False)
            xf, rf = e.optimize(x0, g_tol=1e-9, max_iter=25, lbfgs_memory=mem)
            monkeypatch.setenv("MAMBO_OPT_HOSTLOOP", "1")
            xh, rh = e.optimize(x0, g_tol=1e-9, max_iter=25, lbfgs_memory=mem)
            monkeypatch.delenv("MAMBO_OPT_HOSTLOOP", raising=False)
        assert rf["iterations"] == rh["iterations"] == 25
        assert abs(rf["minimum"] - rh["minimum"]) <= 1e-7 * abs(rh["minimum"])
        assert rel(xf, xh) <= 1e-6
        assert rf["evaluations"] >= rh["evaluations"]      # speculative evaluations behind a halted iteration are counted
        c_chk = o.cost_grad_factored(xf, T, need_grad=False)
        c_chk = c_chk[0] if isinstance(c_chk, tuple) else c_chk
        assert abs(c_chk - rf["minimum"]) <= 1e-10 * abs(c_chk)
    Tp, xs = o.planted_tbt(n, 1, 9)
    xp0 = xs[0] + 0.02 * np.random.default_rng(3).standard_normal(xs[0].shape)
    tt = float(np.sum(Tp * Tp))
    for mem in (0, 8):
        with mb.CSAEngine(Tp) as e:
            xm, res = e.optimize(xp0, g_tol=1e-9, max_iter=800, lbfgs_memory=mem)
        assert res["minimum"] < 1e-9 * tt, (mem, res)        # (the limited-memory run is at 3e-12 tt after 800 iterations, still falling)


def test_native_optimiser_and_oracle_bfgs_converge_to_the_same_fragment(mb, lib):
    """The reference drives Optim.jl's BFGS (HagerZhang line search); the library's optimiser is a strong-Wolfe BFGS and the
    oracle uses scipy's.  Trajectories differ, converged fragments must not: from the same start near a planted fragment
    (N = 10, beyond the N = 6 of the end-to-end test) both minimisers reproduce the same operator, cost and CSA 1-norm."""
    n = 10
    T, xs = o.planted_tbt(n, 1, 21)
    x0 = xs[0] + 0.01 * np.random.default_rng(8).standard_normal(xs[0].shape)
    tt = float(np.sum(T * T))
    with mb.CSAEngine(T) as e:
        xn, rn = e.optimize(x0, g_tol=1e-10, max_iter=3000)
        Wn, _ = e.reconstruct(xn)
    xo, co, ro = o.greedy_step(T, x0=x0, gtol=1e-10, maxiter=3000)
    Wo = o.to_OP(xo, n)[1]
    assert rn["minimum"] <= 1e-14 * tt and co <= 1e-14 * tt
    assert rel(Wn, Wo) <= 1e-6 and rel(Wn, T) <= 1e-6
    l1n = mb.CSA_L1(mb.CSA_x_to_F_FRAG(xn, n))
    l1o = o.CSA_L1(xo[:n * (n + 1) // 2], n)
    assert abs(l1n - l1o) <= 1e-6 * abs(l1o)


def test_greedy_decomposition_matches_oracle_driver(mb, lib):
    """decompose.jl:1-80 end to end at LiH size (N=6 stand-in: planted R=2 + small symmetric noise): same x0s, same
    BFGS (scipy) on both sides -> same fragments, final 1-norm (lcu.jl:139-182) within north_star's 1e-8."""
    n = 6
    T, xs = o.planted_tbt(n, 2, 2, sigma=1e-3)
    rng = np.random.default_rng(3)
    x0s = [xs[k] + 0.05 * rng.standard_normal(xs[k].shape) for k in range(2)] + [o.random_x(n, 9)]
    H = mb.F_OP(([0.0], [0.0], T))
    frags = mb.CSA_greedy_decomposition(H, 2, x0s=x0s, decomp_tol=1e-7, g_tol=1e-11, maxiter=4000)
    oxs, ocosts, t_rem, _ = o.greedy_decomposition(T, 2, x0s=x0s, decomp_tol=1e-7, gtol=1e-11, maxiter=4000)
    assert len(frags) == len(oxs) == 2
    cL = o.cartan_2b_num_params(n)
    l1 = sum(mb.CSA_L1(f) for f in frags)
    l1_o = sum(o.CSA_L1(x[:cL], n) for x in oxs)
    assert abs(l1 - l1_o) <= 1e-8 * abs(l1_o)
    for f, x in zip(frags, oxs):                      # fragment by fragment too
        assert abs(mb.CSA_L1(f) - o.CSA_L1(x[:cL], n)) <= 1e-8 * o.CSA_L1(x[:cL], n)
    # the fragments reproduce the tensor to the same residual as the oracle's
    W = sum(o.reconstruct_factored(f.x(), n)[1] for f in frags)
    assert float(np.sum((T - W) ** 2)) == pytest.approx(float(np.sum(t_rem ** 2)), rel=1e-6, abs=1e-12)


def test_csa_sd_greedy_decomposition_final_norms_match_oracle(mb, lib):
    """decompose.jl:122-205 end to end (N = 5, planted one- and two-body fragments): total CSA 1-norm of the fragments'
    two-body coefficients (lcu.jl:139-182) and one_body_L1 of what is left (lcu.jl:262-267 with ob_correction,
    fermionic.jl:457-479) agree with the oracle's driver to 1e-8."""
    n = 5
    cL = o.cartan_2b_num_params(n)
    T, xs = o.planted_tbt(n, 2, 11, sigma=1e-3)
    rng = np.random.default_rng(4)
    obt = np.zeros((n, n))
    x_sd = []
    for k in range(2):
        lam1 = rng.standard_normal(n)
        U = o.one_body_unitary(xs[k][cL:], n)
        obt += (U * lam1) @ U.T
        x_sd.append(np.concatenate([lam1, xs[k]]))
    obt += 1e-3 * np.diag(rng.standard_normal(n))
    x0s = [x + 0.03 * rng.standard_normal(x.shape) for x in x_sd]
    H = mb.F_OP(([0.0], obt, T))
    frags = mb.CSA_SD_greedy_decomposition(H, 2, x0s=x0s, decomp_tol=1e-9, g_tol=1e-11, maxiter=4000)
    oxs, ocosts, t_rem, o_rem = o.greedy_decomposition(T, 2, obt=obt, flavour="CSA_SD", x0s=x0s, decomp_tol=1e-9, gtol=1e-11,
                                                       maxiter=4000)
    assert len(frags) == len(oxs) == 2
    l1 = sum(mb.CSA_L1(f) for f in frags)
    l1_o = sum(o.CSA_L1(x[n:n + cL], n) for x in oxs)
    assert abs(l1 - l1_o) <= 1e-8 * abs(l1_o)
    # what is left after the fragments, through the engine's own greedy state
    with mb.CSAEngine(T, obt, "CSA_SD") as e:
        for f in frags:
            e.subtract_fragment(f.x())
        Tr, hr = e.get_residual()
    ob_l1 = mb.one_body_L1(mb.F_OP(([0.0], hr, Tr)))
    ob_l1_o = o.one_body_L1(o_rem, t_rem)
    assert abs(ob_l1 - ob_l1_o) <= 1e-8 * max(abs(ob_l1_o), 1.0)


def test_csa_sd_greedy_step_molecule_sizes(mb, lib):
    """H2O / N2 sizes (N = 7, 10; configs[1]) on planted tensors, where the exact minimum is known (cost 0); the molecules
    themselves are in tests/test_molecules_gpu.py."""
    for n in (7, 10):
        T, xs = o.planted_tbt(n, 1, n)
        lam1 = np.random.default_rng(n).standard_normal(n)
        U = o.one_body_unitary(xs[0][o.cartan_2b_num_params(n):], n)
        obt = (U * lam1) @ U.T
        x_true = np.concatenate([lam1, xs[0]])
        x0 = x_true + 0.01 * np.random.default_rng(1).standard_normal(x_true.shape)
        H = mb.F_OP(([0.0], obt, T))
        sol = mb.CSA_SD_greedy_step(H, x0=x0)
        assert sol.minimum < 1e-10 * (float(np.sum(T * T)) + float(np.sum(obt * obt)))


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [20, 36, 44, 60, 70, 84, 92, 110, 116, 124, 130, 140])
def test_large_kernel_buckets_match_oracle(mb, lib, n):
    """The kernel instantiations between the benchmark sizes (every multiple of 8 columns has its own: n-blocks 3, 5, 6, 8, 9, 11,
    12, 14, 15, 16, 17, 18 here; 1, 2, 4, 7, 10, 13, 19, 20, 21 elsewhere in this file): n = 116 and up take the column-split
    panel kernels with k_esum and the shallower operand ring, n = 140 has no T~ look-ahead registers -- the code paths of
    the N = 152 configuration at sizes whose host tensor stays small."""
    T = o.dense_sym_tbt(n, 5)
    x = o.random_x(n, 6)
    c0, g0 = o.cost_grad_factored(x, T)
    with mb.CSAEngine(T) as e:
        c, g = e.cost_grad(x)
        assert e.info()["mode_name"] == "fullsym-packed"
        nc = e.subtract_fragment(x)
    assert abs(c - c0) <= RTOL * abs(c0) and rel(g, g0) <= RTOL
    assert abs(nc - c0) <= RTOL * abs(c0)


@pytest.mark.parametrize("n,two_gemm,reps", [(140, False, 1500), (132, True, 1000), (100, False, 500)])
def test_hot_kernels_are_bit_reproducible_over_many_launches(mb, lib, monkeypatch, n, two_gemm, reps):
    """Regression test of the operand ring of fused::k_ty / k_fused.  The ring slots are read with LDS (generic proxy) and
    refilled by the TMA engine (async proxy); without a proxy fence in front of the consumers' mbarrier arrive the refill could
    overtake operand loads at the ring shapes that refill a slot in the step right after its consumption (N = 129..144 for
    k_ty, N = 121..144 for k_fused): 0.2 - 2 % of the launches then returned a gradient that was off by ~1e-5 relative, which a
    single parity evaluation hardly ever sees.  Every evaluation at a fixed x must return the same bits."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(n)
    G = torch.randn((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
    G = G + G.permute(1, 0, 2, 3)
    G = G + G.permute(0, 1, 3, 2)
    G = (G + G.permute(2, 3, 0, 1)).mul_(1.0 / (n * n)).contiguous()
    x = o.random_x(n, 6)
    if two_gemm:
        monkeypatch.setenv("MAMBO_TWO_GEMM", "1")
    with mb.CSAEngine.from_device(G.data_ptr(), n) as e:
        del G
        d_x = torch.from_numpy(x).to(dev)
        d_out = torch.empty(e.L + 1, dtype=torch.float64, device=dev)
        e.eval_device(d_x.data_ptr(), d_out.data_ptr())
        e.synchronize()
        ref = d_out.clone()
        bad = 0
        for _ in range(reps):
            e.eval_device(d_x.data_ptr(), d_out.data_ptr())
            e.synchronize()
            bad += 0 if torch.equal(d_out, ref) else 1
    assert bad == 0, f"{bad} of {reps} evaluations differ from the first"


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [50, 100])
def test_full_size_properties(mb, lib, n):
    """BASELINE.json sizes.  (1) pointwise parity with the factored oracle at one x; (2) the three symmetry modes of
    the engine agree with each other; (3) planted round trip: subtracting W(x) then evaluating the zero fragment
    gives the cost at x; (4) cost-only evaluation equals the fused evaluation's cost."""
    import torch
    T = o.dense_sym_tbt(n, 0)
    x = o.random_x(n, 1)
    c0, g0 = o.cost_grad_factored(x, T)
    res = {}
    for name, flag in (("full", mb.SYM_AUTO), ("pair", mb.SYM_PAIR)) + ((("general", mb.SYM_GENERAL),) if n <= 50 else ()):
        with mb.CSAEngine(T, sym=flag) as e:
            res[name] = e.cost_grad(x)
            if name == "full":
                assert e.info()["mode_name"] == "fullsym-packed"
                dev = torch.device("cuda", 0)
                d_x = torch.from_numpy(x).to(dev)
                d_c = torch.zeros(1, dtype=torch.float64, device=dev)
                e.cost_device(d_x.data_ptr(), d_c.data_ptr())
                e.synchronize()
                assert abs(float(d_c.item()) - res[name][0]) <= 1e-13 * abs(c0)
                # (5) lanes give the parent's bits; (6) leading eigenpair (initial guess): eigen-residual on the host matrix view
                with e.lane() as ln:
                    cl, gl = ln.cost_grad(x)
                    assert cl == res[name][0] and np.array_equal(gl, res[name][1])
                lam, vec, st = e.leading_eig(max_iter=400)
                v = vec.reshape(-1, order="F")
                Mv = np.reshape(np.asfortranarray(T), (n * n, n * n), order="F") @ v
                assert np.linalg.norm(Mv - lam * v) <= max(10 * st["residual"], 1e-9 * abs(lam)) + 1e-12 * abs(lam)
                assert abs(float(v @ Mv) - lam) <= 1e-10 * abs(lam)
                nc = e.subtract_fragment(x)
                assert abs(nc - c0) <= RTOL * abs(c0)
                assert e.residual_cost() == pytest.approx(c0, rel=1e-11)
    for name, (c, g) in res.items():
        assert abs(c - c0) <= RTOL * abs(c0), name
        assert rel(g, g0) <= RTOL, name


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [152, 160, 168])
def test_largest_configuration_and_last_bucket_against_oracle(mb, lib, n):
    """BASELINE configs[4] size (N = 152, 19 n-blocks) and the last kernel buckets (N = 160: 20 n-blocks; N = 168: 21 n-blocks,
    six-way column split of the panel kernels, the largest N the library takes): the dense-sym tensor is generated on the
    device (4.3 / 5.2 / 6.4 GB) and handed to the library in place (MAMBO_TBT_ON_DEVICE); the oracle evaluates the same array
    once on the host."""
    import torch
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(n)
    G = torch.randn((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
    G = G + G.permute(1, 0, 2, 3)
    G = G + G.permute(0, 1, 3, 2)
    G = (G + G.permute(2, 3, 0, 1)).mul_(1.0 / (n * n)).contiguous()
    # 8-fold symmetric, so the C-ordered torch array read in Julia (column-major) order is the same tensor
    x = o.random_x(n, 1)
    torch.cuda.synchronize()
    with mb.CSAEngine.from_device(G.data_ptr(), n) as e:
        info = e.info()
        c, gr = e.cost_grad(x)
        c1 = e.cost(x + 0.0)
    assert info["mode_name"] == "fullsym-packed" and c1 == c
    T = G.cpu().numpy()
    del G
    torch.cuda.empty_cache()
    c0, g0 = o.cost_grad_factored(x, T)
    rc, rg = abs(c - c0) / abs(c0), rel(gr, g0)
    print(f"N={n}: rows {info['rows']}, cost rel.err {rc:.2e}, gradient max-norm rel.err {rg:.2e}")
    assert rc <= RTOL and rg <= RTOL


@pytest.mark.parametrize("kind", ["dense_sym", "pairsym", "general"])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
def test_one_body_matrix_of_the_resident_residual(mb, lib, kind, flavour):
    """mambo_csa_one_body_matrix (lcu.jl:262-267, fermionic.jl:457-479): obt + 2 sum_r tbt[:,:,r,r] of the device-resident
    residual, before and after a fragment has been subtracted; one_body_L1 from it against the oracle."""
    n = 9
    T = KINDS[kind](n, 21)
    obt = _obt(n, 2)
    obt = (obt + obt.T) / 2
    x = o.random_x(n, 3, flavour)
    with mb.CSAEngine(T, obt if flavour == "CSA_SD" else None, flavour) as e:
        M = e.one_body_matrix(False, None if flavour == "CSA_SD" else obt)
        assert rel(M, obt + o.ob_correction(T)) <= 1e-13
        assert rel(e.one_body_matrix(True, None if flavour == "CSA_SD" else obt), obt + o.ob_correction(T, True)) <= 1e-13
        if kind != "general":      # one_body_L1 diagonalises a symmetric matrix
            assert abs(e.one_body_L1(False, None if flavour == "CSA_SD" else obt) - o.one_body_L1(obt, T)) <= 1e-12 * o.one_body_L1(obt, T)
        e.subtract_fragment(x)
        h, W = o.reconstruct_factored(x, n, flavour)
        o_rem = obt - h if flavour == "CSA_SD" else obt
        M2 = e.one_body_matrix(False, None if flavour == "CSA_SD" else obt)
        assert rel(M2, o_rem + o.ob_correction(T - W)) <= 1e-12
        c, g = e.cost_grad(x)                      # the call leaves the evaluation state intact
    c0, g0 = o.cost_grad_factored(x, T - W, o_rem if flavour == "CSA_SD" else None, flavour)
    assert abs(c - c0) <= RTOL * abs(c0) and rel(g, g0) <= RTOL


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 6, 7, 10, 12])
@pytest.mark.parametrize("flavour", ["CSA", "CSA_SD"])
def test_molecule_sizes_single_kernel_and_main_pipeline_agree(mb, lib, monkeypatch, n, flavour):
    """N <= 12 with an 8-fold symmetric tensor runs as ONE single-CTA kernel (tiny::k_eval); MAMBO_NO_TINY=1 keeps the
    11-launch pipeline.  Both against the oracle, and against each other, at random points and at a planted minimum."""
    T = o.dense_sym_tbt(n, 3)
    obt = _obt(n, 3) if flavour == "CSA_SD" else None
    if obt is not None:
        obt = 0.5 * (obt + obt.T)
    xs = [o.random_x(n, s, flavour) for s in (0, 5)]
    one = mb.CSAEngine(T, obt, flavour)
    monkeypatch.setenv("MAMBO_NO_TINY", "1")
    pipe = mb.CSAEngine(T, obt, flavour)
    monkeypatch.delenv("MAMBO_NO_TINY")
    try:
        for x in xs:
            c1, g1 = one.cost_grad(x)
            assert one.info()["launches_per_eval"] == 1
            c2, g2 = pipe.cost_grad(x)
            assert pipe.info()["launches_per_eval"] > 1
            c0, g0 = o.cost_grad_factored(x, T, obt, flavour)
            for c, g in ((c1, g1), (c2, g2)):
                assert abs(c - c0) <= RTOL * abs(c0)
                assert rel(g, g0) <= RTOL
            assert abs(one.cost(x) - c0) <= RTOL * abs(c0)
        assert abs(one.residual_cost() - pipe.residual_cost()) <= 1e-13 * pipe.residual_cost()
    finally:
        one.close()
        pipe.close()


@pytest.mark.parametrize("n", [13, 16])
def test_single_kernel_evaluation_up_to_its_largest_size(mb, lib, monkeypatch, n):
    """MAMBO_TINY_MAX raises the threshold up to the kernel's limit (N = 16, 136 packed rows, 182 KB of shared memory)."""
    monkeypatch.setenv("MAMBO_TINY_MAX", "16")
    T = o.dense_sym_tbt(n, 1)
    obt = _obt(n, 1)
    obt = 0.5 * (obt + obt.T)
    for flavour, ob in (("CSA", None), ("CSA_SD", obt)):
        x = o.random_x(n, 2, flavour)
        with mb.CSAEngine(T, ob, flavour) as e:
            c, g = e.cost_grad(x)
            assert e.info()["launches_per_eval"] == 1
        c0, g0 = o.cost_grad_factored(x, T, ob, flavour)
        assert abs(c - c0) <= RTOL * abs(c0)
        assert rel(g, g0) <= RTOL


@pytest.mark.parametrize("n,flavour", [(6, "CSA"), (7, "CSA_SD"), (10, "CSA"), (12, "CSA_SD")])
def test_fused_and_graphed_optimiser_iterations_follow_the_same_trajectory(mb, lib, monkeypatch, n, flavour):
    """Single-CTA problems (num_params <= 256) run a device-driven iteration as evaluation + two fused kernels around the pass
    over the inverse Hessian, replayed from a CUDA graph on single-kernel handles.  Same arithmetic in the same order as the
    eight-launch path (MAMBO_OPT_NO_FUSE=1, MAMBO_OPT_NO_GRAPH=1): bit-identical minimiser, minimum and iteration count."""
    T = o.dense_sym_tbt(n, 4)
    obt = None
    if flavour == "CSA_SD":
        obt = _obt(n, 4)
        obt = 0.5 * (obt + obt.T)
    x0 = o.random_x(n, 7, flavour)
    runs = []
    with mb.CSAEngine(T, obt, flavour) as e:
        assert e.info()["num_params"] <= 256
        for env in ({}, {"MAMBO_OPT_NO_GRAPH": "1"}, {"MAMBO_OPT_NO_FUSE": "1", "MAMBO_OPT_NO_GRAPH": "1"}):
            for k in ("MAMBO_OPT_NO_FUSE", "MAMBO_OPT_NO_GRAPH"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            runs.append(e.optimize(x0, g_tol=1e-10, max_iter=120))
    for k in ("MAMBO_OPT_NO_FUSE", "MAMBO_OPT_NO_GRAPH"):
        monkeypatch.delenv(k, raising=False)
    (xa, ra), (xb, rb), (xc, rc) = runs
    assert ra["iterations"] == rb["iterations"] == rc["iterations"] and ra["iterations"] > 20
    assert ra["minimum"] == rb["minimum"] == rc["minimum"]
    assert np.array_equal(xa, xb) and np.array_equal(xa, xc)
    c0, g0 = o.cost_grad_factored(xa, T, obt, flavour)
    assert abs(c0 - ra["minimum"]) <= 1e-10 * abs(c0)
