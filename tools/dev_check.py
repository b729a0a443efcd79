"""Developer parity sweep on a GPU box: CUDA engine vs oracle (factored numpy) for several N / modes."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import mambo_b200 as m
import csa_oracle as o

def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

def run(n, kind, sym, flavour, seed=0):
    T = kind(n, seed)
    obt = None
    if flavour == "CSA_SD":
        obt = np.random.default_rng(7).standard_normal((n, n)) / n
    x = o.random_x(n, seed + 1, flavour)
    c0, g0 = o.cost_grad_factored(x, T, obt, flavour)
    with m.CSAEngine(T, obt, flavour, sym=sym) as e:
        c1, g1 = e.cost_grad(x)
        info = e.info()
        rc = e.residual_cost()
        rc0 = o.L2_partial_cost(T, None, obt, None)
        ok = abs(c1 - c0) / abs(c0) < 1e-10 and rel(g1, g0) < 1e-10
        print(f"N={n:3d} {kind.__name__:14s} {flavour:6s} mode={info['mode_name']:15s} cost_rel={abs(c1-c0)/abs(c0):.2e} grad_rel={rel(g1,g0):.2e} "
              f"rescost_rel={abs(rc-rc0)/rc0:.2e} launches={info['launches_per_eval']} {'OK' if ok else 'FAIL'}", flush=True)
        if n <= 20:
            # subtract + residual + reconstruct round trip
            h1, W1 = o.reconstruct_factored(x, n, flavour)
            Wd, hd = e.reconstruct(x)
            nc = e.subtract_fragment(x)
            Tr, orr = e.get_residual()
            print(f"      reconstruct_rel={rel(Wd, W1):.2e} subtract_cost_rel={abs(nc-c0)/abs(c0):.2e} residual_rel={rel(Tr, T-W1):.2e}"
                  + (f" obt_rel={rel(orr, obt-h1):.2e} hrec_rel={rel(hd,h1):.2e}" if obt is not None else ""), flush=True)
        return ok

if __name__ == "__main__":
    allok = True
    for n in (3, 6, 10, 17, 24):
        for flavour in ("CSA", "CSA_SD"):
            allok &= run(n, o.dense_sym_tbt, m.SYM_AUTO, flavour)
            allok &= run(n, o.dense_sym_tbt, m.SYM_PAIR, flavour)
            allok &= run(n, o.pairsym_tbt, m.SYM_AUTO, flavour)
            allok &= run(n, o.general_tbt, m.SYM_AUTO, flavour)
    for n in (50,):
        allok &= run(n, o.dense_sym_tbt, m.SYM_AUTO, "CSA")
        allok &= run(n, o.dense_sym_tbt, m.SYM_PAIR, "CSA")
    print("ALL OK" if allok else "SOME FAILED")
    # timing at N=100
    n = 100
    t0 = time.time(); T = o.dense_sym_tbt(n, 0); print("gen", time.time() - t0)
    x = o.random_x(n, 1)
    for sym in (m.SYM_FULL, m.SYM_PAIR):
        t0 = time.time()
        with m.CSAEngine(T, sym=sym) as e:
            print("create", time.time() - t0, e.info())
            c, g = e.cost_grad(x)
            xs = [o.random_x(n, 10 + i) for i in range(8)]
            t0 = time.time()
            for xx in xs * 3:
                e.cost_grad(xx)
            dt = (time.time() - t0) / 24
            print(f"N=100 sym={sym} host-API eval {dt*1e3:.3f} ms  -> {1/dt:.1f} evals/s")
            if sym == m.SYM_FULL:
                t0 = time.time(); c0, g0 = o.cost_grad_factored(x, T); print("oracle s", time.time() - t0)
                print("N=100 cost_rel", abs(c - c0) / abs(c0), "grad_rel", rel(g, g0))
