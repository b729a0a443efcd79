#!/usr/bin/env python
"""bench.py -- CSA fragment cost+gradient throughput (BASELINE.json metric) on N GPUs of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference algorithm's CPU path (oracle port) on the host cores

Workload (config.workload): synthetic dense-sym two-body tensor, N=100 (SURVEY.md 8d / BASELINE.json configs[3]),
seed 0; one *step* = one fused cost+gradient evaluation at each of 8 distinct points x (theta ~ U[0,2pi),
lambda ~ N(0,1)/N).  metric = cost+grad evaluations per second, FP64.

  value   : device-resident leg -- x and (cost, grad) live in HBM, CUDA-event timed on the launching stream.
  e2e     : the same evaluations through the public host API (CSAEngine.cost_grad == the C-ABI call the Julia
            shim makes): host x in, H2D, kernels, D2H of cost+grad, inside the timed region.
  N > 1   : multistart replicas: every rank holds the tensor and evaluates its own batch, no
            data-path collective -> "scaling": "weak".  The row-panel-sharded variant (8e.2, NCCL allreduce of
            1+2N^2 doubles per evaluation) is measured as well and reported under "rowpanel".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 8  # evaluation points per step


def workload_name(n: int) -> str:
    """config.workload: the same string on both arms (ours / --impl reference)."""
    return f"dense-sym N={n} seed 0, CSA cost+grad at {BATCH} points/step per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100, help="orbitals of the synthetic tensor")
    ap.add_argument("--sym", default="auto", choices=["auto", "general", "pair", "full"],
                    help="symmetry mode of the engine (auto detects the 8-fold symmetry of the synthetic tensor)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rowpanel", action="store_true")
    ap.add_argument("--cpu-evals", type=int, default=0, help="evaluations timed for the CPU baseline (0 = auto)")
    ap.add_argument("--no-greedy", action="store_true", help="skip the greedy-CSA time-to-tolerance leg")
    ap.add_argument("--no-df", action="store_true", help="skip the double-factorisation leg")
    ap.add_argument("--lanes", type=int, default=0,
                    help="evaluation lanes per GPU (mambo_csa_create_lane); 0 = auto (see main), 1..8")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def _nvml_loop(self) -> bool:
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
        while not self._halt.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append([str(sm), str(mx), str(pw)] + ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self._halt.wait(0.01)
        return True

    def run(self):
        if self._nvml_loop():
            return
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": sorted(reasons),
                "samples": len(self.samples), "power_w_max": max(float(s[2]) for s in self.samples)}


# ------------------------------------------------------------------------------------------------------------
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core, so lift the BLAS pool
    limit at run time and report the thread count actually in force."""
    cores = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=cores)
        blas = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(blas) if blas else cores
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", cores))


def cpu_reference_rate(n: int, T, xs, budget_s: float, max_evals: int):
    """The reference algorithm on the host cores: the oracle's factored restatement (numpy/OpenBLAS, all
    threads).  Returns (evals_per_s, evals_timed, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import csa_oracle as oracle  # noqa: E402  (checker / baseline only -- never on the product path)
    cores = use_all_host_threads()
    oracle.cost_grad_factored(xs[0], T)  # warm-up (BLAS thread pool, page faults)
    t0 = time.perf_counter()
    done = 0
    while done < max_evals:
        oracle.cost_grad_factored(xs[done % len(xs)], T)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, cores


def cpu_packed_rate(n: int, T, xs, budget_s: float, max_evals: int):
    """The engine's own ALGORITHM on the host cores (oracle.cost_grad_packed: packed rows + lean single-GEMM formulation,
    numpy/OpenBLAS): splits the GPU/CPU ratio into its algorithmic (8x fewer flops than the factored port) and hardware
    parts.  The packed residual is prepared once, like the engine's device-resident T~."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import csa_oracle as oracle  # noqa: E402
    cores = use_all_host_threads()
    packed = oracle.pack_tbt(T)
    oracle.cost_grad_packed(xs[0], packed, n)
    t0 = time.perf_counter()
    done = 0
    while done < max_evals:
        oracle.cost_grad_packed(xs[done % len(xs)], packed, n)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, cores


def cpu_literal_fit(sizes=(6, 10, 14, 18), n_target: int = 100):
    """the reference's loop nests as plain single-threaded C (oracle/csa_literal.c: N^6
    reconstruction, dense N^6 del_w_u, uL x N^6 theta contraction -- what Einsum.jl expands to) timed at small N, fitted
    to t = a N^p, and EXTRAPOLATED to the bench size (the literal path needs 8 N^6 bytes = 8 TB of scratch at N = 100 and
    cannot run there)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import csa_oracle as oracle  # noqa: E402
    import csa_literal as lit    # noqa: E402
    pts = {}
    for m in sizes:
        T = oracle.dense_sym_tbt(m, 0)
        x = oracle.random_x(m, 1)
        t0 = time.perf_counter()
        lit.cost_grad_literal_c(x, T)
        pts[m] = time.perf_counter() - t0
    ns = np.log(np.array(list(pts.keys()), dtype=float)[1:])
    ts = np.log(np.array(list(pts.values()))[1:])
    p, loga = np.polyfit(ns, ts, 1)
    t_target = float(np.exp(loga) * n_target ** p)
    return {"what": "oracle/csa_literal.c (reference loop nests, 1 thread) cost+grad, seconds per evaluation",
            "seconds_per_eval": {str(k): v for k, v in pts.items()}, "fitted_exponent": float(p),
            f"extrapolated_seconds_per_eval_N{n_target}": t_target,
            f"extrapolated_evals_per_s_N{n_target}": 1.0 / t_target,
            "note": f"EXTRAPOLATED from N <= {max(sizes)} with t = a N^p (the O(N^8) theta gradient dominates); not runnable at "
                    f"N = {n_target} (8 N^6 B of scratch)", "cores": 1, "kind": "port"}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path (oracle port) timed on this box's host cores, rank 0 only."""
    if rank != 0:
        return
    from mambo_b200 import pkg  # only for the synthetic generator (no GPU work on this arm)
    import importlib
    syn = importlib.import_module("quantummambo_jl_b200.synthetic")
    n = args.n
    T = syn.dense_sym_tbt(n, 0)
    xs = [syn.random_x(n, i) for i in range(BATCH)]
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import csa_oracle as oracle
    cores = use_all_host_threads()
    for _ in range(max(args.warmup, 1)):
        oracle.cost_grad_factored(xs[0], T)
    # each step evaluates all BATCH points, like a step of the GPU arm (about 0.5 s per point at N = 100 on 16 host threads)
    sample = BATCH
    t0 = time.perf_counter()
    for s in range(args.steps):
        for i in range(sample):
            oracle.cost_grad_factored(xs[(s * sample + i) % BATCH], T)
    dt = time.perf_counter() - t0
    evals = args.steps * sample
    v = evals / dt
    line = {
        "impl": "reference", "metric": "csa_cost_grad_evals_per_s", "value": v, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "N": n, "points_per_step": BATCH, "points_timed_per_step": sample},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"all {BATCH} points per step x {args.steps} steps, oracle factored numpy/OpenBLAS, "
                                   f"{cores} threads (the literal reference needs 8*N^6 B = 8 TB scratch at N=100: not runnable)"},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def greedy_time_to_tolerance(mb, syn, device: int, n: int = 50, R: int = 4, tol_rel: float = 1e-2, alpha_max: int = 8,
                             max_iter: int = 1000, cpu_budget_s: float = 25.0):
    """Second half of BASELINE.json's metric: greedy CSA (decompose.jl:1-80) on a planted tensor -- the sum of R random CSA
    fragments plus 1e-9 symmetric noise -- run to a tolerance on the SQUARED-norm residual (decomp_tol, decompose.jl:45),
    every step with Optim's full budget of 1000 BFGS iterations (g_tol 1e-8) from the SVD guess of the residual
    (tbt_svd_1st on the device).  Greedy CSA peels a planted sum only approximately (the planted fragments are not
    orthogonal: ~25 % of the cost per fragment for the first R, then slowly), so the tolerance is stated relative to the
    initial cost and chosen reachable: tol = tol_rel * initial.  The oracle's driver (scipy BFGS on the factored numpy
    port, all host threads) runs the FIRST step of the same problem from the same x0 beside it, bounded in wall time."""
    rng = np.random.default_rng(2000)
    cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
    T = np.zeros((n, n, n, n), order="F")
    with mb.CSAEngine(np.zeros((n, n, n, n)), device=device) as e:
        for _ in range(R):
            W, _ = e.reconstruct(np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)]))
            T += W
    g = rng.standard_normal((n, n, n, n))
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    T = np.asfortranarray(T + 1e-9 * g / 8.0)
    del g

    def run(lbfgs_memory: int):
        trail, evals, its, l1, first = [], 0, 0, 0.0, None
        with mb.CSAEngine(T, device=device) as e:
            c_init = c = e.residual_cost()
            tol = tol_rel * c_init
            e.leading_eig(max_iter=8)                                                     # Lanczos workspace up front
            e.optimize_reserve(lbfgs_memory)
            t0 = time.perf_counter()
            x0_first = None
            while c > tol and len(trail) < alpha_max:
                lam, th = mb.tbt_svd_1st(e)                                               # guesses.jl:51-84 on the resident residual
                x0 = np.concatenate([lam, th])
                if x0_first is None:
                    x0_first = x0.copy()
                t1 = time.perf_counter()
                xm, res = e.optimize(x0, g_tol=1e-8, max_iter=max_iter, lbfgs_memory=lbfgs_memory)
                c = e.subtract_fragment(xm)
                if first is None:
                    first = {"seconds": time.perf_counter() - t1, "evaluations": res["evaluations"],
                             "iterations": res["iterations"], "cost": c}
                trail.append(c)
                evals += res["evaluations"]
                its += res["iterations"]
                l1 += float(mb.CSA_L1(mb.CSA_x_to_F_FRAG(xm, n)))
            dt = time.perf_counter() - t0
        return {"seconds": dt, "fragments": len(trail), "evaluations": evals, "iterations": its, "initial_cost": c_init,
                "tolerance": tol, "reached": bool(c <= tol), "final_cost": c, "cost_trail": trail, "final_l1_norm": l1,
                "ms_per_evaluation_incl_optimizer": dt / max(evals, 1) * 1e3, "first_step": first}, x0_first

    out, x0_first = run(0)
    try:   # HBM roofline of the dense BFGS's pass over the inverse Hessian, at this leg's L and at the bench tensor's (N = 100)
        import ctypes as C
        lib = mb.load()
        peak = None
        try:
            peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        out["bfgs_pass_roofline"] = {}
        for L_ in (n * n, 100 * 100):
            ms_, by_ = C.c_double(), C.c_double()
            if lib.mambo_debug_bfgs_pass(device, L_, 20, C.byref(ms_), C.byref(by_)) == 0:
                gbs = by_.value / (ms_.value * 1e-3) / 1e9
                out["bfgs_pass_roofline"][f"L={L_}"] = {
                    "bound": "hbm", "kernel": "k_sym_update_gemv + k_sym_reduce (rank-2 update of the symmetric inverse Hessian fused with H g)",
                    "achieved": gbs, "peak": peak or 6534.5, "unit": "GB/s", "frac": gbs / (peak or 6534.5), "ms_per_pass": ms_.value,
                    "algorithmic_bytes": by_.value, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peak else "fallback 6534.5 GB/s (file absent)"}
    except Exception as exc:  # noqa: BLE001
        out["bfgs_pass_roofline"] = {"error": repr(exc)}
    out["workload"] = (f"planted N={n} R={R} + 1e-9 noise, greedy CSA to {tol_rel:g} x the initial squared-norm cost, <= {alpha_max} "
                       f"fragments x <= {max_iter} BFGS iterations (native dense BFGS like Optim.jl's), SVD start")
    out["lbfgs_memory_10"], _ = run(10)
    # ---- the oracle's driver on the host cores: first greedy step of the same problem, same x0, bounded in time ------------
    if cpu_budget_s > 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import csa_oracle as oracle  # noqa: E402  (baseline only)
        cores = use_all_host_threads()
        Tc = np.ascontiguousarray(T)
        t0 = time.perf_counter()
        oracle.greedy_step(Tc, x0=x0_first, maxiter=4)           # times an iteration INCLUDING scipy's dense L x L BFGS update
        t_it = (time.perf_counter() - t0) / 5.0
        it_cpu = int(max(5, min(max_iter, cpu_budget_s / t_it)))
        t0 = time.perf_counter()
        _, c_cpu, r_cpu = oracle.greedy_step(Tc, x0=x0_first, maxiter=it_cpu)
        dt_cpu = time.perf_counter() - t0
        first = out["first_step"]
        out["cpu_first_step"] = {
            "what": "oracle.greedy_step (scipy BFGS, factored numpy/OpenBLAS port) on the first fragment from the same x0",
            "cores": cores, "kind": "port", "seconds": dt_cpu, "iterations": int(r_cpu.nit), "evaluations": int(r_cpu.nfev),
            "iteration_budget": it_cpu, "cost": float(c_cpu), "ms_per_evaluation_incl_optimizer": dt_cpu / max(int(r_cpu.nfev), 1) * 1e3,
            "gpu_over_cpu_per_evaluation": (dt_cpu / max(int(r_cpu.nfev), 1)) / (first["seconds"] / max(first["evaluations"], 1)),
            "note": "bounded sample: the CPU run stops at its iteration budget, so compare per-evaluation times (optimiser "
                    "included) and the cost reached, not the wall times"}
    # BASELINE.json configs[2]: 8-way multistart of the first greedy step (random theta guesses), here on ONE GPU:
    # the 8 runs one after the other on one handle versus spread over evaluation lanes that share the residual
    K, ms_iter = 8, 300
    x0s = np.stack([np.concatenate([np.zeros(cL), 2 * np.pi * rng.random(uL)]) for _ in range(K)])
    with mb.CSAEngine(T, device=device) as e:
        lanes = [e] + [e.lane() for _ in range(K - 1)]
        t0 = time.perf_counter()
        _, r1, b1 = mb.multistart(lanes[:1], x0s, max_iter=ms_iter)
        t_seq = time.perf_counter() - t0
        t0 = time.perf_counter()
        _, r8, b8 = mb.multistart(lanes, x0s, max_iter=ms_iter)
        t_lanes = time.perf_counter() - t0
        for ln in lanes[1:]:
            ln.close()
    out["multistart"] = {"workload": f"{K} random-theta starts of the first greedy step, planted N={n} R={R}, <= {ms_iter} BFGS iterations each",
                         "seconds_one_handle": t_seq, "seconds_8_lanes": t_lanes,
                         "evaluations": sum(r["evaluations"] for r in r8), "best_start": b8,
                         "minima_agree": bool(max(abs(a["minimum"] - b["minimum"]) for a, b in zip(r1, r8))
                                              <= 1e-9 * max(abs(r["minimum"]) for r in r1)),
                         "best_minimum": r8[b8]["minimum"]}
    return out



def df_leg(mb, syn, eng, device: int, hbm_peak, n_cpu: int = 50, cpu: bool = True):
    """DF_decomposition (decompose.jl:297-398) + DF_based_greedy (:400-436) of the resident residual on the
    device (Lanczos tridiagonalisation to completion, bisection + twisted factorisation, DMMA back-transformation, batched
    Jacobi for the fragment matrices).  `eng` holds the bench's own N = 100 tensor; the oracle's LAPACK path (the reference
    algorithm: `eigh` of the N^2 x N^2 view) runs beside it at N = n_cpu, where it finishes in seconds."""
    out = {}

    def one(e, tag):
        info = e.info()
        R = info["rows"]
        e.df_decompose(max_frags=1)                                   # warm-up: workspaces, attributes
        t0 = time.perf_counter()
        st = e.df_decompose(verify=False)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        lam, U1, Lm = e.df_based_greedy()
        t_g = time.perf_counter() - t1
        m = st["krylov_dim"]
        # HBM traffic of the tridiagonalisation: one pass over T~ per step plus, twice per step, a dot pass and a projection
        # pass over the j + 1 Krylov vectors found so far
        lanczos_bytes = 8.0 * R * R * m + 4.0 * 8.0 * R * (m * (m + 1) / 2.0)
        gbs = lanczos_bytes / max(st["seconds_lanczos"], 1e-12) / 1e9
        out[tag] = {"N": info["N"], "rows": R, "fragments": st["num_frags"], "krylov_dim": m, "restarts": st["restarts"],
                    "seconds_total": dt, "seconds_lanczos": st["seconds_lanczos"], "seconds_tridiagonal": st["seconds_tridiag"],
                    "seconds_backtransform": st["seconds_backtransform"], "seconds_fragment_eigen": st["seconds_fragments"],
                    "seconds_df_based_greedy": t_g, "ortho_dev": st["ortho_dev_after"], "jacobi_sweeps_max": st["jacobi_sweeps_max"],
                    "roofline": {"bound": "hbm", "kernel": "Lanczos step: lanczos::k_symv + k_dots + df::k_project_split (x2)",
                                 "achieved": gbs, "peak": hbm_peak or 6534.5, "unit": "GB/s", "frac": gbs / (hbm_peak or 6534.5),
                                 "algorithmic_bytes": lanczos_bytes, "traffic": None,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback 6534.5 GB/s (file absent)"}}
        return st

    one(eng, "device")
    if cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import csa_oracle as oracle  # noqa: E402  (baseline only)
        cores = use_all_host_threads()
        Tc = syn.dense_sym_tbt(n_cpu, 0)
        with mb.CSAEngine(Tc, device=device) as e2:
            one(e2, "device_small")
            coeff, _, _, _, _ = e2.df_get(0, out["device_small"]["fragments"], want_L=False, want_U=False)
        t0 = time.perf_counter()
        fr = oracle.DF_decomposition(Tc)
        t_cpu = time.perf_counter() - t0
        c_ref = np.array([f[0] for f in fr])
        out["cpu_small"] = {"what": f"oracle.DF_decomposition (numpy/LAPACK eigh of the N^2 x N^2 view + one eigh per fragment), N={n_cpu}",
                            "cores": cores, "kind": "port", "seconds": t_cpu, "fragments": len(fr),
                            "coeff_parity_rel": float(np.max(np.abs(coeff - c_ref)) / np.max(np.abs(c_ref))) if len(fr) == len(coeff) else None,
                            "gpu_speedup": t_cpu / out["device_small"]["seconds_total"]}
    return out

# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import mambo_b200 as mb
    import importlib
    syn = importlib.import_module("quantummambo_jl_b200.synthetic")

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CSA engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    sym = {"auto": mb.SYM_AUTO, "general": mb.SYM_GENERAL, "pair": mb.SYM_PAIR, "full": mb.SYM_FULL}[args.sym]
    T = syn.dense_sym_tbt(n, 0)                       # same tensor on every rank (multistart replicas)
    xs = [syn.random_x(n, rank * BATCH + i) for i in range(BATCH)]   # distinct start points per rank
    eng = mb.CSAEngine(T, flavour="CSA", device=local_rank, sym=sym)
    info = eng.info()
    L = eng.L

    # ---- measured FP64 peak (cuBLAS DGEMM through torch) : the tensor roofline denominator -------------------
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    dgemm_tflops = 2.0 * 4096 ** 3 / (best * 1e-3) / 1e12
    del a, b

    # ---- evaluation lanes: independent evaluations of the step overlap on the device (mambo_csa_create_lane) ----------
    tiles = (n + 15) // 16
    # default: twice as many lanes as orbital-rotation chains fit on the device at two CTAs per SM (measured at N=100:
    # 2 / 4 / 8 lanes -> 3.3k / 4.35k / 4.48k evals/s; the queued chains start as soon as SMs free up), rounded down to a
    # divisor of the step's BATCH points so that every lane carries the same number of evaluations
    nl = args.lanes if args.lanes > 0 else max(1, min(8, (4 * info["num_sms"]) // (tiles * tiles)))
    nl = min(nl, BATCH)
    if args.lanes <= 0:
        while BATCH % nl:
            nl -= 1
    group = [eng] + [eng.lane() for _ in range(nl - 1)]
    main_stream = torch.cuda.Stream(dev)      # real side streams: handle 0 (legacy default) would mean "own stream"
    torch.cuda.set_stream(main_stream)
    streams = [torch.cuda.Stream(dev) for _ in range(nl)]
    d_x = torch.from_numpy(np.stack(xs)).to(dev)                     # [BATCH, L] in HBM
    d_out = torch.empty(BATCH, L + 1, dtype=torch.float64, device=dev)

    def device_leg(lanes: int, profile: bool):
        """K steps of BATCH evaluations, x and (cost, grad) resident in HBM, round-robin over `lanes` lanes; CUDA events
        on a main stream that the lane streams fork from and join into.  Returns (ms, launches, prof, segs)."""
        act, sts = group[:lanes], streams[:lanes]
        for e, st in zip(act, sts):
            e.set_stream(st.cuda_stream)

        def step():
            for i in range(BATCH):
                act[i % lanes].eval_device(d_x[i].data_ptr(), d_out[i].data_ptr())

        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
        l0 = sum(e.get_profile()["total_launches"] for e in act)
        if profile:
            act[0].set_profiling(True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_stream)
        for st in sts:
            st.wait_stream(main_stream)
        for _ in range(args.steps):
            step()
        for st in sts:
            main_stream.wait_stream(st)
        e1.record(main_stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms_ = e0.elapsed_time(e1)
        segs_ = act[0].get_segments() if profile else None
        prof_ = act[0].get_profile() if profile else None
        if profile:
            act[0].set_profiling(False)
        launches_ = sum(e.get_profile()["total_launches"] for e in act) - l0
        if world > 1:
            t = torch.tensor([ms_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        for e in act:
            e.set_stream(None)
        return ms_, launches_, prof_, segs_

    sampler = ClockSampler(local_rank)
    sampler.start()
    # single lane: every kernel runs alone, so the fused kernel's CUDA-event time is its own duration (roofline)
    ms1, launches1, prof, segs = device_leg(1, True)
    info = eng.info()
    # all lanes: the headline device-resident number
    if nl > 1:
        ms, launches, _, _ = device_leg(nl, False)
    else:
        ms, launches = ms1, launches1
    evals_total = world * BATCH * args.steps
    value = evals_total / (ms * 1e-3)
    value1 = evals_total / (ms1 * 1e-3)
    cost_check = float(d_out[0, 0].item())

    # ---- end-to-end leg: public host API, host buffers, H2D + D2H inside the timed region --------------------------
    def host_leg(lanes: int):
        """The same K x BATCH evaluations through the synchronous host closure (mambo_csa_cost_grad: host x in, H2D,
        kernels, D2H of cost+grad), one caller thread per lane, a fresh x every call (the memo never hits)."""
        act = group[:lanes]
        for k, e in enumerate(act):
            e.cost_grad(xs[k] + 0.0)
        acc = [0.0] * lanes

        def worker(k):
            for s_ in range(args.steps):
                for i in range(k, BATCH, lanes):
                    xi = xs[i].copy()
                    xi[0] += 1e-9 * (s_ + 1)
                    c, g = act[k].cost_grad(xi)
                    acc[k] += c

        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if lanes == 1:
            worker(0)
        else:
            th = [threading.Thread(target=worker, args=(k,)) for k in range(lanes)]
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
        dt_ = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_ = float(t.item())
        return dt_

    # ---- initial-guess leg: leading eigenpair of the resident residual, HBM-bound matvec ---------------
    guess = None
    if rank == 0:
        try:
            eng.set_stream(main_stream.cuda_stream)
            rp = info["rows_padded"]
            d_v = torch.zeros(rp, dtype=torch.float64, device=dev)
            d_v[:info["rows"]] = 1.0 / np.sqrt(info["rows"])
            d_y = torch.zeros(rp, dtype=torch.float64, device=dev)
            for _ in range(5):
                eng.symv_device(d_v.data_ptr(), d_y.data_ptr())
            torch.cuda.synchronize()
            reps = 50
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(main_stream)
            for _ in range(reps):
                eng.symv_device(d_v.data_ptr(), d_y.data_ptr())
            eb.record(main_stream)
            torch.cuda.synchronize()
            mv_ms = ea.elapsed_time(eb) / reps
            eng.set_stream(None)
            eng.leading_eig(max_iter=8)                      # allocates the Lanczos workspace
            t0 = time.perf_counter()
            lam_top, _, st = eng.leading_eig()
            t_eig = time.perf_counter() - t0
            hbm_peak = None
            try:
                hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            mv_bytes = 8.0 * info["rows"] * info["rows"]
            gbs = mv_bytes / (mv_ms * 1e-3) / 1e9
            guess = {"what": "leading eigenpair of the residual's N^2 x N^2 view (tbt_svd_1st, guesses.jl:51-63) by device Lanczos",
                     "seconds": t_eig, "iterations": st["iterations"], "rel_residual": st["residual"] / max(abs(lam_top), 1e-300),
                     "eigenvalue": lam_top,
                     "roofline": {"bound": "hbm", "kernel": "lanczos::k_symv", "achieved": gbs, "peak": hbm_peak or 6534.5,
                                  "unit": "GB/s", "frac": gbs / (hbm_peak or 6534.5),
                                  "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback 6534.5 GB/s (file absent)",
                                  "algorithmic_bytes_per_launch": mv_bytes, "ms_per_launch": mv_ms, "launches_timed": reps,
                                  "traffic": None}}
        except Exception as exc:  # noqa: BLE001  (a side leg must never cost the main line)
            guess = {"error": repr(exc)}
            eng.set_stream(None)

    dfres = None
    if rank == 0 and world == 1 and not args.no_df:
        try:
            hbm = None
            try:
                hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            dfres = df_leg(mb, syn, eng, local_rank, hbm, cpu=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001  (a side leg must never cost the main line)
            dfres = {"error": repr(exc)}

    # caller threads per rank: one per lane, but never more than this rank's share of the host cores (the synchronous closure
    # spins in cudaStreamSynchronize by default: 64 spinning threads on a 32-core host added nothing in round 1).  When the
    # callers of the whole box still outnumber the cores the closures sleep on a blocking-sync event instead.
    cores_total = os.cpu_count() or 1
    host_nl = max(1, min(nl, cores_total // max(world, 1)))
    while BATCH % host_nl:
        host_nl -= 1
    # wait mode per leg: spinning callers (lowest latency) as long as every spinning thread of the box has a core of its own,
    # blocking-sync events otherwise
    def set_wait(threads_per_rank: int) -> bool:
        blk = world * threads_per_rank > max(1, cores_total - 2)
        for e in group:
            e.set_wait_mode(blk)
        return blk

    def batch_leg():
        """The step's BATCH points through ONE C-ABI call per step (mambo_csa_cost_grad_batch over the lanes): host x in,
        H2D, kernels, D2H of every cost+grad inside the timed region, a single caller thread."""
        xb = np.stack(xs)
        mb.cost_grad_batch(group, xb)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc_ = 0.0
        for s_ in range(args.steps):
            xb[:, 0] += 1e-9
            c_, g_ = mb.cost_grad_batch(group, xb)
            acc_ += float(c_[0])
        dt_ = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_ = float(t.item())
        return dt_

    blocking1 = set_wait(1)
    t_e2e1 = host_leg(1)
    blocking = set_wait(host_nl)
    t_e2e = host_leg(host_nl) if host_nl > 1 else t_e2e1
    set_wait(1)
    t_batch = batch_leg()
    sampler.stop()
    e2e_threads = evals_total / t_e2e
    e2e_value1 = evals_total / t_e2e1
    e2e_batch = evals_total / t_batch
    # both are the public host API with host buffers (H2D + D2H inside the timed region): one synchronous closure call per
    # point from one caller thread per lane, or ONE mambo_csa_cost_grad_batch call per step from a single caller thread
    e2e_value, e2e_api = (e2e_batch, "mambo_csa_cost_grad_batch") if e2e_batch >= e2e_threads else (e2e_threads, "mambo_csa_cost_grad")
    for e in group[1:]:
        e.close()
    stream = main_stream

    # ---- row-panel sharded variant (strong scaling, NCCL allreduce of the partial vector) -----------------------
    rowpanel = None
    if world > 1 and not args.no_rowpanel:
        eng.close()
        torch.cuda.empty_cache()
        engs = mb.CSAEngine(T, flavour="CSA", device=local_rank, sym=sym, shard=rank, nshards=world)
        engs.set_stream(stream.cuda_stream)
        plen = engs.info()["partial_len"]
        xs0 = [syn.random_x(n, i) for i in range(BATCH)]              # the same points on every rank
        d_x0 = torch.from_numpy(np.stack(xs0)).to(dev)
        d_part = torch.empty(plen, dtype=torch.float64, device=dev)

        def step_nccl():
            for i in range(BATCH):
                engs.eval_partial_device(d_x0[i].data_ptr(), d_part.data_ptr())
                dist.all_reduce(d_part)
                engs.eval_finish_device(d_part.data_ptr(), d_out[i].data_ptr())

        def step_peer():                      # in-library exchange over NVLink peer memory, no host collective
            for i in range(BATCH):
                engs.eval_sharded_device(d_x0[i].data_ptr(), d_out[i].data_ptr())

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(step):
            for _ in range(3):
                step()
            torch.cuda.synchronize(); dist.barrier()
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            torch.cuda.synchronize(); dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        nccl_ms = timed(step_nccl)
        chk_nccl = d_out[0].clone()
        # multi-rank parity: the sharded evaluation (NCCL variant here, peer variant below) against the oracle on rank 0
        ref0 = None
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import csa_oracle as oracle  # noqa: E402  (checker only)
            use_all_host_threads()
            c_ref, g_ref = oracle.cost_grad_factored(xs0[0], T)
            ref0 = np.concatenate([[c_ref], g_ref])

        def parity(vec):
            v = vec.cpu().numpy()
            return max(abs(v[0] - ref0[0]) / abs(ref0[0]), float(np.max(np.abs(v[1:] - ref0[1:])) / np.max(np.abs(ref0[1:]))))
        par_nccl = parity(chk_nccl) if rank == 0 else None
        rowpanel = {"value": BATCH * args.steps / (nccl_ms * 1e-3), "unit": "evals/s", "scaling": "strong",
                    "exchange": "NCCL allreduce of the partial vector", "doubles_exchanged_per_eval": plen,
                    "ms_per_eval": nccl_ms / (BATCH * args.steps), "parity_rel": par_nccl,
                    "parity_what": "max(cost rel.err, gradient max-norm rel.err) of the sharded evaluation at x[0] against "
                                   "oracle.cost_grad_factored on rank 0 (tolerance 1e-10)"}
        # in-library exchange: a failure here (IPC mapping refused, a peer missing) must never cost the main line
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            handles = [None] * world
            dist.all_gather_object(handles, engs.peer_export())
            for r, hd in enumerate(handles):
                if r != rank:
                    engs.peer_attach(r, hd)
        except Exception as exc:  # noqa: BLE001
            ok.zero_()
            peer_err = repr(exc)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            peer_ms = timed(step_peer)
            engs.synchronize()
            agree = float((d_out[0] - chk_nccl).abs().max().item() / chk_nccl.abs().max().item())
            rowpanel = {"value": BATCH * args.steps / (peer_ms * 1e-3), "unit": "evals/s", "scaling": "strong",
                        "exchange": "NVLink peer-memory push + rank-ordered sum inside the library (k_peer_push / k_peer_sum)",
                        "doubles_exchanged_per_eval": plen, "ms_per_eval": peer_ms / (BATCH * args.steps),
                        "nccl_allreduce_variant": {"value": BATCH * args.steps / (nccl_ms * 1e-3),
                                                   "ms_per_eval": nccl_ms / (BATCH * args.steps), "parity_rel": par_nccl},
                        "peer_vs_nccl_max_rel_diff": agree,
                        "parity_rel": parity(d_out[0]) if rank == 0 else None,
                        "parity_what": "max(cost rel.err, gradient max-norm rel.err) of the sharded evaluation at x[0] against "
                                       "oracle.cost_grad_factored on rank 0 (tolerance 1e-10)"}
            # ---- several sharded evaluations in flight per GPU: lanes of the shard handle, each with its own peer buffers, the
            # peers' flags awaited by the stream (mambo_csa_set_peer_wait) so that no spinning kernel holds an SM ------------------
            try:
                nls = max(1, min(nl, BATCH))
                lanes_s = [engs] + [engs.lane() for _ in range(nls - 1)]
                okl = torch.ones(1, dtype=torch.int32, device=dev)
                try:
                    for ln in lanes_s[1:]:
                        hl = [None] * world
                        dist.all_gather_object(hl, ln.peer_export())
                        for r, hd in enumerate(hl):
                            if r != rank:
                                ln.peer_attach(r, hd)
                    for ln in lanes_s:
                        ln.set_peer_wait(True)
                except Exception:  # noqa: BLE001
                    okl.zero_()
                dist.all_reduce(okl, op=dist.ReduceOp.MIN)
                if int(okl.item()) == 1:
                    lstreams = [torch.cuda.Stream(dev) for _ in range(nls)]
                    for ln, st_ in zip(lanes_s, lstreams):
                        ln.set_stream(st_.cuda_stream)

                    def step_lanes():
                        for i in range(BATCH):
                            lanes_s[i % nls].eval_sharded_device(d_x0[i].data_ptr(), d_out[i].data_ptr())

                    for _ in range(3):
                        step_lanes()
                    torch.cuda.synchronize(); dist.barrier()
                    e0.record(stream)
                    for st_ in lstreams:
                        st_.wait_stream(stream)
                    for _ in range(args.steps):
                        step_lanes()
                    for st_ in lstreams:
                        stream.wait_stream(st_)
                    e1.record(stream)
                    torch.cuda.synchronize(); dist.barrier()
                    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    lanes_ms = float(t.item())
                    for ln in lanes_s:
                        ln.synchronize()
                    rowpanel["lanes"] = {"value": BATCH * args.steps / (lanes_ms * 1e-3), "unit": "evals/s", "lanes_per_gpu": nls,
                                         "ms_per_eval": lanes_ms / (BATCH * args.steps),
                                         "parity_rel": parity(d_out[0]) if rank == 0 else None,
                                         "what": "the same sharded evaluations with lanes_per_gpu of them in flight on every GPU (lanes of the "
                                                 "shard handle, stream-ordered wait for the peers' flags): throughput of the row-panel "
                                                 "partition, where `value` above is its latency-bound single-evaluation rate"}
                    for ln in lanes_s:
                        ln.set_stream(None)
                for ln in lanes_s[1:]:
                    ln.close()
                engs.set_stream(stream.cuda_stream)
            except Exception as exc:  # noqa: BLE001  (a side leg must never cost the main line)
                rowpanel["lanes"] = {"error": repr(exc)}
        else:
            rowpanel["peer_exchange"] = "unavailable on this box (CUDA IPC mapping failed on some rank)"
        engs.close()
        lp = (rowpanel.get("lanes") or {}).get("parity_rel")
        if rank == 0 and lp is not None and not (lp <= 1e-10):
            raise SystemExit(f"row-panel sharded evaluation over lanes disagrees with the oracle: rel {lp:.3e} > 1e-10")
        if rank == 0 and rowpanel.get("parity_rel") is not None and not (rowpanel["parity_rel"] <= 1e-10):
            raise SystemExit(f"row-panel sharded evaluation disagrees with the oracle: rel {rowpanel['parity_rel']:.3e} > 1e-10")

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ----------------------------------------
    cpu = cpu_packed = cpu_lit = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        max_evals = args.cpu_evals or (24 if n >= 100 else 200)
        rate, done, cores = cpu_reference_rate(n, T, xs, budget_s=15.0, max_evals=max_evals)
        cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": f"{done} cost+grad evaluations of the same N={n} workload, oracle factored numpy/OpenBLAS on {cores} threads"}
        try:
            prate, pdone, pcores = cpu_packed_rate(n, T, xs, budget_s=8.0, max_evals=4 * max_evals)
            cpu_packed = {"value": prate, "unit": "evals/s", "cores": pcores, "kind": "port",
                          "sample": f"{pdone} evaluations, oracle.cost_grad_packed: the engine's own algorithm (packed rows, lean "
                                    f"single-GEMM formulation, 8x fewer flops than the factored port) in numpy/OpenBLAS on {pcores} threads"}
        except Exception as exc:  # noqa: BLE001
            cpu_packed = {"error": repr(exc)}
        try:
            cpu_lit = cpu_literal_fit(n_target=n)
        except Exception as exc:  # noqa: BLE001
            cpu_lit = {"error": repr(exc)}

    greedy = None
    if rank == 0 and world == 1 and not args.no_greedy:
        greedy = greedy_time_to_tolerance(mb, syn, local_rank, cpu_budget_s=0.0 if args.no_cpu_baseline else 25.0)

    if rank == 0:
        clocks = sampler.summary()
        fused_ms = prof["fused_ms"] / max(prof["fused_launches"], 1)
        rows = info["rows"]
        passes = 2.0 if info["mode"] == mb.SYM_GENERAL else 1.0
        gemms = info.get("hot_gemms", 2)                        # 1: single-GEMM formulation (k_ty); 2: k_fused
        alg_flops = 2.0 * gemms * rows * rows * n               # per hot-kernel launch, of the algorithm actually run
        exec_flops = info["flops_per_eval"] / passes             # incl. tile padding
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "fused_traffic.json")
        if os.path.exists(tpath):
            try:
                import hashlib
                tj = json.load(open(tpath))
                src = os.path.join(ROOT, "quantummambo.jl_b200", "csrc", "fused_kernel.cuh")
                if tj.get("fused_kernel_cuh_sha256") == hashlib.sha256(open(src, "rb").read()).hexdigest():
                    traffic = tj.get(f"N{n}_{info['mode_name']}")     # ncu capture of exactly this kernel source
            except Exception:
                traffic = None
        roof = {
            "bound": "tensor",
            "kernel": "fused::k_ty (Y = T~ A~, FP64 DMMA.8x8x4)" if gemms == 1 else "fused::k_fused (FP64 DMMA.8x8x4)",
            "achieved": alg_flops / (fused_ms * 1e-3) / 1e12, "peak": dgemm_tflops, "unit": "TFLOP/s",
            "frac": alg_flops / (fused_ms * 1e-3) / 1e12 / dgemm_tflops,
            "peak_source": "cuBLAS DGEMM 4096^3 (torch.matmul f64) measured in this run; MEASURED_PEAKS.json has no FP64 entry"
                           + ("" if os.path.exists(peaks_path) else " (file absent)"),
            "traffic": traffic, "ms_per_launch": fused_ms, "launches_timed": prof["fused_launches"],
            "algorithmic_flops_per_launch": alg_flops, "executed_flops_per_launch": exec_flops,
            "executed_frac": exec_flops / (fused_ms * 1e-3) / 1e12 / dgemm_tflops,
            "unsymmetrised_4N5_equivalent_tflops": 4.0 * n ** 5 / (fused_ms * 1e-3) / 1e12,
            "flops_note": "achieved = 2*gemms*rows^2*N flop of the algorithm actually run (packed rows, single-GEMM formulation) over "
                          "the hot kernel's CUDA-event time; SURVEY 8d's 4N^5 figure over the same time is quoted for comparison only",
            "share_of_step": prof["fused_ms"] / ms1,
            "timed_in": "the single-lane timed region (each kernel runs alone there; with several lanes a fused launch queues "
                        "behind other lanes' kernels between its two events)",
            "segments_ms_per_eval": segs,
        }
        line = {
            "metric": "csa_cost_grad_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "N": n,
                       "points_per_step": BATCH, "lanes_per_gpu": nl, "mode": info["mode_name"], "rows": rows,
                       "tensor_bytes_per_gpu": info["tensor_bytes"], "l2_policy": "inputs larger than L2 (residual "
                       f"{info['tensor_bytes'] / 1e6:.0f} MB streamed once per evaluation; L2 is 126 MB)",
                       "parallelism": "multistart replicas, one per GPU" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": BATCH * L * 8,
                    "d2h_bytes_per_step": BATCH * (L + 1) * 8, "api": e2e_api,
                    "timer": "host perf_counter around the C-ABI calls with host buffers; the better of the two public forms below",
                    "closure_threads": {"value": e2e_threads, "caller_threads_per_gpu": host_nl,
                                        "what": "mambo_csa_cost_grad, one synchronous call per point, one caller thread per lane"},
                    "batch_call": {"value": e2e_batch, "what": "one mambo_csa_cost_grad_batch call per step from a single caller "
                                   "thread: the step's points over the lanes, each lane re-armed as soon as its point is collected"},
                    "single_caller": e2e_value1,
                    "wait_mode": {"closure_threads": "blocking-sync event" if blocking else "spin",
                                  "batch_call / single_caller": "blocking-sync event" if blocking1 else "spin"}},
            "single_lane": {"value": value1, "ms_per_step": ms1 / args.steps,
                            "note": "one evaluation in flight at a time (the sequential BFGS case)"},
            "gpu_launches": int(launches), "launches_per_eval": launches / (BATCH * args.steps),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "cost_check": cost_check,
        }
        if cpu_packed is not None:
            line["cpu_baseline_packed"] = cpu_packed
        if cpu_lit is not None:
            line["cpu_literal"] = cpu_lit
        if rowpanel is not None:
            line["rowpanel"] = rowpanel
        if greedy is not None:
            line["greedy"] = greedy
        if guess is not None:
            line["initial_guess"] = guess
        if dfres is not None:
            line["double_factorisation"] = dfres
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
