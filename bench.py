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
    # each step: a bounded sample of the BATCH-point workload (2 of the 8 points), so K steps stay within minutes
    sample = 2 if n >= 100 else BATCH
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
        "config": {"workload": f"dense-sym N={n} seed 0, CSA cost+grad at {BATCH} points/step", "N": n,
                   "points_per_step": BATCH},
        "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the {BATCH} points per step x {args.steps} steps, oracle factored numpy/OpenBLAS, "
                                   f"{cores} threads (the literal reference needs 8*N^6 B = 8 TB scratch at N=100: not runnable)"},
        "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def greedy_time_to_tolerance(mb, syn, device: int, n: int = 50, R: int = 2, frags: int = 2, max_iter: int = 300):
    """Second half of BASELINE.json's metric: greedy CSA on a planted tensor (sum of R random CSA fragments) with the
    library's device-resident BFGS, bounded to `frags` fragments x `max_iter` iterations so the default bench stays
    short.  Reports wall time, evaluations and the squared-norm cost trail (tolerance is on the squared norm, like
    decomp_tol at decompose.jl:45)."""
    rng = np.random.default_rng(2000)
    cL, uL = n * (n + 1) // 2, n * (n - 1) // 2
    T = np.zeros((n, n, n, n), order="F")
    with mb.CSAEngine(np.zeros((n, n, n, n)), device=device) as e:
        for _ in range(R):
            W, _ = e.reconstruct(np.concatenate([rng.standard_normal(cL), 2 * np.pi * rng.random(uL)]))
            T += W

    def run(lbfgs_memory: int):
        trail, evals = [], 0
        rng_x = np.random.default_rng(2001)
        with mb.CSAEngine(T, device=device) as e:
            c_init = e.residual_cost()
            t0 = time.perf_counter()
            for _ in range(frags):
                x0 = np.concatenate([np.zeros(cL), 2 * np.pi * rng_x.random(uL)])      # decompose.jl:86-94
                xm, res = e.optimize(x0, g_tol=1e-8, max_iter=max_iter, lbfgs_memory=lbfgs_memory)
                trail.append(e.subtract_fragment(xm))
                evals += res["evaluations"]
            dt = time.perf_counter() - t0
        return c_init, dt, evals, trail

    c0, dt, evals, trail = run(0)
    out = {"workload": f"planted N={n} R={R}, {frags} greedy fragments x <= {max_iter} BFGS iterations (native optimiser, dense "
                       "inverse Hessian like Optim.jl's BFGS)",
           "seconds": dt, "evaluations": evals, "initial_cost": c0, "cost_trail": trail,
           "ms_per_evaluation_incl_optimizer": dt / max(evals, 1) * 1e3}
    _, dt_l, evals_l, trail_l = run(10)
    out["lbfgs_memory_10"] = {"seconds": dt_l, "evaluations": evals_l, "cost_trail": trail_l,
                              "ms_per_evaluation_incl_optimizer": dt_l / max(evals_l, 1) * 1e3}
    # BASELINE.json configs[2]: 8-way multistart of the first greedy step (random theta guesses), here on ONE GPU:
    # the 8 runs one after the other on one handle versus spread over evaluation lanes that share the residual
    K = 8
    x0s = np.stack([np.concatenate([np.zeros(cL), 2 * np.pi * rng.random(uL)]) for _ in range(K)])
    with mb.CSAEngine(T, device=device) as e:
        lanes = [e] + [e.lane() for _ in range(K - 1)]
        t0 = time.perf_counter()
        _, r1, b1 = mb.multistart(lanes[:1], x0s, max_iter=max_iter)
        t_seq = time.perf_counter() - t0
        t0 = time.perf_counter()
        _, r8, b8 = mb.multistart(lanes, x0s, max_iter=max_iter)
        t_lanes = time.perf_counter() - t0
        for ln in lanes[1:]:
            ln.close()
    out["multistart"] = {"workload": f"{K} random-theta starts of the first greedy step, planted N={n} R={R}, <= {max_iter} BFGS iterations each",
                         "seconds_one_handle": t_seq, "seconds_8_lanes": t_lanes,
                         "evaluations": sum(r["evaluations"] for r in r8), "best_start": b8,
                         "minima_agree": bool(max(abs(a["minimum"] - b["minimum"]) for a, b in zip(r1, r8))
                                              <= 1e-9 * max(abs(r["minimum"]) for r in r1)),
                         "best_minimum": r8[b8]["minimum"]}
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

    # caller threads per rank: one per lane, but no more than two per host core of this rank's share (the synchronous
    # closure spins in cudaStreamSynchronize; 8 ranks x 8 threads on a 16-core host measured 35 % slower than 8 x 4)
    host_nl = max(1, min(nl, (2 * (os.cpu_count() or 1)) // max(world, 1)))
    while BATCH % host_nl:
        host_nl -= 1
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

    t_e2e1 = host_leg(1)
    t_e2e = host_leg(host_nl) if host_nl > 1 else t_e2e1
    t_batch = batch_leg()
    sampler.stop()
    e2e_value = evals_total / t_e2e
    e2e_value1 = evals_total / t_e2e1
    e2e_batch = evals_total / t_batch
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
        rowpanel = {"value": BATCH * args.steps / (nccl_ms * 1e-3), "unit": "evals/s", "scaling": "strong",
                    "exchange": "NCCL allreduce of the partial vector", "doubles_exchanged_per_eval": plen,
                    "ms_per_eval": nccl_ms / (BATCH * args.steps)}
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
                                                   "ms_per_eval": nccl_ms / (BATCH * args.steps)},
                        "peer_vs_nccl_max_rel_diff": agree}
        else:
            rowpanel["peer_exchange"] = "unavailable on this box (CUDA IPC mapping failed on some rank)"
        engs.close()

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        max_evals = args.cpu_evals or (24 if n >= 100 else 200)
        rate, done, cores = cpu_reference_rate(n, T, xs, budget_s=15.0, max_evals=max_evals)
        cpu = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
               "sample": f"{done} cost+grad evaluations of the same N={n} workload, oracle factored numpy/OpenBLAS on {cores} threads"}

    greedy = None
    if rank == 0 and world == 1 and not args.no_greedy:
        greedy = greedy_time_to_tolerance(mb, syn, local_rank)

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
                traffic = json.load(open(tpath)).get(f"N{n}_{info['mode_name']}")
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
            "config": {"workload": f"dense-sym N={n} seed 0, CSA cost+grad at {BATCH} points/step per GPU", "N": n,
                       "points_per_step": BATCH, "lanes_per_gpu": nl, "mode": info["mode_name"], "rows": rows,
                       "tensor_bytes_per_gpu": info["tensor_bytes"], "l2_policy": "inputs larger than L2 (residual "
                       f"{info['tensor_bytes'] / 1e6:.0f} MB streamed once per evaluation; L2 is 126 MB)",
                       "parallelism": "multistart replicas, one per GPU" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": BATCH * L * 8,
                    "d2h_bytes_per_step": BATCH * (L + 1) * 8,
                    "timer": "host perf_counter around synchronous C-ABI calls (mambo_csa_cost_grad), one caller thread per lane",
                    "caller_threads_per_gpu": host_nl, "single_caller": e2e_value1,
                    "batch_call": {"value": e2e_batch, "what": "one mambo_csa_cost_grad_batch call per step from a single caller "
                                   "thread (the step's points in one wave over the lanes; waves do not overlap)"}},
            "single_lane": {"value": value1, "ms_per_step": ms1 / args.steps,
                            "note": "one evaluation in flight at a time (the sequential BFGS case)"},
            "gpu_launches": int(launches), "launches_per_eval": launches / (BATCH * args.steps),
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "cost_check": cost_check,
        }
        if rowpanel is not None:
            line["rowpanel"] = rowpanel
        if greedy is not None:
            line["greedy"] = greedy
        if guess is not None:
            line["initial_guess"] = guess
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
