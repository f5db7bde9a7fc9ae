#!/usr/bin/env python3
"""bench.py -- MC slew trials/sec (AL-iLQR + rollout) on N B200s, one process per GPU.

A "step" is one pass of the hot path (ts_monte_carlo_run: scoping field -> gramian cutoff ->
fine field table -> eigen-axis/Bryson weights -> AL-iLQR -> TVLQR replay -> slew-time rule,
then the NCCL gather of outcome records + statistics) over one ensemble of synthetic trials.

Workloads (config.workload):
  mc_fixed_orbit  BASELINE configs[2]: 4,096 slews per GPU, random initial attitudes (uniform on
                  S^3), fixed LEO orbit of src/monte_carlo.jl:122-127 (RAAN 0, anomaly 90), 1U
                  inertia, tf 2400 s, cutoff 30, alpha 0.1.                         [default]
  mc_sweep        BASELINE configs[3]: per-trial inclination/altitude/RAAN/anomaly/MJD/IGRF date.
  igrf            BASELINE configs[1]: 10^8 random LEO points through K1 (evals/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...] [--trials T]
For N > 1 launch under torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GM = 3.986004418E14 * (1 / 1000) ** 3
J_1U = np.diag([0.00125] * 3)
QF = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0.0, 0.0])   # monte_carlo.jl:114
SEED = 0x5EED


# --------------------------------------------------------------------------- synthetic ensembles
def make_trials(workload, n, rank, seed=SEED):
    """Synthetic inputs of one rank's shard; reproducible from (seed, rank)."""
    rng = np.random.default_rng([seed, rank])
    q0 = rng.normal(size=(n, 4))
    q0 /= np.linalg.norm(q0, axis=1, keepdims=True)               # uniform on S^3
    x0 = np.concatenate([np.zeros((n, 3)), q0, np.zeros((n, 1))], axis=1)
    xf = np.tile(np.concatenate([[0, 0, 0], QF, [1.0]]), (n, 1))
    Jm = np.tile(J_1U.reshape(-1), (n, 1))
    qn = rng.normal(size=(n, 3)) * (math.pi / 180) ** 2           # TortoiseSat.jl:231
    if workload == "mc_fixed_orbit":
        kep = np.array([[0.0, 400.0 + 6371.0, 96.6, 0.0, 0.0, 90.0]])
        fo = [(GM, 58155.0, 2019.0, (400.0 + 6371.0) * 1000.0, 0.0, 0.0, 0)]
        shared, cutoff = True, 30.0
    else:                                                           # mc_sweep
        alt = rng.uniform(350, 800, size=n)
        kep = np.stack([np.zeros(n), alt + 6371.0, rng.uniform(0, 98, size=n), rng.uniform(0, 360, size=n), np.zeros(n),
                        rng.uniform(0, 360, size=n)], axis=1)
        fo = [(GM, rng.uniform(58155, 58520), 2015 + 5 * rng.random(), (a + 6371.0) * 1000.0, 0.0, 0.0, 0) for a in alt]
        shared, cutoff = False, 100.0
    return dict(kep=kep, fo=fo, x0=x0, xf=xf, Jm=Jm, qn=qn, shared=shared, cutoff=cutoff)


def mc_config(host, tr, n):
    cfg = host.default_mc_config(n, shared_orbit=tr["shared"], run_tvlqr=True, tf=2400.0, N_scope=5000, cutoff=tr["cutoff"], dt=0.2,
                                 alpha=0.1, beta=1e3)
    cfg.tvlqr.noise_mode = 2
    cfg.tvlqr.seed = SEED
    return cfg


def workload_name(workload, n, points):
    if workload == "igrf":
        return "igrf (BASELINE configs[1]): %d random LEO points per GPU, date 2019.0" % points
    return "%s (BASELINE configs[%d]): %d slews per GPU, 1U inertia, tf 2400 s, cutoff %g, alpha 0.1, goal_mask 0x7F, " \
           "AL-iLQR 20x50, TVLQR replay with Philox noise" % (workload, 2 if workload == "mc_fixed_orbit" else 3, n,
                                                              30.0 if workload == "mc_fixed_orbit" else 100.0)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_mc_sample(tr, n_sample, nthreads):
    """The reference algorithm (CPU oracle port; Julia is not available) on the first n_sample trials
    of the workload, OpenMP over trials.  Returns (trials/s, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import slew_setup as S
    t0 = time.time()
    slews = []
    base = None
    for t in range(n_sample):
        k = tr["kep"][0 if tr["shared"] else t]
        f = tr["fo"][0 if tr["shared"] else t]
        if tr["shared"] and base is not None:
            s = S.build_slew(k, J_1U, tr["x0"][t, 3:7], QF, mjd=f[1], igrf_date=f[2], field_radius_m=f[3], tf=2400.0, cutoff=tr["cutoff"],
                             alpha=0.1, t_final=base.t_final)
        else:
            s = S.build_slew(k, J_1U, tr["x0"][t, 3:7], QF, mjd=f[1], igrf_date=f[2], field_radius_m=f[3], tf=2400.0, cutoff=tr["cutoff"],
                             alpha=0.1)
            base = s
        slews.append(s)
    Xs, Us, Ks, out = S.oracle_solve(slews, nthreads=nthreads, want_K=False)
    o, g = S.tvlqr_opts_pair(noise_mode=2, seed=SEED)
    for t, s in enumerate(slews):
        S.oracle_tvlqr(s, Xs[t], Us[t], s.x0 * np.array([1] * 7 + [0]), o, trial=t)
    dt = time.time() - t0
    return n_sample / dt, dt


def cpu_igrf_sample(n, nthreads):
    from oracle import oracle as orc
    rng = np.random.default_rng(SEED)
    lat = np.arcsin(2 * rng.random(n) - 1)
    lon = math.pi * (2 * rng.random(n) - 1)
    r = 6371200.0 + 300000.0 + 900000.0 * rng.random(n)
    t0 = time.time()
    orc.igrf12_batch(2019.0, r, lat, lon, nthreads=nthreads)
    dt = time.time() - t0
    return n / dt, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mc_fixed_orbit", choices=["mc_fixed_orbit", "mc_sweep", "igrf"])
    ap.add_argument("--trials", type=int, default=4096, help="trials per GPU (weak scaling)")
    ap.add_argument("--points", type=int, default=100_000_000, help="igrf workload: points per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = host_cores()
    metric = "igrf12_evals_per_sec" if a.workload == "igrf" else "mc_slew_trials_per_sec"
    unit = "evals/s" if a.workload == "igrf" else "trials/s"

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return 0
        from oracle import oracle as orc
        orc.build()
        vals = []
        if a.workload == "igrf":
            n = 2_000_000
            sample = "%d of the 1e8 points per step, oracle igrf12 (C++ restatement of src/igrf.jl), %d OpenMP threads" % (n, cores)
            for i in range(a.warmup + a.steps):
                v, _ = cpu_igrf_sample(n, cores)
                if i >= a.warmup:
                    vals.append(v)
        else:
            tr = make_trials(a.workload, max(cores, 1), 0)
            n = max(cores, 1)
            sample = "first %d trials of the %d-trial ensemble per step (one per host thread), oracle pipeline " \
                     "(C++ restatement of the reference algorithm; Julia unavailable)" % (n, a.trials)
            steps = max(1, min(a.steps, 2))
            for i in range(min(a.warmup, 0) + steps):
                v, _ = cpu_mc_sample(tr, n, cores)
                vals.append(v)
        v = float(np.mean(vals))
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": 1e3 * (n / v), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload_name(a.workload, a.trials, a.points), "trials_per_gpu": a.trials},
                          "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    import tortoisesat.jl_b200 as tb
    from tortoisesat.jl_b200 import host, parallel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints a "NCCL version ..." banner on STDOUT when the first communicator is created; keep stdout
        # clean for the single JSON line by pointing fd 1 at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    eng = tb.Engine(local_rank)
    peak_fp64 = eng.fp64_peak_tflops()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # K1 throughput (BASELINE's second metric), device-resident inputs, after its own warm-up
    def igrf_run(n, reps):
        g = torch.Generator(device=dev).manual_seed(SEED + rank)
        u = torch.rand(3, n, generator=g, device=dev, dtype=torch.float64)
        lat = torch.asin(2 * u[0] - 1)
        lon = math.pi * (2 * u[1] - 1)
        r = 6371200.0 + 300000.0 + 900000.0 * u[2]
        del u
        o = [torch.empty(n, device=dev, dtype=torch.float64) for _ in range(3)]
        ms = []
        for i in range(3 + reps):
            eng.igrf12_batch(2019.0, r, lat, lon, out=o)
            if i >= 3:
                ms.append(eng.last_kernel_ms())
        return float(np.mean(ms)), (r, lat, lon, o)

    if a.workload == "igrf":
        n = a.points
        ms_k, bufs = igrf_run(n, 1)
        r, lat, lon, o = bufs
        # host copies in PINNED memory (contract: inputs come from pinned host memory); numpy views share it
        pinned = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
        for dst, src in zip(pinned, (r, lat, lon)):
            dst.copy_(src)
        rh, lah, loh = (x.numpy() for x in pinned)
        l0 = eng.launch_count()
        clk = ClockSampler(local_rank)
        for _ in range(a.warmup):
            eng.igrf12_batch(2019.0, r, lat, lon, out=o)
        barrier()
        clk.start()
        t0 = time.perf_counter()
        ks = []
        for _ in range(a.steps):
            eng.igrf12_batch(2019.0, r, lat, lon, out=o)
            ks.append(eng.last_kernel_ms())
        barrier()
        el = max_over_ranks(time.perf_counter() - t0)
        launches = eng.launch_count() - l0 - a.warmup
        kms = max_over_ranks(float(np.mean(ks)))
        # e2e: host buffers through the C ABI (H2D + kernel + D2H inside the call)
        n_e = min(n, 20_000_000)
        outp = [torch.empty(n_e, dtype=torch.float64).pin_memory().numpy() for _ in range(3)]
        for _ in range(2):
            eng.igrf12_batch(2019.0, rh[:n_e], lah[:n_e], loh[:n_e], out=outp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.igrf12_batch(2019.0, rh[:n_e], lah[:n_e], loh[:n_e], out=outp)
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0) / a.steps
        clocks = clk.stop()
        value = world * n / (kms * 1e-3)
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": kms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": workload_name("igrf", 0, n),
                                                  "l2": "inputs+outputs 4.8 GB per GPU > L2"},
                "roofline": {"bound": "fp64", "achieved": 2243.0 * n / (kms * 1e-3) / 1e12, "peak": peak_fp64, "unit": "TFLOP/s",
                             "frac": 2243.0 * n / (kms * 1e-3) / 1e12 / peak_fp64, "traffic": None,
                             "peak_source": "measured live: register-resident DFMA micro-benchmark (ts_fp64_peak_probe)",
                             "hbm_gbs_algorithmic": 48.0 * n / (kms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak},
                "e2e": {"value": world * n_e / e2e_s, "unit": unit, "h2d_bytes_per_step": 24 * n_e, "d2h_bytes_per_step": 24 * n_e,
                        "note": "%d points per call through ts_igrf12_batch with pinned host buffers (H2D + kernel + D2H inside the call)" % n_e},
                "gpu_launches": launches, "clocks": clocks, "wall_s_timed": el}
        if rank == 0 and not a.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build()
            v, dt = cpu_igrf_sample(4_000_000, cores)
            line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                    "sample": "4e6 of the points, oracle igrf12 (C++ restatement of src/igrf.jl), OpenMP"}
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- Monte-Carlo workloads
    n = a.trials
    tr = make_trials(a.workload, n, rank)
    cfg = mc_config(host, tr, n)
    fo = np.zeros(len(tr["fo"]), dtype=host.FIELD_OPTS_DTYPE)
    for i, f in enumerate(tr["fo"]):
        fo[i] = f
    sid = (np.arange(n) + rank * n).astype(np.uint32)

    def step():
        t0 = time.perf_counter()
        out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, tr["x0"], tr["xf"], tr["Jm"], q_noise0=tr["qn"], stream_id=sid)
        t_call = time.perf_counter() - t0
        allout = parallel.gather_outcomes(out, device=dev)          # NCCL all-gather of 64-byte records
        vec = parallel.reduce_stats(parallel.stats_vector(st), device=dev)
        return out, st, allout, vec, t_call

    for _ in range(a.warmup):
        step()
    barrier()
    l0 = eng.launch_count()
    clk = ClockSampler(local_rank)
    clk.start()
    t0 = time.perf_counter()
    dev_ms, call_s, solve_ms, flops_solve, stats_last = [], [], [], [], None
    for _ in range(a.steps):
        out, st, allout, vec, t_call = step()
        dev_ms.append(st.ms_field + st.ms_prep + st.ms_solve + st.ms_tvlqr)
        call_s.append(t_call)
        solve_ms.append(st.ms_solve)
        kn = (out["N"] - 1).astype(np.float64)
        flops_solve.append(float(np.sum(kn * (out["inner_iters"] * 6100.0 + out["ls_rollouts"] * 500.0))))
        stats_last = (out, st, allout, vec)
    barrier()
    wall = max_over_ranks(time.perf_counter() - t0)
    clocks = clk.stop()
    launches = eng.launch_count() - l0
    out, st, allout, vec = stats_last
    dev_s = max_over_ranks(float(np.mean(dev_ms)) * 1e-3)       # device time of the kernels (CUDA events), max over ranks
    e2e_s = wall / a.steps                                         # through the C ABI with host buffers + gather
    total_trials = world * n
    solve_s = float(np.mean(solve_ms)) * 1e-3
    ach = float(np.mean(flops_solve)) / solve_s / 1e12
    # HBM traffic of K3 per knot-iteration, from the ncu --set full captures summarised in profiles/
    # (k3_narrow_r1e: 4736 trials, N = 300, 8 warps/SM: dram read 73.4 GB + write 131.2 GB over 1.24e8 knot-iterations
    #  = 1650 B; k3_wide_r1d, one warp per trial with all 21 candidates written: 2050 B)
    K3_DRAM_BYTES_PER_KNOT_ITER = 1650.0
    knot_iters = float(np.sum((out["N"] - 1).astype(np.float64) * out["inner_iters"]))
    h2d = n * (8 + 8 + 9 + 3) * 8 + n * 4 + len(fo) * (6 * 8 + 56)
    d2h = n * 64 + 8 * len(fo)
    conv = int(vec[1])
    line = {"metric": metric, "value": total_trials / dev_s, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": e2e_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a.workload, n, 0),
                       "trials_per_gpu": n, "knots": float(vec[9] / max(1.0, vec[0] - vec[2])),
                       "l2": "per-GPU working set %.1f GB > L2 (126 MB)" % (n * float(np.max(out["N"])) * 980 / 1e9),
                       "converged": conv, "no_cutoff": int(vec[2]), "mean_inner_iters": float(vec[7] / max(1.0, vec[0] - vec[2])),
                       "mean_ls_rollouts": float(vec[8] / max(1.0, vec[0] - vec[2])),
                       "mean_slew_time_s": float(vec[4] / max(1.0, vec[0] - vec[2])), "fail_slew": int(vec[3])},
            "roofline": {"bound": "fp64", "kernel": "K3 AL-iLQR solve: k3_alilqr_kernel (4 trials per warp) + k3_wide_kernel (one warp per "
                                                      "straggler) for a single wave of trials, k3_queue_kernel for more", "achieved": ach, "peak": peak_fp64, "unit": "TFLOP/s",
                         "frac": ach / peak_fp64, "traffic": K3_DRAM_BYTES_PER_KNOT_ITER * knot_iters,
                         "traffic_source": "1650 B per knot-iteration (ncu --set full, dram__bytes_read+write, capture of "
                                           "tools/k3_small.py 4736: 204.6 GB / 1.24e8 knot-iterations) x this run's knot-iterations",
                         "hbm_gbs_from_traffic": K3_DRAM_BYTES_PER_KNOT_ITER * knot_iters / solve_s / 1e9, "hbm_peak_gbs": hbm_peak,
                         "peak_source": "measured live: register-resident DFMA micro-benchmark (ts_fp64_peak_probe); "
                                        "MEASURED_PEAKS.json has no FP64 row",
                         "flop_model": "6100 FLOP per knot-iteration (rk3 Jacobian 2600 + Riccati step 3500) + 500 per "
                                       "line-search rollout knot (SURVEY 8d), counted from per-trial iteration counters",
                         "kernel_share_of_step": solve_s / float(np.mean(dev_ms) * 1e-3)},
            "e2e": {"value": total_trials / e2e_s, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "note": "ts_monte_carlo_run with host per-trial inputs + outcome D2H + NCCL gather; trajectories stay in HBM"},
            "stage_ms": {"field": st.ms_field, "prep": st.ms_prep, "solve": st.ms_solve, "tvlqr": st.ms_tvlqr},
            "k3_split": dict(zip(("persistent_ms", "straggler_ms", "handed_over"), eng.k3_last_split())),
            "gpu_launches": launches, "clocks": clocks}
    # secondary metric: IGRF-12 evals/s (K1), 1e8 points
    try:
        ms_k, _ = igrf_run(100_000_000 if n >= 1024 else 1_000_000, 3)
        npts = 100_000_000 if n >= 1024 else 1_000_000
        line["igrf12"] = {"evals_per_s": world * npts / (ms_k * 1e-3), "points_per_gpu": npts, "ms": ms_k,
                          "fp64_frac": 2243.0 * npts / (ms_k * 1e-3) / 1e12 / peak_fp64}
    except Exception as ex:  # pragma: no cover
        line["igrf12"] = {"error": str(ex)}
    if rank == 0 and not a.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        ns = max(cores, 1)
        v, dt = cpu_mc_sample(tr, min(ns, n), cores)
        line["cpu_baseline"] = {"value": v, "unit": unit, "cores": cores, "kind": "port",
                                "sample": "first %d trials of this rank's ensemble (one per host thread), %.1f s; C++ restatement "
                                          "of the reference algorithm (Julia unavailable)" % (min(ns, n), dt)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
