"""ncu-sized K3 run on the BENCHMARK ensemble (BASELINE configs[2]: fixed orbit, N = 2044 knots, random attitudes).

  python tools/k3_profile.py <n_trials> [max_outer] [suspend_after] [k3_pair] [k3_wide_occ]
  python tools/k3_profile.py <n_trials> <max_outer> both      (default hand-over, then suspend_after = 3: both kernels in their own regime)

Runs the first n_trials of bench.py's rank-0 ensemble through ts_monte_carlo_run (field -> weights -> K3, no
replay) with the outer-iteration cap lowered so that an `ncu --set full` replay stays short; suspend_after = 3
pushes every trial into k3_wide_kernel (one warp per trial).  Prints the knot-iterations of the run so that
dram__bytes can be quoted per knot-iteration.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench as B
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

eng = tb.Engine(0)
tr = B.make_trials("mc_fixed_orbit", 4096, 0)


def run(n, max_outer, suspend, pair, occ):
    sub = dict(tr)
    for k in ("x0", "xf", "Jm", "qn"):
        sub[k] = tr[k][:n]
    cfg = B.mc_config(host, sub, n)
    cfg.run_tvlqr = 0
    cfg.ilqr.max_outer = max_outer
    if suspend >= 0:
        cfg.ilqr.k3_suspend_after = suspend
    if pair >= 0:
        cfg.ilqr.k3_pair = pair
    cfg.ilqr.k3_wide_occ = occ
    fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
    fo[0] = tr["fo"][0]
    out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, sub["x0"], sub["xf"], sub["Jm"], q_noise0=sub["qn"],
                                  stream_id=np.arange(n).astype(np.uint32))
    ki = (out["N"] - 1).astype(np.float64) * out["inner_iters"]
    ro = float(np.sum((out["N"] - 1).astype(np.float64) * out["ls_rollouts"]))
    print("trials", n, "N", int(out["N"][0]), "max_outer", max_outer, "suspend", suspend, "solve ms", st.ms_solve, "split", eng.k3_last_split())
    print("knot_iterations %.6e rollout_knots %.6e inner mean/max %.1f %d status %s" % (
        float(ki.sum()), ro, out["inner_iters"].mean(), out["inner_iters"].max(), np.bincount(out["status"], minlength=6).tolist()))
    print("cycles per knot-iteration of the slowest trial at 1.965 GHz: %.0f" % (st.ms_solve * 1e-3 * 1.965e9 / (out["inner_iters"].max() * (out["N"][0] - 1))))
    cyc = eng.k3_last_cycles(n)   # SM cycles per trial: backward pass (incl. linearisation), forward passes, linearisation share
    print("SM cycles per knot-iteration, mean over trials: backward %.0f (of which linearisation %.0f) forward %.0f ; pair=%d occ=%d" % (
        np.mean(cyc[:, 0] / ki), np.mean(cyc[:, 2] / ki), np.mean(cyc[:, 1] / ki), cfg.ilqr.k3_pair, cfg.ilqr.k3_wide_occ), flush=True)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
max_outer = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if len(sys.argv) > 3 and sys.argv[3] == "both":
    run(n, max_outer, -1, -1, 0)
    run(n, max_outer, 3, -1, 0)
else:
    run(n, max_outer, int(sys.argv[3]) if len(sys.argv) > 3 else -1, int(sys.argv[4]) if len(sys.argv) > 4 else -1,
        int(sys.argv[5]) if len(sys.argv) > 5 else 0)
