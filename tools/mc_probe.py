"""GPU probe on the BENCHMARK ensembles (bench.make_trials): where the solves end up, and why.

  python tools/mc_probe.py <mc_fixed_orbit|mc_sweep> <n_trials> [--suspend a,b,...] [--pair 0,1] [--occ 0,4,6] [--early 1.0,2.0] [--flips] [--why]

Per run: K3 time and its split between the two kernels, per-status histogram, inner-iteration quantiles.
--why     keeps the trajectories and decomposes c_max of the non-converged trials: which constraint (control bound
          |u| <= 1 or a goal component) holds c_max above the 1e-3 tolerance, and the per-outer progress.
--flips   re-runs the ensemble under each alternative of the SURVEY App. C assumption registry (A1..A7) and counts the
          trials whose status / outer-iteration count change: the spec risk of the unpinned solver details.
Writes gpurun_out/mc_probe_<workload>_<n>.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench as B
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

workload = sys.argv[1] if len(sys.argv) > 1 else "mc_fixed_orbit"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
args = sys.argv[3:]
suspends = [150]
pairs = [1]
occs = [0]
earlies = [None]
for i, a in enumerate(args):
    if a == "--occ":
        occs = [int(x) for x in args[i + 1].split(",")]
    if a == "--suspend":
        suspends = [int(x) for x in args[i + 1].split(",")]
    if a == "--early":
        earlies = [float(x) for x in args[i + 1].split(",")]
    if a == "--pair":
        pairs = [int(x) for x in args[i + 1].split(",")]
eng = tb.Engine(0)
tr = B.make_trials(workload, n, 0)
fo = np.zeros(len(tr["fo"]), dtype=host.FIELD_OPTS_DTYPE)
for i, f in enumerate(tr["fo"]):
    fo[i] = f
sid = np.arange(n).astype(np.uint32)
STAT = ["converged", "max_outer", "cost_blowup", "reg_max", "nan", "no_cutoff"]
res = {"workload": workload, "n": n, "runs": []}


def run(cfg, label):
    out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, tr["x0"], tr["xf"], tr["Jm"], q_noise0=tr["qn"], stream_id=sid)
    act = out["status"] != 5
    it = out["inner_iters"][act].astype(float)
    hist = np.bincount(out["status"], minlength=6).tolist()
    rec = dict(label=label, ms_solve=st.ms_solve, ms_field=st.ms_field, ms_tvlqr=st.ms_tvlqr, split=list(eng.k3_last_split()),
               status=dict(zip(STAT, hist)), inner_q=np.percentile(it, [50, 75, 90, 95, 99, 100]).tolist(), inner_mean=float(it.mean()),
               ls_mean=float(out["ls_rollouts"][act].mean()), outer_q=np.percentile(out["outer_iters"][act], [50, 90, 100]).tolist(),
               N_q=np.percentile(out["N"][act], [0, 50, 90, 100]).tolist(), fail_slew=int(st.n_fail_slew),
               trials_per_s=n / (1e-3 * (st.ms_field + st.ms_prep + st.ms_solve + st.ms_tvlqr)))
    print(json.dumps(rec), flush=True)
    res["runs"].append(rec)
    return out, st


base_out = None
for s in suspends:
    for pr, oc, ea in [(p_, o_, e_) for p_ in pairs for o_ in (occs if p_ == 0 else [0]) for e_ in earlies]:
        cfg = B.mc_config(host, tr, n)
        cfg.ilqr.k3_suspend_after = s
        cfg.ilqr.k3_pair = pr
        cfg.ilqr.k3_wide_occ = oc
        if ea is not None:
            cfg.ilqr.k3_early_factor = ea
        out, st = run(cfg, "suspend=%d pair=%d occ=%d early=%s" % (s, pr, oc, ea))
        if base_out is None:
            base_out = out.copy()
        else:
            same = all(np.array_equal(out[f], base_out[f]) for f in ("status", "outer_iters", "inner_iters", "ls_rollouts"))
            print("   same iteration paths as the first run:", same, "| max |dJ|/|J| %.2e" % float(np.max(np.abs(out["J"] - base_out["J"]) / np.maximum(np.abs(base_out["J"]), 1e-300))), flush=True)

if "--why" in args:
    cfg = B.mc_config(host, tr, n)
    cfg.keep_trajectories = 1
    out, st = run(cfg, "keep_trajectories")
    t = eng.mc_trajectories(n, want=("X", "U"))
    ko = t["knot_offs"]
    xf = tr["xf"]
    rows = []
    for i in np.nonzero(out["status"] == 1)[0]:
        X = t["X"][ko[i]:ko[i + 1]]
        U = t["U"][ko[i]:ko[i + 1] - 1]
        ub = float(np.max(np.abs(U)) - 1.0)
        ge = np.abs(X[-1, :7] - xf[i, :7])
        rows.append((ub, float(ge[:3].max()), float(ge[3:7].max()), float(out["c_max"][i]), int(out["inner_iters"][i])))
    rows = np.array(rows) if rows else np.zeros((0, 5))
    why = {}
    if len(rows):
        dom = np.argmax(rows[:, :3], axis=1)
        why = dict(n_max_outer=int(len(rows)), binding_control_bound=int((dom == 0).sum()), binding_goal_rate=int((dom == 1).sum()),
                   binding_goal_attitude=int((dom == 2).sum()), c_max_q=np.percentile(rows[:, 3], [0, 10, 50, 90, 100]).tolist(),
                   u_excess_q=np.percentile(rows[:, 0], [0, 50, 100]).tolist(), goal_q_err_q=np.percentile(rows[:, 2], [0, 50, 100]).tolist(),
                   inner_q=np.percentile(rows[:, 4], [0, 50, 100]).tolist(),
                   frac_within_2x_tol=float((rows[:, 3] < 2e-3).mean()), frac_within_10x_tol=float((rows[:, 3] < 1e-2).mean()))
    # slew angle vs outcome
    qf = xf[:, 3:7]
    ang = 2 * np.degrees(np.arccos(np.clip(np.abs(np.sum(tr["x0"][:, 3:7] * qf, axis=1)), 0, 1)))
    why["angle_deg_mean_by_status"] = {STAT[k]: float(ang[out["status"] == k].mean()) for k in range(6) if (out["status"] == k).any()}
    why["corr_angle_inner"] = float(np.corrcoef(ang, out["inner_iters"])[0, 1])
    print("why:", json.dumps(why), flush=True)
    res["why"] = why

if "--flips" in args:
    flips = {}
    for flag in ["stage_cost_dt", "a2_active_ge", "a3_grad_over_N", "a4_no_intermediate", "a5_dual_active_only", "a6_penalty_conditional",
                 "a7_carry_cost"]:
        cfg = B.mc_config(host, tr, n)
        setattr(cfg.ilqr, flag, 1)
        out, st = run(cfg, "flip " + flag)
        flips[flag] = dict(status_changed=int((out["status"] != base_out["status"]).sum()),
                           outer_changed=int((out["outer_iters"] != base_out["outer_iters"]).sum()),
                           inner_changed=int((out["inner_iters"] != base_out["inner_iters"]).sum()),
                           converged=int((out["status"] == 0).sum()), converged_default=int((base_out["status"] == 0).sum()))
        print("flip", flag, flips[flag], flush=True)
    # the eigen-axis conjugate "fix" (ADVICE r1): how much does the literal qmult(q_f, q_0) matter for the ensemble?
    cfg = B.mc_config(host, tr, n)
    cfg.eigen_axis_fix = 1
    out, st = run(cfg, "eigen_axis_fix=1")
    flips["eigen_axis_fix"] = dict(status_changed=int((out["status"] != base_out["status"]).sum()),
                                   outer_changed=int((out["outer_iters"] != base_out["outer_iters"]).sum()),
                                   converged=int((out["status"] == 0).sum()), converged_default=int((base_out["status"] == 0).sum()))
    print("flip eigen_axis_fix", flips["eigen_axis_fix"], flush=True)
    res["flips"] = flips

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "mc_probe_%s_%d.json" % (workload, n)), "w"), indent=1)
