#!/usr/bin/env python3
"""Generates tests/golden/alilqr_oracle_frozen.json: outputs of the CPU oracle's AL-iLQR + TVLQR replay on three small
slews, FROZEN so that a later change to the oracle (the anchor every GPU parity test compares against) cannot
go unnoticed.  These are NOT reference (Julia) outputs -- TrajectoryOptimization.jl v0.1.2 is not available here and
the AL-iLQR parity stays "unpinned" (DESIGN.md section 3); they pin the oracle to itself as of round 1.

Run from the repo root:  python tests/golden/gen_alilqr_frozen.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import slew_setup as S  # noqa: E402

CASES = [
    dict(name="1P_5deg_60s", kep=[0, 6578, 96, 0, 0, 90], J="J_1P", axis=[1, 0, 1], angle=5.0, t_final=60.0, goal_mask=0x7F),
    dict(name="1P_20deg_30s", kep=[0, 6578, 96, 0, 0, 90], J="J_1P", axis=[1, 0, 1], angle=20.0, t_final=30.0, goal_mask=0x7F),
    dict(name="3U_3deg_50s_literal_goal", kep=[0, 6871, 51.6, 30, 0, 10], J="J_3U", axis=[0, 1, 0], angle=3.0, t_final=50.0,
         goal_mask=0xFF),
]


def run(case):
    s = S.build_slew(case["kep"], getattr(S, case["J"]), S.quat_axis_angle(case["axis"], case["angle"]), np.array([1.0, 0, 0, 0]),
                     t_final=case["t_final"])
    o = S.orc.default_ilqr_opts()
    o.goal_mask = case["goal_mask"]
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    r = out[0]
    X, U, K = Xs[0], Us[0], Ks[0]
    return dict(case, N=int(s.N), status=int(r["status"]), outer_iters=int(r["outer_iters"]), inner_iters=int(r["inner_iters"]),
                ls_rollouts=int(r["ls_rollouts"]), J_cost=float(r["J"]), c_max=float(r["c_max"]),
                x_final=[float(v) for v in X[-1]], u_first=[float(v) for v in U[0]], u_absmax=float(np.max(np.abs(U))),
                X_sum=float(np.sum(X)), U_sum=float(np.sum(U)), K_frob=float(np.sqrt(np.sum(K * K))),
                Qd=[float(v) for v in s.Qd], Rd=[float(v) for v in s.Rd])


if __name__ == "__main__":
    data = dict(note="CPU-oracle outputs frozen at round 1 (see the generator's docstring); tolerances in tests/test_oracle_frozen.py",
                cases=[run(c) for c in CASES])
    with open(os.path.join(HERE, "alilqr_oracle_frozen.json"), "w") as f:
        json.dump(data, f, indent=1)
    for c in data["cases"]:
        print(c["name"], c["status"], c["outer_iters"], c["inner_iters"], c["ls_rollouts"], c["J_cost"], c["c_max"])
