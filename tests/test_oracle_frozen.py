"""CPU: the oracle's AL-iLQR against its own frozen round-1 outputs (tests/golden/alilqr_oracle_frozen.json).
The oracle is the anchor of every GPU parity test; this guards the anchor itself against silent drift.
(Not a reference pin: see the generator's docstring and DESIGN.md section 3.)"""
import json
import os

import numpy as np
import pytest

import slew_setup as S

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "alilqr_oracle_frozen.json")))["cases"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_alilqr_matches_frozen_outputs(case):
    s = S.build_slew(case["kep"], getattr(S, case["J"]), S.quat_axis_angle(case["axis"], case["angle"]), np.array([1.0, 0, 0, 0]),
                     t_final=case["t_final"])
    assert s.N == case["N"]
    assert np.allclose(s.Qd, case["Qd"], rtol=1e-12, atol=0) and np.allclose(s.Rd, case["Rd"], rtol=1e-12, atol=0)
    o = S.orc.default_ilqr_opts()
    o.goal_mask = case["goal_mask"]
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    r, X, U, K = out[0], Xs[0], Us[0], Ks[0]
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts"):
        assert int(r[f]) == case[f], f
    assert abs(r["J"] - case["J_cost"]) <= 1e-9 * abs(case["J_cost"])
    assert abs(r["c_max"] - case["c_max"]) <= 1e-9 * max(1.0, case["c_max"])
    assert np.allclose(X[-1], case["x_final"], rtol=1e-9, atol=1e-12)
    assert np.allclose(U[0], case["u_first"], rtol=1e-8, atol=1e-12)
    assert abs(np.max(np.abs(U)) - case["u_absmax"]) <= 1e-9 * case["u_absmax"]
    assert abs(np.sqrt(np.sum(K * K)) - case["K_frob"]) <= 1e-8 * case["K_frob"]
