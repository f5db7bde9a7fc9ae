"""GPU: the reference-named single-call functions of the Python host (batch-of-1 routes through the same C ABI)."""
import math

import numpy as np
import pytest

import slew_setup as S
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_solve_slew_and_attitude_simulation_call_shapes(engine):
    """solve_slew = the TrajOpt block of TortoiseSat.jl:145-199 for one slew; attitude_simulation =
    attitude_controller.jl:1-48: reference shapes (state/knot columns), same numbers as the batch entry points."""
    from tortoisesat.jl_b200 import host
    s = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], 10.0), np.array([1.0, 0, 0, 0]), t_final=30.0)
    Xs, Us, Ks, ref = S.oracle_solve([s])
    Q, R, Qf = np.diag(s.Qd), np.diag(s.Rd), np.diag(s.Qfd)
    X, U, K, out = host.solve_slew(s.x0, s.xf, s.J, Q, R, Qf, s.B, s.N, s.dt, N_field=s.index_scale, tf_scope=1.0 / s.clock_rate)
    assert X.shape == (8, s.N) and U.shape == (3, s.N - 1) and K.shape == (3, 8, s.N - 1)
    assert out["status"] == ref[0]["status"] and out["outer_iters"] == ref[0]["outer_iters"]
    assert abs(out["J"] - ref[0]["J"]) <= 1e-6 * abs(ref[0]["J"])
    assert np.max(np.abs(X.T - Xs[0])) < 1e-8
    Qb, Rb, Qfb = host.bryson_weights(s.x0, s.xf, s.J, s.t_final, alpha=10.0, beta=1e3)
    assert np.allclose(np.diag(Qb), s.Qd, rtol=1e-12) and np.allclose(np.diag(Rb), s.Rd, rtol=1e-12) and np.allclose(np.diag(Qfb), s.Qfd, rtol=1e-12)
    # closed-loop replay without noise from the optimised start: tracks the plan
    Ql, Rl, Qfl = np.diag([10.0] * 6), np.diag([7.5e3] * 3), np.diag([1000.0] * 6)
    x0l = s.x0.copy()
    x0l[7] = 0.0
    Xsim, Usim, dX, Kl = host.attitude_simulation(None, None, "rk4", X, U, s.dt, x0l, 0.0, s.t_final, Ql, Rl, Qfl, B_ECI=s.B, J=s.J,
                                                  N_field=s.index_scale, tf_scope=1.0 / s.clock_rate, noise_mode=0)
    assert Xsim.shape[0] == 8 and Usim.shape[0] == 3 and dX.shape[0] == 6 and Kl.shape == (3, 6, s.N - 1)
    assert np.max(np.abs(Xsim[3:7, -1] - X[3:7, Xsim.shape[1] - 1])) < 1e-3
    # the element-wise dynamics under their reference names
    dx = host.DerivFunction(X[:, 3], U[:, 3], s.B, s.J, s.index_scale, 1.0 / s.clock_rate)
    dg = host.gain_simulator(X[:, 3], U[:, 3], s.B, s.J, s.index_scale, 1.0 / s.clock_rate)
    assert dx.shape == (8,) and np.allclose(dx, dg, rtol=1e-12, atol=1e-18) and abs(dx[7] - s.clock_rate) < 1e-15
    assert np.allclose(host.q_inv([1, 2, 3, 4]), [1, -2, -3, -4]) and np.allclose(host.hat([1, 2, 3]) @ [4, 5, 6], np.cross([1, 2, 3], [4, 5, 6]))


def test_monte_carlo_script_call(engine):
    """monte_carlo(number_sims=...) = the loop of monte_carlo.jl:118-262 in one call: the arrays the script leaves in
    globals, reproducible from the seed."""
    from tortoisesat.jl_b200 import host
    ilqr = host.default_ilqr_opts()
    ilqr.max_outer = 6
    r1 = host.monte_carlo(number_sims=6, seed=11, ilqr=ilqr)
    r2 = host.monte_carlo(number_sims=6, seed=11, ilqr=ilqr)
    assert r1["A"].shape == (6, 6) and np.all(r1["A"][:, 1] == 6771.0) and np.all((r1["A"][:, 3] >= 0) & (r1["A"][:, 3] < 360))
    assert r1["t_final"].shape == (6,) and np.all(r1["t_final"] > 0) and np.all(r1["slew_time"] <= r1["t_final"] + 1e-12)
    assert np.array_equal(r1["fails"], (r1["slew_time"] == r1["t_final"]).astype(float))
    assert r1["stats"].n_trials == 6
    for k in ("t_final", "slew_time", "fails"):
        assert np.array_equal(r1[k], r2[k])
    assert np.array_equal(r1["outcomes"]["J"], r2["outcomes"]["J"])
