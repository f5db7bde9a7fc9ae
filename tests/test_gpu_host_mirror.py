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


def test_comparison_controller_on_gpu(engine, orc):
    """K7: batched Psiaki PD closed loop through ts_psiaki_pd_batch, ragged batch, against the oracle
    (rollouts to 1e-10), plus the golden rows of the numpy transliteration and the reference-named wrappers."""
    import json
    import os
    from tortoisesat.jl_b200 import host
    fx = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fixtures.json")))["psiaki"]
    cases = fx["cases"]
    L = orc.lib()
    N_i = [c["N"] for c in cases] + [cases[0]["N"] - 37]
    cs = cases + [dict(cases[0], N=cases[0]["N"] - 37)]
    wg = np.concatenate([np.array(c["w_guess"])[:c["N"]] for c in cs])
    qg = np.concatenate([np.array(c["q_guess"])[:c["N"]] for c in cs])
    B = np.concatenate([np.array(c["B_eci"])[:c["N"]] for c in cs])
    X, M, Qe, offs = engine.psiaki_pd_batch(N_i, np.stack([c["x0"] for c in cs]), wg, qg, B, np.stack([np.array(c["J"]).reshape(-1) for c in cs]),
                                            0.2, cases[0]["C_1"], cases[0]["C_2"])
    for t, c in enumerate(cs):
        if t == 1:
            continue                                                     # different gains: checked through the wrapper below
        N = c["N"]
        Xo, Mo, Qo = np.zeros((N, 7)), np.zeros((N, 3)), np.zeros((N, 4))
        a = [orc.f64(np.array(c[k])[:N] if k != "x0" else c[k]) for k in ("x0", "w_guess", "q_guess", "B_eci")]
        J = orc.f64(c["J"])
        L.orc_psiaki_pd_simulation(N, orc.P(a[0]), orc.P(a[1]), orc.P(a[2]), orc.P(a[3]), orc.P(J), 0.2, c["C_1"], c["C_2"], orc.P(Xo),
                                   orc.P(Mo), orc.P(Qo))
        assert np.max(np.abs(X[offs[t]:offs[t + 1]] - Xo)) < 1e-10
        assert np.max(np.abs(M[offs[t]:offs[t + 1]] - Mo)) <= 1e-9 * np.max(np.abs(Mo)) + 1e-18
        assert np.max(np.abs(Qe[offs[t]:offs[t + 1]] - Qo)) < 1e-10
    c = cases[1]
    x, m, qb = host.psiaki_pd_simulation(c["x0"], np.array(c["w_guess"]).T, np.array(c["q_guess"]).T, np.array(c["B_eci"]).T, np.array(c["J"]),
                                         c["dt"], c["C_1"], c["C_2"])
    for k, row in c["X_rows"].items():
        assert np.max(np.abs(x[:, int(k)] - np.array(row))) < 1e-10
    lin = fx["linear"]
    dx = host.attitude_dynamics_linear(lin["x"], lin["u"], lin["x_linear"], lin["B_B"], np.array(lin["J"]))
    assert np.max(np.abs(dx - np.array(lin["dx"]))) <= 1e-12 * np.max(np.abs(lin["dx"]))
