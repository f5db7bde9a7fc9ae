"""CPU: the K3 kernel SOURCE (ilqr_math.cuh / ilqr_solver.cuh) compiled for the host with
emulated 8-lane teams (tests/hostsim) vs the independent CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import slew_setup as S
from oracle import oracle as orc


@pytest.mark.parametrize("J", [S.J_1P, S.J_3U, np.array([[2e-3, 1e-4, 0], [1e-4, 3e-3, 2e-4], [0, 2e-4, 1e-3]])])
def test_analytic_rk3_jacobian_vs_forward_mode_duals(J):
    hs = S.hostsim()
    rng = np.random.default_rng(0)
    Bt = rng.normal(size=(400, 3)) * 3e-5
    dt = 0.2
    for trial in range(20):
        x = np.concatenate([rng.normal(size=3) * 0.01, rng.normal(size=4) * 0.8, [rng.random() * 0.9]])
        u = rng.normal(size=3)
        d = orc.make_dyn(Bt, 400.0, 1.0 / 5, J)
        A, B = np.zeros((8, 8)), np.zeros((8, 3))
        orc.lib().orc_rk3_jacobian(C.byref(d), orc.P(x), orc.P(u), dt, orc.P(A), orc.P(B))
        c = dt / 5
        rows = [min(int(np.floor(t * 400 + 1)), 400) - 1 for t in (x[7], x[7] + c / 2, x[7] - c + 2 * c)]
        xn, AB = np.zeros(7), np.zeros((7, 10))
        r = [np.ascontiguousarray(Bt[i]) for i in rows]
        hs.hs_rk3_jac7(orc.P(np.ascontiguousarray(J)), orc.P(np.ascontiguousarray(x[:7])), orc.P(u), orc.P(r[0]), orc.P(r[1]),
                       orc.P(r[2]), dt, orc.P(xn), orc.P(AB))
        xo = np.zeros(8)
        orc.lib().orc_rk3_step(C.byref(d), orc.P(x), orc.P(u), dt, orc.P(xo))
        assert np.max(np.abs(xn - xo[:7])) < 1e-15
        assert np.max(np.abs(AB[:, :7] - A[:7, :7])) < 1e-14
        assert np.max(np.abs(AB[:, 7:] - B[:7])) < 1e-12 * max(1.0, np.max(np.abs(B)))
        assert np.all(A[:7, 7] == 0) and np.all(A[7, :7] == 0) and A[7, 7] == 1.0   # clock state decoupled (Q2)
        # the register-resident JVP form the K3 kernel actually uses (column-major output)
        cm = np.zeros((10, 7))
        hs.hs_rk3_jac7_jvp(orc.P(np.ascontiguousarray(J)), orc.P(np.ascontiguousarray(x[:7])), orc.P(u), orc.P(r[0]), orc.P(r[1]),
                           orc.P(r[2]), dt, orc.P(cm))
        assert np.max(np.abs(cm.T[:, :7] - A[:7, :7])) < 1e-14
        assert np.max(np.abs(cm.T[:, 7:] - B[:7])) < 1e-12 * max(1.0, np.max(np.abs(B)))
        if np.count_nonzero(J - np.diag(np.diagonal(J))) == 0:
            # the diagonal-inertia instantiation (K3's *_diag_kernel): the same values with the zero products left out
            cd, xd = np.zeros((10, 7)), np.zeros(7)
            hs.hs_rk3_jac7_jvp_diag(orc.P(np.ascontiguousarray(J)), orc.P(np.ascontiguousarray(x[:7])), orc.P(u), orc.P(r[0]),
                                    orc.P(r[1]), orc.P(r[2]), dt, orc.P(cd), orc.P(xd))
            assert np.max(np.abs(cd - cm)) < 1e-15 and np.max(np.abs(xd - xn)) < 1e-16


CASES = [  # (slew angle deg, horizon s, goal mask, expected status)
    (5.0, 60.0, 0x7F, 0),    # converges
    (20.0, 30.0, 0x7F, 1),   # infeasible in 30 s -> all 20 outer iterations
    (8.0, 40.0, 0xFF, 2),    # literal goal on the clock state (Q2) -> cost exceeds max_cost_value
]


@pytest.mark.parametrize("angle,tfin,mask,status", CASES)
def test_team_solver_matches_oracle(angle, tfin, mask, status):
    s = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], angle), np.array([1.0, 0, 0, 0]), t_final=tfin)
    o = orc.default_ilqr_opts()
    o.goal_mask = mask
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    X, U, K, oc = S.hostsim_solve(s, o)
    ref = out[0]
    assert ref["status"] == status
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "N"):
        assert oc[f] == ref[f], f
    assert abs(oc["J"] - ref["J"]) <= 1e-6 * abs(ref["J"])
    assert abs(oc["c_max"] - ref["c_max"]) <= 1e-6 * max(1.0, ref["c_max"])
    assert np.max(np.abs(X - Xs[0])) < 1e-9
    assert np.max(np.abs(U - Us[0])) < 1e-9


@pytest.mark.parametrize("angle,tfin,mask", [(5.0, 60.0, 0x7F), (20.0, 30.0, 0x7F), (8.0, 40.0, 0xFF)])
def test_wide_team_matches_oracle(angle, tfin, mask):
    """The whole-warp (32-lane) team used for a warp's straggler: 32 knots linearised per chunk, all 21 line-search
    candidates in one batch.  Must take the same iteration path as the sequential oracle."""
    s = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], angle), np.array([1.0, 0, 0, 0]), t_final=tfin)
    o = orc.default_ilqr_opts()
    o.goal_mask = mask
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    X, U, K, oc = S.hostsim_solve(s, o, width=32)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "N"):
        assert oc[f] == out[0][f], f
    assert abs(oc["J"] - out[0]["J"]) <= 1e-6 * abs(out[0]["J"])
    assert np.max(np.abs(X - Xs[0])) < 1e-9 and np.max(np.abs(U - Us[0])) < 1e-9


def test_team_width_does_not_change_results():
    """A trial may start in an 8-lane team and be finished by a whole warp (straggler hand-over): the solver must be
    bit-identical at both widths -- same trajectories, gains, cost and counters."""
    s = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], 20.0), np.array([1.0, 0, 0, 0]), t_final=30.0)
    o = orc.default_ilqr_opts()
    X8, U8, K8, o8 = S.hostsim_solve(s, o, width=8)
    X32, U32, K32, o32 = S.hostsim_solve(s, o, width=32)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "N", "J", "c_max"):
        assert o8[f] == o32[f], f
    assert np.array_equal(X8, X32) and np.array_equal(U8, U32) and np.array_equal(K8, K32)


def test_wide_team_stage_cost_dt_3u_inertia_and_gains():
    """The 30-lane Riccati step of the whole-warp team with a non-spherical inertia, dt-scaled stage cost and a knot
    count that is not a multiple of 32: gains, trajectories and counters against the oracle AND bit-identical to
    the 8-lane team."""
    s = S.build_slew([0, 6871, 51.6, 30, 0, 10], S.J_3U, S.quat_axis_angle([0, 1, 0], 3.0), np.array([1.0, 0, 0, 0]), t_final=50.0)
    o = orc.default_ilqr_opts()
    o.stage_cost_dt = 1
    o.max_outer = 6
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    X8, U8, K8, o8 = S.hostsim_solve(s, o, width=8)
    X32, U32, K32, o32 = S.hostsim_solve(s, o, width=32)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "J", "c_max"):
        assert o8[f] == o32[f], f
    assert np.array_equal(X8, X32) and np.array_equal(U8, U32) and np.array_equal(K8, K32)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts"):
        assert o32[f] == out[0][f], f
    assert abs(o32["J"] - out[0]["J"]) <= 1e-6 * abs(out[0]["J"])
    assert np.max(np.abs(K32 - Ks[0])) <= 1e-8 * np.max(np.abs(Ks[0]))


def test_stage_cost_dt_and_3u_inertia():
    s = S.build_slew([0, 6871, 51.6, 30, 0, 10], S.J_3U, S.quat_axis_angle([0, 1, 0], 3.0), np.array([1.0, 0, 0, 0]), t_final=50.0)
    o = orc.default_ilqr_opts()
    o.stage_cost_dt = 1
    o.max_outer = 6
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    X, U, K, oc = S.hostsim_solve(s, o)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts"):
        assert oc[f] == out[0][f], f
    assert abs(oc["J"] - out[0]["J"]) <= 1e-6 * abs(out[0]["J"])
    assert np.max(np.abs(K - Ks[0])) <= 1e-8 * np.max(np.abs(Ks[0]))


@pytest.mark.parametrize("angle,tfin,Jm", [(5.0, 60.0, "1P"), (20.0, 30.0, "1P"), (3.0, 50.0, "3U")])
def test_quaternion_aware_team_matches_oracle(angle, tfin, Jm):
    """SURVEY 8(f2), ts_ilqr_opts.quat_error: the kernel source's QUAT team (7-state, analytic Jacobians projected in
    place, diagonal c|q|^2 cost block, 30-lane Riccati step) against the oracle's literal dense version of the same
    variant: same iteration path, gains in error coordinates (last two columns zero)."""
    s = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P if Jm == "1P" else S.J_3U, S.quat_axis_angle([1, 0, 1], angle),
                     np.array([1.0, 0, 0, 0]), t_final=tfin)
    o = orc.default_ilqr_opts()
    o.quat_error = 1
    Xs, Us, Ks, out = S.oracle_solve([s], o)
    o0 = orc.default_ilqr_opts()
    _, _, _, out0 = S.oracle_solve([s], o0)
    X, U, K, oc = S.hostsim_solve(s, o, width=32)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "N"):
        assert oc[f] == out[0][f], f
    assert abs(oc["J"] - out[0]["J"]) <= 1e-6 * abs(out[0]["J"])
    assert np.max(np.abs(X - Xs[0])) < 1e-9 and np.max(np.abs(U - Us[0])) < 1e-9
    assert np.max(np.abs(K - Ks[0])) <= 1e-8 * np.max(np.abs(Ks[0]))
    assert np.all(K[:, :, 6:] == 0.0)
    # the 8-lane team (four trials per warp on the GPU) runs the same variant bit for bit: a trial may be handed over
    X8, U8, K8, o8 = S.hostsim_solve(s, o, width=8)
    for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "J", "c_max"):
        assert o8[f] == oc[f], f
    assert np.array_equal(X8, X) and np.array_equal(U8, U) and np.array_equal(K8, K)
    # and it IS a different algorithm: the iteration path differs from the default solver's
    assert (out[0]["inner_iters"], out[0]["ls_rollouts"]) != (out0[0]["inner_iters"], out0[0]["ls_rollouts"])
