"""CPU: pins the C++ oracle against tests/golden/ref_fixtures.json -- values produced by the line-by-line numpy
transliteration of the reference's own sources (tests/golden/gen_ref_fixtures.py, which cites file:line for every
function) and, for AL-iLQR, by an independent second implementation of the SURVEY App. C specification.

Tolerances: 1e-12 relative for closed-form arithmetic (different libm / summation order only), 1e-10 for quantities
behind ~1e4 sequential Euler steps or a 3x3 inverse, 1e-8 on AL-iLQR costs after hundreds of iterations (iteration
COUNTS must be identical)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_KEEP = []


def _p(a):
    """address of a float64 copy of `a` that stays alive (ctypes does not hold references to temporaries)"""
    b = np.ascontiguousarray(a, dtype=np.float64)
    _KEEP.append(b)
    if len(_KEEP) > 4096:
        del _KEEP[:2048]
    return b.ctypes.data


@pytest.fixture(scope="module")
def fx():
    return json.load(open(os.path.join(HERE, "golden", "ref_fixtures.json")))


def close(a, b, rel, abs_=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.max(np.abs(b)), 1e-300)
    assert a.shape == b.shape
    err = np.max(np.abs(a - b))
    assert err <= rel * scale + abs_, (err, scale)


def test_kep_eci_and_orbit_rhs(orc, fx):
    for c in fx["kep_ECI"]:
        rv, kep = orc.kep_eci(c["kep"], c["t0"], c["GM"])
        close(rv[0], c["rv"][0], 1e-13, 1e-9)       # exact zeros of cosd(90) become |x| < 1e-9 km at worst
        close(rv[1], c["rv"][1], 1e-13, 1e-12)
        assert abs(kep[5] - c["kep6_after"]) < 1e-12
    L = orc.lib()
    for c in fx["OrbitPlotter"]:
        dx = np.zeros(6)
        L.orc_orbit_rhs(_p(c["x"]), orc.P(dx))
        close(dx, c["dx"], 1e-14)


def test_igrf12_legendre(orc, fx):
    for c in fx["igrf12"]:
        close(orc.igrf12(c["date"], c["r"], c["lat"], c["lon"]), c["B"], 1e-12)
    for c in fx["legendre"]:
        P = orc.legendre_schmidt(c["theta"], 13)
        close(P, c["P"], 1e-13)
        close(orc.dlegendre_schmidt(c["theta"], P), c["dP"], 1e-13)


def test_magnetic_simulation_gramian_cutoff(orc, fx):
    for c in fx["magnetic_simulation"]:
        N = c["N"]
        B, pos, vel, rc = orc.magnetic_simulation(c["kep"], c["GM"], c["mjd"], c["igrf_date"], c["field_radius_m"], c["t0"], c["tf"], N)
        assert rc == 0
        bscale = np.max(np.abs(B))
        for k, row in c["scope_rows"].items():
            assert np.max(np.abs(B[int(k)] - np.array(row))) < 1e-10 * bscale, k
        for k, p in c["scope_pos"].items():
            close(pos[int(k)], p, 1e-11)
        for k, v in c.get("scope_vel", {}).items():
            close(vel[int(k)], v, 1e-11)
        G = orc.magnetic_gramian(B, (c["tf"] - c["t0"]) / N)
        for k, g in c["gram"].items():
            close(G[int(k)], g, 1e-10)
        assert orc.condition_based_time(G, c["cutoff"]) == c["tf_index"]
        if "fine_rows" in c:
            assert abs(c["tf_index"] * (c["tf"] - c["t0"]) / N - c["t_final"]) < 1e-12
            Nf = c["N_fine"]
            assert Nf == int(math.floor((c["t_final"] - c["t0"]) / 0.2))
            Bf, posf, _, _ = orc.magnetic_simulation(c["kep"], c["GM"], c["mjd"], c["igrf_date"], c["field_radius_m"], c["t0"], c["t_final"], Nf)
            for k, row in c["fine_rows"].items():
                assert np.max(np.abs(Bf[int(k)] - np.array(row))) < 1e-10 * bscale, k
            for k, p in c["fine_pos"].items():
                close(posf[int(k)], p, 1e-11)
    L = orc.lib()
    for c in fx["cond"]:
        got = L.orc_cond_sym3(_p(c["G"]))
        assert abs(got - c["cond"]) <= 1e-9 * c["cond"]


def test_eigen_axis_slew_literal_and_bryson(orc, fx):
    L = orc.lib()
    for c in fx["eigen_axis_slew"]:
        nt = c["nt"]
        t = 0.0 + c["dt"] * np.arange(nt)
        w, q = np.zeros((nt, 3)), np.zeros((nt, 4))
        L.orc_eigen_axis_slew(_p(c["x0"]), _p(c["xf"]), orc.P(t), nt, orc.P(w), orc.P(q))
        for k, row in c["w_rows"].items():
            close(w[int(k)], row, 1e-11, 1e-18)
        for k, row in c["q_rows"].items():
            close(q[int(k)], row, 1e-12)
        Qd, Qfd, Rd = np.zeros(8), np.zeros(8), np.zeros(3)
        L.orc_bryson_weights(orc.P(w), nt, _p(c["J"]), c["dt"], c["alpha"], c["beta"], orc.P(Qd), orc.P(Qfd), orc.P(Rd))
        close(Qd, c["Qd"], 1e-10)
        close(Qfd, c["Qfd"], 1e-10)
        close(Rd, c["Rd"], 1e-9)        # 1/m_max^2 with m_max from a difference of nearly equal rates
    # the literal product differs from the conjugate one exactly when both attitudes are non-identity
    c = fx["eigen_axis_slew"][2]
    nt = c["nt"]
    t = c["dt"] * np.arange(nt)
    w0, q0, w1, q1 = np.zeros((nt, 3)), np.zeros((nt, 4)), np.zeros((nt, 3)), np.zeros((nt, 4))
    L.orc_eigen_axis_slew_mode(_p(c["x0"]), _p(c["xf"]), orc.P(t), nt, orc.P(w0), orc.P(q0), 0)
    L.orc_eigen_axis_slew_mode(_p(c["x0"]), _p(c["xf"]), orc.P(t), nt, orc.P(w1), orc.P(q1), 1)
    assert np.max(np.abs(w0 - w1)) > 1e-6
    close(w0[1], c["w_rows"]["1"], 1e-11)


def test_dynamics_rk3_jacobians(orc, fx):
    d = fx["dynamics"]
    B = np.zeros((max(128, 8), 3))
    B[:128] = np.array(d["B_rows"])
    dyn = orc.make_dyn(B, float(d["N"]), 1.0 / (d["tf"] - d["t0"]), np.array(d["J"]))
    L = orc.lib()
    for c in d["cases"]:
        x, u = orc.f64(c["x"]), orc.f64(c["u"])
        dx = np.zeros(8)
        L.orc_deriv_function(C.byref(dyn), orc.P(x), orc.P(u), orc.P(dx))
        close(dx, c["DerivFunction"], 1e-13)
        L.orc_gain_simulator(C.byref(dyn), orc.P(x), orc.P(u), orc.P(dx))
        close(dx, c["gain_simulator"], 1e-13)
        n9 = np.array(c["noise9"])
        nz = np.concatenate([n9[0:3] * (.38 * math.pi / 180) ** 2, n9[3:6] * (1 * math.pi / 180) ** 2, n9[6:9] * (1E-5) ** 2])
        L.orc_simulator(C.byref(dyn), orc.P(x), orc.P(u), orc.P(nz), orc.P(dx))
        close(dx, c["simulator"], 1e-12)
        dx7 = np.zeros(7)
        L.orc_attitude_dynamics(_p(x[:7]), orc.P(u), _p(c["B_B"]), _p(c["J_ad"]), orc.P(dx7))
        close(dx7, c["attitude_dynamics"], 1e-13)
        xn = np.zeros(8)
        L.orc_rk3_step(C.byref(dyn), orc.P(x), orc.P(u), 0.2, orc.P(xn))
        close(xn, c["rk3"], 1e-14)
        A, Bm = np.zeros((8, 8)), np.zeros((8, 3))
        L.orc_rk3_jacobian(C.byref(dyn), orc.P(x), orc.P(u), 0.2, orc.P(A), orc.P(Bm))
        close(A, c["rk3_A"], 1e-12)
        close(Bm, c["rk3_B"], 1e-12)


def _slew_from_fixture(S, c):
    s = S.build_slew(c["kep"], np.array(c["J"]), np.array(c["x0"][3:7]), np.array(c["xf"][3:7]), t_final=c["t_final"], alpha=c["alpha"])
    assert s.N == c["N"]
    if c["Qd"] is not None:
        assert np.max(np.abs(s.Qd - np.array(c["Qd"])) / np.maximum(np.array(c["Qd"]), 1e-300)) < 1e-10
    return s


def test_alilqr_second_implementation(orc, fx):
    """The oracle's AL-iLQR (C++, dual numbers, hand-rolled Cholesky) against the numpy restatement of the same App. C
    specification (complex-step Jacobians, LAPACK Cholesky): identical iteration paths, J to 1e-8, trajectories to 1e-7."""
    import slew_setup as S
    for c in fx["alilqr"]:
        s = _slew_from_fixture(S, c)
        o = orc.default_ilqr_opts()
        for k, v in c["opts"].items():
            setattr(o, k, v)
        Xs, Us, Ks, out = S.oracle_solve([s], opts=o)
        r, info = out[0], c["info"]
        assert (r["status"], r["outer_iters"], r["inner_iters"], r["ls_rollouts"]) == \
               (info["status"], info["outer_iters"], info["inner_iters"], info["ls_rollouts"]), (c["name"], r, info)
        assert abs(r["J"] - info["J"]) <= 1e-8 * abs(info["J"])
        assert abs(r["c_max"] - info["c_max"]) <= 1e-8
        for k, row in c["X_rows"].items():
            close(Xs[0][int(k)], row, 1e-7)
        for k, row in c["U_rows"].items():
            close(Us[0][int(k)], row, 1e-6, 1e-9)
        if c["K0"] is not None:
            close(Ks[0][0], c["K0"], 1e-6, 1e-9)


def test_quaternion_aware_variant_second_implementation(orc, fx):
    """the oracle's quat_error mode (8 x 8 arrays padded with a zero error-state slot) against the numpy
    transliteration of the reference's own hooks -- quaternion_error and quaternion_expansion of quaternion_toolbox.jl, a
    true 7 x 7 error-state Riccati recursion: identical iteration paths, J to 1e-8, trajectories to 1e-7."""
    import slew_setup as S
    assert len(fx["alilqr_quat"]) >= 2
    for c in fx["alilqr_quat"]:
        s = _slew_from_fixture(S, c)
        o = orc.default_ilqr_opts()
        for k, v in c["opts"].items():
            setattr(o, k, v)
        assert o.quat_error == 1
        Xs, Us, Ks, out = S.oracle_solve([s], opts=o)
        r, info = out[0], c["info"]
        assert (r["status"], r["outer_iters"], r["inner_iters"], r["ls_rollouts"]) == \
               (info["status"], info["outer_iters"], info["inner_iters"], info["ls_rollouts"]), (c["name"], r, info)
        assert abs(r["J"] - info["J"]) <= 1e-8 * abs(info["J"])
        assert abs(r["c_max"] - info["c_max"]) <= 1e-8
        for k, row in c["X_rows"].items():
            close(Xs[0][int(k)], row, 1e-7)
        for k, row in c["U_rows"].items():
            close(Us[0][int(k)], row, 1e-6, 1e-9)
        K0 = np.array(c["K0"])
        close(Ks[0][0], K0, 1e-6, 1e-9)
        assert np.all(K0[:, 6:] == 0.0) and np.all(Ks[0][0][:, 6:] == 0.0)     # gains live in the 6-dim error state


def test_tvlqr_replay_and_postprocessing(orc, fx):
    import slew_setup as S
    L = orc.lib()
    for c in fx["tvlqr"]:
        a = fx["alilqr"][c["alilqr_case"]]
        s = _slew_from_fixture(S, a)
        o, _g = None, None
        o = orc.TvlqrOpts()
        o.dt, o.t0, o.dt_squared, o.seed = 0.2, 0.0, 1, 0
        for i in range(6):
            o.Qd[i], o.Qfd[i] = 10.0, 1000.0
        for i in range(3):
            o.Rd[i] = c["R"]
        noise = None
        if c["noise"] is not None:
            n = np.array(c["noise"])                                      # (N, 4, 9) raw draws of simulator.jl:5,10,22
            noise = np.concatenate([n[..., 0:3] * (.38 * math.pi / 180) ** 2, n[..., 3:6] * (1 * math.pi / 180) ** 2,
                                    n[..., 6:9] * (1E-5) ** 2], axis=-1)
            o.noise_mode = 1
        else:
            o.noise_mode = 0
        X, U = np.array(c["X_lqr"]), np.array(c["U_lqr"])
        Xs, Us, dX, K, ns, slew = S.oracle_tvlqr(s, X, U, np.array(c["x0_lqr"]), o, trial=0, noise=noise)
        assert ns == c["N_sim"]
        for k, row in c["K_rows"].items():
            close(K[int(k)], row, 1e-9, 1e-12)
        for k, row in c["X_sim_rows"].items():
            close(Xs[int(k)], row, 1e-10)
        for k, row in c["U_sim_rows"].items():
            close(Us[int(k)], row, 1e-9, 1e-12)
        for k, row in c["dX_rows"].items():
            close(dX[int(k)], row, 1e-9, 1e-13)
        assert abs(slew - c["slew_time"]) < 1e-9
    for c in fx["mc_postprocess"]:
        X = orc.f64(c["X_sim"])
        qf = orc.f64(c["q_final"])
        s1 = L.orc_mc_slew_time(orc.P(X), X.shape[0], orc.P(qf), c["t_final"], c["time_step"], 0.05, 0.08727, 0, 1)
        s2 = L.orc_mc_slew_time(orc.P(X), X.shape[0], orc.P(qf), c["t_final"], c["time_step"], 0.05, 0.08727, 1, 7)
        assert abs(s1 - c["slew_time"]) < 1e-12 and abs(s2 - c["slew_time_literal_i7"]) < 1e-12
        assert (s1 == c["t_final"]) == bool(c["fail"])


def test_comparison_controller(orc, fx):
    """the oracle's Psiaki PD closed loop and attitude_dynamics_linear against the numpy transliteration of
    comparison/psiaki2005.jl:116-164, psiaki_dynamics.jl:1-26,63-73 and attitude_dynamics.jl:26-48."""
    L = orc.lib()
    for c in fx["psiaki"]["cases"]:
        N = c["N"]
        X, M, Qe = np.zeros((N, 7)), np.zeros((N, 3)), np.zeros((N, 4))
        L.orc_psiaki_pd_simulation(N, _p(c["x0"]), _p(c["w_guess"]), _p(c["q_guess"]), _p(c["B_eci"]), _p(c["J"]), c["dt"], c["C_1"], c["C_2"],
                                   orc.P(X), orc.P(M), orc.P(Qe))
        for k, row in c["X_rows"].items():
            close(X[int(k)], row, 1e-11)
        for k, row in c["M_rows"].items():
            close(M[int(k)], row, 1e-9, 1e-16)
        for k, row in c["q_err_rows"].items():
            close(Qe[int(k)], row, 1e-11)
    c = fx["psiaki"]["linear"]
    dx = np.zeros(7)
    L.orc_attitude_dynamics_linear(_p(c["x"]), _p(c["u"]), _p(c["x_linear"]), _p(c["B_B"]), _p(c["J"]), orc.P(dx))
    close(dx, c["dx"], 1e-13)
