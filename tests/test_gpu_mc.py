"""GPU: slew preparation, K4 TVLQR replay and the fused Monte-Carlo path vs the CPU oracle."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import slew_setup as S
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GM = S.GM
HERE = os.path.dirname(os.path.abspath(__file__))


def _perturb(q0, qn):
    th = np.linalg.norm(qn)
    out = np.zeros(4)
    orc.lib().orc_qmult(orc.P(orc.f64(q0)), orc.P(np.concatenate([[math.cos(th / 2)], qn / th * math.sin(th / 2)])), orc.P(out))
    return out


def test_slew_weights_and_guess(engine):
    from tortoisesat.jl_b200 import host
    rng = np.random.default_rng(3)
    T = 9
    x0 = np.zeros((T, 8))
    xf = np.zeros((T, 8))
    Jm = np.zeros((T, 9))
    tfin = rng.uniform(40, 480, size=T)
    for t in range(T):
        q = rng.normal(size=4)
        x0[t, 3:7] = q / np.linalg.norm(q)
        xf[t, 3:7] = [math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0]
        xf[t, 7] = 1
        Jm[t] = (S.J_1U if t % 2 else S.J_3U).reshape(-1)
    Qd, Qfd, Rd, wg, qg, goffs = engine.slew_weights_batch(x0, xf, Jm, tfin, dt=0.2, alpha=0.1, beta=1e3, want_guess=True)
    L = orc.lib()
    for t in range(T):
        nt = int(goffs[t + 1] - goffs[t])
        tt = 0.2 * np.arange(nt)
        assert nt == int(math.floor(tfin[t] / 0.2 + 1e-9)) + 1
        w_o, q_o = np.zeros((nt, 3)), np.zeros((nt, 4))
        L.orc_eigen_axis_slew(orc.P(orc.f64(x0[t, :7])), orc.P(orc.f64(xf[t, :7])), orc.P(tt), nt, orc.P(w_o), orc.P(q_o))
        Qo, Qfo, Ro = np.zeros(8), np.zeros(8), np.zeros(3)
        L.orc_bryson_weights(orc.P(w_o), nt, orc.P(orc.f64(Jm[t])), 0.2, 0.1, 1e3, orc.P(Qo), orc.P(Qfo), orc.P(Ro))
        assert np.allclose(wg[goffs[t]:goffs[t + 1]], w_o, rtol=1e-9, atol=1e-16)
        assert np.allclose(qg[goffs[t]:goffs[t + 1]], q_o, rtol=0, atol=1e-12)
        assert np.allclose(Qd[t], Qo, rtol=1e-9) and np.allclose(Qfd[t], Qfo, rtol=1e-9) and np.allclose(Rd[t], Ro, rtol=1e-8)
    # config-1 known answers
    x0c = np.concatenate([[0, 0, 0], S.quat_axis_angle([1, 0, 1], 90.0), [0]])
    xfc = np.array([0, 0, 0, 1.0, 0, 0, 0, 1])
    Qd, Qfd, Rd = engine.slew_weights_batch([x0c], [xfc], S.J_1P.reshape(1, 9), [311.04], dt=0.2, alpha=10.0, beta=1e3)
    assert abs(Qd[0, 0] / 317739.65036560915 - 1) < 1e-9 and Qd[0, 3] == 1e4 and abs(Rd[0, 0] / 286.97003750452905 - 1) < 1e-8
    w, q = host.eigen_axis_slew(x0c[:7], xfc[:7], 0.2 * np.arange(1556))
    assert w.shape == (1556, 3) and abs(np.max(np.abs(w)) - 5.610018498990348e-3) < 1e-12


@pytest.fixture(scope="module")
def solved_pair():
    qf = np.array([1.0, 0, 0, 0])
    sl = [S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], 5.0), qf, t_final=60.0),
          S.build_slew([0, 6871, 51.6, 30, 0, 10], S.J_3U, S.quat_axis_angle([0, 1, 0], 3.0), qf, t_final=50.0)]
    Xs, Us, Ks, out = S.oracle_solve(sl, nthreads=2)
    return sl, Xs, Us


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tvlqr_replay_matches_oracle(engine, solved_pair, mode):
    sl, Xs, Us = solved_pair
    rng = np.random.default_rng(7)
    o, g = S.tvlqr_opts_pair(noise_mode=mode, seed=77)
    T = len(sl)
    N_i = [s.N for s in sl]
    Xl = np.concatenate(Xs)
    Ul = np.concatenate([np.vstack([u, np.zeros((1, 3))]) for u in Us])
    x0l = []
    for s in sl:
        x = s.x0.copy()
        x[3:7] = _perturb(s.x0[3:7], rng.normal(size=3) * (math.pi / 180) ** 2)
        x[7] = 0
        x0l.append(x)
    rows = np.array([s.B.shape[0] for s in sl])
    B_offs = np.concatenate([[0], np.cumsum(rows)])[:-1]
    noise = None
    if mode == 1:
        noise = rng.normal(size=(sum(N_i), 4, 9)) * np.array([1e-5] * 3 + [3e-4] * 3 + [1e-10] * 3)
    Xsim, Usim, dX, K, nsim, slew, offs = engine.tvlqr_sim_batch(
        N_i, Xl, Ul, np.stack(x0l), np.stack([s.J.reshape(-1) for s in sl]), np.concatenate([s.B for s in sl]), B_offs, rows,
        [s.index_scale for s in sl], [s.clock_rate for s in sl], [s.t_final for s in sl], np.stack([s.xf[3:7] for s in sl]),
        opts=g, noise=noise, stream_id=[11, 12])
    for t, s in enumerate(sl):
        nz = None if noise is None else noise[offs[t]:offs[t + 1]]
        a = S.oracle_tvlqr(s, Xs[t], Us[t], x0l[t], o, trial=11 + t, noise=nz)
        assert nsim[t] == a[4]
        n = int(nsim[t])
        assert np.max(np.abs(Xsim[offs[t]:offs[t] + n] - a[0])) < 1e-10          # rollouts to 1e-10
        assert np.max(np.abs(Usim[offs[t]:offs[t] + n] - a[1])) < 1e-9
        assert np.max(np.abs(dX[offs[t]:offs[t] + n] - a[2])) < 1e-10
        assert np.max(np.abs(K[offs[t]:offs[t] + s.N - 1] - a[3])) <= 1e-9 * np.max(np.abs(a[3]))
        assert slew[t] == a[5]


def _oracle_trial(kep, fo, x0, xf, J, cfg, qn, sid):
    """The per-trial pipeline of monte_carlo.jl:118-262 / TortoiseSat.jl:58-265 with the oracle."""
    s = S.build_slew(kep, J, x0[3:7], xf[3:7], mjd=fo["mjd"], igrf_date=fo["igrf_date"], field_radius_m=fo["field_radius_m"],
                     t0=cfg.t0, tf=cfg.tf, N_scope=int(cfg.N_scope), cutoff=cfg.cutoff, dt=cfg.dt, alpha=cfg.alpha, beta=cfg.beta)
    Xs, Us, Ks, out = S.oracle_solve([s])
    o, g = S.tvlqr_opts_pair(noise_mode=2, seed=int(cfg.tvlqr.seed), R=float(cfg.tvlqr.Rd[0]))
    x0l = s.x0.copy()
    x0l[3:7] = _perturb(s.x0[3:7], qn)
    x0l[7] = 0
    a = S.oracle_tvlqr(s, Xs[0], Us[0], x0l, o, trial=sid)
    _oracle_trial.last = (Xs[0], Us[0], a)
    return s, out[0], a[5]


def test_fused_monte_carlo_shared_orbit(engine):
    """configs[2] in miniature: fixed LEO orbit (monte_carlo.jl:122-127 with RAAN 0, anomaly 90), random attitudes."""
    from tortoisesat.jl_b200 import host
    rng = np.random.default_rng(21)
    n = 4
    cfg = host.default_mc_config(n, shared_orbit=True, run_tvlqr=True, tf=2400.0, cutoff=30.0, alpha=0.1)
    cfg.tvlqr.noise_mode = 2
    cfg.tvlqr.seed = 4242
    kep = np.array([[0, 6771.0, 96.6, 0.0, 0.0, 90.0]])
    fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
    fo[0] = (GM, 58155.0, 2019.0, 6771000.0, 0, 0, 0)
    qf = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])
    x0 = np.zeros((n, 8))
    xf = np.tile(np.concatenate([[0, 0, 0], qf, [1.0]]), (n, 1))
    for t in range(n):
        dq = S.quat_axis_angle(rng.normal(size=3), rng.uniform(5, 40))
        q = np.zeros(4)
        orc.lib().orc_qmult(orc.P(orc.f64(qf)), orc.P(dq), orc.P(q))
        x0[t, 3:7] = q
    Jm = np.tile(S.J_1U.reshape(-1), (n, 1))
    qn = rng.normal(size=(n, 3)) * (math.pi / 180) ** 2
    out, st = engine.monte_carlo_run(cfg, kep, fo, x0, xf, Jm, q_noise0=qn, stream_id=np.arange(100, 100 + n))
    assert st.n_trials == n and st.n_no_cutoff == 0
    for t in range(n):
        s, ref, slew = _oracle_trial(kep[0], fo[0], x0[t], xf[t], S.J_1U, cfg, qn[t], 100 + t)
        g = out[t]
        assert g["N"] == s.N and abs(g["t_final"] - s.t_final) < 1e-9
        assert g["status"] == ref["status"] and g["outer_iters"] == ref["outer_iters"], (t, g, ref)
        assert abs(g["J"] - ref["J"]) <= 1e-6 * abs(ref["J"])
        assert abs(g["c_max"] - ref["c_max"]) <= 1e-6 * max(1.0, ref["c_max"])
        assert g["slew_time"] == slew, (t, g["slew_time"], slew)
    assert st.flops > 0 and st.ms_solve > 0


def test_fused_monte_carlo_sweep_and_no_cutoff(engine):
    """configs[3] in miniature: per-trial inclination / altitude / RAAN / anomaly / MJD / IGRF date; one
    equatorial trial never reaches the cutoff -> per-trial status NO_CUTOFF, batch unaffected."""
    from tortoisesat.jl_b200 import host
    rng = np.random.default_rng(8)
    n = 4
    cfg = host.default_mc_config(n, shared_orbit=False, run_tvlqr=True, tf=2400.0, cutoff=100.0, alpha=0.1)
    cfg.tvlqr.noise_mode = 2
    cfg.tvlqr.seed = 9
    kep = np.zeros((n, 6))
    fo = np.zeros(n, dtype=host.FIELD_OPTS_DTYPE)
    for t in range(n):
        alt = rng.uniform(350, 800)
        kep[t] = [0, alt + 6371.0, rng.uniform(40, 98), rng.uniform(0, 360), 0, rng.uniform(0, 360)]
        fo[t] = (GM, rng.uniform(58155, 58520), 2015 + 5 * rng.random(), (alt + 6371.0) * 1000, 0, 0, 0)
    cfg.cutoff = 100.0
    qf = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])
    x0 = np.zeros((n, 8))
    xf = np.tile(np.concatenate([[0, 0, 0], qf, [1.0]]), (n, 1))
    for t in range(n):
        dq = S.quat_axis_angle(rng.normal(size=3), rng.uniform(5, 30))
        q = np.zeros(4)
        orc.lib().orc_qmult(orc.P(orc.f64(qf)), orc.P(dq), orc.P(q))
        x0[t, 3:7] = q
    Jm = np.tile(S.J_1U.reshape(-1), (n, 1))
    qn = rng.normal(size=(n, 3)) * (math.pi / 180) ** 2
    out, st = engine.monte_carlo_run(cfg, kep, fo, x0, xf, Jm, q_noise0=qn)
    for t in range(n):
        s, ref, slew = _oracle_trial(kep[t], fo[t], x0[t], xf[t], S.J_1U, cfg, qn[t], t)
        g = out[t]
        assert g["N"] == s.N and abs(g["t_final"] - s.t_final) < 1e-9, (t, g, s.N)
        assert g["status"] == ref["status"] and g["outer_iters"] == ref["outer_iters"], (t, g, ref)
        assert abs(g["J"] - ref["J"]) <= 1e-6 * abs(ref["J"])
        assert g["slew_time"] == slew
    # a cutoff nobody can reach -> every trial NO_CUTOFF, call still succeeds
    cfg2 = host.default_mc_config(n, shared_orbit=False, run_tvlqr=False, tf=2400.0, cutoff=1.0)
    out2, st2 = engine.monte_carlo_run(cfg2, kep, fo, x0, xf, Jm)
    assert np.all(out2["status"] == 5) and st2.n_no_cutoff == n


@pytest.mark.parametrize("quat", [0, 1])
def test_bench_ensemble_sample_matches_oracle(engine, quat):
    """The first 16 trials of bench.py's configs[2] ensemble at FULL size (N = 2044 knots, random attitudes on S^3,
    20 x 50 AL-iLQR iterations; several of them go through the straggler hand-over): identical status and
    outer / inner / line-search counters, converged cost and constraint violation to 1e-6 (north_star's bar).
    quat = 1: the same with the quaternion-aware solver variant (ts_ilqr_opts.quat_error)."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import bench as B
    from tortoisesat.jl_b200 import host
    n = 16
    tr = B.make_trials("mc_fixed_orbit", 4096, 0)
    sub = dict(tr)
    for k in ("x0", "xf", "Jm", "qn"):
        sub[k] = tr[k][:n]
    cfg = B.mc_config(host, sub, n)
    cfg.run_tvlqr = 0
    cfg.ilqr.quat_error = quat
    fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
    fo[0] = tr["fo"][0]
    out, st = engine.monte_carlo_run(cfg, tr["kep"], fo, sub["x0"], sub["xf"], sub["Jm"], q_noise0=sub["qn"])
    assert engine.k3_last_split()[2] > 0                      # the one-warp-per-trial launch was exercised
    slews, base = [], None
    f = tr["fo"][0]
    for t in range(n):
        s = S.build_slew(tr["kep"][0], B.J_1U, tr["x0"][t, 3:7], B.QF, mjd=f[1], igrf_date=f[2], field_radius_m=f[3], tf=2400.0,
                         cutoff=tr["cutoff"], alpha=0.1, **({} if base is None else dict(t_final=base.t_final)))
        base = base or s
        slews.append(s)
    oo = orc.default_ilqr_opts()
    oo.quat_error = quat
    Xs, Us, Ks, ref = S.oracle_solve(slews, oo, nthreads=orc.lib().orc_max_threads(), want_K=False)
    for t in range(n):
        g, r = out[t], ref[t]
        assert g["N"] == slews[t].N == 2044
        for fld in ("status", "outer_iters", "inner_iters", "ls_rollouts"):
            assert g[fld] == r[fld], (t, fld, g, r)
        assert abs(g["J"] - r["J"]) <= 1e-6 * abs(r["J"])
        assert abs(g["c_max"] - r["c_max"]) <= 1e-6 * max(1.0, r["c_max"])


def test_fused_monte_carlo_quaternion_aware(engine):
    """The fused Monte-Carlo call with ts_ilqr_opts.quat_error = 1 (the solver configuration monte_carlo.jl:158,192
    actually requests): field pass -> weights -> QUAT K3 kernels -> TVLQR replay, every trial against the oracle pipeline
    run with the same option."""
    from tortoisesat.jl_b200 import host
    rng = np.random.default_rng(77)
    n = 5
    cfg = host.default_mc_config(n, shared_orbit=True, run_tvlqr=True, tf=2400.0, cutoff=30.0, alpha=0.1)
    cfg.tvlqr.noise_mode, cfg.tvlqr.seed = 2, 99
    cfg.ilqr.quat_error = 1
    kep = np.array([[0, 6771.0, 96.6, 0.0, 0.0, 90.0]])
    fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
    fo[0] = (GM, 58155.0, 2019.0, 6771000.0, 0, 0, 0)
    qf = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])
    x0 = np.zeros((n, 8))
    xf = np.tile(np.concatenate([[0, 0, 0], qf, [1.0]]), (n, 1))
    for t in range(n):
        dq = S.quat_axis_angle(rng.normal(size=3), rng.uniform(5, 40))
        q = np.zeros(4)
        orc.lib().orc_qmult(orc.P(orc.f64(qf)), orc.P(dq), orc.P(q))
        x0[t, 3:7] = q
    Jm = np.tile(S.J_1U.reshape(-1), (n, 1))
    qn = rng.normal(size=(n, 3)) * (math.pi / 180) ** 2
    out, st = engine.monte_carlo_run(cfg, kep, fo, x0, xf, Jm, q_noise0=qn, stream_id=np.arange(n))
    oq = orc.default_ilqr_opts()
    oq.quat_error = 1
    for t in range(n):
        s = S.build_slew(kep[0], S.J_1U, x0[t, 3:7], xf[t, 3:7], mjd=58155.0, igrf_date=2019.0, field_radius_m=6771000.0,
                         t0=cfg.t0, tf=cfg.tf, N_scope=int(cfg.N_scope), cutoff=cfg.cutoff, dt=cfg.dt, alpha=cfg.alpha, beta=cfg.beta)
        Xs, Us, Ks, ref = S.oracle_solve([s], oq)
        o, g_ = S.tvlqr_opts_pair(noise_mode=2, seed=99, R=float(cfg.tvlqr.Rd[0]))
        x0l = s.x0.copy()
        x0l[3:7] = _perturb(s.x0[3:7], qn[t])
        x0l[7] = 0
        a = S.oracle_tvlqr(s, Xs[0], Us[0], x0l, o, trial=t)
        g = out[t]
        assert g["N"] == s.N
        assert g["status"] == ref[0]["status"] and g["outer_iters"] == ref[0]["outer_iters"], (t, g, ref[0])
        assert abs(g["J"] - ref[0]["J"]) <= 1e-6 * abs(ref[0]["J"])
        assert abs(g["c_max"] - ref[0]["c_max"]) <= 1e-6 * max(1.0, ref[0]["c_max"])
        if g["inner_iters"] == ref[0]["inner_iters"]:
            assert g["slew_time"] == a[5], (t, g["slew_time"], a[5])


def test_monte_carlo_trajectories_match_oracle(engine):
    """keep_trajectories = 1: the arrays monte_carlo.jl leaves in globals (states, control_inputs, sim_states,
    sim_control_inputs, B_ECI_total; monte_carlo.jl:52-66,149,200-201,232-233) come back through
    ts_mc_fetch_trajectories and equal the oracle's per-trial pipeline."""
    from tortoisesat.jl_b200 import host
    rng = np.random.default_rng(33)
    n = 3
    cfg = host.default_mc_config(n, shared_orbit=False, run_tvlqr=True, tf=2400.0, cutoff=30.0, alpha=0.1)
    cfg.tvlqr.noise_mode, cfg.tvlqr.seed, cfg.keep_trajectories = 2, 515, 1
    assert cfg.tvlqr.Rd[0] == 0.5e3                                    # monte_carlo.jl:227
    kep = np.zeros((n, 6))
    fo = np.zeros(n, dtype=host.FIELD_OPTS_DTYPE)
    for t in range(n):
        kep[t] = [0, 6771.0, 96.6, rng.uniform(0, 360), 0, rng.uniform(0, 360)]
        fo[t] = (GM, 58155.0, 2019.0, 6771000.0, 0, 0, 0)
    qf = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])
    x0 = np.zeros((n, 8))
    x0[:, 3] = 1.0                                                     # q_0 = identity, as monte_carlo.jl:108-111
    x0[1, 3:7] = S.quat_axis_angle([0.2, 1, -0.4], 50.0)              # and one trial with both attitudes non-identity
    xf = np.tile(np.concatenate([[0, 0, 0], qf, [1.0]]), (n, 1))
    Jm = np.tile(S.J_1U.reshape(-1), (n, 1))
    qn = rng.normal(size=(n, 3)) * (math.pi / 180) ** 2
    out, st = engine.monte_carlo_run(cfg, kep, fo, x0, xf, Jm, q_noise0=qn, stream_id=np.arange(7, 7 + n))
    tr = engine.mc_trajectories(n)
    ko, ro = tr["knot_offs"], tr["row_offs"]
    assert ko[-1] == out["N"].sum() and sum(st.n_status) == n
    for t in range(n):
        s, ref, slew = _oracle_trial(kep[t], fo[t], x0[t], xf[t], S.J_1U, cfg, qn[t], 7 + t)
        Xo, Uo, a = _oracle_trial.last
        assert (out[t]["status"], out[t]["outer_iters"], out[t]["inner_iters"]) == (ref["status"], ref["outer_iters"], ref["inner_iters"])
        N = s.N
        assert ko[t + 1] - ko[t] == N
        assert np.max(np.abs(tr["X"][ko[t]:ko[t + 1]] - Xo)) < 1e-8
        assert np.max(np.abs(tr["U"][ko[t]:ko[t + 1] - 1] - Uo)) < 1e-7
        ns = a[4]
        assert np.max(np.abs(tr["X_sim"][ko[t]:ko[t] + ns] - a[0])) < 1e-7      # replay of a 1e-8-equal optimised slew
        assert np.max(np.abs(tr["U_sim"][ko[t]:ko[t] + ns] - a[1])) < 1e-6
        rows = 2 * N
        Bg = tr["B_eci"][ro[t]:ro[t] + rows]
        used = np.nonzero(np.any(Bg != 0, axis=1))[0]
        assert len(used) > 8 and used[-1] < rows - 1
        assert np.max(np.abs(Bg[used] - s.B[used])) < 1e-10 * np.max(np.abs(s.B))
        assert out[t]["slew_time"] == slew
    # the reference-named wrapper hands the same arrays back in the script's shapes
    r = host.monte_carlo(number_sims=2, seed=3, trajectories=True)
    assert len(r["states"]) == 2 and r["states"][0].shape[0] == 8 and r["control_inputs"][0].shape[0] == 3
    assert r["sim_states"][0].shape == r["states"][0].shape and r["B_ECI_total"][0].shape[1] == 3
    assert r["states"][0].shape[1] == r["outcomes"]["N"][0] == r["control_inputs"][0].shape[1] + 1


def test_sweep_ensemble_sample_matches_oracle(engine):
    """BASELINE configs[3] shape: the first trials of bench.py's magnetic-diversity sweep (per-trial inclination,
    altitude, RAAN, anomaly, MJD, IGRF date; ragged horizons) against the oracle pipeline: same horizon, status,
    outer / inner / line-search counters, J and c_max to 1e-6, same slew time."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    import bench as B
    from tortoisesat.jl_b200 import host
    n = 24
    tr = B.make_trials("mc_sweep", n, 0)
    cfg = B.mc_config(host, tr, n)
    fo = np.zeros(n, dtype=host.FIELD_OPTS_DTYPE)
    for i, f in enumerate(tr["fo"]):
        fo[i] = f
    sid = np.arange(n).astype(np.uint32)
    out, st = engine.monte_carlo_run(cfg, tr["kep"], fo, tr["x0"], tr["xf"], tr["Jm"], q_noise0=tr["qn"], stream_id=sid)
    pick = [t for t in range(n) if out[t]["status"] != 5 and out[t]["N"] <= 2600][:8]    # bounded oracle time
    assert len(pick) >= 6
    import concurrent.futures as cf

    def one(t):
        return _oracle_trial_sweep(tr, fo, cfg, t)
    with cf.ThreadPoolExecutor(8) as ex:                 # the oracle calls release the GIL (ctypes)
        refs = list(ex.map(one, pick))
    n_same = 0
    for t, (s, ref, slew) in zip(pick, refs):
        g = out[t]
        assert g["N"] == s.N and abs(g["t_final"] - s.t_final) < 1e-9, (t, g, s.N)
        assert g["status"] == ref["status"] and g["outer_iters"] == ref["outer_iters"], (t, g, ref)
        assert abs(g["J"] - ref["J"]) <= 1e-6 * abs(ref["J"])
        assert abs(g["c_max"] - ref["c_max"]) <= 1e-6 * max(1.0, ref["c_max"])
        if g["inner_iters"] == ref["inner_iters"] and g["ls_rollouts"] == ref["ls_rollouts"]:
            n_same += 1
            assert g["slew_time"] == slew
    assert n_same >= len(pick) - 1


def _oracle_trial_sweep(tr, fo, cfg, t):
    s = S.build_slew(tr["kep"][t], S.J_1U, tr["x0"][t, 3:7], tr["xf"][t, 3:7], mjd=fo[t]["mjd"], igrf_date=fo[t]["igrf_date"],
                     field_radius_m=fo[t]["field_radius_m"], t0=cfg.t0, tf=cfg.tf, N_scope=int(cfg.N_scope), cutoff=cfg.cutoff,
                     dt=cfg.dt, alpha=cfg.alpha, beta=cfg.beta)
    Xs, Us, Ks, out = S.oracle_solve([s], want_K=False)
    o = S.oracle_tvlqr_opts(noise_mode=2, seed=int(cfg.tvlqr.seed), R=float(cfg.tvlqr.Rd[0]))
    x0l = s.x0.copy()
    x0l[3:7] = _perturb(s.x0[3:7], tr["qn"][t])
    x0l[7] = 0
    a = S.oracle_tvlqr(s, Xs[0], Us[0], x0l, o, trial=t)
    return s, out[0], a[5]


def test_tvlqr_tracking_batch_matches_oracle(engine):
    """BASELINE configs[4] shape: batched TVLQR tracking (simulator.jl / gain_simulator.jl through attitude_simulation)
    of optimised slews with Philox disturbance draws, stand-alone through ts_tvlqr_sim_batch with DEVICE-side noise
    generation, on a batch large enough to fill several warps: every trial against the oracle replay."""
    qf = np.array([1.0, 0, 0, 0])
    base = [S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], a), qf, t_final=tf_)
            for a, tf_ in ((3.0, 40.0), (5.0, 60.0), (2.0, 30.0))]
    Xs, Us, Ks, out = S.oracle_solve(base, nthreads=3)
    n = 96
    rng = np.random.default_rng(91)
    idx = [i % 3 for i in range(n)]
    o, g = S.tvlqr_opts_pair(noise_mode=2, seed=2024)
    N_i = [base[i].N for i in idx]
    Xl = np.concatenate([Xs[i] for i in idx])
    Ul = np.concatenate([np.vstack([Us[i], np.zeros((1, 3))]) for i in idx])
    x0l = []
    for i in idx:
        x = base[i].x0.copy()
        x[3:7] = _perturb(base[i].x0[3:7], rng.normal(size=3) * (math.pi / 180) ** 2)
        x[7] = 0
        x0l.append(x)
    rows = np.array([base[i].B.shape[0] for i in idx])
    B_offs = np.concatenate([[0], np.cumsum(rows)])[:-1]
    Xsim, Usim, dX, K, nsim, slew, offs = engine.tvlqr_sim_batch(
        N_i, Xl, Ul, np.stack(x0l), np.stack([base[i].J.reshape(-1) for i in idx]), np.concatenate([base[i].B for i in idx]), B_offs, rows,
        [base[i].index_scale for i in idx], [base[i].clock_rate for i in idx], [base[i].t_final for i in idx],
        np.stack([base[i].xf[3:7] for i in idx]), opts=g, stream_id=np.arange(1000, 1000 + n))
    for t, i in enumerate(idx):
        a = S.oracle_tvlqr(base[i], Xs[i], Us[i], x0l[t], o, trial=1000 + t)
        k = int(nsim[t])
        assert k == a[4]
        assert np.max(np.abs(Xsim[offs[t]:offs[t] + k] - a[0])) < 1e-10
        assert np.max(np.abs(Usim[offs[t]:offs[t] + k] - a[1])) < 1e-9
        assert slew[t] == a[5]


def test_multi_gpu_handle_matches_single(engine):
    """ts_create_multi: trials sharded round-robin over the devices of one node, outcomes gathered with NCCL; results
    must not depend on the device count (Philox streams are keyed by the global trial id)."""
    import torch
    from tortoisesat.jl_b200 import host
    ndev = torch.cuda.device_count()
    rng = np.random.default_rng(5)
    n = 10
    cfg = host.default_mc_config(n, shared_orbit=True, run_tvlqr=True, tf=2400.0, cutoff=30.0, alpha=0.1)
    cfg.tvlqr.noise_mode, cfg.tvlqr.seed = 2, 99
    cfg.ilqr.max_outer = 4
    kep = np.array([[0, 6771.0, 96.6, 0.0, 0.0, 90.0]])
    fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
    fo[0] = (GM, 58155.0, 2019.0, 6771000.0, 0, 0, 0)
    q0 = rng.normal(size=(n, 4))
    q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
    x0 = np.concatenate([np.zeros((n, 3)), q0, np.zeros((n, 1))], axis=1)
    xf = np.tile(np.concatenate([[0, 0, 0], [math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0], [1.0]]), (n, 1))
    Jm = np.tile(S.J_1U.reshape(-1), (n, 1))
    qn = rng.normal(size=(n, 3)) * (math.pi / 180) ** 2
    ref, st_ref = engine.monte_carlo_run(cfg, kep, fo, x0, xf, Jm, q_noise0=qn, stream_id=np.arange(n))
    for devs in ([0], list(range(min(ndev, 2)))):
        m = host.MultiEngine(devs)
        assert m.device_count() == len(devs)
        out, st = m.monte_carlo_run(cfg, kep, fo, x0, xf, Jm, q_noise0=qn)
        m.close()
        for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "N", "J", "c_max", "t_final", "slew_time"):
            assert np.array_equal(out[f], ref[f]), (devs, f)
        assert st.n_trials == n and st.n_converged == st_ref.n_converged and abs(st.sum_slew_time - st_ref.sum_slew_time) < 1e-9
