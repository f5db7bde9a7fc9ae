"""Shared test/bench helpers: builds slew problems (the per-trial pipeline of
reference src/TortoiseSat.jl:49-199) with the CPU oracle, and thin ctypes access
to the host lane-emulator of the K3 kernel source."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GM = 3.986004418E14 * (1 / 1000) ** 3
J_1P = np.diag([0.0001041667] * 3)
J_1U = np.diag([0.00125] * 3)
J_3U = np.diag([0.020833, 0.020833, 0.0041666])

_HS = None


def hostsim():
    """Compiles (g++) and loads tests/hostsim/libhostsim.so."""
    global _HS
    if _HS is None:
        d = os.path.join(HERE, "hostsim")
        lib = os.path.join(d, "libhostsim.so")
        root = os.path.dirname(HERE)
        srcs = [os.path.join(d, "hostsim.cpp")] + [os.path.join(root, "tortoisesat.jl_b200", "csrc", f)
                                                   for f in ("ilqr_solver.cuh", "ilqr_math.cuh", "tvlqr_solver.cuh")]
        srcs = [s for s in srcs if os.path.exists(s)]
        if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
            r = subprocess.run(["/usr/bin/g++", "-O2", "-std=c++20", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                                "-o", lib, os.path.join(d, "hostsim.cpp")], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stderr)
        _HS = C.CDLL(lib)
        _HS.hs_alilqr_solve.argtypes = [C.c_int64] + [C.c_void_p] * 7 + [C.c_int64, C.c_double, C.c_double, C.c_double,
                                                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                          C.c_void_p, C.c_void_p]
        _HS.hs_alilqr_solve_w.argtypes = [C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_int64, C.c_double, C.c_double, C.c_double,
                                                                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                                                     C.c_void_p, C.c_void_p]
        _HS.hs_rk3_jac7.argtypes = [C.c_void_p] * 6 + [C.c_double, C.c_void_p, C.c_void_p]
        _HS.hs_rk3_jac7_jvp.argtypes = [C.c_void_p] * 6 + [C.c_double, C.c_void_p]
        _HS.hs_rk3_jac7_jvp_diag.argtypes = [C.c_void_p] * 6 + [C.c_double, C.c_void_p, C.c_void_p]
        _HS.hs_rk4_jac7.argtypes = [C.c_void_p] * 7 + [C.c_double, C.c_void_p, C.c_void_p]
        _HS.hs_dyn_f.argtypes = [C.c_void_p] * 5
        _HS.hs_philox4x32_10.argtypes = [C.c_void_p] * 3
        _HS.hs_tvlqr_noise.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        _HS.hs_tvlqr.argtypes = [C.c_int64] + [C.c_void_p] * 5 + [C.c_int64, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                                                   C.c_uint32] + [C.c_void_p] * 7
        _HS.hs_tvlqr.restype = C.c_int64
    return _HS


def quat_axis_angle(axis, deg):
    a = np.asarray(axis, dtype=float)
    a = a / np.linalg.norm(a)
    h = math.radians(deg) / 2
    return np.concatenate([[math.cos(h)], a * math.sin(h)])


class Slew:
    """One trial's solver inputs (everything ts_alilqr_solve_batch needs)."""
    pass


def build_slew(kep, J, q0, qf, *, mjd=58155.0, igrf_date=2019.0, field_radius_m=6771000.0, t0=0.0, tf=5400.0, N_scope=5000,
               cutoff=50.0, dt=0.2, alpha=10.0, beta=1e3, t_final=None, w0=(0.0, 0.0, 0.0), rows_needed_only=True):
    """Field scoping -> cutoff -> fine table -> eigen-axis guess -> Bryson weights
    (TortoiseSat.jl:58-89,119-168), all with the oracle.  `t_final` overrides the
    gramian-derived horizon (used to make short test problems)."""
    s = Slew()
    kep = np.asarray(kep, dtype=float)
    if t_final is None:
        B0, _, _, _ = orc.magnetic_simulation(kep, GM, mjd, igrf_date, field_radius_m, t0, tf, N_scope)
        idx = orc.condition_based_time(orc.magnetic_gramian(B0, (tf - t0) / N_scope), cutoff)
        if idx == 0:
            raise RuntimeError("no cutoff")
        t_final = idx * (tf - t0) / N_scope
    N = int(math.floor((t_final - t0) / dt))
    B, _, _, _ = orc.magnetic_simulation(kep, GM, mjd, igrf_date, field_radius_m, t0, t_final, N)
    s.N, s.dt, s.t_final, s.B = N, dt, t_final, B
    s.index_scale = float(N)
    s.clock_rate = 1.0 / (tf - t0)
    s.J = np.asarray(J, dtype=float)
    s.x0 = np.concatenate([w0, q0, [t0]]).astype(float)
    s.xf = np.concatenate([[0, 0, 0], qf, [1.0]]).astype(float)
    nt = int(math.floor((t_final - t0) / dt + 1e-9)) + 1
    t = t0 + dt * np.arange(nt)
    w_g = np.zeros((nt, 3))
    q_g = np.zeros((nt, 4))
    L = orc.lib()
    L.orc_eigen_axis_slew(orc.P(orc.f64(s.x0[:7])), orc.P(orc.f64(s.xf[:7])), orc.P(t), nt, orc.P(w_g), orc.P(q_g))
    s.Qd, s.Qfd, s.Rd = np.zeros(8), np.zeros(8), np.zeros(3)
    L.orc_bryson_weights(orc.P(w_g), nt, orc.P(orc.f64(s.J)), dt, alpha, beta, orc.P(s.Qd), orc.P(s.Qfd), orc.P(s.Rd))
    s.w_guess, s.q_guess, s.t = w_g, q_g, t
    return s


def oracle_solve(slews, opts=None, nthreads=1, want_K=True):
    """orc_alilqr_solve_batch on a list of Slew objects -> (X list, U list, K list, outcomes)."""
    L = orc.lib()
    T = len(slews)
    N_i = np.array([s.N for s in slews], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(N_i)]).astype(np.int64)
    rows = np.array([s.B.shape[0] for s in slews], dtype=np.int64)
    B_offs = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    Ball = np.ascontiguousarray(np.concatenate([s.B for s in slews]))
    pk = lambda name, w: np.ascontiguousarray(np.stack([getattr(s, name).reshape(-1) for s in slews])).reshape(T, w)
    x0, xf, Jm, Qd, Qfd, Rd = pk("x0", 8), pk("xf", 8), pk("J", 9), pk("Qd", 8), pk("Qfd", 8), pk("Rd", 3)
    isc = np.array([s.index_scale for s in slews])
    cr = np.array([s.clock_rate for s in slews])
    X = np.zeros((int(offs[-1]), 8))
    U = np.zeros((int(offs[-1]), 3))
    K = np.zeros((int(offs[-1]), 24)) if want_K else None
    out = np.zeros(T, dtype=orc.OUTCOME_DTYPE)
    o = opts if opts is not None else orc.default_ilqr_opts()
    L.orc_alilqr_solve_batch(T, orc.P(N_i), orc.P(offs), orc.P(x0), orc.P(xf), orc.P(Jm), orc.P(Qd), orc.P(Qfd), orc.P(Rd),
                             orc.P(Ball), orc.P(B_offs), orc.P(rows), orc.P(isc), orc.P(cr), slews[0].dt, None, C.byref(o),
                             orc.P(X), orc.P(U), orc.P(K) if want_K else None, out.ctypes.data, nthreads)
    Xs = [X[offs[t]:offs[t + 1]] for t in range(T)]
    Us = [U[offs[t]:offs[t + 1] - 1] for t in range(T)]
    Ks = [K[offs[t]:offs[t + 1] - 1].reshape(-1, 3, 8) for t in range(T)] if want_K else None
    return Xs, Us, Ks, out


def hostsim_solve(s, opts=None, width=8):
    hs = hostsim()
    o = opts if opts is not None else orc.default_ilqr_opts()
    X = np.zeros((s.N, 8))
    U = np.zeros((s.N, 3))
    K = np.zeros((s.N, 24))
    out = np.zeros(1, dtype=orc.OUTCOME_DTYPE)
    B = np.ascontiguousarray(s.B)
    hs.hs_alilqr_solve_w(width, s.N, orc.P(orc.f64(s.x0)), orc.P(orc.f64(s.xf)), orc.P(orc.f64(s.J.reshape(-1))), orc.P(orc.f64(s.Qd)),
                       orc.P(orc.f64(s.Qfd)), orc.P(orc.f64(s.Rd)), orc.P(B), B.shape[0], s.index_scale, s.clock_rate, s.dt,
                       None, C.addressof(o), orc.P(X), orc.P(U), orc.P(K), out.ctypes.data)
    return X, U[:-1], K[:-1].reshape(-1, 3, 8), out[0]


def oracle_tvlqr_opts(noise_mode=0, seed=0, dt=0.2, dt_squared=1, R=7.5e3):
    """Oracle TvlqrOpts with the constants of TortoiseSat.jl:251-260 (R = 7.5e3) / monte_carlo.jl:216-227 (R = 0.5e3).
    Touches only the oracle (bench.py's CPU legs must not load the product library)."""
    o = orc.TvlqrOpts()
    o.dt, o.t0, o.dt_squared, o.noise_mode, o.seed = dt, 0.0, dt_squared, noise_mode, seed
    for i in range(6):
        o.Qd[i], o.Qfd[i] = 10.0, 1000.0
    for i in range(3):
        o.Rd[i] = R
    return o


def tvlqr_opts_pair(noise_mode=0, seed=0, dt=0.2, literal=0, dt_squared=1, R=7.5e3):
    """(oracle TvlqrOpts, product TvlqrOpts) with the same constants; GPU tests only (loads the product library)."""
    import tortoisesat.jl_b200 as tb
    g = tb.host.default_tvlqr_opts()
    g.dt, g.noise_mode, g.seed, g.literal_postproc, g.dt_squared = dt, noise_mode, seed, literal, dt_squared
    for i in range(3):
        g.Rd[i] = R
    o = oracle_tvlqr_opts(noise_mode, seed, dt, dt_squared, R)
    for i in range(6):
        assert (o.Qd[i], o.Qfd[i]) == (g.Qd[i], g.Qfd[i])
    return o, g


def oracle_tvlqr(s, X, U, x0_lqr, o, trial=0, noise=None, literal=0):
    """orc_attitude_simulation + orc_mc_slew_time for one Slew; X (N,8), U (N-1,3)."""
    L = orc.lib()
    d = orc.make_dyn(s.B, s.index_scale, s.clock_rate, s.J)
    o.tf = s.t_final
    N = s.N
    Xs, Us, dX, K = np.zeros((N, 8)), np.zeros((N, 3)), np.zeros((N, 6)), np.zeros((N - 1, 3, 6))
    X = orc.f64(X)
    U = orc.f64(U)
    ns = L.orc_attitude_simulation(C.byref(d), C.byref(o), N, orc.P(X), orc.P(U), orc.P(orc.f64(x0_lqr)),
                                   None if noise is None else orc.P(orc.f64(noise)), trial, orc.P(Xs), orc.P(Us), orc.P(dX), orc.P(K))
    slew = L.orc_mc_slew_time(orc.P(Xs), ns, orc.P(orc.f64(s.xf[3:7])), s.t_final, o.dt, 0.05, 0.08727, literal, trial + 1)
    return Xs[:ns], Us[:ns], dX[:ns], K, ns, slew


def hostsim_tvlqr(s, X, U, x0_lqr, g, trial=0, noise=None):
    hs = hostsim()
    N = s.N
    Xs, Us, dX, K = np.zeros((N, 8)), np.zeros((N, 3)), np.zeros((N, 6)), np.zeros((N, 3, 6))
    slew = np.zeros(1)
    B = np.ascontiguousarray(s.B)
    Up = np.zeros((N, 3))
    Up[:N - 1] = U
    ns = hs.hs_tvlqr(N, orc.P(orc.f64(X)), orc.P(Up), orc.P(orc.f64(x0_lqr)), orc.P(orc.f64(s.J.reshape(-1))), orc.P(B), B.shape[0],
                     s.index_scale, s.clock_rate, s.t_final, orc.P(orc.f64(s.xf[3:7])), trial, C.addressof(g),
                     None if noise is None else orc.P(orc.f64(noise)), orc.P(Xs), orc.P(Us), orc.P(dX), orc.P(K), orc.P(slew))
    return Xs[:ns], Us[:ns], dX[:ns], K[:N - 1], ns, slew[0]
