"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int64)


def build(force=False):
    lib = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".inc"))]
    if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        r = subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return lib


class FieldOpts(C.Structure):
    _fields_ = [("GM", C.c_double), ("mjd", C.c_double), ("igrf_date", C.c_double), ("field_radius_m", C.c_double),
                ("t0", C.c_double), ("tf", C.c_double), ("N", C.c_int64)]


class Dyn(C.Structure):
    _fields_ = [("B_eci", C.c_void_p), ("B_rows", C.c_int64), ("index_scale", C.c_double), ("clock_rate", C.c_double),
                ("J", C.c_double * 9)]


class IlqrOpts(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("max_outer", "max_inner", "max_linesearch", "dJ_counter_limit", "stage_cost_dt",
                                         "goal_mask")] + \
               [(k, C.c_double) for k in ("cost_tol", "cost_tol_intermediate", "grad_tol", "grad_tol_intermediate",
                                          "constraint_tol", "penalty_initial", "penalty_scaling", "penalty_max", "dual_max",
                                          "ls_lower", "ls_upper", "bp_reg_increase", "bp_reg_max", "bp_reg_min", "bp_reg_fp",
                                          "max_cost_value", "max_state_value", "max_control_value", "u_max", "u_min")] + \
               [(k, C.c_int32) for k in ("a2_active_ge", "a3_grad_over_N", "a4_no_intermediate", "a5_dual_active_only",
                                         "a6_penalty_conditional", "a7_carry_cost")] + \
               [("constraint_decrease_ratio", C.c_double), ("k3_suspend_after", C.c_int32), ("k3_tail_share", C.c_int32),
                ("k3_early_factor", C.c_double), ("k3_pair", C.c_int32), ("k3_wide_occ", C.c_int32),
                ("quat_error", C.c_int32), ("k3_generic_inertia", C.c_int32)]


class Outcome(C.Structure):
    _fields_ = [("status", C.c_int32), ("outer_iters", C.c_int32), ("inner_iters", C.c_int32), ("ls_rollouts", C.c_int32),
                ("N", C.c_int64), ("J", C.c_double), ("c_max", C.c_double), ("t_final", C.c_double),
                ("slew_time", C.c_double), ("flops", C.c_double)]


class TvlqrOpts(C.Structure):
    _fields_ = [("dt", C.c_double), ("t0", C.c_double), ("tf", C.c_double), ("Qd", C.c_double * 6), ("Qfd", C.c_double * 6),
                ("Rd", C.c_double * 3), ("dt_squared", C.c_int32), ("noise_mode", C.c_int32), ("seed", C.c_uint64)]


class McConfig(C.Structure):
    """orc_mc_config (oracle.cpp): the oracle's own layout of the Monte-Carlo configuration."""
    _fields_ = [("n_trials", C.c_int64), ("shared_orbit", C.c_int32), ("run_tvlqr", C.c_int32), ("t0", C.c_double),
                ("tf", C.c_double), ("N_scope", C.c_int64), ("cutoff", C.c_double), ("dt", C.c_double), ("alpha", C.c_double),
                ("beta", C.c_double), ("eigen_axis_fix", C.c_int32), ("keep_trajectories", C.c_int32), ("ilqr", IlqrOpts),
                ("tvlqr", TvlqrOpts)]


OUTCOME_DTYPE = np.dtype([("status", "<i4"), ("outer_iters", "<i4"), ("inner_iters", "<i4"), ("ls_rollouts", "<i4"),
                          ("N", "<i8"), ("J", "<f8"), ("c_max", "<f8"), ("t_final", "<f8"), ("slew_time", "<f8"),
                          ("flops", "<f8")])


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_igrf12.argtypes = [C.c_double] * 4 + [dp]
        L.orc_igrf12syn.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, dp]
        L.orc_igrf12_batch.argtypes = [C.c_double, C.c_int64] + [C.c_void_p] * 6 + [C.c_int]
        L.orc_legendre_schmidt.argtypes = [C.c_double, C.c_int, C.c_void_p]
        L.orc_dlegendre_schmidt.argtypes = [C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_kep_eci.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.orc_orbit_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_sind.argtypes = [C.c_double]
        L.orc_sind.restype = C.c_double
        L.orc_cosd.argtypes = [C.c_double]
        L.orc_cosd.restype = C.c_double
        L.orc_magnetic_simulation.argtypes = [C.c_void_p, C.POINTER(FieldOpts), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_magnetic_gramian.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
        L.orc_cond_sym3.argtypes = [C.c_void_p]
        L.orc_cond_sym3.restype = C.c_double
        L.orc_condition_based_time.argtypes = [C.c_void_p, C.c_int64, C.c_double]
        L.orc_condition_based_time.restype = C.c_int64
        L.orc_qmult.argtypes = [C.c_void_p] * 3
        L.orc_qrot.argtypes = [C.c_void_p] * 3
        for f in ("orc_deriv_function", "orc_gain_simulator"):
            getattr(L, f).argtypes = [C.POINTER(Dyn), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_simulator.argtypes = [C.POINTER(Dyn), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_attitude_dynamics.argtypes = [C.c_void_p] * 5
        L.orc_rk3_step.argtypes = [C.POINTER(Dyn), C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_rk3_jacobian.argtypes = [C.POINTER(Dyn), C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_eigen_axis_slew.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_eigen_axis_slew_mode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_bryson_weights.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.orc_ilqr_default_opts.argtypes = [C.POINTER(IlqrOpts)]
        L.orc_alilqr_solve_batch.argtypes = [C.c_int64] + [C.c_void_p] * 13 + [C.c_double, C.c_void_p, C.POINTER(IlqrOpts),
                                                                               C.c_void_p, C.c_void_p, C.c_void_p,
                                                                               C.c_void_p, C.c_int]
        L.orc_attitude_simulation.argtypes = [C.POINTER(Dyn), C.POINTER(TvlqrOpts), C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.orc_attitude_simulation.restype = C.c_int64
        L.orc_mc_slew_time.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_int64]
        L.orc_mc_slew_time.restype = C.c_double
        L.orc_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.orc_tvlqr_noise.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_psiaki_pd_simulation.argtypes = [C.c_int64] + [C.c_void_p] * 5 + [C.c_double] * 3 + [C.c_void_p] * 3
        L.orc_attitude_dynamics_linear.argtypes = [C.c_void_p] * 6
        L.orc_mc_run.argtypes = [C.POINTER(McConfig)] + [C.c_void_p] * 9 + [C.c_int]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def P(a):
    return a.ctypes.data if a is not None else None


# ------------------------------------------------------------------ wrappers
def igrf12(date, r_m, lat, lon):
    o = (C.c_double * 3)()
    rc = lib().orc_igrf12(date, r_m, lat, lon, o)
    if rc:
        raise ValueError("igrf12 domain error %d" % rc)
    return np.array(o[:])


def igrf12syn(isv, date, itype, alt, colat, elong):
    o = (C.c_double * 4)()
    rc = lib().orc_igrf12syn(isv, date, itype, alt, colat, elong, o)
    if rc:
        raise ValueError("igrf12syn domain error %d" % rc)
    return np.array(o[:])


def igrf12_batch(date, r_m, lat, lon, nthreads=1):
    r_m, lat, lon = f64(r_m), f64(lat), f64(lon)
    n = r_m.shape[0]
    Bn, Be, Bd = np.empty(n), np.empty(n), np.empty(n)
    rc = lib().orc_igrf12_batch(date, n, P(r_m), P(lat), P(lon), P(Bn), P(Be), P(Bd), nthreads)
    return Bn, Be, Bd, rc


def legendre_schmidt(theta, nmax):
    Pm = np.zeros((nmax + 1, nmax + 1))
    lib().orc_legendre_schmidt(theta, nmax, P(Pm))
    return Pm


def dlegendre_schmidt(theta, Pm):
    nmax = Pm.shape[0] - 1
    dPm = np.zeros_like(Pm)
    lib().orc_dlegendre_schmidt(theta, nmax, P(f64(Pm)), P(dPm))
    return dPm


def kep_eci(kep, t0, GM):
    k = f64(np.array(kep, dtype=np.float64).copy())
    out = np.zeros(6)
    lib().orc_kep_eci(P(k), t0, GM, P(out))
    return out.reshape(2, 3), k


def magnetic_simulation(kep, GM, mjd, igrf_date, field_radius_m, t0, tf, N):
    o = FieldOpts(GM, mjd, igrf_date, field_radius_m, t0, tf, N)
    B = np.zeros((2 * N, 3))
    pos = np.zeros((2 * N + 1, 3))
    vel = np.zeros((2 * N + 1, 3))
    k = f64(np.array(kep, dtype=np.float64).copy())
    rc = lib().orc_magnetic_simulation(P(k), C.byref(o), P(B), P(pos), P(vel))
    return B, pos, vel, rc


def magnetic_gramian(B, dt):
    B = f64(B)
    G = np.zeros((B.shape[0], 3, 3))
    lib().orc_magnetic_gramian(P(B), B.shape[0], dt, P(G))
    return G


def condition_based_time(G, cutoff):
    G = f64(G)
    return int(lib().orc_condition_based_time(P(G), G.shape[0], cutoff))


def make_dyn(B_eci, index_scale, clock_rate, J):
    B_eci = f64(B_eci)
    d = Dyn()
    d.B_eci = B_eci.ctypes.data
    d.B_rows = B_eci.shape[0]
    d.index_scale = index_scale
    d.clock_rate = clock_rate
    d.J[:] = list(f64(J).ravel())
    d._keep = B_eci
    return d


def default_ilqr_opts():
    o = IlqrOpts()
    lib().orc_ilqr_default_opts(C.byref(o))
    return o


FIELD_OPTS_DTYPE = np.dtype([("GM", "<f8"), ("mjd", "<f8"), ("igrf_date", "<f8"), ("field_radius_m", "<f8"), ("t0", "<f8"),
                             ("tf", "<f8"), ("N", "<i8")])


def mc_config(n_trials, shared_orbit=True, run_tvlqr=True, t0=0.0, tf=2400.0, N_scope=5000, cutoff=30.0, dt=0.2, alpha=0.1, beta=1e3,
              noise_mode=2, seed=0, R_lqr=0.5e3):
    """Constants of src/monte_carlo.jl:37-78,169-171,216-227."""
    cfg = McConfig()
    cfg.n_trials, cfg.shared_orbit, cfg.run_tvlqr = n_trials, int(shared_orbit), int(run_tvlqr)
    cfg.t0, cfg.tf, cfg.N_scope, cfg.cutoff, cfg.dt, cfg.alpha, cfg.beta = t0, tf, N_scope, cutoff, dt, alpha, beta
    lib().orc_ilqr_default_opts(C.byref(cfg.ilqr))
    cfg.tvlqr.dt, cfg.tvlqr.t0, cfg.tvlqr.dt_squared, cfg.tvlqr.noise_mode, cfg.tvlqr.seed = dt, t0, 1, noise_mode, seed
    for i in range(6):
        cfg.tvlqr.Qd[i], cfg.tvlqr.Qfd[i] = 10.0, 1000.0
    for i in range(3):
        cfg.tvlqr.Rd[i] = R_lqr
    return cfg


def mc_run(cfg, kep6, fopts, x0, xf, Jmat, q_noise0=None, stream_id=None, nthreads=1):
    """The whole per-trial pipeline (orc_mc_run), OpenMP over trials.  Returns (outcomes, per-trial CPU seconds)."""
    n = int(cfg.n_trials)
    kep6 = f64(np.atleast_2d(kep6))
    fopts = np.ascontiguousarray(fopts, dtype=FIELD_OPTS_DTYPE)
    x0, xf, Jmat = f64(np.asarray(x0).reshape(n, 8)), f64(np.asarray(xf).reshape(n, 8)), f64(np.asarray(Jmat).reshape(n, 9))
    qn = None if q_noise0 is None else f64(np.asarray(q_noise0).reshape(n, 3))
    sid = None if stream_id is None else np.ascontiguousarray(stream_id, dtype=np.uint32)
    out = np.zeros(n, dtype=OUTCOME_DTYPE)
    secs = np.zeros(n)
    lib().orc_mc_run(C.byref(cfg), P(kep6), fopts.ctypes.data, P(x0), P(xf), P(Jmat), P(qn), None if sid is None else sid.ctypes.data,
                     out.ctypes.data, P(secs), int(nthreads))
    return out, secs
