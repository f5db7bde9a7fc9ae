#!/usr/bin/env python3
"""Golden-vector generator: a line-by-line numpy transliteration of the reference sources that ARE in
/root/reference (TEST INFRASTRUCTURE -- never imported by the product).

Every function below follows ONE reference function and cites its file:line; the arithmetic is written in the
reference's operation order with plain IEEE doubles (python floats / numpy float64), so the values this script
freezes into tests/golden/ref_fixtures.json are what the Julia scripts compute on the same inputs, up to the
last-bit differences of libm (sin/cos/acos) and of LLVM's muladd contraction.  Jacobians that the reference takes
with ForwardDiff (exact forward-mode derivatives) are taken here with the complex-step method (h = 1e-30, exact
to round-off).  `tests/test_ref_fixtures.py` checks the C++ oracle against the JSON (no GPU, no /root/reference
needed at test time); this script is re-run only when a fixture is added:

    python tests/golden/gen_ref_fixtures.py [/root/reference]

Section A  transliteration of in-tree reference code (kep_ECI.jl, OrbitPlotter.jl, magnetic_toolbox.jl, igrf.jl,
           legendre.jl, dlegendre.jl, eigen_axis_slew.jl, TortoiseSat.jl:157-168, monte_carlo.jl:165-176,237-262,
           DerivFunction.jl, gain_simulator.jl, simulator.jl, attitude_dynamics.jl, attitude_controller.jl)
Section B  an independent second implementation (numpy, dense 8-state, complex-step Jacobians) of the AL-iLQR
           specification frozen in SURVEY.md Appendix C.  TrajectoryOptimization.jl v0.1.2 is NOT in
           /root/reference, so this pins the oracle's AL-iLQR to a second reading of the same spec, not to Julia
           output ("parity unpinned" vs the real package stays true and is said so in DESIGN.md).
"""
import json
import math
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# ======================================================================================================
# Section A -- transliteration
# ======================================================================================================


def _cnorm(v):
    """norm(v) for real or complex-step vectors (no conjugation, so that the imaginary part carries d/dh)."""
    return np.sqrt(np.sum(v * v))


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def qmult(q1, q2):
    """src/qmult.jl:1-3 (scalar-first Hamilton product; only q[1] and q[2:4] are read)."""
    q1 = np.asarray(q1)
    q2 = np.asarray(q2)
    return np.concatenate([[q1[0] * q2[0] - np.sum(q1[1:4] * q2[1:4])], q1[0] * q2[1:4] + q2[0] * q1[1:4] + cross(q1[1:4], q2[1:4])])


def qrot(q, r):
    """src/qrot.jl:1-3."""
    return r + 2 * cross(q[1:4], cross(q[1:4], r) + q[0] * r)


def q_inv(q):
    """src/attitude_controller.jl:164-166."""
    return np.concatenate([[q[0]], -q[1:4]])


def hat(x):
    """src/magnetic_toolbox.jl:142-146."""
    return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])


def _trig_deg(fn, deg):
    import mpmath
    mpmath.mp.dps = 40
    return float(fn(mpmath.mpf(deg) * mpmath.pi / 180))


def sind(x):
    """Julia Base.sind: exact argument reduction in degrees (sind(180) == 0), then sin/cos of the reduced angle."""
    import mpmath
    rx = math.copysign(math.fmod(x, 360.0), x)
    arx = abs(rx)
    if rx == 0.0:
        return rx
    if arx < 45:
        return _trig_deg(mpmath.sin, rx)
    if arx <= 135:
        return math.copysign(_trig_deg(mpmath.cos, 90.0 - arx), rx)
    if arx == 180:
        return math.copysign(0.0, rx)
    if arx < 225:
        return _trig_deg(mpmath.sin, (180.0 - arx) * (1.0 if rx > 0 else -1.0))
    if arx <= 315:
        return -math.copysign(_trig_deg(mpmath.cos, 270.0 - arx), rx)
    return _trig_deg(mpmath.sin, rx - math.copysign(360.0, rx))


def cosd(x):
    """Julia Base.cosd (cosd(90) == 0 exactly)."""
    import mpmath
    rx = abs(math.fmod(x, 360.0))
    if rx <= 45:
        return _trig_deg(mpmath.cos, rx)
    if rx < 135:
        return _trig_deg(mpmath.sin, 90.0 - rx)
    if rx <= 225:
        return -_trig_deg(mpmath.cos, 180.0 - rx)
    if rx < 315:
        return _trig_deg(mpmath.sin, rx - 270.0)
    return _trig_deg(mpmath.cos, 360.0 - rx)


def R_z(angle):
    """src/kep_ECI.jl:37-42 (degrees)."""
    return np.array([[cosd(angle), sind(angle), 0], [-sind(angle), cosd(angle), 0], [0, 0, 1]])


def R_x(angle):
    """src/kep_ECI.jl:44-49 (degrees)."""
    return np.array([[1, 0, 0], [0, cosd(angle), sind(angle)], [0, -sind(angle), cosd(angle)]])


def kep_ECI(kep_elements, t0, GM):
    """src/kep_ECI.jl:1-35.  Mutates kep_elements[5] like the reference (:7-8)."""
    A = kep_elements
    A[5] = math.fmod(kep_elements[5] + t0 * math.sqrt(GM / kep_elements[1] ** 3), 360)
    E = np.zeros(101)
    E[0] = A[5] / 180 * math.pi
    for i in range(100):
        E[i + 1] = E[i] - (E[i] - A[0] * math.sin(E[i]) - A[5] / 180 * math.pi) / (1 - A[0] * math.cos(E[i]))
    nu = 2 * math.degrees(math.atan2(math.sqrt(1 + A[0]) * math.sin(E[-1] / 2), math.sqrt(1 - A[0]) * math.cos(E[-1] / 2)))
    r_c = A[1] * (1 - A[0] * math.cos(E[-1]))
    o = r_c * np.array([cosd(nu), sind(nu), 0])
    o_dot = math.sqrt(GM * A[1]) / r_c * np.array([-math.sin(E[-1]), math.sqrt(1 - A[0] ** 2) * math.cos(E[-1]), 0])
    M = R_z(-A[3]) @ R_x(-A[2]) @ R_z(-A[4])
    return np.stack([M @ o, M @ o_dot])


def OrbitPlotter(x):
    """src/OrbitPlotter.jl:1-52 (the GMST/lat/long block :18-22 is dead code; the literal "J2" term :40-42)."""
    r = x[0:3]
    v = x[3:6]
    GM = 3.986004418E14 * (1 / 1000) ** 3
    nr = np.linalg.norm(r)
    f_grav = GM / (nr ** 2) * -r / nr
    J2 = 0.0010826359
    f_J2 = np.array([J2 * r[0] / nr ** 7 * (6 * r[2] - 1.5 * (r[0] ** 2 + r[1] ** 2)),
                     J2 * r[1] / nr ** 7 * (6 * r[2] - 1.5 * (r[0] ** 2 + r[1] ** 2)),
                     J2 * r[2] / nr ** 7 * (3 * r[2] - 4.5 * (r[0] ** 2 + r[1] ** 2))])
    a = f_grav + f_J2
    return np.concatenate([v, a])


def euler_solve(u0, dt, nsteps):
    """DiffEqBase.solve(prob, dt=dt, adaptive=false, Euler()) (src/magnetic_toolbox.jl:52-54): u += dt*f(u)."""
    sol = np.zeros((6, nsteps + 1))
    u = np.array(u0, dtype=float)
    sol[:, 0] = u
    for k in range(nsteps):
        u = u + dt * OrbitPlotter(u)
        sol[:, k + 1] = u
    return sol


def Rz(theta):
    """src/magnetic_toolbox.jl:136-140 (radians)."""
    return np.array([[math.cos(theta), math.sin(theta), 0], [-math.sin(theta), math.cos(theta), 0], [0, 0, 1]])


def parse_table(name):
    text = open(os.path.join(REF, "src", "igrf12_coefs.jl")).read()
    m = re.search(r"const\s+%s\s*=\s*\[(.*?)\n\]" % name, text, re.S)
    rows = []
    for line in m.group(1).splitlines():
        line = line.split("#")[0].strip()
        if line:
            rows.append([float(t) for t in line.split()])
    return np.array(rows)


_G = _H = None


def legendre_schmidt(phi, n_max):
    """src/legendre.jl:254-292 with ph_term = false (1-based P[n+1,m+1] -> P[n,m])."""
    P = np.zeros((n_max + 1, n_max + 1))
    c = math.cos(phi)
    s = math.sqrt(1 - c ** 2)
    P[0, 0] = 1
    P[1, 0] = +c
    P[1, 1] = -s
    P[1, 1] *= -1
    for n in range(2, n_max + 1):
        for m in range(0, n):
            aux = (n - m) * (n + m)
            a_nm = math.sqrt(((2 * n - 1) * (2 * n - 1)) / aux)
            b_nm = math.sqrt(((n + m - 1) * (n - m - 1)) / aux)
            P[n, m] = a_nm * c * P[n - 1, m] - b_nm * P[n - 2, m]
        P[n, n] = +s * math.sqrt((2 * n - 1) / (2 * n)) * P[n - 1, n - 1]
    return P


def dlegendre_schmidt(phi, P):
    """src/dlegendre.jl:221-309 (reached through the Schmidt alias :411-419), ph_term = false."""
    rows = P.shape[0]
    dP = np.zeros_like(P)
    phi = math.fmod(phi, 2 * math.pi)
    if phi < 0:
        phi += 2 * math.pi
    fact = -1 if phi > math.pi else 1
    for n in range(1, rows):
        for m in range(0, n + 1):
            if m == 0:
                aux = math.sqrt(n * (n + 1) / 2)
                a_nm = +0.5 * aux
                b_nm = -0.5 * aux
                dP[n, 0] = -a_nm * P[n, 1] + b_nm * P[n, 1]
            elif m == 1:
                a_nm = +0.5 * math.sqrt(2 * n * (n + 1))
                b_nm = -0.5 * math.sqrt((n + 2) * (n - 1))
                dP[n, 1] = a_nm * P[n, 0] + b_nm * (P[n, 2] if n >= 2 else 0.0)   # P is 14x14 in the reference: P[2,3] = 0
            elif n != m:
                a_nm = +0.5 * math.sqrt((n + m) * (n - m + 1))
                b_nm = -0.5 * math.sqrt((n + m + 1) * (n - m))
                dP[n, m] = a_nm * P[n, m - 1] + b_nm * P[n, m + 1]
            else:
                a_nm = +0.5 * math.sqrt((n + m) * (n - m + 1))
                dP[n, m] = a_nm * P[n, m - 1]
            dP[n, m] *= fact
    return dP


def igrf12(date, r, lam, Om):
    """src/igrf.jl:70-274 (geocentric).  Returns [north, east, down] in nT."""
    global _G, _H
    if _G is None:
        _G, _H = parse_table("G_igrf12"), parse_table("H_igrf12")
    G, H = _G, _H
    theta = math.pi / 2 - lam
    phi = Om if Om >= 0 else 2 * math.pi + Om
    r = r / 1000
    idx = int(math.floor((date - 1900) * 0.2 + 1)) if date < 2020 else 24
    epoch = 1900 + (idx - 1) * 5
    dt_ = date - epoch
    n_max = 10 if epoch < 1995 else 13
    P = legendre_schmidt(theta, n_max)
    dP = dlegendre_schmidt(theta, P)
    a = 6371.2
    sin_p, cos_p = math.sin(1 * phi), math.cos(1 * phi)
    ratio = a / r
    fact = ratio
    dVr = dVt = dVp = 0.0
    kg = kh = 0
    c0 = idx + 2 - 1          # 1-based column idx+2 -> 0-based
    for n in range(1, n_max + 1):
        aux_r = aux_t = aux_p = 0.0
        Gnm_e0 = G[kg, c0]
        if date < 2015:
            dG = (G[kg, c0 + 1] - Gnm_e0) / 5
        else:
            dG = G[kg, 26]
        Gnm = Gnm_e0 + dG * dt_
        kg += 1
        aux_r += -(n + 1) / r * Gnm * P[n, 0]
        aux_t += Gnm * dP[n, 0]
        sin_m1, sin_m2 = 0.0, -sin_p
        cos_m1, cos_m2 = 1.0, +cos_p
        for m in range(1, n + 1):
            sin_m = 2 * cos_p * sin_m1 - sin_m2
            cos_m = 2 * cos_p * cos_m1 - cos_m2
            Gnm_e0 = G[kg, c0]
            Hnm_e0 = H[kh, c0]
            if date < 2015:
                dG = (G[kg, c0 + 1] - Gnm_e0) / 5
                dH = (H[kh, c0 + 1] - Hnm_e0) / 5
            else:
                dG = G[kg, 26]
                dH = H[kh, 26]
            Gnm = Gnm_e0 + dG * dt_
            Hnm = Hnm_e0 + dH * dt_
            kg += 1
            kh += 1
            GcHs = Gnm * cos_m + Hnm * sin_m
            GsHc = Gnm * sin_m - Hnm * cos_m
            aux_r += -(n + 1) / r * GcHs * P[n, m]
            aux_t += GcHs * dP[n, m]
            aux_p += (-m * GsHc * dP[n, m]) if theta == 0 else (-m * GsHc * P[n, m])
            sin_m2, sin_m1 = sin_m1, sin_m
            cos_m2, cos_m1 = cos_m1, cos_m
        fact *= ratio
        aux_r *= fact
        aux_p *= fact
        aux_t *= fact
        dVr += aux_r
        dVp += aux_p
        dVt += aux_t
    dVr *= a
    dVp *= a
    dVt *= a
    x = +1 / r * dVt
    y = (-1 / r * dVp) if theta == 0 else (-1 / (r * math.sin(theta)) * dVp)
    z = dVr
    return np.array([x, y, z])


def magnetic_simulation(Kep, GM, MJD, R_E, alt, t0, tf, N, igrf_date=2019):
    """src/magnetic_toolbox.jl:33-106.  Returns B (2N x 3, Tesla), pos, vel (3 x (2N+1))."""
    pos_0 = kep_ECI(np.array(Kep, dtype=float), t0, GM)
    u0 = np.concatenate([pos_0[0, :], pos_0[1, :]])
    dt = (tf - t0) / N
    sol = euler_solve(u0, dt, 2 * N)
    pos = sol[0:3, :]
    vel = sol[3:6, :]
    step = (tf - t0) / N
    lat = np.zeros(2 * N)
    lon = np.zeros(2 * N)
    GMST = np.zeros(2 * N)
    for i in range(2 * N):
        t = t0 + i * step                                                     # t0:(tf-t0)/N:2*tf
        GMST[i] = (280.4606 + 360.9856473 * (t / 24 / 60 / 60 + MJD) - 51544.5) / 180 * math.pi
        pe = Rz(GMST[i]) @ pos[:, i]
        lat[i] = math.asin(pe[2] / np.linalg.norm(pe))
        lon[i] = math.atan2(pe[1], pe[0])
    B = np.zeros((2 * N, 3))
    NED_to_ENU = np.array([[0, 1, 0], [1, 0, 0], [0, 0, -1]])
    for i in range(2 * N - 1):
        b = igrf12(igrf_date, (alt + R_E) * 1000, lat[i], lon[i]) / 1.e9
        R_ENU_to_XYZ = np.array([[-math.sin(lon[i]), -math.sin(lat[i]) * math.cos(lon[i]), math.cos(lat[i]) * math.cos(lon[i])],
                                 [math.cos(lon[i]), -math.sin(lat[i]) * math.sin(lon[i]), math.cos(lat[i]) * math.sin(lon[i])],
                                 [0, math.cos(lat[i]), math.sin(lat[i])]])
        B[i, :] = ((Rz(GMST[i]).T @ R_ENU_to_XYZ) @ NED_to_ENU) @ b
    return B, pos, vel


def magnetic_gramian(B_N, dt):
    """src/magnetic_toolbox.jl:1-12 (first term without dt)."""
    n = B_N.shape[0]
    G = np.zeros((n, 3, 3))
    G[0] = hat(B_N[0]) @ hat(B_N[0]).T
    for i in range(1, n):
        G[i] = G[i - 1] + hat(B_N[i]) @ hat(B_N[i]).T * dt
    return G


def condition_based_time(B_gram, cutoff):
    """src/magnetic_toolbox.jl:14-31 (cond = 2-norm condition number; first index below the cutoff, 1-based, else 0)."""
    for i in range(B_gram.shape[0]):
        s = np.linalg.svd(B_gram[i], compute_uv=False)
        c = np.inf if s[-1] == 0 else s[0] / s[-1]
        if c < cutoff:
            return i + 1
    return 0


def julia_range(t0, dt, tf):
    """t0:dt:tf as Julia builds it: length = floor((tf-t0)/dt) + 1 (with the range code's guard against the
    quotient landing a hair below an integer), elements t0 + k*dt."""
    n = int(math.floor((tf - t0) / dt + 1e-9)) + 1
    return t0 + dt * np.arange(n)


def eigen_axis_slew(x0, xf, t):
    """src/eigen_axis_slew.jl:1-38.  NOTE :16 builds qmult([q2;-q2[2:4]], q1): a 7-vector whose entries 5..7 are
    never read by qmult, i.e. the literal product is qmult(q2, q1) -- NOT conj(q2) (x) q1."""
    q1 = np.asarray(x0[3:7], dtype=float)
    q2 = np.asarray(xf[3:7], dtype=float)
    q_e = qmult(np.concatenate([q2, -q2[1:4]]), q1)
    theta_f = 2 * math.acos(q_e[0])
    axis = -q_e[1:4] / (math.sin(theta_f / 2))
    alpha = math.pi / t[-1]
    theta = theta_f * 1 / 2 * (np.ones(len(t)) - np.cos(alpha * t))
    d_theta = list(np.diff(theta) / (t[1] - t[0]))
    d_theta.append(d_theta[-1])
    w_guess = np.zeros((len(t), 3))
    for i in range(len(t)):
        w_guess[i, :] = d_theta[i] * axis
    q_guess = np.zeros((len(t), 4))
    for i in range(len(t)):
        q_guess[i, :] = qmult(q1, np.concatenate([[math.cos(theta[i] / 2)], axis * math.sin(theta[i] / 2)]))
    return w_guess, q_guess


def bryson_weights(w_guess, J, dt, alpha, beta):
    """src/TortoiseSat.jl:157-168 (alpha = 10) == src/monte_carlo.jl:165-176 (alpha = 0.1, dt = time_step[i])."""
    X13 = w_guess.T
    w_max = np.max(np.abs(X13))
    tau_max = np.max(J @ np.diff(X13, axis=1) / dt)
    m_max = tau_max / 1.e-5 * 1.e2
    Qd = np.zeros(8)
    Qfd = np.zeros(8)
    Qd[0:3] = alpha / w_max ** 2
    Qfd[0:3] = (alpha / w_max ** 2) * 10
    Qd[3:7] = alpha * beta
    Qfd[3:7] = alpha * beta * 10
    Rd = np.ones(3) * (1 / m_max ** 2)
    return Qd, Qfd, Rd, w_max, tau_max, m_max


class Globals:
    """The untyped globals the reference's dynamics read: B_ECI, N, p.J, tf, t0 (src/DerivFunction.jl:28,41,44)."""

    def __init__(self, B_ECI, N, J, tf, t0=0.0):
        self.B_ECI, self.N, self.J, self.tf, self.t0 = np.asarray(B_ECI, dtype=float), N, np.asarray(J, dtype=float), tf, t0
        self.Jinv = np.linalg.inv(self.J)


def _row(g, t):
    i = int(math.floor((t * g.N + 1).real)) if isinstance(t, complex) or np.iscomplexobj(t) else int(math.floor(t * g.N + 1))
    return g.B_ECI[i - 1, :]


def DerivFunction(g, x, u):
    """src/DerivFunction.jl:1-48."""
    omega = x[0:3]
    q = x[3:7] / _cnorm(x[3:7])
    t = x[7]
    q_dot = 0.5 * qmult(q, np.concatenate([[0], omega]))
    B_B = qrot(q, _row(g, t))
    tau_c = cross(u[0:3] * 1.e-2, B_B)
    omega_dot = g.Jinv @ (tau_c - cross(omega, g.J @ omega))
    return np.concatenate([omega_dot, q_dot, [1 / (g.tf - g.t0)]])


def gain_simulator(g, x, u):
    """src/gain_simulator.jl:1-53 (u/100 instead of u*1e-2, quirk Q8)."""
    omega = x[0:3]
    q = x[3:7] / _cnorm(x[3:7])
    t = x[7]
    q_dot = 0.5 * qmult(q, np.concatenate([[0], omega]))
    B_B = qrot(q, _row(g, t) * 1)
    tau_c = cross(u[0:3] / 100, B_B)
    omega_dot = g.Jinv @ (tau_c - cross(omega, g.J @ omega))
    return np.concatenate([omega_dot, q_dot, [1 / (g.tf - g.t0)]])


def simulator(g, x, u, noise9):
    """src/simulator.jl:1-42 with the three random draws supplied: noise9 = [randn(3); randn(3,1); rand(3)]."""
    omega_noise = noise9[0:3] * (.38 * math.pi / 180) ** 2
    omega = x[0:3] + omega_noise
    q_noise = noise9[3:6] * (1 * math.pi / 180) ** 2
    th = np.linalg.norm(q_noise)
    r_noise = q_noise / th
    q = qmult(x[3:7] / np.linalg.norm(x[3:7]), np.concatenate([[math.cos(th / 2)], r_noise * math.sin(th / 2)]))
    t = x[7]
    q_dot = 0.5 * qmult(q, np.concatenate([[0], omega]))
    B_N_noise = noise9[6:9] * (1E-5) ** 2
    B_B = qrot(q, _row(g, t) + B_N_noise)
    tau_c = cross(u[0:3] / 100, B_B)
    omega_dot = g.Jinv @ (tau_c - cross(omega, g.J @ omega))
    return np.concatenate([omega_dot, q_dot, [1 / (g.tf - g.t0)]])


def attitude_dynamics(x, u, B_B, J):
    """src/attitude_dynamics.jl:2-24."""
    omega = x[0:3]
    q = x[3:7] / np.linalg.norm(x[3:7])
    q_dot = 0.5 * qmult(q, np.concatenate([[0], omega]))
    tau_c = cross(u[0:3], B_B)
    omega_dot = np.linalg.inv(J) @ (tau_c - cross(omega, J @ omega))
    return np.concatenate([omega_dot, q_dot])


def rk3(f, dt):
    """src/attitude_controller.jl:178-187 (== TrajectoryOptimization rk3, ZOH)."""
    def fd(x, u):
        k1 = f(x, u) * dt
        k2 = f(x + k1 / 2, u) * dt
        k3 = f(x - k1 + 2 * k2, u) * dt
        return x + (k1 + 4 * k2 + k3) / 6
    return fd


def rk4(f, dt):
    """src/attitude_controller.jl:122-132; f(stage, x, u) so that `simulator` can draw per-stage noise (quirk Q7)."""
    def fd(x, u):
        k1 = f(0, x, u) * dt
        k2 = f(1, x + k1 / 2, u) * dt
        k3 = f(2, x + k2 / 2, u) * dt
        k4 = f(3, x + k3, u) * dt
        return x + (k1 + 2 * k2 + 2 * k3 + k4) / 6
    return fd


def rk4_aug(f, n, m):
    """src/attitude_controller.jl:134-145 composed with f_augmented! (:148-150): dt = S[end]^2 (quirk Q6);
    the augmented derivative is zero in the control and dt slots."""
    def f_aug(S):
        return np.concatenate([f(S[0:n], S[n:n + m]), np.zeros(m + 1, dtype=S.dtype)])

    def fd(S):
        dt = S[-1] ** 2
        k1 = f_aug(S) * dt
        k2 = f_aug(S + k1 / 2) * dt
        k3 = f_aug(S + k2 / 2) * dt
        k4 = f_aug(S + k3) * dt
        return S + (k1 + 2 * k2 + 2 * k3 + k4) / 6
    return fd


def jacobian_cs(fun, S, h=1e-30):
    """ForwardDiff.jacobian (exact forward mode) by the complex-step method."""
    n = len(S)
    Jd = np.zeros((n, n))
    for j in range(n):
        Sc = np.array(S, dtype=complex)
        Sc[j] += 1j * h
        Jd[:, j] = np.imag(fun(Sc)) / h
    return Jd


def attitude_lqr(g, dt, X_lqr, U_lqr, Q_lqr, R_lqr, Qf_lqr):
    """src/attitude_controller.jl:95-119 (Jacobians) + :50-93 (projection + Riccati).  X_lqr 8 x N, U_lqr 3 x (N-1)."""
    n, m, N = X_lqr.shape[0], U_lqr.shape[0], X_lqr.shape[1]
    fd_aug_gains = rk4_aug(lambda x, u: gain_simulator(g, x, u), n, m)
    Aq = np.zeros((n - 1, n - 1, N))
    Bq = np.zeros((n - 1, m, N))
    for k in range(N - 1):
        Sd = np.concatenate([X_lqr[:, k], U_lqr[:, k], [dt]])
        Jd = jacobian_cs(fd_aug_gains, Sd)
        Aq[:, :, k] = Jd[0:n - 1, 0:n - 1]
        Bq[:, :, k] = Jd[0:n - 1, n:n + m]
    A = np.zeros((6, 6, N))
    B = np.zeros((6, 3, N))
    for k in range(N - 1):
        qk = X_lqr[3:7, k]
        sk, vk = qk[0], qk[1:4]
        qn = X_lqr[3:7, k + 1]
        sn, vn = qn[0], qn[1:4]
        Gk = np.vstack([-vk, sk * np.eye(3) + hat(vk)])
        Gn = np.vstack([-vn, sn * np.eye(3) + hat(vn)])
        perm_Gn = np.zeros((6, 7))
        perm_Gk = np.zeros((7, 6))
        perm_Gn[0:3, 0:3] = np.eye(3)
        perm_Gn[3:6, 3:7] = Gn.T
        perm_Gk[0:3, 0:3] = np.eye(3)
        perm_Gk[3:7, 3:6] = Gk
        A[:, :, k] = perm_Gn @ Aq[:, :, k] @ perm_Gk
        B[:, :, k] = perm_Gn @ Bq[:, :, k]
    S = np.zeros((6, 6, N))
    K = np.zeros((3, 6, N - 1))
    S[:, :, N - 1] = Qf_lqr
    for k in range(N - 2, -1, -1):
        K[:, :, k] = np.linalg.inv(R_lqr + B[:, :, k].T @ S[:, :, k + 1] @ B[:, :, k]) @ (B[:, :, k].T @ S[:, :, k + 1] @ A[:, :, k])
        AK = A[:, :, k] - B[:, :, k] @ K[:, :, k]
        S[:, :, k] = Q_lqr + K[:, :, k].T @ R_lqr @ K[:, :, k] + AK.T @ S[:, :, k + 1] @ AK
    return K


def attitude_simulation(g, X_lqr, U_lqr, dt_lqr, x0_lqr, t0, tf, Q_lqr, R_lqr, Qf_lqr, noise=None):
    """src/attitude_controller.jl:1-48 with integration = :rk4.  noise: (N_sim-1, 4, 9) draws of `simulator`, or None
    for a noise-free replay (simulator with zero noise is NOT gain_simulator: 0/0 in r_noise -- so noise = None uses
    gain_simulator as f!, which is what a zero-noise run means)."""
    dt = dt_lqr
    t_sim = julia_range(t0, dt, tf)
    if len(t_sim) > X_lqr.shape[1]:
        t_sim = julia_range(t0, dt, tf - dt)
    N_sim = len(t_sim)
    K = attitude_lqr(g, dt, X_lqr, U_lqr, Q_lqr, R_lqr, Qf_lqr)
    X_sim = np.zeros((8, N_sim))
    X_sim[:, 0] = x0_lqr
    U_sim = np.zeros((3, N_sim))
    dX = np.zeros((6, N_sim))
    for k in range(N_sim - 1):
        dX[0:3, k] = X_sim[0:3, k] - X_lqr[0:3, k]
        dX[3:6, k] = qmult(q_inv(X_lqr[3:7, k]), X_sim[3:7, k])[1:4]
        U_sim[:, k] = U_lqr[:, k] - K[:, :, k] @ dX[:, k]
        if noise is None:
            fd = rk4(lambda s, x, u: gain_simulator(g, x, u), dt)
        else:
            fd = rk4(lambda s, x, u, _k=k: simulator(g, x, u, noise[_k, s]), dt)
        X_sim[:, k + 1] = fd(X_sim[:, k], U_sim[:, k])
    return X_sim, U_sim, dX, K


def mc_postprocess(sim_states_i, q_final, t_final_i, time_step_i, slew_limits, trial_i_1based=None):
    """src/monte_carlo.jl:237-262 for one trial.  trial_i_1based = None uses column j for the rate (the evident
    intent); an integer reproduces the literal `sim_states[i][1:3,i]` of :247 (quirk Q12)."""
    slew_time = t_final_i
    for j in range(1, sim_states_i.shape[1] + 1):
        col = j if trial_i_1based is None else trial_i_1based
        omega_norm = np.linalg.norm(sim_states_i[0:3, col - 1])
        eq = qmult(q_inv(q_final), sim_states_i[3:7, j - 1])
        error_angle = 2 * math.acos(min(eq[0], 1.))
        if j > 10 and omega_norm < slew_limits[0] and error_angle < slew_limits[1] and slew_time == t_final_i:
            slew_time = time_step_i * j
    return slew_time, (1 if slew_time == t_final_i else 0)


def attitude_dynamics_linear(x, u, x_linear, B_B, J):
    """src/attitude_dynamics.jl:26-48."""
    omega = x[0:3]
    q = x[3:7] / np.linalg.norm(x[3:7])
    q_dot = 0.5 * qmult(q, np.concatenate([[0], x_linear[3:6]]))
    tau_c = cross(u[0:3], B_B)
    omega_dot = np.linalg.inv(J) @ (tau_c - cross(omega, J @ omega))
    return np.concatenate([omega_dot, q_dot])


def psiaki_controller(C_1, C_2, J, q, w_bar, B_meas, m_limit=None):
    """src/comparison/psiaki_dynamics.jl:1-26 (the m_limit clamp is commented out in the reference)."""
    T_req = -(C_1 * w_bar + C_2 * np.linalg.inv(J) @ q[1:4])
    return cross(B_meas, T_req) / (np.linalg.norm(B_meas) ** 2)


def rk4_psiaki(f, x, dt, u, B_B, J):
    """src/comparison/psiaki_dynamics.jl:63-73."""
    f1 = f(x, u, B_B, J)
    f2 = f(x + .5 * f1 * dt, u, B_B, J)
    f3 = f(x + .5 * f2 * dt, u, B_B, J)
    f4 = f(x + f3 * dt, u, B_B, J)
    return x + 1 / 6 * (f1 + 2 * f2 + 2 * f3 + f4) * dt


def psiaki_pd_simulation(x0, w_guess, q_guess, B_ECI, J, dt, C_1, C_2):
    """src/comparison/psiaki2005.jl:116-164.  w_guess 3 x N, q_guess 4 x N, B_ECI 3 x N (columns = steps)."""
    N = w_guess.shape[1]
    x = np.zeros((7, N))
    x[:, 0] = x0
    m_all = np.zeros((3, N))
    qbar = np.zeros((4, N))
    xd = attitude_dynamics(x[:, 0], np.zeros(3), B_ECI[:, 0], J)          # :124
    x[:, 1] = x[:, 0] + dt * xd                                            # :125
    for i in range(1, N - 1):                                              # for i = 2:length(t)-1
        B_meas = qrot(q_inv(x[3:7, i]), B_ECI[:, i])                       # :141
        w_bar = w_guess[:, i] - x[0:3, i]                                  # :145
        qbar[:, i] = qmult(x[3:7, i], q_guess[:, i])                       # :152
        m = psiaki_controller(C_1, C_2, J, qbar[:, i], w_bar, B_meas)      # :157
        x[:, i + 1] = rk4_psiaki(attitude_dynamics, x[:, i], dt, m, B_meas, J)   # :161
        x[3:7, i + 1] = x[3:7, i + 1] / np.linalg.norm(x[3:7, i + 1])      # :162
        m_all[:, i] = m
    return x, m_all, qbar


# ======================================================================================================
# Section B -- AL-iLQR, second implementation of SURVEY.md Appendix C (dense 8-state, numpy)
# ======================================================================================================
ALILQR_DEFAULTS = dict(max_outer=20, max_inner=50, max_linesearch=20, dJ_counter_limit=10, stage_cost_dt=0, goal_mask=0x7F,
                       cost_tol=1e-4, cost_tol_intermediate=1e-3, grad_tol=1e-5, grad_tol_intermediate=1e-5, constraint_tol=1e-3,
                       penalty_initial=1.0, penalty_scaling=10.0, penalty_max=1e8, dual_max=1e8, ls_lower=1e-8, ls_upper=10.0,
                       bp_reg_increase=1.6, bp_reg_max=1e8, bp_reg_min=1e-8, bp_reg_fp=10.0, max_cost_value=1e8,
                       max_state_value=1e8, max_control_value=1e8, u_max=1.0, u_min=-1.0,
                       # assumption registry: 0 = the frozen default, 1 = the named alternative
                       a2_active_ge=0, a3_grad_over_N=0, a4_no_intermediate=0, a5_dual_active_only=0, a6_penalty_conditional=0,
                       a7_carry_cost=0, constraint_decrease_ratio=0.25)


def quaternion_error(X1, X2):
    """src/quaternion_toolbox.jl:63-75 (MRP of the error quaternion; 7 entries, the last one stays zero)."""
    dx = np.zeros(7)
    dx[0:3] = X1[0:3] - X2[0:3]
    q_e = qmult(q_inv(X2[3:7]), X1[3:7])
    dx[3:6] = q_e[1:4] / (1 + q_e[0])
    return dx


def perm_Gk(x):
    """src/quaternion_toolbox.jl:22-35: perm_Gk (8 x 7) = [I3 0; 0 G(q); 0 0] with G(q) = [-v'; s I + hat(v)] of the raw
    state quaternion; perm_Gn of the same state is its transpose."""
    s_, v = x[3], x[4:7]
    P = np.zeros((8, 7))
    P[0:3, 0:3] = np.eye(3)
    P[3:7, 3:6] = np.vstack([-v, s_ * np.eye(3) + hat(v)])
    return P


def alilqr_solve(g, x0, xf, Qd, Qfd, Rd, N, dt, opts=None, U0=None):
    """AL-iLQR per SURVEY.md App. C on the problem of src/TortoiseSat.jl:145-146,169,178-199.  Returns X (N x 8),
    U ((N-1) x 3), K ((N-1) x 3 x 8), info dict."""
    o = dict(ALILQR_DEFAULTS)
    if opts:
        o.update(opts)
    n, m = 8, 3
    sc = dt if o["stage_cost_dt"] else 1.0
    step = rk3(lambda x, u: DerivFunction(g, x, u), dt)
    gm = np.array([(o["goal_mask"] >> i) & 1 for i in range(n)], dtype=bool)
    Q, Qf, R = np.diag(Qd), np.diag(Qfd), np.diag(Rd)
    U = np.zeros((N - 1, m)) if U0 is None else np.array(U0, dtype=float)
    X = np.zeros((N, n))
    X[0] = x0
    for k in range(N - 1):
        X[k + 1] = step(X[k], U[k])
    lam_b = np.zeros((N - 1, 6))
    mu_b = np.full((N - 1, 6), o["penalty_initial"])
    lam_g = np.zeros(n)
    mu_g = np.full(n, o["penalty_initial"])

    def cons(Uk):
        return np.concatenate([Uk - o["u_max"], o["u_min"] - Uk])

    def active(c, lam):
        return ((c >= 0.0) if o["a2_active_ge"] else (c > 0.0)) | (lam > 0.0)

    def al_cost(Xt, Ut):
        Jc, cmax = 0.0, 0.0
        for k in range(N - 1):
            e = Xt[k] - xf
            Jc += (0.5 * e @ Q @ e + 0.5 * Ut[k] @ R @ Ut[k]) * sc
            c = cons(Ut[k])
            act = active(c, lam_b[k])
            Jc += lam_b[k] @ c + 0.5 * np.sum(np.where(act, mu_b[k], 0.0) * c * c)
            cmax = max(cmax, float(np.max(np.maximum(c, 0.0))))
        e = Xt[N - 1] - xf
        Jc += 0.5 * e @ Qf @ e
        Jc += np.sum(np.where(gm, lam_g * e + 0.5 * mu_g * e * e, 0.0))
        if gm.any():
            cmax = max(cmax, float(np.max(np.abs(e[gm]))))
        return Jc, cmax

    def jac(xk, uk):
        S0 = np.concatenate([xk, uk])
        Jd = np.zeros((n, n + m))
        for j in range(n + m):
            Sc = np.array(S0, dtype=complex)
            Sc[j] += 1e-30j
            Jd[:, j] = np.imag(step(Sc[0:n], Sc[n:n + m])) / 1e-30
        return Jd[:, 0:n], Jd[:, n:]

    rho = drho = 0.0

    def reg_inc():
        nonlocal rho, drho
        drho = max(drho * o["bp_reg_increase"], o["bp_reg_increase"])
        rho = max(rho * drho, o["bp_reg_min"])

    def reg_dec():
        nonlocal rho, drho
        drho = min(drho / o["bp_reg_increase"], 1.0 / o["bp_reg_increase"])
        rho = rho * drho * (1.0 if rho * drho > o["bp_reg_min"] else 0.0)

    K = np.zeros((N - 1, m, n))
    d = np.zeros((N - 1, m))
    qa = bool(o.get("quat_error", 0))   # quaternion-aware variant (monte_carlo.jl:158,192): 7-dim error state
    status, outer, inner_total, ls_total = 1, 0, 0, 0
    J, c_max, c_max_prev = 0.0, 0.0, np.inf
    J_carry = None
    inner_per_outer = []
    for oi in range(1, o["max_outer"] + 1):
        outer = oi
        last = oi == o["max_outer"]
        inter = (not last) and not o["a4_no_intermediate"]
        ctol = o["cost_tol_intermediate"] if inter else o["cost_tol"]
        gtol = o["grad_tol_intermediate"] if inter else o["grad_tol"]
        rho = drho = 0.0
        J_prev, _ = al_cost(X, U)
        if o["a7_carry_cost"] and J_carry is not None:
            J_prev = J_carry
        J = J_prev
        dJ_zero, abort, it_used = 0, False, 0
        for it in range(1, o["max_inner"] + 1):
            inner_total += 1
            it_used = it
            AB = [jac(X[k], U[k]) for k in range(N - 1)]
            # ---- backward pass (App. C step 3)
            restarts = 0
            while True:
                e = X[N - 1] - xf
                Sxx = Qf + np.diag(np.where(gm, mu_g, 0.0))
                Sx = Qf @ e + np.where(gm, lam_g + mu_g * e, 0.0)
                if qa:   # quaternion_expansion(cost, xN), quaternion_toolbox.jl:40-52
                    EN = perm_Gk(X[N - 1])
                    Sxx = EN.T @ Sxx @ EN
                    Sx = EN.T @ Sx
                dV1 = dV2 = 0.0
                ok = True
                for k in range(N - 2, -1, -1):
                    A, B = AB[k]
                    c = cons(U[k])
                    act = active(c, lam_b[k])
                    Imu = np.where(act, mu_b[k], 0.0)
                    lx = sc * (Q @ (X[k] - xf))
                    lxx = sc * Q
                    if qa:   # quaternion_expansion(cost, x, u) and the dynamics in the same coordinates
                        E0, E1 = perm_Gk(X[k]), perm_Gk(X[k + 1])
                        A, B = E1.T @ A @ E0, E1.T @ B
                        lx, lxx = E0.T @ lx, E0.T @ lxx @ E0
                    lu = sc * (R @ U[k]) + (lam_b[k, 0:3] + Imu[0:3] * c[0:3]) - (lam_b[k, 3:6] + Imu[3:6] * c[3:6])
                    luu = sc * R + np.diag(Imu[0:3] + Imu[3:6])
                    Qx = lx + A.T @ Sx
                    Qu = lu + B.T @ Sx
                    Qxx = lxx + A.T @ Sxx @ A
                    Quu = luu + B.T @ Sxx @ B
                    Qux = B.T @ Sxx @ A
                    Qr = 0.5 * (Quu + Quu.T) + rho * np.eye(m)
                    try:
                        L = np.linalg.cholesky(Qr)
                    except np.linalg.LinAlgError:
                        ok = False
                        break
                    Kk = -np.linalg.solve(L.T, np.linalg.solve(L, Qux))
                    dk = -np.linalg.solve(L.T, np.linalg.solve(L, Qu))
                    K[k, :, :Kk.shape[1]], d[k] = Kk, dk
                    Sx = Qx + Kk.T @ Quu @ dk + Kk.T @ Qu + Qux.T @ dk
                    Sxx = Qxx + Kk.T @ Quu @ Kk + Kk.T @ Qux + Qux.T @ Kk
                    Sxx = 0.5 * (Sxx + Sxx.T)
                    dV1 += dk @ Qu
                    dV2 += 0.5 * dk @ Quu @ dk
                if ok:
                    break
                reg_inc()
                restarts += 1
                if rho > o["bp_reg_max"] or restarts > 200:
                    break
            if not ok:
                status, abort = 3, True
                break
            reg_dec()
            # ---- forward pass / line search (App. C step 4)
            alpha, z, Jn, it_ls, accepted = 1.0, -1.0, np.inf, 0, True
            Xb, Ub = None, None
            while (z <= o["ls_lower"] or z > o["ls_upper"]) and (Jn >= J_prev):
                if it_ls > o["max_linesearch"]:
                    accepted = False
                    Jn, _ = al_cost(X, U)
                    reg_inc()
                    rho += o["bp_reg_fp"]
                    break
                ls_total += 1
                Xb = np.zeros_like(X)
                Ub = np.zeros_like(U)
                Xb[0] = x0
                okr = True
                for k in range(N - 1):
                    dxk = quaternion_error(Xb[k], X[k]) if qa else Xb[k] - X[k]
                    Ub[k] = U[k] + K[k][:, :len(dxk)] @ dxk + alpha * d[k]
                    Xb[k + 1] = step(Xb[k], Ub[k])
                    if not (np.max(np.abs(Xb[k + 1])) < o["max_state_value"]) or not (np.max(np.abs(Ub[k])) < o["max_control_value"]):
                        okr = False
                        break
                if not okr:
                    it_ls += 1
                    alpha /= 2.0
                    continue
                Jn, _ = al_cost(Xb, Ub)
                expected = -alpha * (dV1 + alpha * dV2)
                z = (J_prev - Jn) / expected if expected > 0 else -1.0
                it_ls += 1
                alpha /= 2.0
            if accepted:
                X, U = Xb, Ub
            if not (Jn == Jn):
                status, abort = 4, True
                break
            if Jn > o["max_cost_value"]:
                J, status, abort = Jn, 2, True
                break
            dJ = abs(Jn - J_prev)
            J_prev = Jn
            J = Jn
            dJ_zero = dJ_zero + 1 if dJ == 0 else 0
            grad = float(np.sum(np.max(np.abs(d) / (np.abs(U) + 1.0), axis=1))) / (N if o["a3_grad_over_N"] else (N - 1))
            if (0.0 < dJ < ctol) or grad < gtol or dJ_zero > o["dJ_counter_limit"]:
                break
        inner_per_outer.append(it_used)
        J, c_max = al_cost(X, U)
        J_carry = J
        if abort:
            break
        # ---- outer update: duals (A5), penalties (A6)
        for k in range(N - 1):
            c = cons(U[k])
            act = active(c, lam_b[k])
            lnew = np.clip(lam_b[k] + mu_b[k] * c, -o["dual_max"], o["dual_max"])
            if o["a5_dual_active_only"]:
                lnew = np.where(act, lnew, lam_b[k])
            lam_b[k] = np.maximum(0.0, lnew)
        e = X[N - 1] - xf
        lam_g = np.where(gm, np.clip(lam_g + mu_g * e, -o["dual_max"], o["dual_max"]), lam_g)
        grow = (not o["a6_penalty_conditional"]) or (c_max > o["constraint_decrease_ratio"] * c_max_prev)
        if grow:
            mu_b = np.minimum(mu_b * o["penalty_scaling"], o["penalty_max"])
            mu_g = np.where(gm, np.minimum(mu_g * o["penalty_scaling"], o["penalty_max"]), mu_g)
        c_max_prev = c_max
        if c_max < o["constraint_tol"]:
            status = 0
            break
    info = dict(status=status, outer_iters=outer, inner_iters=inner_total, ls_rollouts=ls_total, J=float(J), c_max=float(c_max),
                inner_per_outer=inner_per_outer)
    return X, U, K.copy(), info


# ======================================================================================================
# fixtures
# ======================================================================================================
def L(a):
    return np.asarray(a, dtype=float).tolist()


def main():
    rng = np.random.default_rng(20260118)
    GM = 3.986004418E14 * (1 / 1000) ** 3
    fx = {"_generator": "tests/golden/gen_ref_fixtures.py (numpy transliteration of /root/reference/src; see its docstring)"}

    # ---- kep_ECI + OrbitPlotter
    fx["kep_ECI"] = []
    for kep, t0 in [([0, 6578, 96, 0, 0, 90], 0.0), ([0, 6771, 96.6, 123.4, 0, 271.8], 0.0), ([0.1, 7000, 51.6, 40, 30, 75], 100.0),
                    ([0.01, 6900, 28.5, 359.0, 181.0, 12.0], 2400.0)]:
        k = np.array(kep, dtype=float)
        rv = kep_ECI(k, t0, GM)
        fx["kep_ECI"].append(dict(kep=kep, t0=t0, GM=GM, rv=L(rv), kep6_after=float(k[5])))
    fx["OrbitPlotter"] = []
    for _ in range(4):
        x = np.concatenate([rng.normal(size=3) * 4000 + np.array([0, 0, 5000.0]), rng.normal(size=3) * 5])
        fx["OrbitPlotter"].append(dict(x=L(x), dx=L(OrbitPlotter(x))))

    # ---- igrf12 / legendre / dlegendre (transliteration of the vendored copy; tables parsed from the reference)
    fx["igrf12"] = []
    for date, r, lat, lon in [(2019, 6771000.0, 0.0, 0.0), (2019, 6771000.0, 0.5, -2.0), (2019, 6771000.0, -1.2, 3.0),
                              (2017.5, 6871200.0, 0.9, 1.0), (1987.25, 7000000.0, -0.3, -0.7), (2019, 6771000.0, math.pi / 2, 0.3),
                              (2019, 6771000.0, -math.pi / 2, 0.3), (2003.4, 6671200.0, 1.1, -3.1), (1900.0, 6471200.0, 0.01, 0.0)]:
        fx["igrf12"].append(dict(date=date, r=r, lat=lat, lon=lon, B=L(igrf12(date, r, lat, lon))))
    fx["legendre"] = []
    for th in [0.0, 0.3, 1.2, math.pi / 2, 2.9, math.pi]:
        P = legendre_schmidt(th, 13)
        fx["legendre"].append(dict(theta=th, P=L(P), dP=L(dlegendre_schmidt(th, P))))

    # ---- magnetic_simulation / gramian / cutoff: TortoiseSat.jl config (R_E = 6178 typo :31, alt 400, 1P) ...
    ms = []
    R_E_p, alt = 6371.0, 400.0                                   # p.R_E, global alt (magnetic_toolbox.jl:44,81)
    B0, pos0, vel0 = magnetic_simulation([0, 400 + 6178, 96, 0, 0, 90], GM, 58155.0, R_E_p, alt, 0.0, 5400.0, 5000)
    G0 = magnetic_gramian(B0, 5400.0 / 5000)
    idx0 = condition_based_time(G0, 50)
    t_final = idx0 * (5400.0 - 0.0) / 5000
    N1 = int(math.floor((t_final - 0.0) / 0.2))
    B1, pos1, vel1 = magnetic_simulation([0, 400 + 6178, 96, 0, 0, 90], GM, 58155.0, R_E_p, alt, 0.0, t_final, N1)
    ms.append(dict(name="TortoiseSat.jl:58-89", kep=[0, 6578, 96, 0, 0, 90], GM=GM, mjd=58155.0, igrf_date=2019.0,
                   field_radius_m=(alt + R_E_p) * 1000, t0=0.0, tf=5400.0, N=5000, cutoff=50, tf_index=idx0, t_final=t_final,
                   N_fine=N1, scope_rows={str(i): L(B0[i]) for i in (0, 1, 2, 287, 4999, 9998, 9999)},
                   scope_pos={str(i): L(pos0[:, i]) for i in (0, 1, 5000, 10000)},
                   scope_vel={str(i): L(vel0[:, i]) for i in (0, 1, 5000, 10000)},
                   gram={str(i): L(G0[i]) for i in (0, 1, 287, 288)},
                   fine_rows={str(i): L(B1[i]) for i in (0, 1, 2, 89, 90, N1, 2 * N1 - 2, 2 * N1 - 1)},
                   fine_pos={str(i): L(pos1[:, i]) for i in (1, N1)}))
    # ... and the monte_carlo.jl orbit (a = 6771, i = 96.6, tf = 2400, cutoff 30) with two RAAN/anomaly draws
    for raan, nu in [(0.0, 90.0), (211.7, 33.3)]:
        Bm, pm, vm = magnetic_simulation([0, 6771.0, 96.6, raan, 0, nu], GM, 58155.0, 6371.0, 400.0, 0.0, 2400.0, 5000)
        Gm = magnetic_gramian(Bm, 2400.0 / 5000)
        im = condition_based_time(Gm, 30)
        ms.append(dict(name="monte_carlo.jl:122-140", kep=[0, 6771.0, 96.6, raan, 0, nu], GM=GM, mjd=58155.0, igrf_date=2019.0,
                       field_radius_m=6771000.0, t0=0.0, tf=2400.0, N=5000, cutoff=30, tf_index=im, t_final=im * 2400.0 / 5000,
                       scope_rows={str(i): L(Bm[i]) for i in (0, 1, 2, im - 1, 9998)},
                       scope_pos={str(i): L(pm[:, i]) for i in (0, 1, 10000)},
                       gram={str(i): L(Gm[i]) for i in (0, 1, im - 1)}))
    fx["magnetic_simulation"] = ms
    fx["cond"] = [dict(G=L(Gm[i]), cond=float(np.linalg.cond(Gm[i]))) for i in (1, 5, 50, im - 1, 3000)]

    # ---- eigen_axis_slew + Bryson: config 1 (xf identity), the MC script pair (x0 identity), both non-identity
    J1P, J1U, J3U = np.diag([0.0001041667] * 3), np.diag([0.00125] * 3), np.diag([0.020833, 0.020833, 0.0041666])
    r101 = np.array([1, 0, 1]) / math.sqrt(2)
    q45 = np.concatenate([[cosd(45)], r101 * sind(45)])
    qF = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])
    qa = rng.normal(size=4)
    qa /= np.linalg.norm(qa)
    qb = rng.normal(size=4)
    qb /= np.linalg.norm(qb)
    fx["eigen_axis_slew"] = []
    for name, q0, qf, tfin, Jm, alpha in [("TortoiseSat.jl:119-168", q45, np.array([1.0, 0, 0, 0]), t_final, J1P, 10.0),
                                          ("monte_carlo.jl:108-176", np.array([1.0, 0, 0, 0]), qF, 326.4, J1U, 0.1),
                                          ("both non-identity (bench ensemble shape)", qa, qF, 408.96, J1U, 0.1),
                                          ("both non-identity, 3U", qa, qb, 120.0, J3U, 10.0)]:
        t = julia_range(0.0, 0.2, tfin)
        x0 = np.concatenate([[0, 0, 0], q0])
        xf = np.concatenate([[0, 0, 0], qf])
        w, q = eigen_axis_slew(x0, xf, t)
        Qd, Qfd, Rd, wm, tm, mm = bryson_weights(w, Jm, 0.2, alpha, 1e3)
        fx["eigen_axis_slew"].append(dict(name=name, x0=L(x0), xf=L(xf), t_final=tfin, dt=0.2, nt=len(t), J=L(Jm), alpha=alpha, beta=1e3,
                                          w_rows={str(i): L(w[i]) for i in (0, 1, len(t) // 2, len(t) - 2, len(t) - 1)},
                                          q_rows={str(i): L(q[i]) for i in (0, 1, len(t) // 2, len(t) - 1)},
                                          Qd=L(Qd), Qfd=L(Qfd), Rd=L(Rd), w_max=wm, tau_max=tm, m_max=mm))

    # ---- dynamics, rk3, Jacobians on the config-1 field table
    g = Globals(B1, N1, J1P, 5400.0, 0.0)
    fx["dynamics"] = dict(B_rows=L(B1[:128]), N=N1, tf=5400.0, t0=0.0, J=L(J1P), cases=[])
    for _ in range(5):
        x = np.concatenate([rng.normal(size=3) * 0.01, rng.normal(size=4) * (1 + 0.05 * rng.normal()), [rng.uniform(0, 0.05)]])
        u = rng.uniform(-1.5, 1.5, size=3)
        n9 = np.concatenate([rng.normal(size=6), rng.random(3)])
        step = rk3(lambda xx, uu: DerivFunction(g, xx, uu), 0.2)
        S0 = np.concatenate([x, u])
        Jd = np.zeros((8, 11))
        for j in range(11):
            Sc = np.array(S0, dtype=complex)
            Sc[j] += 1e-30j
            Jd[:, j] = np.imag(step(Sc[:8], Sc[8:])) / 1e-30
        fx["dynamics"]["cases"].append(dict(x=L(x), u=L(u), noise9=L(n9), DerivFunction=L(DerivFunction(g, x, u)),
                                            gain_simulator=L(gain_simulator(g, x, u)), simulator=L(simulator(g, x, u, n9)),
                                            attitude_dynamics=L(attitude_dynamics(x[:7], u, B1[3], J3U)), B_B=L(B1[3]), J_ad=L(J3U),
                                            rk3=L(step(x, u)), rk3_A=L(Jd[:, :8]), rk3_B=L(Jd[:, 8:])))

    # ---- AL-iLQR (Section B) on short slews, then the TVLQR replay of the first one (attitude_controller.jl)
    fx["alilqr"] = []
    tv = None
    ax2 = np.array([0.3, -1, 0.5]) / np.linalg.norm([0.3, -1, 0.5])
    cases = [("5 deg about [1,0,1], 1P, 60 s", np.concatenate([[cosd(2.5)], r101 * sind(2.5)]), 60.0, J1P, 10.0, {}),
             ("same, stage cost x dt (A1)", np.concatenate([[cosd(2.5)], r101 * sind(2.5)]), 60.0, J1P, 10.0, {"stage_cost_dt": 1}),
             ("2 deg about [0.3,-1,0.5], 1P, 30 s", np.concatenate([[cosd(1.0)], ax2 * sind(1.0)]), 30.0, J1P, 10.0, {}),
             ("same, literal goal on the clock state (Q2)", np.concatenate([[cosd(1.0)], ax2 * sind(1.0)]), 30.0, J1P, 10.0, {"goal_mask": 0xFF})]
    if "--slow" in sys.argv:   # ~25 min of pure Python: a slew that is infeasible in 30 s and runs all 20 x 50 iterations
        cases.append(("20 deg about [1,0,1], 1P, 30 s", np.concatenate([[cosd(10)], r101 * sind(10)]), 30.0, J1P, 10.0, {}))
    for name, q0, tfin, Jm, alpha, opts in cases:
        Nn = int(math.floor(tfin / 0.2))
        Bn, _, _ = magnetic_simulation([0, 6578, 96, 0, 0, 90], GM, 58155.0, 6371.0, 400.0, 0.0, tfin, Nn)
        gg = Globals(Bn, Nn, Jm, 5400.0, 0.0)
        x0 = np.concatenate([[0, 0, 0], q0, [0.0]])
        xf = np.concatenate([[0, 0, 0], [1.0, 0, 0, 0], [1.0]])
        t = julia_range(0.0, 0.2, tfin)
        w, _q = eigen_axis_slew(x0[:7], xf[:7], t)
        Qd, Qfd, Rd, *_ = bryson_weights(w, Jm, 0.2, alpha, 1e3)
        X, U, K, info = alilqr_solve(gg, x0, xf, Qd, Qfd, Rd, Nn, 0.2, opts)
        print("alilqr", name, info, flush=True)
        fx["alilqr"].append(dict(name=name, kep=[0, 6578, 96, 0, 0, 90], t_final=tfin, N=Nn, J=L(Jm), alpha=alpha, x0=L(x0), xf=L(xf), opts=opts,
                                 Qd=L(Qd), Qfd=L(Qfd), Rd=L(Rd), info=info, X_rows={str(i): L(X[i]) for i in (0, 1, Nn // 2, Nn - 1)},
                                 U_rows={str(i): L(U[i]) for i in (0, 1, Nn // 2, Nn - 2)}, K0=L(K[0]), U_absmax=float(np.max(np.abs(U)))))
        if tv is None:
            tv = (gg, X, U, x0, xf, tfin, Nn)

    if "--slow" not in sys.argv:
        # counts-only record of the slow case (a run of this script with --slow: 776 inner iterations, 8917 rollouts)
        q0 = np.concatenate([[cosd(10)], r101 * sind(10)])
        fx["alilqr"].append(dict(name="20 deg about [1,0,1], 1P, 30 s (counts only; --slow regenerates the rows)", kep=[0, 6578, 96, 0, 0, 90],
                                 t_final=30.0, N=150, J=L(J1P), alpha=10.0, x0=L(np.concatenate([[0, 0, 0], q0, [0.0]])),
                                 xf=[0, 0, 0, 1.0, 0, 0, 0, 1.0], opts={}, Qd=None, Qfd=None, Rd=None,
                                 info=dict(status=1, outer_iters=20, inner_iters=776, ls_rollouts=8917, J=1883234.8722420842,
                                           c_max=0.06295110632761834,
                                           inner_per_outer=[16, 14, 14, 11, 6, 24, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 50, 41, 50]),
                                 X_rows={}, U_rows={}, K0=None, U_absmax=None))
    gg, X, U, x0, xf, tfin, Nn = tv
    Q_lqr = np.diag([10.0] * 6)
    Qf_lqr = Q_lqr * 100
    fx["tvlqr"] = []
    qn = rng.normal(size=3) * (math.pi / 180) ** 2
    th = np.linalg.norm(qn)
    x0_lqr = np.zeros(8)
    x0_lqr[0:3] = x0[0:3]
    x0_lqr[3:7] = qmult(x0[3:7], np.concatenate([[math.cos(th / 2)], qn / th * math.sin(th / 2)]))
    for name, Rl, with_noise in [("R = 7.5e3 (TortoiseSat.jl:260), noise-free", 7.5e3, False),
                                 ("R = 0.5e3 (monte_carlo.jl:227), simulator noise", 0.5e3, True)]:
        noise = None
        if with_noise:
            noise = np.concatenate([rng.normal(size=(Nn, 4, 6)), rng.random((Nn, 4, 3))], axis=2)
        Xs, Us, dX, Kt = attitude_simulation(gg, X.T, U.T, 0.2, x0_lqr, 0.0, tfin, Q_lqr, np.eye(3) * Rl, Qf_lqr, noise)
        slew, fail = mc_postprocess(Xs, xf[3:7], tfin, 0.2, (0.05, 0.08727))
        fx["tvlqr"].append(dict(name=name, alilqr_case=0, X_lqr=L(X), U_lqr=L(U), x0_lqr=L(x0_lqr), R=Rl, t_final=tfin, N=Nn,
                                noise=None if noise is None else L(noise), N_sim=Xs.shape[1],
                                X_sim_rows={str(i): L(Xs[:, i]) for i in (0, 1, 2, Xs.shape[1] // 2, Xs.shape[1] - 1)},
                                U_sim_rows={str(i): L(Us[:, i]) for i in (0, 1, Xs.shape[1] - 2)},
                                dX_rows={str(i): L(dX[:, i]) for i in (0, 1, Xs.shape[1] - 2)},
                                K_rows={str(i): L(Kt[:, :, i]) for i in (0, 1, Nn // 2, Nn - 2)}, slew_time=slew, fail=fail))
    # MC post-processing rule on synthetic replays (one that settles, one that does not)
    fx["mc_postprocess"] = []
    for settle in (True, False):
        ns = 200
        Xs = np.zeros((8, ns))
        for j in range(ns):
            ang = 1.0 * math.exp(-j / 20.0) if settle else 1.0
            Xs[0:3, j] = [0.2 * math.exp(-j / 15.0) if settle else 0.2, 0, 0]
            Xs[3:7, j] = qmult(qF, np.array([math.cos(ang / 2), math.sin(ang / 2), 0, 0]))
        s1, f1 = mc_postprocess(Xs, qF, 40.0, 0.2, (0.05, 0.08727))
        s2, f2 = mc_postprocess(Xs, qF, 40.0, 0.2, (0.05, 0.08727), trial_i_1based=7)
        fx["mc_postprocess"].append(dict(X_sim=L(Xs.T), q_final=L(qF), t_final=40.0, time_step=0.2, slew_time=s1, fail=f1,
                                         slew_time_literal_i7=s2, fail_literal_i7=f2))

    out = os.path.join(HERE, "ref_fixtures.json")
    json.dump(fx, open(out, "w"))
    print("wrote", out, os.path.getsize(out), "bytes")
    add_psiaki_section()


def add_quat_section():
    """Quaternion-aware AL-iLQR fixtures; `--only quat` regenerates just this section."""
    path = os.path.join(HERE, "ref_fixtures.json")
    fx = json.load(open(path))
    GM = 3.986004418E14 * (1 / 1000) ** 3
    out = []
    for c in (fx["alilqr"][0], fx["alilqr"][2]):
        tfin, Nn = c["t_final"], c["N"]
        Bn, _, _ = magnetic_simulation(c["kep"], GM, 58155.0, 6371.0, 400.0, 0.0, tfin, Nn)
        gg = Globals(Bn, Nn, np.array(c["J"]).reshape(3, 3), 5400.0, 0.0)
        opts = dict(c["opts"], quat_error=1)
        X, U, K, info = alilqr_solve(gg, np.array(c["x0"]), np.array(c["xf"]), np.array(c["Qd"]), np.array(c["Qfd"]), np.array(c["Rd"]),
                                     Nn, 0.2, opts)
        print("alilqr quat", c["name"], info, flush=True)
        out.append(dict(c, name=c["name"] + " -- quaternion_error / quaternion_expansion", opts=opts, info=info,
                        X_rows={str(i): L(X[i]) for i in (0, 1, Nn // 2, Nn - 1)}, U_rows={str(i): L(U[i]) for i in (0, 1, Nn // 2, Nn - 2)},
                        K0=L(K[0]), U_absmax=float(np.max(np.abs(U)))))
    fx["alilqr_quat"] = out
    json.dump(fx, open(path, "w"))


def add_psiaki_section():
    """Comparison-controller fixtures; `--only psiaki` regenerates just this section."""
    out = os.path.join(HERE, "ref_fixtures.json")
    fx = json.load(open(out))
    rng = np.random.default_rng(77)
    GM = 3.986004418E14 * (1 / 1000) ** 3
    J1U = np.diag([0.00125] * 3)
    J3U = np.diag([0.020833, 0.020833, 0.0041666])
    cases = []
    for name, Jm, tfin, C1, C2 in [("psiaki2005.jl:116-164 constants (1U, C_1 = 1e-6, C_2 = 1e-9)", J1U, 60.0, 1E-6, 1E-9),
                                   ("3U, stronger gains", J3U, 40.0, 2E-4, 3E-6)]:
        t = julia_range(0.0, 0.2, tfin)
        N = len(t)
        B, _, _ = magnetic_simulation([0, 6771.0, 96.6, 40.0, 0, 120.0], GM, 58155.0, 6371.0, 400.0, 0.0, tfin, N)
        B_ECI = B[:N].T                                                   # psiaki2005.jl:73-74
        q_0 = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])         # :99
        x0g = np.concatenate([[0, 0, 0], q_0])
        xfg = np.concatenate([[0, 0, 0], [1.0, 0, 0, 0]])
        wg, qg = eigen_axis_slew(x0g, xfg, t)                              # :118
        x0 = np.concatenate([wg[0], qg[0]])                                # x = [w_guess; q_guess] (:119), column 1
        X, M, Qe = psiaki_pd_simulation(x0, wg.T, qg.T, B_ECI, Jm, 0.2, C1, C2)
        cases.append(dict(name=name, J=L(Jm), dt=0.2, C_1=C1, C_2=C2, N=N, x0=L(x0), w_guess=L(wg), q_guess=L(qg), B_eci=L(B_ECI.T),
                          X_rows={str(i): L(X[:, i]) for i in (0, 1, 2, N // 2, N - 1)},
                          M_rows={str(i): L(M[:, i]) for i in (1, 2, N // 2, N - 2)},
                          q_err_rows={str(i): L(Qe[:, i]) for i in (1, N // 2, N - 2)}))
    x = np.concatenate([rng.normal(size=3) * 0.01, rng.normal(size=4)])
    xl = rng.normal(size=7) * 0.02
    u = rng.normal(size=3) * 0.1
    Bb = rng.normal(size=3) * 3e-5
    fx["psiaki"] = dict(cases=cases, linear=dict(x=L(x), u=L(u), x_linear=L(xl), B_B=L(Bb), J=L(J3U),
                                                 dx=L(attitude_dynamics_linear(x, u, xl, Bb, J3U))))
    json.dump(fx, open(out, "w"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "quat":
        add_quat_section()
        sys.exit(0)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "psiaki":
        add_psiaki_section()
    else:
        main()
