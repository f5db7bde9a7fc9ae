"""GPU: the reference's small building blocks as element-wise batch ops vs the CPU oracle
(SURVEY 8a rows a2, a3, a6, a7, a13, a14, a15, a17)."""
import ctypes as C
import math

import numpy as np
import pytest

import slew_setup as S
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GM = S.GM


def test_kep_eci_and_orbit_rhs(engine):
    rng = np.random.default_rng(2)
    n = 64
    kep = np.stack([rng.uniform(0, 0.2, n), rng.uniform(6600, 8000, n), rng.uniform(0, 180, n), rng.uniform(-360, 720, n),
                    rng.uniform(0, 360, n), rng.uniform(0, 360, n)], axis=1)
    kep[0] = [0, 6578, 96, 0, 0, 90]
    t0 = rng.uniform(0, 100, n)
    t0[0] = 0
    rv = engine.kep_eci_batch(kep, t0, GM)
    for i in range(n):
        ref, _ = orc.kep_eci(kep[i], t0[i], GM)
        assert np.max(np.abs(rv[i] - ref.reshape(-1)) / np.array([7000] * 3 + [8] * 3)) < 1e-12, i
    assert abs(rv[0, 0]) < 1e-11 and abs(rv[0, 4]) < 1e-14   # cosd(nu ~ 90): x of r and y of v vanish to round-off
    dx = engine.orbit_rhs_batch(rv)
    L = orc.lib()
    for i in range(n):
        o = np.zeros(6)
        L.orc_orbit_rhs(orc.P(np.ascontiguousarray(rv[i])), orc.P(o))
        assert np.allclose(dx[i], o, rtol=1e-13, atol=0)
    from tortoisesat.jl_b200 import host
    k = [0, 6578, 96, 0, 0, 450.0]
    out = host.kep_ECI(k, 0.0, GM)
    assert out.shape == (2, 3) and k[5] == 90.0      # mutates its argument like the reference (Q13)


def test_legendre_dlegendre(engine, orc):
    th = np.array([0.0, 1e-9, 0.3, 1.1, math.pi / 2, 2.9, math.pi - 1e-9, math.pi])
    for nmax in (10, 13):
        P, dP = engine.legendre_schmidt_batch(th, nmax)
        for i, t in enumerate(th):
            Po = orc.legendre_schmidt(t, nmax)
            dPo = orc.dlegendre_schmidt(t, Po)
            assert np.max(np.abs(P[i] - Po)) < 1e-13
            assert np.max(np.abs(dP[i] - dPo)) < 1e-12
    from tortoisesat.jl_b200 import host
    assert abs(np.sum(host.legendre(0.7, 13)[5, :6] ** 2) - 1) < 1e-13
    assert host.dlegendre(0.7, 13).shape == (14, 14)


@pytest.mark.parametrize("J", [S.J_1P, S.J_3U, np.array([[2e-3, 1e-4, 0], [1e-4, 3e-3, 2e-4], [0, 2e-4, 1e-3]])])
def test_dynamics_and_rk3(engine, J):
    rng = np.random.default_rng(4)
    n = 50
    Bt = rng.normal(size=(300, 3)) * 3e-5
    x = np.concatenate([rng.normal(size=(n, 3)) * 0.02, rng.normal(size=(n, 4)), rng.random((n, 1)) * 0.9], axis=1)
    u = rng.normal(size=(n, 3))
    d = orc.make_dyn(Bt, 300.0, 1.0 / 2400, J)
    L = orc.lib()
    for mode, fn in ((0, L.orc_deriv_function), (1, L.orc_gain_simulator)):
        dx = engine.dynamics_batch(mode, x, u, Bt, J, index_scale=300.0, clock_rate=1.0 / 2400)
        for i in range(n):
            o = np.zeros(8)
            fn(C.byref(d), orc.P(np.ascontiguousarray(x[i])), orc.P(np.ascontiguousarray(u[i])), orc.P(o))
            assert np.max(np.abs(dx[i] - o)) <= 1e-12 * max(1.0, np.max(np.abs(o))), (mode, i)
    BB = rng.normal(size=(n, 3)) * 3e-5
    dx7 = engine.dynamics_batch(2, x[:, :7], u, BB, J)
    for i in range(n):
        o = np.zeros(7)
        L.orc_attitude_dynamics(orc.P(np.ascontiguousarray(x[i, :7])), orc.P(np.ascontiguousarray(u[i])), orc.P(np.ascontiguousarray(BB[i])),
                                orc.P(np.ascontiguousarray(J)), orc.P(o))
        assert np.max(np.abs(dx7[i] - o)) <= 1e-12 * max(1.0, np.max(np.abs(o)))
    xn = engine.rk3_step_batch(x, u, Bt, J, 300.0, 1.0 / 2400, 0.2)
    for i in range(n):
        o = np.zeros(8)
        L.orc_rk3_step(C.byref(d), orc.P(np.ascontiguousarray(x[i])), orc.P(np.ascontiguousarray(u[i])), 0.2, orc.P(o))
        assert np.max(np.abs(xn[i] - o)) < 1e-13
        assert xn[i, 7] == o[7]                        # clock state accumulated bit-exactly (quirk Q1)


def test_empty_batches_and_argument_errors(engine):
    import tortoisesat.jl_b200 as tb
    assert engine.kep_eci_batch(np.zeros((0, 6))).shape == (0, 6)
    assert engine.orbit_rhs_batch(np.zeros((0, 6))).shape == (0, 6)
    with pytest.raises(tb.TortoiseError):
        engine.legendre_schmidt_batch([0.1], 14)
    with pytest.raises(tb.TortoiseError):
        engine.alilqr_solve_batch(N_i=[1], x0=np.zeros((1, 8)), xf=np.zeros((1, 8)), Jmat=np.eye(3).reshape(1, 9), Qd=np.zeros((1, 8)),
                                  Qfd=np.zeros((1, 8)), Rd=np.ones((1, 3)), B_eci=np.zeros((4, 3)), B_offs=[0], B_rows=[4],
                                  index_scale=[1.0], clock_rate=[1.0], dt=0.2)
    # empty Monte-Carlo
    from tortoisesat.jl_b200 import host
    cfg = host.default_mc_config(0)
    out, st = engine.monte_carlo_run(cfg, np.zeros((1, 6)), np.zeros(1, dtype=host.FIELD_OPTS_DTYPE), np.zeros((0, 8)), np.zeros((0, 8)),
                                     np.zeros((0, 9)))
    assert out.shape == (0,) and st.n_trials == 0
