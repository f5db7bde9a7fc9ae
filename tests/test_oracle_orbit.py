"""CPU: pins the oracle's orbit / field-table layer against the survey's provisional
known answers (produced by an independent Python restatement) and
against internal identities."""
import math

import numpy as np

GM = 3.986004418E14 * (1 / 1000) ** 3


def test_sind_cosd_exact_zeros(orc):
    L = orc.lib()
    assert L.orc_cosd(90.0) == 0.0 and L.orc_cosd(270.0) == 0.0 and L.orc_cosd(-90.0) == 0.0
    assert L.orc_sind(180.0) == 0.0 and L.orc_sind(360.0) == 0.0
    for x in np.linspace(-720, 720, 97):
        assert abs(L.orc_sind(x) - math.sin(math.radians(x))) < 1e-14
        assert abs(L.orc_cosd(x) - math.cos(math.radians(x))) < 1e-14


def test_kep_eci_config1(orc):
    rv, kep = orc.kep_eci([0, 6578, 96, 0, 0, 90], 0.0, GM)
    assert rv[0, 0] == 0.0 or abs(rv[0, 0]) < 1e-12
    assert abs(rv[0, 1] - (-687.588231374625)) < 1e-9
    assert abs(rv[0, 2] - 6541.965027732502) < 1e-9
    assert abs(rv[1, 0] - (-7.784342809549734)) < 1e-12
    assert abs(rv[1, 1]) < 1e-12 and abs(rv[1, 2]) < 1e-12


def test_kep_eci_eccentric_consistency(orc):
    # vis-viva and angular momentum for an eccentric orbit (treating Kep[6] as mean anomaly, quirk Q13)
    a, e = 7000.0, 0.1
    rv, _ = orc.kep_eci([e, a, 51.6, 40.0, 30.0, 75.0], 0.0, GM)
    r, v = rv[0], rv[1]
    assert abs(np.dot(v, v) - GM * (2 / np.linalg.norm(r) - 1 / a)) < 1e-9
    h = np.linalg.norm(np.cross(r, v))
    assert abs(h - math.sqrt(GM * a * (1 - e * e))) < 1e-7
    assert abs(np.cross(r, v)[2] / h - math.cos(math.radians(51.6))) < 1e-12


def test_config1_scoping_and_fine_pass(orc):
    kep = [0, 6578, 96, 0, 0, 90]
    B, pos, vel, rc = orc.magnetic_simulation(kep, GM, 58155.0, 2019.0, 6771000.0, 0.0, 5400.0, 5000)
    assert rc == 0
    assert np.allclose(B[0], [-1.40943518e-06, 8.09895982e-06, -4.63002796e-05], rtol=2e-8)
    assert np.all(B[-1] == 0.0)
    G = orc.magnetic_gramian(B, 5400.0 / 5000)
    idx = orc.condition_based_time(G, 50)
    assert idx == 288
    t_final = idx * 5400.0 / 5000
    assert abs(t_final - 311.04) < 1e-12
    N = int(math.floor(t_final / 0.2))
    assert N == 1555
    Bf, posf, _, _ = orc.magnetic_simulation(kep, GM, 58155.0, 2019.0, 6771000.0, 0.0, t_final, N)
    exp = np.array([[-1.4094351791427393e-06, 8.0989598203005258e-06, -4.6300279590337312e-05],
                    [-1.3954016661905592e-06, 8.0976700450329386e-06, -4.6301470674009637e-05],
                    [-1.3813676889614736e-06, 8.0963798385175541e-06, -4.6302658637865654e-05]])
    assert np.max(np.abs(Bf[:3] - exp) / np.linalg.norm(exp, axis=1, keepdims=True)) < 1e-10
    assert np.allclose(posf[1], [-1.5570688022390502, -687.58823137462502, 6541.9650277325018], rtol=1e-13)


def test_gramian_identity_and_cond(orc):
    rng = np.random.default_rng(3)
    B = rng.normal(size=(50, 3)) * 3e-5
    dt = 1.08
    G = orc.magnetic_gramian(B, dt)
    acc = np.zeros((3, 3))
    for i, b in enumerate(B):
        H = np.dot(b, b) * np.eye(3) - np.outer(b, b)   # hat*hat' = |b|^2 I - b b'
        acc = H if i == 0 else acc + H * dt
        assert np.allclose(G[i], acc, rtol=1e-12, atol=1e-24)
        if i > 3:
            assert abs(orc.lib().orc_cond_sym3(orc.P(np.ascontiguousarray(G[i]))) / np.linalg.cond(G[i]) - 1) < 1e-9
    assert orc.condition_based_time(G, 1.0) == 0   # never reached -> 0 (magnetic_toolbox.jl:23)
