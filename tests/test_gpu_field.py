"""GPU: K2 (kep_ECI + Euler orbit + field table + gramian + cutoff) vs the CPU oracle."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GM = 3.986004418E14 * (1 / 1000) ** 3
TOL = 1e-10  # north_star: field vectors to 1e-10 relative


def _opts(tb, rows):
    o = np.zeros(len(rows), dtype=tb.host.FIELD_OPTS_DTYPE)
    for i, r in enumerate(rows):
        o[i] = r
    return o


def test_config1_field_table_and_cutoff(engine, orc):
    import tortoisesat.jl_b200 as tb
    kep = np.array([[0, 6578, 96, 0, 0, 90.0]])
    o = _opts(tb, [(GM, 58155.0, 2019.0, 6771000.0, 0.0, 5400.0, 5000)])
    B, offs, pos, vel = engine.magnetic_simulation_batch(kep, o, want_pos=True)
    Bo, poso, velo, _ = orc.magnetic_simulation(kep[0], GM, 58155.0, 2019.0, 6771000.0, 0.0, 5400.0, 5000)
    assert B.shape == (10000, 3) and np.all(B[-1] == 0)
    err = np.linalg.norm(B[:-1] - Bo[:-1], axis=1) / np.linalg.norm(Bo[:-1], axis=1)
    assert err.max() < TOL, err.max()
    assert np.max(np.abs(pos - poso)) / 6578.0 < 1e-12
    assert np.max(np.abs(vel - velo)) / 7.7 < 1e-12
    idx = engine.condition_cutoff_batch(B, [0], [10000], [5400.0 / 5000], [50.0])
    assert idx[0] == 288 == orc.condition_based_time(orc.magnetic_gramian(Bo, 1.08), 50)
    G = engine.magnetic_gramian_batch(B, [0], [10000], [1.08])
    Go = orc.magnetic_gramian(Bo, 1.08)
    assert np.max(np.abs(G - Go)) / np.max(np.abs(Go)) < 1e-10
    assert engine.condition_based_time_batch(G, [0], [10000], [50.0])[0] == 288
    assert engine.condition_based_time_batch(G, [0], [10000], [1.0])[0] == 0


def test_ragged_sweep_trials(engine, orc):
    """config-4 style randomisation: inclination/altitude/RAAN/anomaly/MJD/IGRF date per trial,
    ragged N, including a pre-1995 (degree-10) date and an eccentric orbit."""
    import tortoisesat.jl_b200 as tb
    rng = np.random.default_rng(11)
    T = 12
    kep = np.zeros((T, 6))
    rows = []
    for t in range(T):
        alt = rng.uniform(350, 800)
        kep[t] = [0.0 if t % 3 else 0.01 * t, alt + 6371.0, rng.uniform(0, 98), rng.uniform(0, 360), 0.0 if t % 2 else 33.0,
                  rng.uniform(0, 360)]
        date = 2015 + 5 * rng.random() if t != 5 else 1988.4
        rows.append((GM, rng.uniform(58155, 58520), date, (alt + 6371.0) * 1000.0, 0.0, rng.uniform(300, 900), int(rng.integers(40, 400))))
    o = _opts(tb, rows)
    B, offs, pos, vel = engine.magnetic_simulation_batch(kep, o, want_pos=True)
    for t in range(T):
        r = rows[t]
        Bo, poso, velo, _ = orc.magnetic_simulation(kep[t], *r[:6], r[6])
        Bt = B[offs[t]:offs[t + 1]]
        assert Bt.shape == Bo.shape
        err = np.linalg.norm(Bt[:-1] - Bo[:-1], axis=1) / np.linalg.norm(Bo[:-1], axis=1)
        assert err.max() < TOL, (t, err.max())
        pt = pos[offs[t] + t: offs[t + 1] + t + 1]
        assert np.max(np.abs(pt - poso)) / 7000.0 < 1e-11
    rws = (offs[1:] - offs[:-1]).astype(np.int64)
    dts = np.array([(r[5] - r[4]) / r[6] for r in rows])
    cut = np.full(T, 100.0)
    idx = engine.condition_cutoff_batch(B, offs[:-1], rws, dts, cut)
    for t in range(T):
        Bo = B[offs[t]:offs[t + 1]]
        assert idx[t] == orc.condition_based_time(orc.magnetic_gramian(Bo, dts[t]), 100.0)


def test_rows_limit_and_reference_named_wrappers(engine, orc):
    import tortoisesat.jl_b200 as tb
    from tortoisesat.jl_b200 import host
    kep = np.array([[0, 6771, 96.6, 10.0, 0, 20.0]])
    o = _opts(tb, [(GM, 58155.0, 2019.0, 6771000.0, 0.0, 300.0, 1500)])
    full, _ = engine.magnetic_simulation_batch(kep, o)
    lim, _ = engine.magnetic_simulation_batch(kep, o, rows_limit=[90])
    assert np.array_equal(lim[:90], full[:90]) and np.all(lim[90:] == 0)
    p = host.input_parameters("1P", [0, 6578, 96, 0, 0, 90], 12345.0)
    assert p.MJD == 58155.0 and p.alt == 400.0            # quirks Q10
    B, pos, vel = host.magnetic_simulation(p, 0.0, 311.04, 1555, None)
    Bo, poso, _, _ = orc.magnetic_simulation(p.Kep, GM, 58155.0, 2019.0, 6771000.0, 0.0, 311.04, 1555)
    assert pos.shape == (3, 3111) and B.shape == (3110, 3)
    assert (np.linalg.norm(B[:-1] - Bo[:-1], axis=1) / np.linalg.norm(Bo[:-1], axis=1)).max() < TOL
    G = host.magnetic_gramian(B, 0.2)
    assert G.shape == (3, 3, 3110)
    assert host.condition_based_time(G, 1e9) >= 1
