"""CPU: the K4 source (tvlqr_solver.cuh, philox.cuh) compiled for the host vs the oracle."""
import ctypes as C

import numpy as np
import pytest

import slew_setup as S
from oracle import oracle as orc


def test_philox_known_answer_and_agreement():
    hs = S.hostsim()
    L = orc.lib()
    # Random123 known-answer vectors for Philox4x32-10
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kats:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o1, o2 = np.zeros(4, dtype=np.uint32), np.zeros(4, dtype=np.uint32)
        hs.hs_philox4x32_10(c.ctypes.data, k.ctypes.data, o1.ctypes.data)
        L.orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o2.ctypes.data)
        assert tuple(int(v) for v in o1) == exp
        assert tuple(int(v) for v in o2) == exp
    a, b = np.zeros(9), np.zeros(9)
    for (seed, tr, st, sg) in [(0, 0, 0, 0), (12345678901234, 7, 99, 3), (2**63 + 5, 4095, 2043, 1)]:
        hs.hs_tvlqr_noise(seed, tr, st, sg, orc.P(a))
        L.orc_tvlqr_noise(seed, tr, st, sg, orc.P(b))
        assert np.allclose(a, b, rtol=1e-14, atol=0)
    # statistics of the stream: N(0,1)*scale and U(0,1)*scale
    z = np.zeros((4000, 9))
    for i in range(4000):
        hs.hs_tvlqr_noise(99, 3, i, i % 4, orc.P(a))
        z[i] = a
    s_w, s_q, s_b = (.38 * np.pi / 180) ** 2, (np.pi / 180) ** 2, 1e-10
    assert abs(z[:, :3].std() / s_w - 1) < 0.05 and abs(z[:, 3:6].std() / s_q - 1) < 0.05
    assert abs(z[:, 6:].mean() / s_b - 0.5) < 0.03 and z[:, 6:].min() >= 0


@pytest.fixture(scope="module")
def solved():
    s = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], 5.0), np.array([1.0, 0, 0, 0]), t_final=60.0)
    Xs, Us, Ks, out = S.oracle_solve([s])
    return s, Xs[0], Us[0]


@pytest.mark.parametrize("mode", [0, 2, 1])
def test_tvlqr_source_matches_oracle(solved, mode):
    s, X, U = solved
    o, g = S.tvlqr_opts_pair(noise_mode=mode, seed=2026)
    rng = np.random.default_rng(1)
    qn = rng.normal(size=3) * (np.pi / 180) ** 2
    th = np.linalg.norm(qn)
    x0l = s.x0.copy()
    qq = np.zeros(4)
    orc.lib().orc_qmult(orc.P(orc.f64(s.x0[3:7])), orc.P(np.concatenate([[np.cos(th / 2)], qn / th * np.sin(th / 2)])), orc.P(qq))
    x0l[3:7] = qq
    x0l[7] = 0.0
    noise = None
    if mode == 1:
        noise = rng.normal(size=(s.N, 4, 9)) * np.array([1e-5] * 3 + [3e-4] * 3 + [1e-10] * 3)
    a = S.oracle_tvlqr(s, X, U, x0l, o, trial=5, noise=noise)
    b = S.hostsim_tvlqr(s, X, U, x0l, g, trial=5, noise=noise)
    assert a[4] == b[4] == s.N
    assert np.max(np.abs(a[3] - b[3])) <= 1e-10 * np.max(np.abs(a[3]))      # gains
    assert np.max(np.abs(a[0] - b[0])) < 1e-10                               # X_sim (north_star: rollouts 1e-10)
    assert np.max(np.abs(a[1] - b[1])) < 1e-9
    assert np.max(np.abs(a[2] - b[2])) < 1e-10
    assert a[5] == b[5]


def test_dt_quirk_flag_changes_gains(solved):
    s, X, U = solved
    o1, g1 = S.tvlqr_opts_pair()
    o2, g2 = S.tvlqr_opts_pair(dt_squared=0)
    k1 = S.hostsim_tvlqr(s, X, U, s.x0 * np.array([1] * 7 + [0]), g1)[3]
    k2 = S.hostsim_tvlqr(s, X, U, s.x0 * np.array([1] * 7 + [0]), g2)[3]
    assert np.max(np.abs(k2)) > 5 * np.max(np.abs(k1))      # quirk Q6: dt^2 linearisation -> tiny gains
    ko = S.oracle_tvlqr(s, X, U, s.x0 * np.array([1] * 7 + [0]), o2)[3]
    assert np.max(np.abs(ko - k2)) <= 1e-10 * np.max(np.abs(ko))
