"""CPU: the assumption registry of SURVEY.md App. C as named switches (ts_ilqr_opts.a2_* ... a7_*, stage_cost_dt = A1).

TrajectoryOptimization.jl v0.1.2 is not in /root/reference, so every solver detail its call sites do not pin is a
switch whose default is the frozen reading and whose alternative is the other plausible reading.  Each test flips ONE
switch and checks (1) the kernel source (host lane-emulator, same code the GPU runs) follows the oracle under the
alternative too -- identical iteration path, J to 1e-9 -- and (2) what the flip does to the solve, which is the
measured sensitivity quoted in DESIGN.md section 3:
  A1, A4, A6, A7 change the iteration path (never the problem being solved);
  A2, A3, A5 are neutral on this problem class: u = +-1 is never hit exactly (A2), the Todorov gradient criterion never
  fires before the cost criterion (A3), and an inactive inequality has lambda = 0 and c < 0, so lambda + mu*c < 0 is
  projected back to 0 whether or not the update is applied (A5 is algebraically a no-op).
The GPU twin of this test is tests/test_gpu_ilqr.py::test_assumption_flips_on_gpu."""
import numpy as np
import pytest

import slew_setup as S

FLAGS = ["stage_cost_dt", "a2_active_ge", "a3_grad_over_N", "a4_no_intermediate", "a5_dual_active_only", "a6_penalty_conditional",
         "a7_carry_cost"]
NEUTRAL = {"a2_active_ge", "a3_grad_over_N", "a5_dual_active_only"}


@pytest.fixture(scope="module")
def slew():
    return S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], 2.0), np.array([1.0, 0, 0, 0]), t_final=30.0)


@pytest.fixture(scope="module")
def default_outcome(orc, slew):
    return S.oracle_solve([slew])[3][0]


def path(o):
    return int(o["status"]), int(o["outer_iters"]), int(o["inner_iters"]), int(o["ls_rollouts"])


@pytest.mark.parametrize("flag", FLAGS)
def test_flip(orc, slew, default_outcome, flag):
    o = orc.default_ilqr_opts()
    assert getattr(o, flag) == 0
    setattr(o, flag, 1)
    Xo, Uo, Ko, out = S.oracle_solve([slew], opts=o)
    ro = out[0]
    Xh, Uh, Kh, rh = S.hostsim_solve(slew, opts=o)
    assert path(rh) == path(ro), (flag, rh, ro)
    assert abs(rh["J"] - ro["J"]) <= 1e-9 * abs(ro["J"]) and abs(rh["c_max"] - ro["c_max"]) <= 1e-9
    assert np.max(np.abs(Xh[:, :7] - Xo[0][:, :7])) < 1e-8
    if flag in NEUTRAL:
        assert path(ro) == path(default_outcome) and ro["J"] == default_outcome["J"], flag
    else:
        assert path(ro) != path(default_outcome) or ro["J"] != default_outcome["J"], flag
        assert ro["status"] == default_outcome["status"] == 0        # same problem, still solved


def test_defaults_are_the_frozen_spec(orc):
    o = orc.default_ilqr_opts()
    assert [getattr(o, f) for f in FLAGS] == [0] * len(FLAGS)
    assert (o.max_outer, o.max_inner, o.max_linesearch, o.goal_mask) == (20, 50, 20, 0x7F)
    assert (o.penalty_initial, o.penalty_scaling, o.penalty_max, o.constraint_tol) == (1.0, 10.0, 1e8, 1e-3)
    assert o.constraint_decrease_ratio == 0.25
