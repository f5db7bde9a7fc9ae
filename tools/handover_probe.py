"""GPU probe: determinism of the straggler hand-over (k3_wide_kernel) on a mixed ensemble."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import slew_setup as S
import tortoisesat.jl_b200 as tb
from test_gpu_ilqr import _gpu_opts

eng = tb.Engine(0)
rng = np.random.default_rng(77)
qf = np.array([np.sqrt(2) / 2, np.sqrt(2) / 2, 0, 0])
base = S.build_slew([0, 6771, 96.6, 0, 0, 90], S.J_1U, qf, qf, t_final=40.0, tf=2400.0, alpha=0.1)
n = 192
x0 = np.tile(base.x0, (n, 1))
for i in range(n):
    dq = S.quat_axis_angle(rng.normal(size=3), rng.uniform(0.2, 2.5) if i % 3 else rng.uniform(60, 170))
    x0[i, 3:7] = np.array([qf[0] * dq[0] - qf[1:] @ dq[1:], *(qf[0] * dq[1:] + dq[0] * qf[1:] + np.cross(qf[1:], dq[1:]))])
Qd, Qfd, Rd = eng.slew_weights_batch(x0, np.tile(base.xf, (n, 1)), np.tile(base.J.reshape(-1), (n, 1)), [base.t_final] * n,
                                     dt=0.2, alpha=0.1, beta=1e3)
args = dict(N_i=[base.N] * n, x0=x0, xf=np.tile(base.xf, (n, 1)), Jmat=np.tile(base.J.reshape(-1), (n, 1)), Qd=Qd, Qfd=Qfd, Rd=Rd,
            B_eci=base.B, B_offs=[0] * n, B_rows=[base.B.shape[0]] * n, index_scale=[base.index_scale] * n,
            clock_rate=[base.clock_rate] * n, dt=base.dt, want_K=False)
go = _gpu_opts(tb, S.orc.default_ilqr_opts())
go.max_outer = 8
ref = None
for tail, susp in [("1", "0"), ("0", "0"), ("1", "250"), ("1", "250"), ("0", "250"), ("1", "3"), ("1", "3"), ("1", "100"), ("0", "100")]:
    os.environ["TS_K3_TAIL"] = tail
    os.environ["TS_K3_SUSPEND"] = susp
    X, U, K, out, offs = eng.alilqr_solve_batch(**args, opts=go)
    if ref is None:
        ref = (X.copy(), out.copy())
        print("ref status", np.bincount(out["status"], minlength=5).tolist(), "inner max", out["inner_iters"].max())
        continue
    bad = [i for i in range(n) if any(out[f][i] != ref[1][f][i] for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "J"))]
    print("tail", tail, "suspend", susp, "differing trials", bad[:10], "max|dX|", float(np.max(np.abs(X - ref[0]))))
    for i in bad[:4]:
        print("   trial", i, "ref", [ref[1][f][i] for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "J")],
              "got", [out[f][i] for f in ("status", "outer_iters", "inner_iters", "ls_rollouts", "J")])
