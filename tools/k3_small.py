"""Small K3 run for ncu: 32 trials, N = 300."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import slew_setup as S
import tortoisesat.jl_b200 as tb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
suspend = int(os.environ.get("K3_SUSPEND", "-1"))   # tool-level knob: passed to the library through ts_ilqr_opts.k3_suspend_after
same = len(sys.argv) > 2 and sys.argv[2] == "same"
tfin = float(sys.argv[3]) if len(sys.argv) > 3 else 60.0   # horizon: N = tfin / 0.2   # identical trials: every warp of an SM stays in the same phase
eng = tb.Engine(0)
rng = np.random.default_rng(5)
qf = np.array([1.0, 0, 0, 0])
base = S.build_slew([0, 6578, 96, 0, 0, 90], S.J_1P, S.quat_axis_angle([1, 0, 1], 5.0), qf, t_final=tfin)
x0 = np.tile(base.x0, (n, 1))
for i in range(n):
    if i == 0 or not same:
        x0[i, 3:7] = S.quat_axis_angle(rng.normal(size=3), rng.uniform(2, 6))
    else:
        x0[i] = x0[0]
args = dict(N_i=[base.N] * n, x0=x0, xf=np.tile(base.xf, (n, 1)), Jmat=np.tile(base.J.reshape(-1), (n, 1)), Qd=np.tile(base.Qd, (n, 1)),
            Qfd=np.tile(base.Qfd, (n, 1)), Rd=np.tile(base.Rd, (n, 1)), B_eci=base.B, B_offs=[0] * n, B_rows=[base.B.shape[0]] * n,
            index_scale=[base.index_scale] * n, clock_rate=[base.clock_rate] * n, dt=base.dt, want_K=False)
opts = tb.host.default_ilqr_opts()
if suspend >= 0:
    opts.k3_suspend_after = suspend
X, U, K, out, offs = eng.alilqr_solve_batch(**args, opts=opts)
ms = eng.last_kernel_ms()
its = out["inner_iters"]
i = int(np.argmax(its))
cyc = eng.k3_last_cycles(n)   # SM cycles per trial: backward pass, forward pass, linearisation share
print("slowest trial cycles: backward %.3g (linearise %.3g) forward %.3g ; per knot-iter: bwd %.0f (lin %.0f) fwd %.0f" % (
    cyc[i, 0], cyc[i, 2], cyc[i, 1], cyc[i, 0] / (its[i] * base.N), cyc[i, 2] / (its[i] * base.N), cyc[i, 1] / (its[i] * base.N)))
print("K3 split (persistent ms, straggler ms, handed over):", eng.k3_last_split())
print("n", n, "same" if same else "random", "N", base.N, "kernel ms", ms, "status", np.bincount(out["status"], minlength=5).tolist(), "inner mean/max", its.mean(), its.max(),
      "ls mean", out["ls_rollouts"].mean(), "cycles/knot-iter (max trial, 1.9GHz)", ms * 1e-3 * 1.9e9 / (its.max() * base.N))
