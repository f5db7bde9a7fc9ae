"""GPU probe: K3 throughput on the configs[2] ensemble (fixed orbit, random attitudes)."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import slew_setup as S
from oracle import oracle as orc
import tortoisesat.jl_b200 as tb

ntr = [int(a) for a in sys.argv[1:]] or [256, 1024, 4096]
eng = tb.Engine(0)
rng = np.random.default_rng(2026)
qf = np.array([math.sqrt(2) / 2, math.sqrt(2) / 2, 0, 0])
kep = [0, 6771.0, 96.6, 0.0, 0.0, 90.0]
base = S.build_slew(kep, S.J_1U, np.array([1.0, 0, 0, 0]), qf, tf=2400.0, cutoff=30.0, alpha=0.1)
print("N", base.N, "t_final", base.t_final, flush=True)
T = max(ntr)
q0 = rng.normal(size=(T, 4)); q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
L = orc.lib()
Qd = np.zeros((T, 8)); Qfd = np.zeros((T, 8)); Rd = np.zeros((T, 3))
nt = base.t.shape[0]
t0 = time.time()
for i in range(T):
    x0 = np.concatenate([[0, 0, 0], q0[i]])
    w_g = np.zeros((nt, 3)); q_g = np.zeros((nt, 4))
    L.orc_eigen_axis_slew(orc.P(x0), orc.P(orc.f64(base.xf[:7])), orc.P(base.t), nt, orc.P(w_g), orc.P(q_g))
    L.orc_bryson_weights(orc.P(w_g), nt, orc.P(orc.f64(base.J)), base.dt, 0.1, 1e3, orc.P(Qd[i]), orc.P(Qfd[i]), orc.P(Rd[i]))
print("weights host s", time.time() - t0, flush=True)
x0 = np.concatenate([np.zeros((T, 3)), q0, np.zeros((T, 1))], axis=1)
xf = np.tile(base.xf, (T, 1))
res = {}
for n in ntr:
    args = dict(N_i=[base.N] * n, x0=x0[:n], xf=xf[:n], Jmat=np.tile(base.J.reshape(-1), (n, 1)), Qd=Qd[:n], Qfd=Qfd[:n], Rd=Rd[:n],
                B_eci=base.B, B_offs=[0] * n, B_rows=[base.B.shape[0]] * n, index_scale=[base.index_scale] * n,
                clock_rate=[base.clock_rate] * n, dt=base.dt, want_K=False)
    t0 = time.time()
    X, U, K, out, offs = eng.alilqr_solve_batch(**args)
    wall = time.time() - t0
    ms = eng.last_kernel_ms()
    st = np.bincount(out["status"], minlength=5)
    print(n, "kernel ms", ms, "wall s", wall, "trials/s", n / (ms * 1e-3), "status", st.tolist(), "outer mean", out["outer_iters"].mean(),
          "inner mean/max", out["inner_iters"].mean(), out["inner_iters"].max(), "ls mean", out["ls_rollouts"].mean(), flush=True)
    print("   K3 split: persistent %.0f ms, straggler kernel %.0f ms, %d trials handed over" % eng.k3_last_split(),
          "| inner-iteration quantiles 50/75/90/95/99:", np.percentile(out["inner_iters"], [50, 75, 90, 95, 99]).tolist(), flush=True)
    # angle correlation: could the slew angle predict the long trials?
    qfv = np.asarray(base.xf[3:7]); dots = np.abs(q0[:n] @ qfv); ang = 2 * np.degrees(np.arccos(np.clip(dots, 0, 1)))
    itn = out["inner_iters"].astype(float)
    order_a = np.argsort(-ang); order_i = np.argsort(-itn)
    top = min(1184, n // 3)
    print("   angle vs iterations: corr %.3f | of the %d longest trials, %d are among the %d largest angles | mean iters by angle quartile:" % (
        np.corrcoef(ang, itn)[0, 1], top, len(set(order_a[:top]) & set(order_i[:top])), top),
        [round(float(itn[order_a[k * n // 4:(k + 1) * n // 4]].mean()), 1) for k in range(4)], flush=True)
    its = out["inner_iters"].astype(float)
    i = int(np.argmax(its))
    kn = its * base.N
    print("   slowest trial per knot-iter: bwd %.0f (lin %.0f) fwd %.0f | median trial: bwd %.0f (lin %.0f) fwd %.0f" % (
        out["t_final"][i] / kn[i], out["flops"][i] / kn[i], out["slew_time"][i] / kn[i], np.median(out["t_final"] / kn),
        np.median(out["flops"] / kn), np.median(out["slew_time"] / kn)), flush=True)
    res[n] = dict(ms=ms, trials_per_s=n / (ms * 1e-3), status=st.tolist(), inner_mean=float(out["inner_iters"].mean()),
                  inner_max=int(out["inner_iters"].max()), ls_mean=float(out["ls_rollouts"].mean()))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mc_probe.json", "w"), indent=1)
