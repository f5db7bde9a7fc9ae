"""GPU probe: the first n trials of the benchmark ensemble (configs[2], N = 2044) through ts_monte_carlo_run vs the CPU oracle.

  python tools/parity_sample.py [n] [--quat]      --quat: the quaternion-aware solver variant (ts_ilqr_opts.quat_error = 1)
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench as B
import slew_setup as S
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
quat = "--quat" in sys.argv
eng = tb.Engine(0)
tr = B.make_trials("mc_fixed_orbit", 4096, 0)
sub = dict(tr)
for k in ("x0", "xf", "Jm", "qn"):
    sub[k] = tr[k][:n]
cfg = B.mc_config(host, sub, n)
cfg.run_tvlqr = 0
cfg.ilqr.quat_error = 1 if quat else 0
fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
fo[0] = tr["fo"][0]
out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, sub["x0"], sub["xf"], sub["Jm"], q_noise0=sub["qn"], stream_id=np.arange(n).astype(np.uint32))
print("gpu solve ms", st.ms_solve, "split", eng.k3_last_split())
t0 = time.time()
slews = []
base = None
for t in range(n):
    f = tr["fo"][0]
    s = S.build_slew(tr["kep"][0], B.J_1U, tr["x0"][t, 3:7], B.QF, mjd=f[1], igrf_date=f[2], field_radius_m=f[3], tf=2400.0, cutoff=tr["cutoff"],
                     alpha=0.1, **({} if base is None else dict(t_final=base.t_final)))
    base = base or s
    slews.append(s)
oo = S.orc.default_ilqr_opts()
oo.quat_error = 1 if quat else 0
Xs, Us, Ks, ref = S.oracle_solve(slews, oo, nthreads=S.orc.lib().orc_max_threads(), want_K=False)
print("oracle s", time.time() - t0)
nbad = 0
for t in range(n):
    g, r = out[t], ref[t]
    same = all(g[f] == r[f] for f in ("status", "outer_iters", "inner_iters", "ls_rollouts"))
    dj = abs(g["J"] - r["J"]) / abs(r["J"]); dc = abs(g["c_max"] - r["c_max"])
    flag = "" if (same and dj < 1e-6 and dc < 1e-6) else "  <-- differs"
    nbad += bool(flag)
    print(t, "gpu", [int(g[f]) for f in ("status", "outer_iters", "inner_iters", "ls_rollouts")], "ref",
          [int(r[f]) for f in ("status", "outer_iters", "inner_iters", "ls_rollouts")], "dJ/J %.2e dc %.2e" % (dj, dc), flag)
print("differing:", nbad, "of", n)
