"""ncu-sized run of everything EXCEPT the AL-iLQR solve at benchmark sizes: field pass (K2), slew preparation, TVLQR
replay (K4: linearise / Riccati / stage records / replay) on 1184 fixed-orbit benchmark trials with N = 2044 knots
(K3 cut to one outer x three inner iterations: K4 replays whatever trajectory it is given), then K1 on 2e7 points.

  ncu --set full -k regex:"k1_|k2|k4|k_slew" python tools/k_other_profile.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench as B
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1184
eng = tb.Engine(0)
tr = B.make_trials("mc_fixed_orbit", 4096, 0)
sub = dict(tr)
for k in ("x0", "xf", "Jm", "qn"):
    sub[k] = tr[k][:n]
cfg = B.mc_config(host, sub, n)
cfg.ilqr.max_outer, cfg.ilqr.max_inner = 1, 3
fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
fo[0] = tr["fo"][0]
for rep in range(2 if "--twice" in sys.argv else 1):   # (the first call of a process pays the scratch allocations)
    out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, sub["x0"], sub["xf"], sub["Jm"], q_noise0=sub["qn"], stream_id=np.arange(n).astype(np.uint32))
    print("trials", n, "N", int(out["N"][0]), "ms field %.3f prep %.3f solve %.3f tvlqr %.3f" % (st.ms_field, st.ms_prep, st.ms_solve, st.ms_tvlqr))
m = 20_000_000
rng = np.random.default_rng(1)
lat, lon, r = np.arcsin(2 * rng.random(m) - 1), np.pi * (2 * rng.random(m) - 1), 6771000.0 + rng.random(m)
eng.igrf12_batch(2019.0, r, lat, lon)
print("K1", m, "points: kernel ms", eng.last_kernel_ms())
