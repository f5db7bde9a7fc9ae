"""compute-sanitizer-sized run of the whole hot path: a 24-trial ragged Monte-Carlo (field -> weights -> K3 with lane lending,
parking after 3 iterations and the one-warp-per-trial launch -> K4 with pre-generated noise), then K1 / K6 on a few points.

  compute-sanitizer --tool memcheck  python tools/sanitize_small.py
  compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench as B
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
eng = tb.Engine(0)
tr = B.make_trials("mc_sweep", n, 0)
cfg = B.mc_config(host, tr, n)
cfg.tf, cfg.N_scope, cfg.cutoff = 600.0, 300, 1e9          # short horizons (N ~ 10..60): the sanitizers slow kernels 50-100x
cfg.ilqr.max_outer, cfg.ilqr.max_inner = 3, 8
cfg.ilqr.k3_suspend_after = 3
cfg.keep_trajectories = 1
fo = B.field_opts_array(tr)
out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, tr["x0"], tr["xf"], tr["Jm"], q_noise0=tr["qn"], stream_id=np.arange(n).astype(np.uint32))
tj = eng.mc_trajectories(n)
print("status", np.bincount(out["status"], minlength=6).tolist(), "N", out["N"].tolist(), "handed over", eng.k3_last_split()[2],
      "knots", int(tj["knot_offs"][-1]))
rng = np.random.default_rng(1)
m = 1000
lat, lon, r = np.arcsin(2 * rng.random(m) - 1), np.pi * (2 * rng.random(m) - 1), 6771000.0 + rng.random(m)
eng.igrf12_batch(2019.0, r, lat, lon)
eng.igrf12syn_batch(0, 2019.0, 2, r / 1000, np.degrees(np.pi / 2 - lat), np.degrees(lon) % 360)
print("ok")
