"""How well does the state of a trial at the hand-over predict how long it will still run?

  python tools/park_predictor.py [n_trials]

Runs the benchmark ensemble (BASELINE configs[2]) with the default hand-over, then joins the parked trials' counters at
parking (ts_k3_last_parked) with their final iteration counts.  Prints, for express sets of the E parked trials with the
lowest outer count at parking, how many of the trials that end above a given total iteration count they contain.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import bench as B
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = tb.Engine(0)
tr = B.make_trials("mc_fixed_orbit", n, 0)
cfg = B.mc_config(host, tr, n)
cfg.run_tvlqr = 0
fo = np.zeros(1, dtype=host.FIELD_OPTS_DTYPE)
fo[0] = tr["fo"][0]
out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, tr["x0"], tr["xf"], tr["Jm"], q_noise0=tr["qn"], stream_id=np.arange(n).astype(np.uint32))
ti, ou, inn = eng.k3_last_parked(n)
final = out["inner_iters"][ti]
print("parked", len(ti), "outer-at-park histogram", np.bincount(ou).tolist())
for o in range(ou.max() + 1):
    m = ou == o
    if m.any():
        print("  outer %2d at park: %4d trials, final inner iterations quantiles 10/50/90/100: %s" % (
            o, m.sum(), np.percentile(final[m], [10, 50, 90, 100]).astype(int).tolist()))
# rank as the park-order kernel does: remaining budget = (max_outer - outer) * max_inner + (max_inner - it) -> lowest outer first
key = ou.astype(np.int64) * 1000 + inn
order = np.argsort(key, kind="stable")
res = {}
for E in (148, 296, 444, 592):
    ex = np.zeros(len(ti), dtype=bool)
    ex[order[:E]] = True
    row = {}
    for thr in (600, 700, 800, 900, 950):
        long_ = final >= thr
        row[str(thr)] = [int((long_ & ex).sum()), int(long_.sum())]
    # makespan model: express trials at t_e ms per iteration, the others at t_d
    res[str(E)] = row
    print("express set %3d (lowest outer count first): covered / all trials ending at >= thr iterations:" % E, row)
    print("      longest NOT covered: %d iterations" % (final[~ex].max() if (~ex).any() else 0))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "park_predictor.json"), "w"))
