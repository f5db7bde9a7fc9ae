"""GPU probe: the magnetic-diversity sweep (ragged horizons) with the straggler hand-over at several allowances."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench as B
import tortoisesat.jl_b200 as tb
from tortoisesat.jl_b200 import host

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
eng = tb.Engine(0)
tr = B.make_trials("mc_sweep", n, 0)
cfg = B.mc_config(host, tr, n)
fo = np.zeros(len(tr["fo"]), dtype=host.FIELD_OPTS_DTYPE)
for i, f in enumerate(tr["fo"]):
    fo[i] = f
sid = np.arange(n).astype(np.uint32)
for susp in sys.argv[2:] or ["0", "200"]:
    os.environ["TS_K3_SUSPEND"] = susp
    out, st = eng.monte_carlo_run(cfg, tr["kep"], fo, tr["x0"], tr["xf"], tr["Jm"], q_noise0=tr["qn"], stream_id=sid)
    N = out["N"].astype(float); it = out["inner_iters"].astype(float)
    ki = N * it
    print("suspend", susp, "solve ms %.0f" % st.ms_solve, "split", eng.k3_last_split(), "status", np.bincount(out["status"], minlength=6).tolist())
    print("   N quantiles 0/50/90/99/100:", np.percentile(N, [0, 50, 90, 99, 100]).tolist(), "| iters quantiles 50/90/99/100:",
          np.percentile(it, [50, 90, 99, 100]).tolist())
    j = np.argsort(-ki)[:5]
    print("   largest N x iters:", [(int(N[i]), int(it[i]), int(out["status"][i])) for i in j], "| max N trial iters", int(it[np.argmax(N)]),
          "| sum knot-iters %.3g" % ki.sum(), flush=True)
