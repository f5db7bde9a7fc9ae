"""Run on the GPU box: FP64 peak probe + K1 throughput at 10^7 / 10^8 points."""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tortoisesat.jl_b200 as tb

eng = tb.Engine(0)
print("device", eng.device_info())
peak = eng.fp64_peak_tflops()
print("fp64 peak TFLOP/s", peak)
res = {"fp64_peak_tflops": peak}
for n in (10_000_000, 100_000_000):
    g = torch.Generator(device="cuda").manual_seed(0x5EED)
    u = torch.rand(3, n, generator=g, device="cuda", dtype=torch.float64)
    lat = torch.asin(2 * u[0] - 1); lon = math.pi * (2 * u[1] - 1); r = 6371200.0 + 300000.0 + 900000.0 * u[2]
    del u
    o = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
    for _ in range(3):
        eng.igrf12_batch(2019.0, r, lat, lon, out=o)
    ms = []
    for _ in range(5):
        eng.igrf12_batch(2019.0, r, lat, lon, out=o)
        ms.append(eng.last_kernel_ms())
    best = min(ms); med = sorted(ms)[len(ms)//2]
    print(n, "ms", ms, "evals/s", n / (med * 1e-3), "TFLOP/s(2243/pt)", 2243 * n / (med * 1e-3) / 1e12)
    res["k1_%d" % n] = {"ms_median": med, "ms_best": best, "evals_per_s": n / (med * 1e-3)}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
