"""K1 run for ncu: 2e7 points."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tortoisesat.jl_b200 as tb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
eng = tb.Engine(0)
g = torch.Generator(device="cuda").manual_seed(0x5EED)
u = torch.rand(3, n, generator=g, device="cuda", dtype=torch.float64)
lat = torch.asin(2 * u[0] - 1); lon = math.pi * (2 * u[1] - 1); r = 6371200.0 + 300000.0 + 900000.0 * u[2]
o = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
for _ in range(3):
    eng.igrf12_batch(2019.0, r, lat, lon, out=o)
print("n", n, "ms", eng.last_kernel_ms(), "evals/s", n / (eng.last_kernel_ms() * 1e-3))
