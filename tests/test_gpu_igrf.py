"""GPU: K1 batched IGRF-12 through the C ABI vs the CPU oracle / golden vectors."""
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-10  # north_star: field vectors to 1e-10 relative (norm-relative: igrf.jl:270 divides east by sin(theta))


def relerr(a, b):
    a = np.stack(a, -1)
    b = np.stack(b, -1)
    return np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)


def leo_points(n, seed):
    rng = np.random.default_rng(seed)
    lat = np.arcsin(2 * rng.random(n) - 1)
    lon = math.pi * (2 * rng.random(n) - 1)
    r = 6371200.0 + 300000.0 + 900000.0 * rng.random(n)
    return r, lat, lon


def test_golden_vectors(engine):
    pts = json.load(open(os.path.join(HERE, "golden", "igrf12_golden.json")))["points"]
    for p in pts:
        bn, be, bd = engine.igrf12_batch(p["date"], [p["r_m"]], [p["lat"]], [p["lon"]])
        ref = np.array(p["B_ned_nT"])
        got = np.array([bn[0], be[0], bd[0]])
        assert np.linalg.norm(got - ref) <= TOL * np.linalg.norm(ref), p


@pytest.mark.parametrize("date", [2019.0, 2016.25, 2003.7, 1987.25, 1900.0, 2025.0])
def test_vs_oracle_random_leo(engine, orc, date):
    r, lat, lon = leo_points(200_000, 0x5EED)
    got = engine.igrf12_batch(date, r, lat, lon)
    ref = orc.igrf12_batch(date, r, lat, lon, nthreads=orc.lib().orc_max_threads())[:3]
    e = relerr(got, ref)
    assert e.max() <= TOL, (date, e.max())


def test_poles_and_edges(engine, orc):
    lat = np.array([math.pi / 2, -math.pi / 2, math.pi / 2 - 1e-9, -math.pi / 2 + 1e-9, 1e-3 - math.pi / 2, 0.0, 0.0, 0.0])
    lon = np.array([0.3, 0.3, -1.0, 2.0, 0.5, math.pi, -math.pi, 0.0])
    r = np.full(lat.shape, 6771000.0)
    got = engine.igrf12_batch(2019.0, r, lat, lon)
    ref = orc.igrf12_batch(2019.0, r, lat, lon)[:3]
    assert relerr(got, ref).max() <= TOL
    assert abs(got[1][0] - 161.20551673) < 1e-6          # theta == 0 branch (igrf.jl:235,270)
    assert got[1][1] == 0.0                               # s == 0 at the south pole -> east = -0.0


def test_empty_and_single(engine):
    bn, be, bd = engine.igrf12_batch(2019.0, np.zeros(0), np.zeros(0), np.zeros(0))
    assert bn.shape == (0,)
    from tortoisesat.jl_b200 import host
    b = host.igrf12(2019, 6771000.0, 0.0, 0.0)
    assert np.allclose(b, [22718.46738684, -2084.32350217, -11834.1244487], atol=2e-8)


def test_domain_errors(engine):
    import tortoisesat.jl_b200 as tb
    one = np.array([7.0e6])
    with pytest.raises(tb.TortoiseError):
        engine.igrf12_batch(1899.0, one, np.zeros(1), np.zeros(1))
    with pytest.raises(tb.TortoiseError):
        engine.igrf12_batch(2019.0, one, np.array([1.6]), np.zeros(1))
    with pytest.raises(tb.TortoiseError):
        engine.igrf12_batch(2019.0, one, np.zeros(1), np.array([3.2]))


def test_linearity_in_coefficients_full_size(engine):
    """Size-independent property at 10^7 points on device-resident inputs: the field at
    date d is the epoch-2015 field plus (d-2015) x the SV field, i.e. affine in date."""
    import torch
    n = 10_000_000
    g = torch.Generator(device="cuda").manual_seed(1)
    u = torch.rand(3, n, generator=g, device="cuda", dtype=torch.float64)
    lat = torch.asin(2 * u[0] - 1)
    lon = math.pi * (2 * u[1] - 1)
    r = 6371200.0 + 300000.0 + 900000.0 * u[2]
    outs = []
    for d in (2015.0, 2017.0, 2019.0):
        o = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
        engine.igrf12_batch(d, r, lat, lon, out=o)
        outs.append(torch.stack(o))
    torch.cuda.synchronize()
    mid = 0.5 * (outs[0] + outs[2])
    err = (outs[1] - mid).norm(dim=0) / outs[1].norm(dim=0)
    assert float(err.max()) < 1e-12


def test_igrf_data_map_call_shape(engine, orc):
    """igrf_data (magnetic_toolbox.jl:108-127): the lat/long map behind the reference's mag_field(i,j,c) call shape
    (monte_carlo.jl:90-96) -- node values in Tesla, 1-based indices, periodic extrapolation; includes both pole rows."""
    import tortoisesat.jl_b200 as tb
    n = 64
    mf = tb.host.igrf_data(400, 2019, n=n)
    assert mf.shape == (n, n, 3)
    lat = np.linspace(-math.pi / 2, math.pi / 2, n)
    lon = np.linspace(-math.pi, math.pi, n)
    for (i, j) in [(1, 1), (n, n), (n, 7), (17, 33), (1, 40)]:
        ref = np.array(orc.igrf12_batch(2019.0, [(400 + 6378) * 1000.0], [lat[i - 1]], [lon[j - 1]])[:3]).ravel() / 1e9
        got = np.array([mf(i, j, c) for c in (1, 2, 3)])
        assert np.linalg.norm(got - ref) <= TOL * np.linalg.norm(ref), (i, j)
    assert mf(n + 5, 3 - n, 4) == mf(5, 3, 1)            # extrapolate(..., Periodic())
    assert np.asarray(mf)[4, 2, 0] == mf(5, 3, 1)
    with pytest.raises(ValueError):
        mf(1.5, 2, 1)


def test_host_pipeline_matches_device_path(engine):
    """Host-pointer calls go through a chunked, double-buffered copy/compute pipeline (chunks of 2^22 points on two
    streams): several chunks plus a ragged tail must give bit-identical results to the device-resident path,
    and a domain error in a late chunk must still be reported."""
    import torch
    import tortoisesat.jl_b200 as tb
    n = 2 * (1 << 22) + 12345
    r, lat, lon = leo_points(n, 7)
    host_out = engine.igrf12_batch(2019.0, r, lat, lon)
    dev_in = [torch.from_numpy(x).cuda() for x in (r, lat, lon)]
    o = [torch.empty(n, device="cuda", dtype=torch.float64) for _ in range(3)]
    engine.igrf12_batch(2019.0, *dev_in, out=o)
    torch.cuda.synchronize()
    for h, d in zip(host_out, o):
        assert np.array_equal(h, d.cpu().numpy())
    lat[-3] = 1.6
    with pytest.raises(tb.TortoiseError):
        engine.igrf12_batch(2019.0, r, lat, lon)


@pytest.mark.parametrize("date,isv,itype", [(2019.0, 0, 2), (2019.0, 0, 1), (2017.3, 1, 2), (2003.7, 0, 2), (1987.25, 0, 1), (1996.0, 1, 1)])
def test_igrf12syn_batch_vs_oracle(engine, orc, date, isv, itype):
    """ts_igrf12syn_batch (K6) -- the Fortran-style twin igrf12syn(isv,date,itype,alt,colat,elong) of igrf.jl:335-534 --
    against the oracle's restatement, and (geocentric main field) against igrf12 itself as the reference's own
    cross-check does (igrf.jl:283-287)."""
    rng = np.random.default_rng(12)
    n = 20_000
    colat = np.degrees(np.arccos(2 * rng.random(n) - 1))
    elong = 360.0 * rng.random(n)
    alt = (6371.2 + 300 + 900 * rng.random(n)) if itype == 2 else (300 + 900 * rng.random(n))
    colat[:3] = [0.0, 180.0, 90.0]                      # poles: the st == 0 branch of igrf.jl:511-515
    x, y, z, f = engine.igrf12syn_batch(isv, date, itype, alt, colat, elong)
    ref = np.array([orc.igrf12syn(isv, date, itype, alt[i], colat[i], elong[i]) for i in range(0, n, 97)] +
                   [orc.igrf12syn(isv, date, itype, alt[i], colat[i], elong[i]) for i in range(3)])
    got = np.stack([x, y, z, f], -1)
    got = np.concatenate([got[0:n:97], got[:3]])
    scale = np.abs(ref[:, 3:4])
    assert np.max(np.abs(got - ref) / scale) <= 1e-11
    if isv == 0 and itype == 2:
        lat = np.radians(90.0 - colat[3:])
        lon = np.radians(elong[3:])
        lon = np.where(lon > math.pi, lon - 2 * math.pi, lon)
        bn, be, bd = engine.igrf12_batch(date, alt[3:] * 1000.0, lat, lon)
        e = np.linalg.norm(np.stack([bn - x[3:], be - y[3:], bd - z[3:]], -1), axis=-1) / f[3:]
        assert e.max() < 1e-9                          # two implementations, two coefficient tables
    with pytest.raises(Exception):
        engine.igrf12syn_batch(0, 2031.0, 2, [6771.0], [90.0], [0.0])
