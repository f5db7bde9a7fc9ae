#!/usr/bin/env python3
"""Generates tests/golden/igrf12_golden.json: IGRF-12 field vectors computed by an
INDEPENDENT high-precision evaluation (mpmath, 40 digits) of the published
spherical-harmonic definition

  V = a * sum_n (a/r)^(n+1) sum_m (g cos m phi + h sin m phi) P_n^m(cos theta)   (Schmidt semi-normalised)
  B_north = (1/r) dV/dtheta,  B_east = -(1/(r sin theta)) dV/dphi,  B_down = dV/dr

with Legendre functions from mpmath.legenp and d/dtheta by mpmath.diff -- no
recursion shared with the oracle or the kernels.  Coefficients are parsed from
the reference's own table (src/igrf12_coefs.jl), linear interpolation between
epochs / secular variation after 2015 exactly as igrf.jl:170-179 does.

Run in the build container (needs /root/reference):  python tests/golden/gen_igrf_golden.py
"""
import json
import math
import re
import sys
from pathlib import Path

import mpmath as mp

mp.mp.dps = 40
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")


def parse(name):
    text = (REF / "src/igrf12_coefs.jl").read_text()
    m = re.search(r"const\s+%s\s*=\s*\[(.*?)\n\]" % name, text, re.S)
    rows = {}
    for line in m.group(1).splitlines():
        line = line.split("#")[0].strip()
        if line:
            t = line.split()
            rows[(int(t[0]), int(t[1]))] = [mp.mpf(x) for x in t[2:]]
    return rows


G, H = parse("G_igrf12"), parse("H_igrf12")


def coef(tab, n, m, date):
    if (n, m) not in tab:
        return mp.mpf(0)
    row = tab[(n, m)]
    idx = int(math.floor((date - 1900) * 0.2 + 1)) if date < 2020 else 24
    epoch = 1900 + (idx - 1) * 5
    dt = mp.mpf(date) - epoch
    if date < 2015:
        return row[idx - 1] + (row[idx] - row[idx - 1]) / 5 * dt
    return row[idx - 1] + row[24] * dt


def schmidt(n, m, x):
    # mpmath.legenp includes the Condon-Shortley phase; remove it.
    p = mp.legenp(n, m, x, type=2) * (-1) ** m
    k = mp.sqrt((2 if m else 1) * mp.factorial(n - m) / mp.factorial(n + m))
    return p * k


def field(date, r_m, lat, lon):
    a = mp.mpf("6371.2")
    r = mp.mpf(r_m) / 1000
    theta = mp.pi / 2 - mp.mpf(lat)
    phi = mp.mpf(lon)
    idx = int(math.floor((date - 1900) * 0.2 + 1)) if date < 2020 else 24
    nmax = 10 if (1900 + (idx - 1) * 5) < 1995 else 13
    dVr = dVt = dVp = mp.mpf(0)
    for n in range(1, nmax + 1):
        for m in range(0, n + 1):
            g, h = coef(G, n, m, date), coef(H, n, m, date)
            P = schmidt(n, m, mp.cos(theta))
            dP = mp.diff(lambda th: schmidt(n, m, mp.cos(th)), theta)
            rad = a * (a / r) ** (n + 1)
            ang = g * mp.cos(m * phi) + h * mp.sin(m * phi)
            dang = m * (-g * mp.sin(m * phi) + h * mp.cos(m * phi))
            dVr += -(n + 1) / r * rad * ang * P
            dVt += rad * ang * dP
            dVp += rad * dang * P
    return [float(dVt / r), float(-dVp / (r * mp.sin(theta))), float(dVr)]


def main():
    import random
    rnd = random.Random(20261018)
    pts = [(2019.0, 6771000.0, 0.0, 0.0), (2019.0, 6771000.0, 0.5, -2.0), (2019.0, 6771000.0, -1.2, 3.0),
           (2017.5, 6871200.0, 0.9, 1.0), (1987.25, 7000000.0, -0.3, -0.7), (2015.0, 6771000.0, 1.5, 0.1),
           (2024.5, 7171200.0, -1.55, -3.1), (1900.0, 6500000.0, 0.1, 3.1), (1994.999, 6771000.0, 0.7, 2.2),
           (1995.0, 6771000.0, 0.7, 2.2), (2010.3, 6671200.0, -0.9, -1.3)]
    for _ in range(21):
        pts.append((round(rnd.uniform(1900, 2025), 3), round(rnd.uniform(6.5e6, 8.0e6), 1),
                    round(math.asin(rnd.uniform(-1, 1)), 6), round(rnd.uniform(-math.pi, math.pi), 6)))
    out = []
    for (d, r, la, lo) in pts:
        out.append({"date": d, "r_m": r, "lat": la, "lon": lo, "B_ned_nT": field(d, r, la, lo)})
        print(out[-1])
    Path(__file__).with_name("igrf12_golden.json").write_text(json.dumps(
        {"how": "tests/golden/gen_igrf_golden.py (mpmath, 40 digits, independent of oracle and kernels)", "points": out},
        indent=1))


if __name__ == "__main__":
    main()
