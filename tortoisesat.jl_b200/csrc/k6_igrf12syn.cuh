// K6 -- batched igrf12syn: the Fortran-style twin of igrf12 that the reference vendors next to it
// (src/igrf.jl:335-534, coefficient vector gh_igrf12 of src/igrf12syn_coefs.jl:38-474).  The reference keeps it as
// the independent cross-check of igrf12 (igrf.jl:283-287); here it is the drop-in for callers that use the
// (isv, date, itype, alt, colat, elong) call shape.  One point per thread; the fused p/q recursion of the original
// is kept in its own operation order (1-based work arrays p[105], q[105], cl[13], sl[13] live in local memory -- this
// is a correctness / drop-in path, K1 is the throughput path).
#pragma once
#include "common.cuh"

namespace ts {

struct SynEpoch {   // epoch bookkeeping of igrf.jl:369-423, evaluated once per call on the host
  int ll, nc, kmx;
  double t, tc;
};

inline SynEpoch igrf12syn_epoch(int isv, double date) {
  SynEpoch e;
  if (date < 2015) {
    double t = 0.2 * (date - 1900);
    int ll = (int)floor(t);
    t = t - ll;
    if (date < 1995) {
      e.nc = 120;
      ll = e.nc * ll;
      e.kmx = 66;
    } else {
      e.nc = 195;
      ll = (int)floor(0.2 * (date - 1995));
      ll = 120 * 19 + e.nc * ll;
      e.kmx = 105;
    }
    e.ll = ll;
    e.t = t;
    e.tc = 1 - t;
    if (isv == 1) {
      e.t = +0.2;
      e.tc = -0.2;
    }
  } else {
    e.t = date - 2015;
    e.tc = 1.0;
    if (isv == 1) {
      e.t = 1.0;
      e.tc = 0.0;
    }
    e.ll = 3060;
    e.nc = 195;
    e.kmx = 105;
  }
  return e;
}

__global__ void __launch_bounds__(128) k6_igrf12syn(const double* __restrict__ gh, SynEpoch ep, int itype, int64_t npts,
                                                    const double* __restrict__ alt_, const double* __restrict__ colat_,
                                                    const double* __restrict__ elong_, double* __restrict__ xo,
                                                    double* __restrict__ yo, double* __restrict__ zo, double* __restrict__ fo) {
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= npts) return;
  const double PI = 3.141592653589793;
  const double alt = alt_[i0], colat = colat_[i0], elong = elong_[i0];
  double p[106], q[106], cl[14], sl[14];
  double x = 0.0, y = 0.0, z = 0.0;
  double r = alt;
  double ct = cos(colat * PI / 180);
  double st = sin(colat * PI / 180);
  cl[1] = cos(elong * PI / 180);
  sl[1] = sin(elong * PI / 180);
  double cd = 1.0, sd = 0.0;
  int l = 1, m = 1, n = 0;
  if (itype != 2) {  // geodetic -> geocentric, WGS-84 (igrf.jl:436-451)
    const double a2 = 40680631.6, b2 = 40408296.0;
    double one = a2 * (st * st);
    const double two = b2 * (ct * ct);
    const double three = one + two;
    const double rho = sqrt(three);
    r = sqrt(alt * (alt + 2 * rho) + (a2 * one + b2 * two) / three);
    cd = (alt + rho) / r;
    sd = (a2 - b2) / rho * ct * st / r;
    one = ct;
    ct = ct * cd - st * sd;
    st = st * cd + one * sd;
  }
  const double ratio = 6371.2 / r;
  double rr = ratio * ratio;
  p[1] = 1.0;
  p[3] = st;
  q[1] = 0.0;
  q[3] = ct;
  double fn = 0.0, gn = 0.0;
  for (int k = 2; k <= ep.kmx; ++k) {
    if (n < m) {
      m = 0;
      n = n + 1;
      rr = rr * ratio;
      fn = n;
      gn = n - 1;
    }
    const double fm = m;
    if (m == n) {
      if (k != 3) {
        const double one = sqrt(1 - 0.5 / fm);
        const int j = k - n - 1;
        p[k] = one * st * p[j];
        q[k] = one * (st * q[j] + ct * p[j]);
        cl[m] = cl[m - 1] * cl[1] - sl[m - 1] * sl[1];
        sl[m] = sl[m - 1] * cl[1] + cl[m - 1] * sl[1];
      }
    } else {
      const double gmm = (double)(m * m);
      const double one = sqrt(fn * fn - gmm);
      const double two = sqrt(gn * gn - gmm) / one;
      const double three = (fn + gn) / one;
      const int i = k - n;
      const int j = i - n + 1;
      p[k] = three * ct * p[i] - two * p[j];
      q[k] = three * (ct * q[i] - st * p[i]) - two * q[j];
    }
    const int lm = ep.ll + l;   // 1-based index into gh
    const double one = (ep.tc * gh[lm - 1] + ep.t * gh[lm + ep.nc - 1]) * rr;
    if (m != 0) {
      const double two = (ep.tc * gh[lm] + ep.t * gh[lm + ep.nc]) * rr;
      const double three = one * cl[m] + two * sl[m];
      x = x + three * q[k];
      z = z - (fn + 1) * three * p[k];
      if (st != 0)
        y = y + (one * sl[m] - two * cl[m]) * fm * p[k] / st;
      else
        y = y + (one * sl[m] - two * cl[m]) * q[k] * ct;
      l = l + 2;
    } else {
      x = x + one * q[k];
      z = z - (fn + 1) * one * p[k];
      l = l + 1;
    }
    m = m + 1;
  }
  const double one = x;
  x = x * cd + z * sd;
  z = z * cd - one * sd;
  xo[i0] = x;
  yo[i0] = y;
  zo[i0] = z;
  fo[i0] = sqrt(x * x + y * y + z * z);
}

}  // namespace ts
