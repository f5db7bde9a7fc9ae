// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this code.
//
// CPU restatement (scalar, double precision, literal operation order) of the
// IGRF-12 stack vendored by the reference:
//   legendre  (Schmidt quasi-normalised)  reference src/legendre.jl:254-292
//   dlegendre (Schmidt == fully-norm. routine) reference src/dlegendre.jl:221-309
//   igrf12  geocentric                     reference src/igrf.jl:70-274
//   igrf12syn (Fortran-style cross-check)  reference src/igrf.jl:335-534
// Coefficient data: oracle/igrf12_tables.inc (generated, provenance inside).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

#include "igrf12_tables.inc"

constexpr int IGRF_NMAX = 13;
constexpr int IGRF_DIM = IGRF_NMAX + 1;  // 14x14 work matrices, like the reference

// reference src/legendre.jl:254-292 (ph_term = false path).  P is (nmax+1)^2,
// row-major P[n*(nmax+1)+m]; entries above the diagonal stay 0.
inline void legendre_schmidt(double theta, int nmax, double* P) {
  const int d = nmax + 1;
  std::memset(P, 0, sizeof(double) * d * d);
  const double c = std::cos(theta);
  const double s = std::sqrt(1 - c * c);  // two roundings: c*c, then 1-(.)
  P[0 * d + 0] = 1;
  P[1 * d + 0] = +c;
  P[1 * d + 1] = -s;
  P[1 * d + 1] *= -1;  // (!ph_term)
  for (int n = 2; n <= nmax; ++n) {
    for (int m = 0; m <= n - 1; ++m) {
      const int64_t aux = (int64_t)(n - m) * (n + m);
      const double a_nm = std::sqrt((double)((2 * n - 1) * (2 * n - 1)) / (double)aux);
      const double b_nm = std::sqrt((double)((n + m - 1) * (n - m - 1)) / (double)aux);
      P[n * d + m] = a_nm * c * P[(n - 1) * d + m] - b_nm * P[(n - 2) * d + m];
    }
    P[n * d + n] = +s * std::sqrt((double)(2 * n - 1) / (double)(2 * n)) * P[(n - 1) * d + (n - 1)];
  }
}

// reference src/dlegendre.jl:221-309 reached through the Schmidt alias :411-419.
inline void dlegendre_schmidt(double theta, int nmax, const double* P, double* dP) {
  const int d = nmax + 1;
  std::memset(dP, 0, sizeof(double) * d * d);
  const double twopi = 2 * M_PI;
  double ph = std::fmod(theta, twopi);
  if (ph < 0) ph += twopi;  // Julia mod() is floored
  const double fact = (ph > M_PI) ? -1.0 : 1.0;
  for (int n = 1; n <= nmax; ++n) {
    for (int m = 0; m <= n; ++m) {
      double v;
      if (m == 0) {
        const double aux = std::sqrt((double)(n * (n + 1)) / 2.0);
        const double a_nm = +0.5 * aux;
        const double b_nm = -0.5 * aux;
        v = -a_nm * P[n * d + 1] + b_nm * P[n * d + 1];
      } else if (m == 1) {
        const double a_nm = +0.5 * std::sqrt((double)(2 * n * (n + 1)));
        const double b_nm = -0.5 * std::sqrt((double)((n + 2) * (n - 1)));
        // (n=1,m=1) reads the upper-triangle zero P[1][2] with b_nm = -0.0
        const double pup = (2 <= nmax) ? P[n * d + 2] : 0.0;
        v = a_nm * P[n * d + 0] + b_nm * pup;
      } else if (n != m) {
        const double a_nm = +0.5 * std::sqrt((double)((n + m) * (n - m + 1)));
        const double b_nm = -0.5 * std::sqrt((double)((n + m + 1) * (n - m)));
        v = a_nm * P[n * d + m - 1] + b_nm * P[n * d + m + 1];
      } else {
        const double a_nm = +0.5 * std::sqrt((double)((n + m) * (n - m + 1)));
        v = a_nm * P[n * d + m - 1];
      }
      v *= fact;
      dP[n * d + m] = v;
    }
  }
}

// Error codes mirror the reference's three error() branches (igrf.jl:80-88).
enum { IGRF_OK = 0, IGRF_ERR_DATE = -1, IGRF_ERR_LAT = -2, IGRF_ERR_LON = -3 };

// reference src/igrf.jl:70-274.  r in metres, lat/lon in rad.  out = (N,E,D) nT.
inline int igrf12(double date, double r, double lat, double lon, double out[3]) {
  if ((date < 1900) || (date > 2025)) return IGRF_ERR_DATE;
  if ((lat < -M_PI / 2) || (lat > M_PI / 2)) return IGRF_ERR_LAT;
  if ((lon < -M_PI) || (lon > M_PI)) return IGRF_ERR_LON;

  const double theta = M_PI / 2 - lat;
  const double phi = (lon >= 0) ? lon : 2 * M_PI + lon;
  r /= 1000;

  const int idx = (date < 2020) ? (int)std::floor((date - 1900) * 0.2 + 1) : 24;
  const int epoch = 1900 + (idx - 1) * 5;
  const double dt = date - epoch;
  const int n_max = (epoch < 1995) ? 10 : 13;

  double P[IGRF_DIM * IGRF_DIM], dP[IGRF_DIM * IGRF_DIM];
  legendre_schmidt(theta, n_max, P);
  dlegendre_schmidt(theta, n_max, P, dP);
  const int d = n_max + 1;

  const double a = 6371.2;
  const double sin_phi = std::sin(1 * phi);
  const double cos_phi = std::cos(1 * phi);
  const double ratio = a / r;
  double fact = ratio;

  double dVr = 0, dVt = 0, dVp = 0;
  int kg = 0, kh = 0;  // 0-based rows of G/H; columns: idx+2 (1-based) -> idx-1 in the 25-col table
  const int c0 = idx - 1;

  for (int n = 1; n <= n_max; ++n) {
    double aux_dVr = 0, aux_dVt = 0, aux_dVp = 0;
    double Gnm_e0 = TS_IGRF12_G[kg][c0], dG, dH;
    if (date < 2015) {
      const double Gnm_e1 = TS_IGRF12_G[kg][c0 + 1];
      dG = (Gnm_e1 - Gnm_e0) / 5;
    } else {
      dG = TS_IGRF12_G[kg][24];
    }
    double Gnm = Gnm_e0 + dG * dt;
    kg += 1;
    aux_dVr += -(n + 1) / r * Gnm * P[n * d + 0];
    aux_dVt += Gnm * dP[n * d + 0];

    double sin_mphi = +sin_phi;
    double sin_m_1 = 0.0;
    double sin_m_2 = -sin_phi;
    double cos_mphi = +cos_phi;
    double cos_m_1 = 1.0;
    double cos_m_2 = +cos_phi;

    for (int m = 1; m <= n; ++m) {
      sin_mphi = 2 * cos_phi * sin_m_1 - sin_m_2;
      cos_mphi = 2 * cos_phi * cos_m_1 - cos_m_2;

      Gnm_e0 = TS_IGRF12_G[kg][c0];
      const double Hnm_e0 = TS_IGRF12_H[kh][c0];
      if (date < 2015) {
        const double Gnm_e1 = TS_IGRF12_G[kg][c0 + 1];
        const double Hnm_e1 = TS_IGRF12_H[kh][c0 + 1];
        dG = (Gnm_e1 - Gnm_e0) / 5;
        dH = (Hnm_e1 - Hnm_e0) / 5;
      } else {
        dG = TS_IGRF12_G[kg][24];
        dH = TS_IGRF12_H[kh][24];
      }
      Gnm = Gnm_e0 + dG * dt;
      const double Hnm = Hnm_e0 + dH * dt;
      kg += 1;
      kh += 1;

      const double GcHs = Gnm * cos_mphi + Hnm * sin_mphi;
      const double GsHc = Gnm * sin_mphi - Hnm * cos_mphi;

      aux_dVr += -(n + 1) / r * GcHs * P[n * d + m];
      aux_dVt += GcHs * dP[n * d + m];
      aux_dVp += (theta == 0) ? -m * GsHc * dP[n * d + m] : -m * GsHc * P[n * d + m];

      sin_m_2 = sin_m_1;
      sin_m_1 = sin_mphi;
      cos_m_2 = cos_m_1;
      cos_m_1 = cos_mphi;
    }
    fact *= ratio;
    aux_dVr *= fact;
    aux_dVp *= fact;
    aux_dVt *= fact;
    dVr += aux_dVr;
    dVp += aux_dVp;
    dVt += aux_dVt;
  }
  // skip rows of degrees 11..13 when n_max = 10: nothing more to do (kg/kh unused after)
  dVr *= a;
  dVp *= a;
  dVt *= a;

  out[0] = +1 / r * dVt;
  out[1] = (theta == 0) ? -1 / r * dVp : -1 / (r * std::sin(theta)) * dVp;
  out[2] = dVr;
  return IGRF_OK;
}

// reference src/igrf.jl:335-534 -- Fortran igrf12syn, 1-based arrays kept.
// out = (x north, y east, z down, f total).
inline int igrf12syn(int isv, double date, int itype, double alt, double colat, double elong,
                     double out[4]) {
  if ((date < 1900) || (date > 2025)) return IGRF_ERR_DATE;
  auto gh = [](int i1) -> double { return TS_IGRF12_GH[i1 - 1]; };  // 1-based accessor
  int fn = 0, gn = 0, kmx = 0, ll = 0, nc = 0, nmx = 0;
  double x = 0, y = 0, z = 0, t = 0, tc = 0;
  double cl[14] = {0}, sl[14] = {0}, p[106] = {0}, q[106] = {0};
  if (date < 2015) {
    t = 0.2 * (date - 1900);
    ll = (int)std::floor(t);
    t = t - ll;
    if (date < 1995) {
      nmx = 10;
      nc = 120;
      ll = nc * ll;
      kmx = 66;
    } else {
      nmx = 13;
      nc = 195;
      ll = (int)std::floor(0.2 * (date - 1995));
      ll = 120 * 19 + nc * ll;
      kmx = 105;
    }
    tc = 1 - t;
    if (isv == 1) {
      t = +0.2;
      tc = -0.2;
    }
  } else {
    t = date - 2015;
    tc = 1.0;
    if (isv == 1) {
      t = 1.0;
      tc = 0.0;
    }
    ll = 3060;
    nmx = 13;
    nc = 195;
    kmx = 105;
  }
  (void)nmx;
  double r = alt;
  double ct = std::cos(colat * M_PI / 180);
  double st = std::sin(colat * M_PI / 180);
  cl[1] = std::cos(elong * M_PI / 180);
  sl[1] = std::sin(elong * M_PI / 180);
  double cd = 1.0, sd = 0.0;
  int l = 1, m = 1, n = 0;
  double one, two, three;
  if (itype != 2) {
    const double a2 = 40680631.6, b2 = 40408296.0;
    one = a2 * st * st;
    two = b2 * ct * ct;
    three = one + two;
    const double rho = std::sqrt(three);
    r = std::sqrt(alt * (alt + 2 * rho) + (a2 * one + b2 * two) / three);
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
  for (int k = 2; k <= kmx; ++k) {
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
        one = std::sqrt(1 - 0.5 / fm);
        const int j = k - n - 1;
        p[k] = one * st * p[j];
        q[k] = one * (st * q[j] + ct * p[j]);
        cl[m] = cl[m - 1] * cl[1] - sl[m - 1] * sl[1];
        sl[m] = sl[m - 1] * cl[1] + cl[m - 1] * sl[1];
      }
    } else {
      const double gmm = (double)(m * m);
      one = std::sqrt((double)(fn * fn) - gmm);
      two = std::sqrt((double)(gn * gn) - gmm) / one;
      three = (double)(fn + gn) / one;
      const int i = k - n;
      const int j = i - n + 1;
      p[k] = three * ct * p[i] - two * p[j];
      q[k] = three * (ct * q[i] - st * p[i]) - two * q[j];
    }
    const int lm = ll + l;
    one = (tc * gh(lm) + t * gh(lm + nc)) * rr;
    if (m != 0) {
      two = (tc * gh(lm + 1) + t * gh(lm + nc + 1)) * rr;
      three = one * cl[m] + two * sl[m];
      x = x + three * q[k];
      z = z - (fn + 1) * three * p[k];
      if (st != 0) {
        y = y + (one * sl[m] - two * cl[m]) * fm * p[k] / st;
      } else {
        y = y + (one * sl[m] - two * cl[m]) * q[k] * ct;
      }
      l = l + 2;
    } else {
      x = x + one * q[k];
      z = z - (fn + 1) * one * p[k];
      l = l + 1;
    }
    m = m + 1;
  }
  one = x;
  x = x * cd + z * sd;
  z = z * cd - one * sd;
  out[0] = x;
  out[1] = y;
  out[2] = z;
  out[3] = std::sqrt(x * x + y * y + z * z);
  return IGRF_OK;
}

}  // namespace orc
