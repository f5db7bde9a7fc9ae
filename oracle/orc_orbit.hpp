// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_igrf.hpp header).
//
// CPU restatement of the reference's orbit + geomagnetic environment layer:
//   kep_ECI, R_z, R_x                reference src/kep_ECI.jl:1-49
//   OrbitPlotter (ODE right-hand side) reference src/OrbitPlotter.jl:1-52
//   explicit Euler, fixed step       reference src/magnetic_toolbox.jl:51-56
//                                    (OrdinaryDiffEq v5.32.0 Euler: u += dt*f(u))
//   magnetic_simulation              reference src/magnetic_toolbox.jl:33-106
//   magnetic_gramian                 reference src/magnetic_toolbox.jl:1-12
//   condition_based_time             reference src/magnetic_toolbox.jl:14-31
//   Rz, hat                          reference src/magnetic_toolbox.jl:136-146
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_igrf.hpp"

namespace orc {

// Julia's sind/cosd reduce the argument in degrees exactly (so cosd(90)==0,
// sind(180)==0) and evaluate sin/cos on a reduced angle converted to radians
// in extended precision.  Restated with x87 long double for the conversion.
inline double sind(double x) {
  double rx = std::copysign(std::fmod(x, 360.0), x);
  const double arx = std::fabs(rx);
  auto d2r = [](double deg) -> long double { return (long double)deg * (3.14159265358979323846264338327950288L / 180.0L); };
  if (rx == 0.0) return rx;
  if (arx < 45) return (double)sinl(d2r(rx));
  if (arx <= 135) return std::copysign((double)cosl(d2r(90.0 - arx)), rx);
  if (arx == 180) return std::copysign(0.0, rx);
  if (arx < 225) return (double)sinl(d2r((180.0 - arx) * (rx < 0 ? -1.0 : 1.0)));
  if (arx <= 315) return -std::copysign((double)cosl(d2r(270.0 - arx)), rx);
  return (double)sinl(d2r(rx - std::copysign(360.0, rx)));
}
inline double cosd(double x) {
  const double rx = std::fabs(std::fmod(x, 360.0));
  auto d2r = [](double deg) -> long double { return (long double)deg * (3.14159265358979323846264338327950288L / 180.0L); };
  if (rx <= 45) return (double)cosl(d2r(rx));
  if (rx < 135) return (double)sinl(d2r(90.0 - rx));
  if (rx <= 225) return -(double)cosl(d2r(180.0 - rx));
  if (rx < 315) return (double)sinl(d2r(rx - 270.0));
  return (double)cosl(d2r(360.0 - rx));
}

struct Mat3 {
  double a[3][3];
};
inline Mat3 matmul(const Mat3& A, const Mat3& B) {
  Mat3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = A.a[i][0] * B.a[0][j];
      s += A.a[i][1] * B.a[1][j];
      s += A.a[i][2] * B.a[2][j];
      C.a[i][j] = s;
    }
  return C;
}
inline void matvec(const Mat3& A, const double v[3], double out[3]) {
  for (int i = 0; i < 3; ++i) {
    double s = A.a[i][0] * v[0];
    s += A.a[i][1] * v[1];
    s += A.a[i][2] * v[2];
    out[i] = s;
  }
}
inline Mat3 transpose(const Mat3& A) {
  Mat3 T;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T.a[i][j] = A.a[j][i];
  return T;
}
// reference src/kep_ECI.jl:37-49 (degrees)
inline Mat3 R_z_deg(double ang) {
  return Mat3{{{cosd(ang), sind(ang), 0}, {-sind(ang), cosd(ang), 0}, {0, 0, 1}}};
}
inline Mat3 R_x_deg(double ang) {
  return Mat3{{{1, 0, 0}, {0, cosd(ang), sind(ang)}, {0, -sind(ang), cosd(ang)}}};
}
// reference src/magnetic_toolbox.jl:136-140 (radians)
inline Mat3 Rz(double th) {
  return Mat3{{{std::cos(th), std::sin(th), 0}, {-std::sin(th), std::cos(th), 0}, {0, 0, 1}}};
}
inline double norm3(const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// reference src/kep_ECI.jl:1-35.  kep = [e, a km, i deg, RAAN deg, argp deg, "nu" deg].
// Mutates kep[5] like the reference (:7-8).  out = [r(3) km ; v(3) km/s].
inline void kep_ECI(double kep[6], double t0, double GM, double out[6]) {
  kep[5] = std::fmod(kep[5] + t0 * std::sqrt(GM / (kep[1] * kep[1] * kep[1])), 360.0);
  double E = kep[5] / 180 * M_PI;
  for (int i = 0; i < 100; ++i) {
    E = E - (E - kep[0] * std::sin(E) - kep[5] / 180 * M_PI) / (1 - kep[0] * std::cos(E));
  }
  const double nu = 2 * (std::atan2(std::sqrt(1 + kep[0]) * std::sin(E / 2), std::sqrt(1 - kep[0]) * std::cos(E / 2)) *
                         (180.0 / M_PI));  // 2*atand(y,x)
  const double r_c = kep[1] * (1 - kep[0] * std::cos(E));
  const double o[3] = {r_c * cosd(nu), r_c * sind(nu), r_c * 0.0};
  const double f = std::sqrt(GM * kep[1]) / r_c;
  const double od[3] = {f * -std::sin(E), f * (std::sqrt(1 - kep[0] * kep[0]) * std::cos(E)), f * 0.0};
  const Mat3 R = matmul(matmul(R_z_deg(-kep[3]), R_x_deg(-kep[2])), R_z_deg(-kep[4]));
  matvec(R, o, out);
  matvec(R, od, out + 3);
}

// reference src/OrbitPlotter.jl:1-52 -- 2-body + the literal "J2" expression (:40-42).
inline void orbit_rhs(const double x[6], double dx[6]) {
  const double GM = 3.986004418E14 * ((1.0 / 1000) * (1.0 / 1000) * (1.0 / 1000));
  const double* r = x;
  const double nr = norm3(r);
  const double J2 = 0.0010826359;
  const double nr7 = std::pow(nr, 7);
  const double rxy = r[0] * r[0] + r[1] * r[1];
  const double fJ2[3] = {J2 * r[0] / nr7 * (6 * r[2] - 1.5 * rxy), J2 * r[1] / nr7 * (6 * r[2] - 1.5 * rxy),
                         J2 * r[2] / nr7 * (3 * r[2] - 4.5 * rxy)};
  const double g = GM / (nr * nr);
  for (int i = 0; i < 3; ++i) {
    dx[i] = x[3 + i];
    dx[3 + i] = (g * -r[i] / nr) + fJ2[i];
  }
}

struct FieldOpts {
  double GM;              // km^3/s^2 (p.GM)
  double mjd;             // p.MJD
  double igrf_date;       // 2019 in the reference (magnetic_toolbox.jl:81)
  double field_radius_m;  // (alt+R_E)*1000 in the reference (quirk Q3)
  double t0, tf;
  int64_t N;
};

// reference src/magnetic_toolbox.jl:33-106.  Outputs: B (2N x 3 row-major, Tesla;
// last row stays 0), pos/vel ((2N+1) x 3 row-major by sample; may be null).
inline int magnetic_simulation(const double kep_in[6], const FieldOpts& o, double* B, double* pos, double* vel) {
  double kep[6];
  for (int i = 0; i < 6; ++i) kep[i] = kep_in[i];
  double u[6];
  kep_ECI(kep, o.t0, o.GM, u);
  const int64_t N2 = 2 * o.N;
  const double dt = (o.tf - o.t0) / (double)o.N;
  std::vector<double> P((size_t)(N2 + 1) * 3);
  for (int64_t i = 0; i <= N2; ++i) {
    for (int c = 0; c < 3; ++c) {
      P[(size_t)i * 3 + c] = u[c];
      if (pos) pos[(size_t)i * 3 + c] = u[c];
      if (vel) vel[(size_t)i * 3 + c] = u[3 + c];
    }
    double du[6];
    orbit_rhs(u, du);
    for (int c = 0; c < 6; ++c) u[c] = u[c] + dt * du[c];
  }
  for (int64_t i = 0; i < N2; ++i) B[(size_t)i * 3 + 0] = B[(size_t)i * 3 + 1] = B[(size_t)i * 3 + 2] = 0.0;
  int rc = 0;
  for (int64_t i = 0; i < N2 - 1; ++i) {
    // t = t0:(tf-t0)/N:2tf ; element i (0-based) = t0 + i*dt
    const double t = o.t0 + (double)i * dt;
    const double GMST = (280.4606 + 360.9856473 * (t / 24 / 60 / 60 + o.mjd) - 51544.5) / 180 * M_PI;
    const Mat3 ROT = Rz(GMST);
    double pe[3];
    matvec(ROT, &P[(size_t)i * 3], pe);
    const double lat = std::asin(pe[2] / norm3(pe));
    const double lon = std::atan2(pe[1], pe[0]);
    double bned[3];
    const int e = igrf12(o.igrf_date, o.field_radius_m, lat, lon, bned);
    if (e) rc = e;
    for (int c = 0; c < 3; ++c) bned[c] /= 1.e9;
    const Mat3 R_ENU_XYZ = {{{-std::sin(lon), -std::sin(lat) * std::cos(lon), std::cos(lat) * std::cos(lon)},
                             {std::cos(lon), -std::sin(lat) * std::sin(lon), std::cos(lat) * std::sin(lon)},
                             {0, std::cos(lat), std::sin(lat)}}};
    const Mat3 NED_ENU = {{{0, 1, 0}, {1, 0, 0}, {0, 0, -1}}};
    // left-to-right: ((Rz' * R_ENU_to_XYZ) * NED_to_ENU) * B
    const Mat3 M = matmul(matmul(transpose(ROT), R_ENU_XYZ), NED_ENU);
    matvec(M, bned, &B[(size_t)i * 3]);
  }
  return rc;
}

// hat(x)*hat(x)' accumulated as the reference does (magnetic_toolbox.jl:1-12):
// G_1 = hat(B_1)hat(B_1)', G_i = G_{i-1} + hat(B_i)hat(B_i)'*dt.
inline void hat_hatT(const double b[3], double out[3][3]) {
  const double H[3][3] = {{0, -b[2], b[1]}, {b[2], 0, -b[0]}, {-b[1], b[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = H[i][0] * H[j][0];
      s += H[i][1] * H[j][1];
      s += H[i][2] * H[j][2];
      out[i][j] = s;
    }
}
inline void magnetic_gramian(const double* B, int64_t rows, double dt, double* G /*rows x 9*/) {
  double acc[3][3];
  hat_hatT(B, acc);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[i * 3 + j] = acc[i][j];
  for (int64_t k = 1; k < rows; ++k) {
    double h[3][3];
    hat_hatT(B + k * 3, h);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        acc[i][j] = acc[i][j] + h[i][j] * dt;
        G[k * 9 + i * 3 + j] = acc[i][j];
      }
  }
}

// 2-norm condition number of a symmetric 3x3 matrix (Julia cond() = ratio of
// extreme singular values; for a symmetric matrix |eigenvalues|).  Cyclic
// Jacobi eigenvalue iteration.  Singular -> +inf.
inline double cond_sym3(const double Gin[9]) {
  double A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = 0.5 * (Gin[i * 3 + j] + Gin[j * 3 + i]);
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    const double dia = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-36 * dia || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        const double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A*J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J'*A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
      }
  }
  const double e0 = std::fabs(A[0][0]), e1 = std::fabs(A[1][1]), e2 = std::fabs(A[2][2]);
  const double mx = std::max(e0, std::max(e1, e2)), mn = std::min(e0, std::min(e1, e2));
  if (mn == 0.0) return INFINITY;
  return mx / mn;
}

// reference src/magnetic_toolbox.jl:14-31: first 1-based index with cond < cutoff, else 0.
inline int64_t condition_based_time(const double* G, int64_t rows, double cutoff) {
  for (int64_t i = 0; i < rows; ++i)
    if (cond_sym3(G + i * 9) < cutoff) return i + 1;
  return 0;
}

}  // namespace orc
