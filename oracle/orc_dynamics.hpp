// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_igrf.hpp header).
//
// CPU restatement of the reference's attitude-dynamics layer, templated on the
// scalar so that a forward-mode dual number reproduces what ForwardDiff v0.10.9
// does inside TrajectoryOptimization / attitude_lqr:
//   qmult, qrot                reference src/qmult.jl:1-3, src/qrot.jl:1-3
//   DerivFunction (8 state)    reference src/DerivFunction.jl:1-48
//   gain_simulator             reference src/gain_simulator.jl:1-53
//   simulator (noisy truth)    reference src/simulator.jl:1-42
//   attitude_dynamics (7 st.)  reference src/attitude_dynamics.jl:2-24
//   rk3 / rk4 ZOH discretisers reference src/attitude_controller.jl:122-132,178-187
#pragma once
#include <cmath>
#include <cstdint>

namespace orc {

template <int NP>
struct Dual {
  double v;
  double d[NP];
  Dual() : v(0) {
    for (int i = 0; i < NP; ++i) d[i] = 0;
  }
  Dual(double x) : v(x) {
    for (int i = 0; i < NP; ++i) d[i] = 0;
  }
};
template <int NP>
inline Dual<NP> operator+(const Dual<NP>& a, const Dual<NP>& b) {
  Dual<NP> r;
  r.v = a.v + b.v;
  for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int NP>
inline Dual<NP> operator-(const Dual<NP>& a, const Dual<NP>& b) {
  Dual<NP> r;
  r.v = a.v - b.v;
  for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int NP>
inline Dual<NP> operator-(const Dual<NP>& a) {
  Dual<NP> r;
  r.v = -a.v;
  for (int i = 0; i < NP; ++i) r.d[i] = -a.d[i];
  return r;
}
template <int NP>
inline Dual<NP> operator*(const Dual<NP>& a, const Dual<NP>& b) {
  Dual<NP> r;
  r.v = a.v * b.v;
  for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int NP>
inline Dual<NP> operator/(const Dual<NP>& a, const Dual<NP>& b) {
  Dual<NP> r;
  r.v = a.v / b.v;
  for (int i = 0; i < NP; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
template <int NP>
inline Dual<NP> operator*(double a, const Dual<NP>& b) { return Dual<NP>(a) * b; }
template <int NP>
inline Dual<NP> operator*(const Dual<NP>& a, double b) { return a * Dual<NP>(b); }
template <int NP>
inline Dual<NP> operator/(const Dual<NP>& a, double b) { return a / Dual<NP>(b); }
template <int NP>
inline Dual<NP> operator+(const Dual<NP>& a, double b) { return a + Dual<NP>(b); }
template <int NP>
inline Dual<NP> operator-(const Dual<NP>& a, double b) { return a - Dual<NP>(b); }
template <int NP>
inline Dual<NP> sqrt(const Dual<NP>& a) {
  Dual<NP> r;
  r.v = std::sqrt(a.v);
  for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] / (2 * r.v);
  return r;
}
inline double value_of(double x) { return x; }
template <int NP>
inline double value_of(const Dual<NP>& x) { return x.v; }

using std::sqrt;

// reference src/qmult.jl:1-3 -- Hamilton product, scalar first.
template <class T>
inline void qmult(const T a[4], const T b[4], T out[4]) {
  out[0] = a[0] * b[0] - (a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
  // q1[1]*q2[2:4] + q2[1]*q1[2:4] + cross(q1[2:4],q2[2:4])
  out[1] = a[0] * b[1] + b[0] * a[1] + (a[2] * b[3] - a[3] * b[2]);
  out[2] = a[0] * b[2] + b[0] * a[2] + (a[3] * b[1] - a[1] * b[3]);
  out[3] = a[0] * b[3] + b[0] * a[3] + (a[1] * b[2] - a[2] * b[1]);
}
template <class T>
inline void cross3(const T a[3], const T b[3], T out[3]) {
  out[0] = a[1] * b[2] - a[2] * b[1];
  out[1] = a[2] * b[0] - a[0] * b[2];
  out[2] = a[0] * b[1] - a[1] * b[0];
}
// reference src/qrot.jl:1-3: r + 2*cross(v, cross(v,r) + s*r)
template <class T, class R>
inline void qrot(const T q[4], const R r[3], T out[3]) {
  T rr[3] = {T(r[0]), T(r[1]), T(r[2])};
  T c1[3];
  cross3(q + 1, rr, c1);
  T w[3] = {c1[0] + q[0] * rr[0], c1[1] + q[0] * rr[1], c1[2] + q[0] * rr[2]};
  T c2[3];
  cross3(q + 1, w, c2);
  for (int i = 0; i < 3; ++i) out[i] = rr[i] + 2.0 * c2[i];
}
inline void q_inv(const double q[4], double out[4]) {
  out[0] = q[0];
  out[1] = -q[1];
  out[2] = -q[2];
  out[3] = -q[3];
}

// Everything the reference keeps in untyped globals (B_ECI, N, p.J, tf, t0).
struct DynCtx {
  const double* B_eci;  // field table, rows x 3 (Tesla)
  int64_t B_rows;
  double index_scale;  // global N in floor(Int, t*N+1)            (quirk Q1)
  double clock_rate;   // 1/(tf-t0) with the *scoping* tf          (quirk Q1)
  double J[9];         // inertia, row-major
  double Jinv[9];      // inv(p.J), row-major
};

inline const double* field_row(const DynCtx& c, double t) {
  int64_t idx = (int64_t)std::floor(t * c.index_scale + 1);  // 1-based
  if (idx < 1) idx = 1;
  if (idx > c.B_rows) idx = c.B_rows;
  return c.B_eci + (idx - 1) * 3;
}

// Shared body of DerivFunction / gain_simulator.  u_mode: 0 -> u*1e-2
// (DerivFunction.jl:37), 1 -> u/100 (gain_simulator.jl:42)            (quirk Q8)
template <class T>
inline void deriv_common(const DynCtx& c, const T x[8], const T u[3], int u_mode, T dx[8]) {
  T nq = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  T q[4] = {x[3] / nq, x[4] / nq, x[5] / nq, x[6] / nq};
  T w4[4] = {T(0.0), x[0], x[1], x[2]};
  T qd[4];
  qmult(q, w4, qd);
  const double* Bn = field_row(c, value_of(x[7]));  // floor() => zero derivative
  T BB[3];
  qrot(q, Bn, BB);
  T us[3];
  for (int i = 0; i < 3; ++i) us[i] = (u_mode == 0) ? (u[i] * 1.e-2) : (u[i] / 100.0);
  T tau[3];
  cross3(us, BB, tau);
  T Jw[3];
  for (int i = 0; i < 3; ++i) Jw[i] = c.J[i * 3 + 0] * x[0] + c.J[i * 3 + 1] * x[1] + c.J[i * 3 + 2] * x[2];
  T wJw[3];
  cross3(x, Jw, wJw);
  T rhs[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  for (int i = 0; i < 3; ++i) dx[i] = c.Jinv[i * 3 + 0] * rhs[0] + c.Jinv[i * 3 + 1] * rhs[1] + c.Jinv[i * 3 + 2] * rhs[2];
  for (int i = 0; i < 4; ++i) dx[3 + i] = 0.5 * qd[i];
  dx[7] = T(c.clock_rate);
}
// reference src/DerivFunction.jl:1-48
template <class T>
inline void DerivFunction(const DynCtx& c, const T x[8], const T u[3], T dx[8]) {
  deriv_common(c, x, u, 0, dx);
}
// reference src/gain_simulator.jl:1-53
template <class T>
inline void gain_simulator(const DynCtx& c, const T x[8], const T u[3], T dx[8]) {
  deriv_common(c, x, u, 1, dx);
}
// reference src/simulator.jl:1-42.  noise[9] = (omega_noise(3), q_noise(3), B_noise(3))
// already scaled ((.38pi/180)^2, (pi/180)^2, (1e-5)^2); noise == nullptr -> no
// perturbation at all (the reference would NaN on an exactly-zero q_noise).
inline void simulator(const DynCtx& c, const double x[8], const double u[3], const double* noise, double dx[8]) {
  double om[3] = {x[0], x[1], x[2]};
  const double nq = std::sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  double q[4] = {x[3] / nq, x[4] / nq, x[5] / nq, x[6] / nq};
  const double* Bn = field_row(c, x[7]);
  double Bf[3] = {Bn[0], Bn[1], Bn[2]};
  if (noise) {
    for (int i = 0; i < 3; ++i) om[i] = x[i] + noise[i];
    const double th = std::sqrt(noise[3] * noise[3] + noise[4] * noise[4] + noise[5] * noise[5]);
    const double qn[4] = {std::cos(th / 2), noise[3] / th * std::sin(th / 2), noise[4] / th * std::sin(th / 2),
                          noise[5] / th * std::sin(th / 2)};
    double q2[4];
    qmult(q, qn, q2);
    for (int i = 0; i < 4; ++i) q[i] = q2[i];
    for (int i = 0; i < 3; ++i) Bf[i] = Bn[i] + noise[6 + i];
  }
  const double w4[4] = {0.0, om[0], om[1], om[2]};
  double qd[4];
  qmult(q, w4, qd);
  double BB[3];
  qrot(q, Bf, BB);
  const double us[3] = {u[0] / 100, u[1] / 100, u[2] / 100};
  double tau[3];
  cross3(us, BB, tau);
  double Jw[3];
  for (int i = 0; i < 3; ++i) Jw[i] = c.J[i * 3 + 0] * om[0] + c.J[i * 3 + 1] * om[1] + c.J[i * 3 + 2] * om[2];
  double wJw[3];
  cross3(om, Jw, wJw);
  const double rhs[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  for (int i = 0; i < 3; ++i) dx[i] = c.Jinv[i * 3 + 0] * rhs[0] + c.Jinv[i * 3 + 1] * rhs[1] + c.Jinv[i * 3 + 2] * rhs[2];
  for (int i = 0; i < 4; ++i) dx[3 + i] = 0.5 * qd[i];
  dx[7] = c.clock_rate;
}

// reference src/attitude_dynamics.jl:2-24 (7 state, B_B and J passed in, no 1e-2)
inline void attitude_dynamics(const double x[7], const double u[3], const double BB[3], const double J[9],
                              const double Jinv[9], double dx[7]) {
  const double nq = std::sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  const double q[4] = {x[3] / nq, x[4] / nq, x[5] / nq, x[6] / nq};
  const double w4[4] = {0.0, x[0], x[1], x[2]};
  double qd[4];
  qmult(q, w4, qd);
  double tau[3];
  cross3(u, BB, tau);
  double Jw[3];
  for (int i = 0; i < 3; ++i) Jw[i] = J[i * 3 + 0] * x[0] + J[i * 3 + 1] * x[1] + J[i * 3 + 2] * x[2];
  double wJw[3];
  cross3(x, Jw, wJw);
  const double rhs[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  for (int i = 0; i < 3; ++i) dx[i] = Jinv[i * 3 + 0] * rhs[0] + Jinv[i * 3 + 1] * rhs[1] + Jinv[i * 3 + 2] * rhs[2];
  for (int i = 0; i < 4; ++i) dx[3 + i] = 0.5 * qd[i];
}

// reference src/attitude_controller.jl:178-187 (same tableau TrajOpt's rk3 uses)
template <class T, class F>
inline void rk3_step(F&& f, const T x[8], const T u[3], double dt, T xn[8]) {
  T k1[8], k2[8], k3[8], xs[8];
  f(x, u, k1);
  for (int i = 0; i < 8; ++i) k1[i] = k1[i] * dt;
  for (int i = 0; i < 8; ++i) xs[i] = x[i] + k1[i] / 2.0;
  f(xs, u, k2);
  for (int i = 0; i < 8; ++i) k2[i] = k2[i] * dt;
  for (int i = 0; i < 8; ++i) xs[i] = x[i] - k1[i] + 2.0 * k2[i];
  f(xs, u, k3);
  for (int i = 0; i < 8; ++i) k3[i] = k3[i] * dt;
  for (int i = 0; i < 8; ++i) xn[i] = x[i] + (k1[i] + 4.0 * k2[i] + k3[i]) / 6.0;
}
// reference src/attitude_controller.jl:122-132.  f(stage, x, u, k) lets the
// caller feed per-stage noise (simulator redraws in every stage, quirk Q7).
template <class T, class F>
inline void rk4_step(F&& f, const T x[8], const T u[3], double dt, T xn[8]) {
  T k1[8], k2[8], k3[8], k4[8], xs[8];
  f(0, x, u, k1);
  for (int i = 0; i < 8; ++i) k1[i] = k1[i] * dt;
  for (int i = 0; i < 8; ++i) xs[i] = x[i] + k1[i] / 2.0;
  f(1, xs, u, k2);
  for (int i = 0; i < 8; ++i) k2[i] = k2[i] * dt;
  for (int i = 0; i < 8; ++i) xs[i] = x[i] + k2[i] / 2.0;
  f(2, xs, u, k3);
  for (int i = 0; i < 8; ++i) k3[i] = k3[i] * dt;
  for (int i = 0; i < 8; ++i) xs[i] = x[i] + k3[i];
  f(3, xs, u, k4);
  for (int i = 0; i < 8; ++i) k4[i] = k4[i] * dt;
  for (int i = 0; i < 8; ++i) xn[i] = x[i] + (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
}

// 3x3 inverse via Gauss-Jordan with partial pivoting (inv(p.J)); for the
// diagonal presets this yields exactly 1/J_ii.
inline bool inv3(const double A[9], double out[9]) {
  double M[3][6];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      M[i][j] = A[i * 3 + j];
      M[i][3 + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < 3; ++c) {
    int p = c;
    for (int r = c + 1; r < 3; ++r)
      if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
    if (M[p][c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 6; ++j) {
        const double t = M[c][j];
        M[c][j] = M[p][j];
        M[p][j] = t;
      }
    const double piv = M[c][c];
    for (int j = 0; j < 6; ++j) M[c][j] /= piv;
    for (int r = 0; r < 3; ++r) {
      if (r == c) continue;
      const double f = M[r][c];
      if (f == 0.0) continue;
      for (int j = 0; j < 6; ++j) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i * 3 + j] = M[i][3 + j];
  return true;
}

}  // namespace orc
