// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_igrf.hpp header).
//
// CPU restatement of the paper's comparison controller:
//   psiaki_controller          reference src/comparison/psiaki_dynamics.jl:1-26
//   rk4_psiaki                 reference src/comparison/psiaki_dynamics.jl:63-73
//   attitude_dynamics_linear   reference src/attitude_dynamics.jl:26-48
//   the closed loop            reference src/comparison/psiaki2005.jl:116-164
// Pinned by tests/golden/ref_fixtures.json ("psiaki" section: numpy transliteration of those lines).
#pragma once
#include "orc_dynamics.hpp"

namespace orc {

inline void attitude_dynamics_linear(const double x[7], const double u[3], const double xl[7], const double BB[3], const double J[9],
                                     const double Jinv[9], double dx[7]) {
  const double nq = std::sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  const double q[4] = {x[3] / nq, x[4] / nq, x[5] / nq, x[6] / nq};
  const double w4[4] = {0.0, xl[3], xl[4], xl[5]};   // q_dot = 0.5*qmult(q,[0; x_linear[4:6]])  (:37)
  double qd[4], tau[3], Jw[3], wJw[3];
  qmult(q, w4, qd);
  cross3(u, BB, tau);
  for (int i = 0; i < 3; ++i) Jw[i] = J[i * 3 + 0] * x[0] + J[i * 3 + 1] * x[1] + J[i * 3 + 2] * x[2];
  cross3(x, Jw, wJw);
  const double r[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  for (int i = 0; i < 3; ++i) dx[i] = Jinv[i * 3 + 0] * r[0] + Jinv[i * 3 + 1] * r[1] + Jinv[i * 3 + 2] * r[2];
  for (int i = 0; i < 4; ++i) dx[3 + i] = 0.5 * qd[i];
}

inline void psiaki_controller(double C1, double C2, const double Jinv[9], const double q[4], const double w[3], const double Bm[3],
                              double m[3]) {
  double T[3];
  for (int i = 0; i < 3; ++i) {
    const double jq = Jinv[i * 3 + 0] * q[1] + Jinv[i * 3 + 1] * q[2] + Jinv[i * 3 + 2] * q[3];
    T[i] = -(C1 * w[i] + C2 * jq);
  }
  double c[3];
  cross3(Bm, T, c);
  const double nb = std::sqrt(Bm[0] * Bm[0] + Bm[1] * Bm[1] + Bm[2] * Bm[2]);
  for (int i = 0; i < 3; ++i) m[i] = c[i] / (nb * nb);
}

inline void rk4_psiaki(const double x[7], double dt, const double u[3], const double BB[3], const double J[9], const double Jinv[9],
                       double xn[7]) {
  double f1[7], f2[7], f3[7], f4[7], xs[7];
  attitude_dynamics(x, u, BB, J, Jinv, f1);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + .5 * f1[i] * dt;
  attitude_dynamics(xs, u, BB, J, Jinv, f2);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + .5 * f2[i] * dt;
  attitude_dynamics(xs, u, BB, J, Jinv, f3);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + f3[i] * dt;
  attitude_dynamics(xs, u, BB, J, Jinv, f4);
  for (int i = 0; i < 7; ++i) xn[i] = x[i] + 1.0 / 6 * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]) * dt;
}

// X: N x 7, M: N x 3, Qe: N x 4 (row-major by step); w_guess N x 3, q_guess N x 4, B N x 3.
inline void psiaki_pd_simulation(int64_t N, const double x0[7], const double* w_guess, const double* q_guess, const double* B, const double J[9],
                                 double dt, double C1, double C2, double* X, double* M, double* Qe) {
  double Jinv[9];
  inv3(J, Jinv);
  double x[7];
  for (int i = 0; i < 7; ++i) X[i] = x[i] = x0[i];
  for (int64_t i = 0; i < N * 3; ++i) M[i] = 0.0;
  for (int64_t i = 0; i < N * 4; ++i) Qe[i] = 0.0;
  if (N < 2) return;
  const double z[3] = {0, 0, 0};
  double dx[7];
  attitude_dynamics(x, z, B, J, Jinv, dx);
  for (int i = 0; i < 7; ++i) X[7 + i] = x[i] = x[i] + dt * dx[i];
  for (int64_t k = 1; k < N - 1; ++k) {
    const double qi[4] = {x[3], -x[4], -x[5], -x[6]};
    double Bm[3], wbar[3], qbar[4], m[3], xn[7];
    qrot(qi, B + k * 3, Bm);
    for (int i = 0; i < 3; ++i) wbar[i] = w_guess[k * 3 + i] - x[i];
    qmult(x + 3, q_guess + k * 4, qbar);
    psiaki_controller(C1, C2, Jinv, qbar, wbar, Bm, m);
    rk4_psiaki(x, dt, m, Bm, J, Jinv, xn);
    const double nq = std::sqrt(xn[3] * xn[3] + xn[4] * xn[4] + xn[5] * xn[5] + xn[6] * xn[6]);
    for (int i = 3; i < 7; ++i) xn[i] = xn[i] / nq;
    for (int i = 0; i < 7; ++i) X[(k + 1) * 7 + i] = x[i] = xn[i];
    for (int i = 0; i < 3; ++i) M[k * 3 + i] = m[i];
    for (int i = 0; i < 4; ++i) Qe[k * 4 + i] = qbar[i];
  }
}

}  // namespace orc
