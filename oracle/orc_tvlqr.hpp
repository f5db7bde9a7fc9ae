// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_igrf.hpp header).
//
// CPU restatement of the closed-loop TVLQR replay and the Monte-Carlo glue:
//   attitude_simulation        reference src/attitude_controller.jl:1-48
//   attitude_lqr (Jacobians)   reference src/attitude_controller.jl:95-119
//   attitude_lqr (Riccati)     reference src/attitude_controller.jl:50-93
//   rk4 augmented (dt = S[end]^2, quirk Q6)  reference src/attitude_controller.jl:134-145
//   eigen_axis_slew            reference src/eigen_axis_slew.jl:1-38
//   Bryson weights             reference src/TortoiseSat.jl:157-168, src/monte_carlo.jl:165-176
//   MC post-processing         reference src/monte_carlo.jl:237-262,312-323
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_dynamics.hpp"

namespace orc {

// length(a:s:b) for Float64 ranges as used for t_sim / t_total.
inline int64_t range_len(double a, double s, double b) {
  if (b < a) return 0;
  return (int64_t)std::floor((b - a) / s + 1e-9) + 1;
}

struct TvlqrOpts {
  double dt;           // dt_lqr
  double t0, tf;       // t[1], t[end]
  double Qd[6], Qfd[6], Rd[3];
  int dt_squared = 1;  // quirk Q6: linearise with dt^2
};

// X_lqr: N x 8, U_lqr: (N-1) x 3 (row-major by knot).  noise: (N_sim-1) x 4 x 9
// scaled perturbations (nullable -> noise-free).  Outputs: X_sim N_sim x 8,
// U_sim N_sim x 3, dX N_sim x 6, K (N-1) x 3 x 6.  Returns N_sim.
inline int64_t attitude_simulation(const DynCtx& dyn, const TvlqrOpts& o, int64_t N, const double* X_lqr, const double* U_lqr,
                                   const double x0[8], const double* noise, double* X_sim, double* U_sim, double* dX,
                                   double* Kout) {
  constexpr int n = 8, m = 3;
  int64_t N_sim = range_len(o.t0, o.dt, o.tf);
  if (N_sim > N) N_sim = range_len(o.t0, o.dt, o.tf - o.dt);
  if (N_sim > N) N_sim = N;

  // ---- gains: Jacobians of rk4(gain_simulator) with step dt^2, G(q) projection, Riccati
  const double hlin = o.dt_squared ? o.dt * o.dt : o.dt;
  std::vector<double> A((size_t)N * 36, 0.0), B((size_t)N * 18, 0.0);
  using D = Dual<10>;
  for (int64_t k = 0; k < N - 1; ++k) {
    D x[n], u[m], xn[n];
    for (int i = 0; i < n; ++i) x[i] = D(X_lqr[k * n + i]);
    for (int i = 0; i < 7; ++i) x[i].d[i] = 1.0;
    for (int i = 0; i < m; ++i) {
      u[i] = D(U_lqr[k * m + i]);
      u[i].d[7 + i] = 1.0;
    }
    rk4_step<D>([&](int, const D* xx, const D* uu, D* dx) { gain_simulator<D>(dyn, xx, uu, dx); }, x, u, hlin, xn);
    double Aq[7][7], Bq[7][3];
    for (int i = 0; i < 7; ++i) {
      for (int j = 0; j < 7; ++j) Aq[i][j] = xn[i].d[j];
      for (int j = 0; j < 3; ++j) Bq[i][j] = xn[i].d[7 + j];
    }
    const double* qk = X_lqr + k * n + 3;
    const double* qn = X_lqr + (k + 1) * n + 3;
    auto Gmat = [](const double* q, double G[4][3]) {
      const double s = q[0], v0 = q[1], v1 = q[2], v2 = q[3];
      G[0][0] = -v0; G[0][1] = -v1; G[0][2] = -v2;
      // s*I + hat(v)
      G[1][0] = s;   G[1][1] = -v2; G[1][2] = v1;
      G[2][0] = v2;  G[2][1] = s;   G[2][2] = -v0;
      G[3][0] = -v1; G[3][1] = v0;  G[3][2] = s;
    };
    double Gk[4][3], Gn[4][3];
    Gmat(qk, Gk);
    Gmat(qn, Gn);
    double pGn[6][7] = {{0}}, pGk[7][6] = {{0}};
    for (int i = 0; i < 3; ++i) pGn[i][i] = 1.0, pGk[i][i] = 1.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) pGn[3 + i][3 + j] = Gn[j][i];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) pGk[3 + i][3 + j] = Gk[i][j];
    double T1[6][7];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 7; ++j) {
        double s = 0;
        for (int l = 0; l < 7; ++l) s += pGn[i][l] * Aq[l][j];
        T1[i][j] = s;
      }
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int l = 0; l < 7; ++l) s += T1[i][l] * pGk[l][j];
        A[k * 36 + i * 6 + j] = s;
      }
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int l = 0; l < 7; ++l) s += pGn[i][l] * Bq[l][j];
        B[k * 18 + i * 3 + j] = s;
      }
    }
  }
  std::vector<double> K((size_t)(N - 1) * 18);
  double S[6][6] = {{0}};
  for (int i = 0; i < 6; ++i) S[i][i] = o.Qfd[i];
  for (int64_t k = N - 2; k >= 0; --k) {
    const double* Ak = &A[k * 36];
    const double* Bk = &B[k * 18];
    double SB[6][3], SA[6][6], BSB[9], BSA[3][6];
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int l = 0; l < 6; ++l) s += S[i][l] * Bk[l * 3 + j];
        SB[i][j] = s;
      }
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int l = 0; l < 6; ++l) s += S[i][l] * Ak[l * 6 + j];
        SA[i][j] = s;
      }
    }
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int l = 0; l < 6; ++l) s += Bk[l * 3 + i] * SB[l][j];
        BSB[i * 3 + j] = s + ((i == j) ? o.Rd[i] : 0.0);
      }
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int l = 0; l < 6; ++l) s += Bk[l * 3 + i] * SA[l][j];
        BSA[i][j] = s;
      }
    }
    double Minv[9];
    inv3(BSB, Minv);
    double Kk[3][6];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int l = 0; l < 3; ++l) s += Minv[i * 3 + l] * BSA[l][j];
        Kk[i][j] = s;
        K[k * 18 + i * 6 + j] = s;
      }
    double Acl[6][6];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int l = 0; l < 3; ++l) s += Bk[i * 3 + l] * Kk[l][j];
        Acl[i][j] = Ak[i * 6 + j] - s;
      }
    double SAcl[6][6], Sn[6][6];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int l = 0; l < 6; ++l) s += S[i][l] * Acl[l][j];
        SAcl[i][j] = s;
      }
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = (i == j) ? o.Qd[i] : 0.0;
        double kr = 0;
        for (int l = 0; l < 3; ++l) kr += Kk[l][i] * o.Rd[l] * Kk[l][j];
        s += kr;
        double t = 0;
        for (int l = 0; l < 6; ++l) t += Acl[l][i] * SAcl[l][j];
        Sn[i][j] = s + t;
      }
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) S[i][j] = Sn[i][j];
  }
  if (Kout)
    for (size_t i = 0; i < K.size(); ++i) Kout[i] = K[i];

  // ---- closed-loop simulation with rk4(simulator)
  for (int64_t i = 0; i < N_sim * n; ++i) X_sim[i] = 0;
  for (int64_t i = 0; i < N_sim * m; ++i) U_sim[i] = 0;
  for (int64_t i = 0; i < N_sim * 6; ++i) dX[i] = 0;
  for (int i = 0; i < n; ++i) X_sim[i] = x0[i];
  for (int64_t k = 0; k < N_sim - 1; ++k) {
    double* dx = dX + k * 6;
    for (int i = 0; i < 3; ++i) dx[i] = X_sim[k * n + i] - X_lqr[k * n + i];
    double qi[4], qe[4];
    q_inv(X_lqr + k * n + 3, qi);
    qmult(qi, X_sim + k * n + 3, qe);
    for (int i = 0; i < 3; ++i) dx[3 + i] = qe[1 + i];
    for (int i = 0; i < m; ++i) {
      double s = 0;
      for (int j = 0; j < 6; ++j) s += K[k * 18 + i * 6 + j] * dx[j];
      U_sim[k * m + i] = U_lqr[k * m + i] - s;
    }
    const double* nz = noise ? noise + k * 36 : nullptr;
    rk4_step<double>(
        [&](int stage, const double* xx, const double* uu, double* dxx) {
          simulator(dyn, xx, uu, nz ? nz + stage * 9 : nullptr, dxx);
        },
        X_sim + k * n, U_sim + k * m, o.dt, X_sim + (k + 1) * n);
  }
  return N_sim;
}

// reference src/eigen_axis_slew.jl:1-38.  t has nt entries; outputs nt x 3, nt x 4.
// :16 reads `q_e = qmult([q2;-q2[2:4]],q1)`: the first argument is a 7-vector of which qmult (qmult.jl:1-3) reads only
// entries 1 and 2:4, i.e. the LITERAL product is qmult(q2, q1) -- no conjugate.  conj_fix = 0 reproduces that;
// conj_fix = 1 is the evident intent conj(q2) (x) q1 (identical whenever one of the two attitudes is the identity up to
// the sign of the axis, which is why the shipped scripts never notice).
inline void eigen_axis_slew(const double x0[7], const double xf[7], const double* t, int64_t nt, double* w_guess,
                            double* q_guess, int conj_fix = 0) {
  const double* q1 = x0 + 3;
  const double* q2 = xf + 3;
  const double sg = conj_fix ? -1.0 : 1.0;
  const double q2c[4] = {q2[0], sg * q2[1], sg * q2[2], sg * q2[3]};
  double qe[4];
  qmult(q2c, q1, qe);
  const double theta_f = 2 * std::acos(qe[0]);
  const double sh = std::sin(theta_f / 2);
  const double axis[3] = {-qe[1] / sh, -qe[2] / sh, -qe[3] / sh};
  const double alpha = M_PI / t[nt - 1];
  std::vector<double> theta(nt), dth(nt);
  for (int64_t i = 0; i < nt; ++i) theta[i] = theta_f * 1 / 2 * (1.0 - std::cos(alpha * t[i]));
  for (int64_t i = 0; i + 1 < nt; ++i) dth[i] = (theta[i + 1] - theta[i]) / (t[1] - t[0]);
  dth[nt - 1] = dth[nt - 2];
  for (int64_t i = 0; i < nt; ++i) {
    for (int c = 0; c < 3; ++c) w_guess[i * 3 + c] = dth[i] * axis[c];
    const double qa[4] = {std::cos(theta[i] / 2), axis[0] * std::sin(theta[i] / 2), axis[1] * std::sin(theta[i] / 2),
                          axis[2] * std::sin(theta[i] / 2)};
    qmult(q1, qa, q_guess + i * 4);
  }
}

// Bryson's-rule weights.  reference src/TortoiseSat.jl:157-168 (alpha=10) and
// src/monte_carlo.jl:165-176 (alpha=0.1); beta = 1e3.  w_guess nt x 3.
inline void bryson_weights(const double* w_guess, int64_t nt, const double J[9], double dt, double alpha, double beta,
                           double Qd[8], double Qfd[8], double Rd[3]) {
  double w_max = 0;
  for (int64_t i = 0; i < nt * 3; ++i) w_max = std::max(w_max, std::fabs(w_guess[i]));
  double tau_max = -INFINITY;  // signed maximum, as the reference computes it
  for (int64_t i = 0; i + 1 < nt; ++i) {
    double dw[3];
    for (int c = 0; c < 3; ++c) dw[c] = w_guess[(i + 1) * 3 + c] - w_guess[i * 3 + c];
    for (int r = 0; r < 3; ++r) {
      double s = J[r * 3 + 0] * dw[0];
      s += J[r * 3 + 1] * dw[1];
      s += J[r * 3 + 2] * dw[2];
      tau_max = std::max(tau_max, s / dt);
    }
  }
  const double m_max = tau_max / 1.e-5 * 1.e2;
  for (int i = 0; i < 3; ++i) {
    Qd[i] = alpha / (w_max * w_max);
    Qfd[i] = alpha / (w_max * w_max) * 10;
  }
  for (int i = 3; i < 7; ++i) {
    Qd[i] = alpha * beta;
    Qfd[i] = alpha * beta * 10;
  }
  Qd[7] = Qfd[7] = 0;
  for (int i = 0; i < 3; ++i) Rd[i] = 1 / (m_max * m_max);
}

// reference src/monte_carlo.jl:237-262.  X_sim: N_sim x 8.  Returns slew_time
// (== t_final when the trial "fails").  literal=1 keeps the `[1:3,i]` column bug
// (quirk Q12: column = 1-based trial index i).
inline double mc_slew_time(const double* X_sim, int64_t N_sim, const double q_final[4], double t_final, double time_step,
                           double w_limit, double ang_limit, int literal, int64_t trial_index_1based) {
  double slew = t_final;
  double qi[4];
  q_inv(q_final, qi);
  for (int64_t j = 1; j <= N_sim; ++j) {
    int64_t col = literal ? trial_index_1based : j;
    if (col > N_sim) col = N_sim;  // out-of-bounds in the reference; clamp
    const double* w = X_sim + (col - 1) * 8;
    const double wn = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double qe[4];
    qmult(qi, X_sim + (j - 1) * 8 + 3, qe);
    const double ang = 2 * std::acos(std::min(qe[0], 1.0));
    if (j > 10 && wn < w_limit && ang < ang_limit && slew == t_final) slew = time_step * (double)j;
  }
  return slew;
}

}  // namespace orc
