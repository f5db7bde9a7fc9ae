// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_igrf.hpp header).
//
// CPU restatement of the Augmented-Lagrangian iLQR solve the reference obtains
// from TrajectoryOptimization.jl v0.1.2 (Manifest.toml:994-998; NOT vendored in
// /root/reference).  Call sites restated: reference src/TortoiseSat.jl:145-146
// (Model + rk3), :169 (LQRObjective), :178-188 (BoundConstraint u in [-1,1] on
// knots 1..N-1, goal_constraint(xf) on knot N), :190-199 (Problem, U0, AL solver
// with 50 inner / 20 outer iterations, solve!).
//
// PARITY UNPINNED: the package source is absent and the reference ships no
// golden vectors for this path, so the algorithm below is the frozen
// specification of SURVEY.md Appendix C (assumptions A1..A10), implemented
// literally with the full 8-state model and forward-mode dual-number Jacobians
// (what ForwardDiff does).  The CUDA path is checked against THIS.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "orc_dynamics.hpp"

namespace orc {

struct IlqrOpts {
  int max_outer = 20;                // opts_al.iterations          (TortoiseSat.jl:196)
  int max_inner = 50;                // opts_al.opts_uncon.iterations (TortoiseSat.jl:195)
  int max_linesearch = 20;           // iterations_linesearch
  double cost_tol = 1e-4;            // cost_tolerance
  double cost_tol_intermediate = 1e-3;
  double grad_tol = 1e-5;            // gradient_norm_tolerance
  double grad_tol_intermediate = 1e-5;
  double constraint_tol = 1e-3;      // constraint_tolerance
  double penalty_initial = 1.0;
  double penalty_scaling = 10.0;
  double penalty_max = 1e8;
  double dual_max = 1e8;
  double ls_lower = 1e-8, ls_upper = 10.0;
  double bp_reg_increase = 1.6, bp_reg_max = 1e8, bp_reg_min = 1e-8, bp_reg_fp = 10.0;
  double max_cost_value = 1e8;
  double max_state_value = 1e8, max_control_value = 1e8;
  int dJ_counter_limit = 10;
  int stage_cost_dt = 0;             // A1: 0 = stage cost not scaled by dt (default)
  int goal_mask = 0x7F;              // Q2: bit i set -> terminal equality on state i (0xFF = literal)
  double u_max = 1.0, u_min = -1.0;  // BoundConstraint(n,m,u_max=1,u_min=-1)
  // Assumption registry of SURVEY.md App. C: 0 = the frozen default, 1 = the named alternative (flip-tested)
  int a2_active_ge = 0;            // A2: inequality active when c >= 0 (default: c > 0) or lambda > 0
  int a3_grad_over_N = 0;          // A3: Todorov gradient averaged over N (default: N-1 controls)
  int a4_no_intermediate = 0;      // A4: final tolerances on every outer iteration (default: intermediate on all but the last)
  int a5_dual_active_only = 0;     // A5: dual update only on active inequalities (default: all, then max(0,.))
  int a6_penalty_conditional = 0;  // A6: penalties scaled only if c_max > ratio * previous c_max (default: every outer iteration)
  int a7_carry_cost = 0;           // A7: J_prev of an inner solve = last cost under the OLD multipliers (default: re-evaluated)
  double constraint_decrease_ratio = 0.25;
  // Quaternion-aware variant: what reference src/monte_carlo.jl:158,192 asks of its forked solver with
  // Model(DerivFunction, n, m, quaternion_error, quaternion_expansion) and opts.sat_att = true.  The fork is not in
  // /root/reference, so the way the two hooks enter the solver is FROZEN HERE (parity unpinned), from the hooks' own
  // bodies (src/quaternion_toolbox.jl:15-75) and from the same authors' TVLQR, which projects the same way
  // (src/attitude_controller.jl:59-81):
  //   QA1 state difference in the feedback law: dx = quaternion_error(x, xbar) = [w - wbar; MRP(q_inv(qbar) (x) q); 0]
  //       (7-vector: the clock state has no entry -- quaternion_toolbox.jl:63-75 fills 1:6 of zeros(7))
  //   QA2 expansion in error coordinates with E(x) = perm_Gk = [I3 0; 0 G(q); 0 0] (8x7), G(q) = [-v'; s I + hat(v)]:
  //       A_e = E(x_{k+1})' A E(x_k), B_e = E(x_{k+1})' B, l_x -> E(x_k)' l_x, l_xx -> E(x_k)' l_xx E(x_k)
  //       (quaternion_toolbox.jl:15-38: perm_Gn * cost.Q * perm_Gk etc.; G from the raw, un-normalised quaternion)
  //   QA3 terminal expansion likewise with E(x_N) (quaternion_toolbox.jl:40-52); constraint terms of the AL follow
  //       their cost terms through the same projection
  // Cost, constraints, multiplier updates and line search are those of the default solver.  Gains come out in error
  // coordinates (3 x 7, stored 3 x 8 with a zero last column).
  int quat_error = 0;
};

enum {
  ST_CONVERGED = 0,       // c_max < constraint_tolerance
  ST_MAX_OUTER = 1,       // ran all outer iterations
  ST_COST_BLOWUP = 2,     // J > max_cost_value (TrajOpt would error())
  ST_REG_MAX = 3,         // regularisation exceeded bp_reg_max repeatedly
  ST_NAN = 4
};

struct IlqrOutcome {
  int32_t status, outer_iters, inner_iters, ls_rollouts;
  int64_t N;
  double J, c_max, t_final, slew_time, flops;
};

struct IlqrProblem {
  int64_t N;       // knots
  double dt;
  double x0[8], xf[8];
  double Qd[8], Qfd[8], Rd[3];  // diagonal LQR weights
  DynCtx dyn;
};

namespace detail {
constexpr int n = 8, m = 3, nb = 6;

struct Work {
  int64_t N;
  std::vector<double> X, U, Xb, Ub, K, d, A, B, lam_b, mu_b;
  double lam_g[n], mu_g[n];
  Work(int64_t N_) : N(N_), X(N_ * n), U((N_ - 1) * m), Xb(N_ * n), Ub((N_ - 1) * m), K((N_ - 1) * m * n),
                     d((N_ - 1) * m), A((N_ - 1) * n * n), B((N_ - 1) * n * m), lam_b((N_ - 1) * nb), mu_b((N_ - 1) * nb) {}
};

inline void step(const IlqrProblem& p, const double* x, const double* u, double* xn) {
  rk3_step<double>([&](const double* xx, const double* uu, double* dx) { DerivFunction<double>(p.dyn, xx, uu, dx); }, x, u,
                   p.dt, xn);
}

inline void bound_c(const IlqrOpts& o, const double* u, double c[nb]) {
  for (int i = 0; i < 3; ++i) {
    c[i] = u[i] - o.u_max;
    c[3 + i] = o.u_min - u[i];
  }
}

// AL cost of a trajectory; also returns c_max.
inline double al_cost(const IlqrProblem& p, const IlqrOpts& o, const Work& w, const double* X, const double* U, double* c_max_out) {
  double Jc = 0, cmax = 0;
  const double sc = o.stage_cost_dt ? p.dt : 1.0;
  for (int64_t k = 0; k < w.N - 1; ++k) {
    const double* x = X + k * n;
    const double* u = U + k * m;
    double l = 0;
    for (int i = 0; i < n; ++i) {
      const double e = x[i] - p.xf[i];
      l += 0.5 * p.Qd[i] * e * e;
    }
    for (int i = 0; i < m; ++i) l += 0.5 * p.Rd[i] * u[i] * u[i];
    Jc += l * sc;
    double c[nb];
    bound_c(o, u, c);
    for (int i = 0; i < nb; ++i) {
      const double lam = w.lam_b[k * nb + i], mu = w.mu_b[k * nb + i];
      const bool act = (o.a2_active_ge ? (c[i] >= 0.0) : (c[i] > 0.0)) || (lam > 0.0);
      Jc += lam * c[i] + (act ? 0.5 * mu * c[i] * c[i] : 0.0);
      cmax = std::max(cmax, std::max(0.0, c[i]));
    }
  }
  const double* x = X + (w.N - 1) * n;
  for (int i = 0; i < n; ++i) {
    const double e = x[i] - p.xf[i];
    Jc += 0.5 * p.Qfd[i] * e * e;
    if (o.goal_mask & (1 << i)) {
      Jc += w.lam_g[i] * e + 0.5 * w.mu_g[i] * e * e;
      cmax = std::max(cmax, std::fabs(e));
    }
  }
  if (c_max_out) *c_max_out = cmax;
  return Jc;
}

// QA2: E(x) as an 8 x 8 array whose last column is zero (the error state has 7 entries; entry 7 is the always-zero
// slot quaternion_error leaves at the end).
inline void quat_E(const double* x, double E[n][n]) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) E[i][j] = 0.0;
  for (int i = 0; i < 3; ++i) E[i][i] = 1.0;
  const double s = x[3], v1 = x[4], v2 = x[5], v3 = x[6];
  const double G[4][3] = {{-v1, -v2, -v3}, {s, -v3, v2}, {v3, s, -v1}, {-v2, v1, s}};
  for (int r = 0; r < 4; ++r)
    for (int j = 0; j < 3; ++j) E[3 + r][3 + j] = G[r][j];
}
// QA1: quaternion_error(X1, X2), quaternion_toolbox.jl:63-75, padded to 8 entries
inline void quat_error(const double* X1, const double* X2, double dx[n]) {
  for (int i = 0; i < n; ++i) dx[i] = 0.0;
  for (int i = 0; i < 3; ++i) dx[i] = X1[i] - X2[i];
  const double qi[4] = {X2[3], -X2[4], -X2[5], -X2[6]};
  double qe[4];
  qmult<double>(qi, X1 + 3, qe);
  for (int i = 0; i < 3; ++i) dx[3 + i] = qe[1 + i] / (1.0 + qe[0]);
}

inline void jacobians(const IlqrProblem& p, Work& w) {
  using D = Dual<n + m>;
  for (int64_t k = 0; k < w.N - 1; ++k) {
    D x[n], u[m], xn[n];
    for (int i = 0; i < n; ++i) {
      x[i] = D(w.X[k * n + i]);
      x[i].d[i] = 1.0;
    }
    for (int i = 0; i < m; ++i) {
      u[i] = D(w.U[k * m + i]);
      u[i].d[n + i] = 1.0;
    }
    rk3_step<D>([&](const D* xx, const D* uu, D* dx) { DerivFunction<D>(p.dyn, xx, uu, dx); }, x, u, p.dt, xn);
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) w.A[k * n * n + i * n + j] = xn[i].d[j];
      for (int j = 0; j < m; ++j) w.B[k * n * m + i * m + j] = xn[i].d[n + j];
    }
  }
}

inline bool chol3(const double Ain[9], double L[9]) {
  for (int i = 0; i < 9; ++i) L[i] = 0;
  for (int j = 0; j < 3; ++j) {
    double s = Ain[j * 3 + j];
    for (int k = 0; k < j; ++k) s -= L[j * 3 + k] * L[j * 3 + k];
    if (!(s > 0.0)) return false;
    L[j * 3 + j] = std::sqrt(s);
    for (int i = j + 1; i < 3; ++i) {
      double t = Ain[i * 3 + j];
      for (int k = 0; k < j; ++k) t -= L[i * 3 + k] * L[j * 3 + k];
      L[i * 3 + j] = t / L[j * 3 + j];
    }
  }
  return true;
}
inline void chol3_solve(const double L[9], const double b[3], double x[3]) {
  double y[3];
  for (int i = 0; i < 3; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * 3 + k] * y[k];
    y[i] = s / L[i * 3 + i];
  }
  for (int i = 2; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < 3; ++k) s -= L[k * 3 + i] * x[k];
    x[i] = s / L[i * 3 + i];
  }
}

struct Reg {
  double rho = 0, drho = 0;
};
inline void reg_increase(const IlqrOpts& o, Reg& r) {
  r.drho = std::max(r.drho * o.bp_reg_increase, o.bp_reg_increase);
  r.rho = std::max(r.rho * r.drho, o.bp_reg_min);
}
inline void reg_decrease(const IlqrOpts& o, Reg& r) {
  r.drho = std::min(r.drho / o.bp_reg_increase, 1.0 / o.bp_reg_increase);
  r.rho = r.rho * r.drho * ((r.rho * r.drho > o.bp_reg_min) ? 1.0 : 0.0);
}

// Backward Riccati sweep (App. C step 3).  Returns false if regularisation ran away.
inline bool backward_pass(const IlqrProblem& p, const IlqrOpts& o, Work& w, Reg& reg, double dV[2]) {
  const double sc = o.stage_cost_dt ? p.dt : 1.0;
  int restarts = 0;
restart:
  dV[0] = dV[1] = 0;
  double Sxx[n][n] = {{0}}, Sx[n];
  {
    const double* x = &w.X[(w.N - 1) * n];
    for (int i = 0; i < n; ++i) {
      const double e = x[i] - p.xf[i];
      Sxx[i][i] = p.Qfd[i];
      Sx[i] = p.Qfd[i] * e;
      if (o.goal_mask & (1 << i)) {
        Sxx[i][i] += w.mu_g[i];
        Sx[i] += w.lam_g[i] + w.mu_g[i] * e;
      }
    }
    if (o.quat_error) {  // QA3: E_N' Sxx E_N, E_N' Sx
      double E[n][n], T[n][n], Sp[n][n], sp[n];
      quat_E(x, E);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double t = 0;
          for (int l = 0; l < n; ++l) t += Sxx[i][l] * E[l][j];
          T[i][j] = t;
        }
      for (int i = 0; i < n; ++i) {
        double t = 0;
        for (int l = 0; l < n; ++l) t += E[l][i] * Sx[l];
        sp[i] = t;
        for (int j = 0; j < n; ++j) {
          double u = 0;
          for (int l = 0; l < n; ++l) u += E[l][i] * T[l][j];
          Sp[i][j] = u;
        }
      }
      for (int i = 0; i < n; ++i) {
        Sx[i] = sp[i];
        for (int j = 0; j < n; ++j) Sxx[i][j] = Sp[i][j];
      }
    }
  }
  for (int64_t k = w.N - 2; k >= 0; --k) {
    const double* A = &w.A[k * n * n];
    const double* B = &w.B[k * n * m];
    const double* x = &w.X[k * n];
    const double* u = &w.U[k * m];
    double lx[n], lu[m], luu[m], lxx[n][n];
    for (int i = 0; i < n; ++i) lx[i] = sc * p.Qd[i] * (x[i] - p.xf[i]);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) lxx[i][j] = (i == j) ? sc * p.Qd[i] : 0.0;
    double Ae[n * n], Be[n * m];
    if (o.quat_error) {  // QA2: everything that multiplies a state difference moves to error coordinates
      double E0[n][n], E1[n][n], T[n][n];
      quat_E(x, E0);
      quat_E(&w.X[(k + 1) * n], E1);
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double t = 0;
          for (int l = 0; l < n; ++l) t += A[i * n + l] * E0[l][j];
          T[i][j] = t;
        }
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
          double t = 0;
          for (int l = 0; l < n; ++l) t += E1[l][i] * T[l][j];
          Ae[i * n + j] = t;
        }
        for (int j = 0; j < m; ++j) {
          double t = 0;
          for (int l = 0; l < n; ++l) t += E1[l][i] * B[l * m + j];
          Be[i * m + j] = t;
        }
      }
      double lxp[n], L1[n][n], L2[n][n];
      for (int i = 0; i < n; ++i) {
        double t = 0;
        for (int l = 0; l < n; ++l) t += E0[l][i] * lx[l];
        lxp[i] = t;
        for (int j = 0; j < n; ++j) {
          double v = 0;
          for (int l = 0; l < n; ++l) v += lxx[i][l] * E0[l][j];
          L1[i][j] = v;
        }
      }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double v = 0;
          for (int l = 0; l < n; ++l) v += E0[l][i] * L1[l][j];
          L2[i][j] = v;
        }
      for (int i = 0; i < n; ++i) {
        lx[i] = lxp[i];
        for (int j = 0; j < n; ++j) lxx[i][j] = L2[i][j];
      }
      A = Ae;
      B = Be;
    }
    double c[nb];
    bound_c(o, u, c);
    for (int i = 0; i < m; ++i) {
      lu[i] = sc * p.Rd[i] * u[i];
      luu[i] = sc * p.Rd[i];
      const double lp = w.lam_b[k * nb + i], mp = w.mu_b[k * nb + i];
      const double ln = w.lam_b[k * nb + 3 + i], mn = w.mu_b[k * nb + 3 + i];
      const bool ap = (o.a2_active_ge ? (c[i] >= 0.0) : (c[i] > 0.0)) || (lp > 0.0);
      const bool an = (o.a2_active_ge ? (c[3 + i] >= 0.0) : (c[3 + i] > 0.0)) || (ln > 0.0);
      lu[i] += (lp + (ap ? mp * c[i] : 0.0)) - (ln + (an ? mn * c[3 + i] : 0.0));
      luu[i] += (ap ? mp : 0.0) + (an ? mn : 0.0);
    }
    // Qx = lx + A'Sx ; Qu = lu + B'Sx
    double Qx[n], Qu[m], SA[n][n], SB[n][m], Qxx[n][n], Quu[m][m], Qux[m][n];
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int i = 0; i < n; ++i) s += A[i * n + j] * Sx[i];
      Qx[j] = lx[j] + s;
    }
    for (int j = 0; j < m; ++j) {
      double s = 0;
      for (int i = 0; i < n; ++i) s += B[i * m + j] * Sx[i];
      Qu[j] = lu[j] + s;
    }
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += Sxx[i][l] * A[l * n + j];
        SA[i][j] = s;
      }
      for (int j = 0; j < m; ++j) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += Sxx[i][l] * B[l * m + j];
        SB[i][j] = s;
      }
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += A[l * n + i] * SA[l][j];
        Qxx[i][j] = s + lxx[i][j];
      }
    for (int i = 0; i < m; ++i) {
      for (int j = 0; j < m; ++j) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += B[l * m + i] * SB[l][j];
        Quu[i][j] = s + ((i == j) ? luu[i] : 0.0);
      }
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += B[l * m + i] * SA[l][j];
        Qux[i][j] = s;
      }
    }
    double Qr[9], L[9];
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Qr[i * 3 + j] = 0.5 * (Quu[i][j] + Quu[j][i]) + ((i == j) ? reg.rho : 0.0);
    if (!chol3(Qr, L)) {
      reg_increase(o, reg);
      if (reg.rho > o.bp_reg_max || ++restarts > 200) return false;
      goto restart;
    }
    double Kk[m][n], dk[m];
    for (int j = 0; j < n; ++j) {
      const double b[3] = {-Qux[0][j], -Qux[1][j], -Qux[2][j]};
      double xcol[3];
      chol3_solve(L, b, xcol);
      for (int i = 0; i < m; ++i) Kk[i][j] = xcol[i];
    }
    {
      const double b[3] = {-Qu[0], -Qu[1], -Qu[2]};
      chol3_solve(L, b, dk);
    }
    for (int i = 0; i < m; ++i) {
      for (int j = 0; j < n; ++j) w.K[k * m * n + i * n + j] = Kk[i][j];
      w.d[k * m + i] = dk[i];
    }
    // Sx = Qx + K'Quu d + K'Qu + Qux'd ; Sxx = Qxx + K'Quu K + K'Qux + Qux'K
    double Quud[m], QuuK[m][n];
    for (int i = 0; i < m; ++i) {
      double s = 0;
      for (int l = 0; l < m; ++l) s += Quu[i][l] * dk[l];
      Quud[i] = s;
      for (int j = 0; j < n; ++j) {
        double t = 0;
        for (int l = 0; l < m; ++l) t += Quu[i][l] * Kk[l][j];
        QuuK[i][j] = t;
      }
    }
    double Sxn[n], Sxxn[n][n];
    for (int i = 0; i < n; ++i) {
      double s = Qx[i];
      for (int l = 0; l < m; ++l) s += Kk[l][i] * Quud[l];
      for (int l = 0; l < m; ++l) s += Kk[l][i] * Qu[l];
      for (int l = 0; l < m; ++l) s += Qux[l][i] * dk[l];
      Sxn[i] = s;
      for (int j = 0; j < n; ++j) {
        double t = Qxx[i][j];
        for (int l = 0; l < m; ++l) t += Kk[l][i] * QuuK[l][j];
        for (int l = 0; l < m; ++l) t += Kk[l][i] * Qux[l][j];
        for (int l = 0; l < m; ++l) t += Qux[l][i] * Kk[l][j];
        Sxxn[i][j] = t;
      }
    }
    for (int i = 0; i < n; ++i) {
      Sx[i] = Sxn[i];
      for (int j = 0; j < n; ++j) Sxx[i][j] = 0.5 * (Sxxn[i][j] + Sxxn[j][i]);
    }
    for (int l = 0; l < m; ++l) {
      dV[0] += dk[l] * Qu[l];
      dV[1] += 0.5 * dk[l] * Quud[l];
    }
  }
  reg_decrease(o, reg);
  return true;
}

// rollout with feedback; false if states/controls exceed the divergence bound (A10).
inline bool rollout(const IlqrProblem& p, const IlqrOpts& o, Work& w, double alpha) {
  for (int i = 0; i < n; ++i) w.Xb[i] = p.x0[i];
  for (int64_t k = 0; k < w.N - 1; ++k) {
    double dx[n];
    if (o.quat_error)
      quat_error(&w.Xb[k * n], &w.X[k * n], dx);
    else
      for (int i = 0; i < n; ++i) dx[i] = w.Xb[k * n + i] - w.X[k * n + i];
    for (int i = 0; i < m; ++i) {
      double s = w.U[k * m + i];
      for (int j = 0; j < n; ++j) s += w.K[k * m * n + i * n + j] * dx[j];
      s += alpha * w.d[k * m + i];
      w.Ub[k * m + i] = s;
    }
    step(p, &w.Xb[k * n], &w.Ub[k * m], &w.Xb[(k + 1) * n]);
    double mx = 0, mu = 0;
    for (int i = 0; i < n; ++i) mx = std::max(mx, std::fabs(w.Xb[(k + 1) * n + i]));
    for (int i = 0; i < m; ++i) mu = std::max(mu, std::fabs(w.Ub[k * m + i]));
    if (!(mx < o.max_state_value) || !(mu < o.max_control_value)) return false;  // also catches NaN
  }
  return true;
}
}  // namespace detail

// X: N x 8, U: (N-1) x 3, K: (N-1) x 3 x 8 (nullable), row-major by knot.
inline void alilqr_solve(const IlqrProblem& p, const IlqrOpts& o, const double* U0, double* Xout, double* Uout, double* Kout,
                         IlqrOutcome* out) {
  using namespace detail;
  Work w(p.N);
  const int64_t N = p.N;
  for (int64_t i = 0; i < (N - 1) * m; ++i) w.U[i] = U0 ? U0[i] : 0.0;
  for (int64_t i = 0; i < (N - 1) * nb; ++i) {
    w.lam_b[i] = 0;
    w.mu_b[i] = o.penalty_initial;
  }
  for (int i = 0; i < n; ++i) {
    w.lam_g[i] = 0;
    w.mu_g[i] = o.penalty_initial;
  }
  // initial rollout (open loop)
  for (int i = 0; i < n; ++i) w.X[i] = p.x0[i];
  for (int64_t k = 0; k < N - 1; ++k) step(p, &w.X[k * n], &w.U[k * m], &w.X[(k + 1) * n]);

  int status = ST_MAX_OUTER, outer = 0, inner_total = 0, ls_total = 0;
  double J = 0, c_max = 0, c_max_prev = INFINITY;
  for (int oi = 1; oi <= o.max_outer; ++oi) {
    outer = oi;
    const bool last = (oi == o.max_outer) || o.a4_no_intermediate;
    const double ctol = last ? o.cost_tol : o.cost_tol_intermediate;
    const double gtol = last ? o.grad_tol : o.grad_tol_intermediate;
    Reg reg;
    double J_prev = al_cost(p, o, w, w.X.data(), w.U.data(), nullptr);
    if (o.a7_carry_cost && oi > 1) J_prev = J;  // A7 alternative: the cost carried over from the previous inner solve
    J = J_prev;
    int dJ_zero = 0;
    bool abort_trial = false;
    for (int it = 1; it <= o.max_inner; ++it) {
      ++inner_total;
      jacobians(p, w);
      double dV[2];
      if (!backward_pass(p, o, w, reg, dV)) {
        status = ST_REG_MAX;
        abort_trial = true;
        break;
      }
      // forward pass / line search
      double alpha = 1.0, z = -1.0, expected = 0.0, Jn = INFINITY;
      int iter = 0;
      bool accepted = true;
      while ((z <= o.ls_lower || z > o.ls_upper) && (Jn >= J_prev)) {
        if (iter > o.max_linesearch) {
          accepted = false;
          Jn = al_cost(p, o, w, w.X.data(), w.U.data(), nullptr);
          z = 0;
          alpha = 0;
          expected = 0;
          reg_increase(o, reg);
          reg.rho += o.bp_reg_fp;
          break;
        }
        ++ls_total;
        const bool ok = rollout(p, o, w, alpha);
        if (!ok) {
          ++iter;
          alpha /= 2.0;
          continue;
        }
        Jn = al_cost(p, o, w, w.Xb.data(), w.Ub.data(), nullptr);
        expected = -alpha * (dV[0] + alpha * dV[1]);
        z = (expected > 0) ? (J_prev - Jn) / expected : -1.0;
        ++iter;
        alpha /= 2.0;
      }
      if (accepted) {
        w.X = w.Xb;
        w.U = w.Ub;
      }
      if (!(Jn == Jn)) {
        status = ST_NAN;
        abort_trial = true;
        break;
      }
      if (Jn > o.max_cost_value) {
        J = Jn;
        status = ST_COST_BLOWUP;
        abort_trial = true;
        break;
      }
      const double dJ = std::fabs(Jn - J_prev);
      J_prev = Jn;
      J = Jn;
      if (dJ == 0) ++dJ_zero; else dJ_zero = 0;
      // Todorov gradient with the updated controls (A3: mean over N-1 knots)
      double g = 0;
      for (int64_t k = 0; k < N - 1; ++k) {
        double mxv = 0;
        for (int i = 0; i < m; ++i) mxv = std::max(mxv, std::fabs(w.d[k * m + i]) / (std::fabs(w.U[k * m + i]) + 1.0));
        g += mxv;
      }
      g /= (double)(o.a3_grad_over_N ? N : N - 1);
      if ((0.0 < dJ && dJ < ctol) || g < gtol || dJ_zero > o.dJ_counter_limit) break;
    }
    J = al_cost(p, o, w, w.X.data(), w.U.data(), &c_max);
    if (abort_trial) break;
    // dual + penalty update (A5, A6) with constraint values at the final trajectory
    const bool grow = !o.a6_penalty_conditional || (c_max > o.constraint_decrease_ratio * c_max_prev);
    c_max_prev = c_max;
    for (int64_t k = 0; k < N - 1; ++k) {
      double c[nb];
      bound_c(o, &w.U[k * m], c);
      for (int i = 0; i < nb; ++i) {
        const double l0 = w.lam_b[k * nb + i];
        const bool act = (o.a2_active_ge ? (c[i] >= 0.0) : (c[i] > 0.0)) || (l0 > 0.0);
        double l = l0 + w.mu_b[k * nb + i] * c[i];
        l = std::min(std::max(l, -o.dual_max), o.dual_max);
        if (o.a5_dual_active_only && !act) l = l0;
        w.lam_b[k * nb + i] = std::max(0.0, l);
        if (grow) w.mu_b[k * nb + i] = std::min(w.mu_b[k * nb + i] * o.penalty_scaling, o.penalty_max);
      }
    }
    for (int i = 0; i < n; ++i) {
      if (!(o.goal_mask & (1 << i))) continue;
      const double e = w.X[(N - 1) * n + i] - p.xf[i];
      double l = w.lam_g[i] + w.mu_g[i] * e;
      w.lam_g[i] = std::min(std::max(l, -o.dual_max), o.dual_max);
      if (grow) w.mu_g[i] = std::min(w.mu_g[i] * o.penalty_scaling, o.penalty_max);
    }
    if (c_max < o.constraint_tol) {
      status = ST_CONVERGED;
      break;
    }
  }
  for (int64_t i = 0; i < N * n; ++i) Xout[i] = w.X[i];
  for (int64_t i = 0; i < (N - 1) * m; ++i) Uout[i] = w.U[i];
  if (Kout)
    for (int64_t i = 0; i < (N - 1) * m * n; ++i) Kout[i] = w.K[i];
  out->status = status;
  out->outer_iters = outer;
  out->inner_iters = inner_total;
  out->ls_rollouts = ls_total;
  out->N = N;
  out->J = J;
  out->c_max = c_max;
  out->t_final = 0;
  out->slew_time = 0;
  out->flops = 0;
}

}  // namespace orc
