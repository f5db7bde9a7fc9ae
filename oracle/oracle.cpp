// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI exports of the CPU restatement (orc_*.hpp) of the reference's hot
// path.  Loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs -- never by the product library.
//
// PARITY STATUS: the reference has no tests, fixtures or golden vectors
//  and cannot run here (no Julia).  IGRF is pinned by the
// reference's own two independent implementations + tables agreeing (igrf12 vs
// igrf12syn) and by an independent numpy restatement (tests/golden/).  Orbit,
// dynamics, AL-iLQR and TVLQR are "parity unpinned" against real Julia output.
#include <omp.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "orc_igrf.hpp"
#include "orc_orbit.hpp"
#include "orc_dynamics.hpp"
#include "orc_ilqr.hpp"
#include "orc_tvlqr.hpp"
#include "orc_philox.hpp"

using namespace orc;

extern "C" {

// ---------------------------------------------------------------- IGRF
void orc_legendre_schmidt(double theta, int nmax, double* P) { legendre_schmidt(theta, nmax, P); }
void orc_dlegendre_schmidt(double theta, int nmax, const double* P, double* dP) { dlegendre_schmidt(theta, nmax, P, dP); }
int orc_igrf12(double date, double r_m, double lat, double lon, double* out3) { return igrf12(date, r_m, lat, lon, out3); }
int orc_igrf12syn(int isv, double date, int itype, double alt, double colat, double elong, double* out4) {
  return igrf12syn(isv, date, itype, alt, colat, elong, out4);
}
int orc_igrf12_batch(double date, int64_t n, const double* r_m, const double* lat, const double* lon, double* Bn, double* Be,
                     double* Bd, int nthreads) {
  int rc = 0;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double o[3];
    const int e = igrf12(date, r_m[i], lat[i], lon[i], o);
    if (e) {
#pragma omp atomic write
      rc = e;
      o[0] = o[1] = o[2] = NAN;
    }
    Bn[i] = o[0];
    Be[i] = o[1];
    Bd[i] = o[2];
  }
  return rc;
}

// ---------------------------------------------------------------- orbit / field
void orc_kep_eci(double* kep6, double t0, double GM, double* out6) { kep_ECI(kep6, t0, GM, out6); }
void orc_orbit_rhs(const double* x6, double* dx6) { orbit_rhs(x6, dx6); }
double orc_sind(double x) { return sind(x); }
double orc_cosd(double x) { return cosd(x); }

struct orc_field_opts {
  double GM, mjd, igrf_date, field_radius_m, t0, tf;
  int64_t N;
};
int orc_magnetic_simulation(const double* kep6, const orc_field_opts* o, double* B, double* pos, double* vel) {
  FieldOpts f{o->GM, o->mjd, o->igrf_date, o->field_radius_m, o->t0, o->tf, o->N};
  return magnetic_simulation(kep6, f, B, pos, vel);
}
void orc_magnetic_gramian(const double* B, int64_t rows, double dt, double* G) { magnetic_gramian(B, rows, dt, G); }
double orc_cond_sym3(const double* G9) { return cond_sym3(G9); }
int64_t orc_condition_based_time(const double* G, int64_t rows, double cutoff) { return condition_based_time(G, rows, cutoff); }

// ---------------------------------------------------------------- dynamics
struct orc_dyn {
  const double* B_eci;
  int64_t B_rows;
  double index_scale, clock_rate;
  double J[9];
};
static DynCtx make_ctx(const orc_dyn* d) {
  DynCtx c;
  c.B_eci = d->B_eci;
  c.B_rows = d->B_rows;
  c.index_scale = d->index_scale;
  c.clock_rate = d->clock_rate;
  for (int i = 0; i < 9; ++i) c.J[i] = d->J[i];
  inv3(c.J, c.Jinv);
  return c;
}
void orc_qmult(const double* a, const double* b, double* out) { qmult(a, b, out); }
void orc_qrot(const double* q, const double* r, double* out) { qrot(q, r, out); }
void orc_deriv_function(const orc_dyn* d, const double* x8, const double* u3, double* dx8) {
  DynCtx c = make_ctx(d);
  DerivFunction<double>(c, x8, u3, dx8);
}
void orc_gain_simulator(const orc_dyn* d, const double* x8, const double* u3, double* dx8) {
  DynCtx c = make_ctx(d);
  gain_simulator<double>(c, x8, u3, dx8);
}
void orc_simulator(const orc_dyn* d, const double* x8, const double* u3, const double* noise9, double* dx8) {
  DynCtx c = make_ctx(d);
  simulator(c, x8, u3, noise9, dx8);
}
void orc_attitude_dynamics(const double* x7, const double* u3, const double* BB, const double* J9, double* dx7) {
  double Jinv[9];
  inv3(J9, Jinv);
  attitude_dynamics(x7, u3, BB, J9, Jinv, dx7);
}
void orc_rk3_step(const orc_dyn* d, const double* x8, const double* u3, double dt, double* xn8) {
  DynCtx c = make_ctx(d);
  rk3_step<double>([&](const double* xx, const double* uu, double* dx) { DerivFunction<double>(c, xx, uu, dx); }, x8, u3, dt, xn8);
}
// A (8x8 row-major), B (8x3 row-major) of the rk3 step by forward-mode duals.
void orc_rk3_jacobian(const orc_dyn* d, const double* x8, const double* u3, double dt, double* A, double* B) {
  DynCtx c = make_ctx(d);
  using D = Dual<11>;
  D x[8], u[3], xn[8];
  for (int i = 0; i < 8; ++i) {
    x[i] = D(x8[i]);
    x[i].d[i] = 1;
  }
  for (int i = 0; i < 3; ++i) {
    u[i] = D(u3[i]);
    u[i].d[8 + i] = 1;
  }
  rk3_step<D>([&](const D* xx, const D* uu, D* dx) { DerivFunction<D>(c, xx, uu, dx); }, x, u, dt, xn);
  for (int i = 0; i < 8; ++i) {
    for (int j = 0; j < 8; ++j) A[i * 8 + j] = xn[i].d[j];
    for (int j = 0; j < 3; ++j) B[i * 3 + j] = xn[i].d[8 + j];
  }
}

// ---------------------------------------------------------------- slew prep
void orc_eigen_axis_slew(const double* x0, const double* xf, const double* t, int64_t nt, double* w, double* q) {
  eigen_axis_slew(x0, xf, t, nt, w, q, 0);
}
// conj_fix = 1: conj(q_f) (x) q_0 instead of the reference's literal qmult(q_f, q_0) (eigen_axis_slew.jl:16)
void orc_eigen_axis_slew_mode(const double* x0, const double* xf, const double* t, int64_t nt, double* w, double* q, int conj_fix) {
  eigen_axis_slew(x0, xf, t, nt, w, q, conj_fix);
}
void orc_bryson_weights(const double* w, int64_t nt, const double* J9, double dt, double alpha, double beta, double* Qd,
                        double* Qfd, double* Rd) {
  bryson_weights(w, nt, J9, dt, alpha, beta, Qd, Qfd, Rd);
}

// ---------------------------------------------------------------- AL-iLQR
struct orc_ilqr_opts {
  int32_t max_outer, max_inner, max_linesearch, dJ_counter_limit, stage_cost_dt, goal_mask;
  double cost_tol, cost_tol_intermediate, grad_tol, grad_tol_intermediate, constraint_tol;
  double penalty_initial, penalty_scaling, penalty_max, dual_max;
  double ls_lower, ls_upper, bp_reg_increase, bp_reg_max, bp_reg_min, bp_reg_fp;
  double max_cost_value, max_state_value, max_control_value, u_max, u_min;
  int32_t a2_active_ge, a3_grad_over_N, a4_no_intermediate, a5_dual_active_only, a6_penalty_conditional, a7_carry_cost;
  double constraint_decrease_ratio;
  // launch-scheme fields of the product's ts_ilqr_opts (same layout; meaningless on the CPU)
  int32_t k3_suspend_after, k3_tail_share;
  double k3_early_factor;
};
static IlqrOpts make_opts(const orc_ilqr_opts* s) {
  IlqrOpts o;
  if (!s) return o;
  o.max_outer = s->max_outer; o.max_inner = s->max_inner; o.max_linesearch = s->max_linesearch;
  o.dJ_counter_limit = s->dJ_counter_limit; o.stage_cost_dt = s->stage_cost_dt; o.goal_mask = s->goal_mask;
  o.cost_tol = s->cost_tol; o.cost_tol_intermediate = s->cost_tol_intermediate;
  o.grad_tol = s->grad_tol; o.grad_tol_intermediate = s->grad_tol_intermediate;
  o.constraint_tol = s->constraint_tol; o.penalty_initial = s->penalty_initial;
  o.penalty_scaling = s->penalty_scaling; o.penalty_max = s->penalty_max; o.dual_max = s->dual_max;
  o.ls_lower = s->ls_lower; o.ls_upper = s->ls_upper; o.bp_reg_increase = s->bp_reg_increase;
  o.bp_reg_max = s->bp_reg_max; o.bp_reg_min = s->bp_reg_min; o.bp_reg_fp = s->bp_reg_fp;
  o.max_cost_value = s->max_cost_value; o.max_state_value = s->max_state_value;
  o.max_control_value = s->max_control_value; o.u_max = s->u_max; o.u_min = s->u_min;
  o.a2_active_ge = s->a2_active_ge; o.a3_grad_over_N = s->a3_grad_over_N; o.a4_no_intermediate = s->a4_no_intermediate;
  o.a5_dual_active_only = s->a5_dual_active_only; o.a6_penalty_conditional = s->a6_penalty_conditional;
  o.a7_carry_cost = s->a7_carry_cost; o.constraint_decrease_ratio = s->constraint_decrease_ratio;
  return o;
}
void orc_ilqr_default_opts(orc_ilqr_opts* s) {
  IlqrOpts o;
  s->max_outer = o.max_outer; s->max_inner = o.max_inner; s->max_linesearch = o.max_linesearch;
  s->dJ_counter_limit = o.dJ_counter_limit; s->stage_cost_dt = o.stage_cost_dt; s->goal_mask = o.goal_mask;
  s->cost_tol = o.cost_tol; s->cost_tol_intermediate = o.cost_tol_intermediate;
  s->grad_tol = o.grad_tol; s->grad_tol_intermediate = o.grad_tol_intermediate;
  s->constraint_tol = o.constraint_tol; s->penalty_initial = o.penalty_initial;
  s->penalty_scaling = o.penalty_scaling; s->penalty_max = o.penalty_max; s->dual_max = o.dual_max;
  s->ls_lower = o.ls_lower; s->ls_upper = o.ls_upper; s->bp_reg_increase = o.bp_reg_increase;
  s->bp_reg_max = o.bp_reg_max; s->bp_reg_min = o.bp_reg_min; s->bp_reg_fp = o.bp_reg_fp;
  s->max_cost_value = o.max_cost_value; s->max_state_value = o.max_state_value;
  s->max_control_value = o.max_control_value; s->u_max = o.u_max; s->u_min = o.u_min;
  s->a2_active_ge = o.a2_active_ge; s->a3_grad_over_N = o.a3_grad_over_N; s->a4_no_intermediate = o.a4_no_intermediate;
  s->a5_dual_active_only = o.a5_dual_active_only; s->a6_penalty_conditional = o.a6_penalty_conditional;
  s->a7_carry_cost = o.a7_carry_cost; s->constraint_decrease_ratio = o.constraint_decrease_ratio;
  s->k3_suspend_after = 150; s->k3_tail_share = 1; s->k3_early_factor = 2.0;
}

// Batched solve.  Per trial t: N_i[t] knots, ragged arrays addressed through
// offs[t] (in knots): X + offs*8, U + offs*3, K + offs*24.  Field tables through
// B_offs[t] (in rows) with B_rows[t] rows.  x0/xf are 8 per trial, Qd/Qfd 8, Rd 3,
// Jmat 9 per trial; index_scale/clock_rate per trial.
void orc_alilqr_solve_batch(int64_t n_trials, const int64_t* N_i, const int64_t* offs, const double* x0, const double* xf,
                            const double* Jmat, const double* Qd, const double* Qfd, const double* Rd, const double* B_eci,
                            const int64_t* B_offs, const int64_t* B_rows, const double* index_scale, const double* clock_rate,
                            double dt, const double* U0, const orc_ilqr_opts* opts, double* X, double* U, double* K,
                            IlqrOutcome* out, int nthreads) {
  const IlqrOpts o = make_opts(opts);
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int64_t t = 0; t < n_trials; ++t) {
    IlqrProblem p;
    p.N = N_i[t];
    p.dt = dt;
    for (int i = 0; i < 8; ++i) {
      p.x0[i] = x0[t * 8 + i];
      p.xf[i] = xf[t * 8 + i];
      p.Qd[i] = Qd[t * 8 + i];
      p.Qfd[i] = Qfd[t * 8 + i];
    }
    for (int i = 0; i < 3; ++i) p.Rd[i] = Rd[t * 3 + i];
    p.dyn.B_eci = B_eci + B_offs[t] * 3;
    p.dyn.B_rows = B_rows[t];
    p.dyn.index_scale = index_scale[t];
    p.dyn.clock_rate = clock_rate[t];
    for (int i = 0; i < 9; ++i) p.dyn.J[i] = Jmat[t * 9 + i];
    inv3(p.dyn.J, p.dyn.Jinv);
    alilqr_solve(p, o, U0 ? U0 + offs[t] * 3 : nullptr, X + offs[t] * 8, U + offs[t] * 3, K ? K + offs[t] * 24 : nullptr,
                 &out[t]);
  }
}

// ---------------------------------------------------------------- TVLQR
struct orc_tvlqr_opts {
  double dt, t0, tf;
  double Qd[6], Qfd[6], Rd[3];
  int32_t dt_squared;
  int32_t noise_mode;  // 0 none, 1 explicit array, 2 philox(seed, trial)
  uint64_t seed;
};
int64_t orc_attitude_simulation(const orc_dyn* d, const orc_tvlqr_opts* o, int64_t N, const double* X_lqr, const double* U_lqr,
                                const double* x0, const double* noise, uint32_t trial, double* X_sim, double* U_sim, double* dX,
                                double* K) {
  DynCtx c = make_ctx(d);
  TvlqrOpts t;
  t.dt = o->dt; t.t0 = o->t0; t.tf = o->tf; t.dt_squared = o->dt_squared;
  for (int i = 0; i < 6; ++i) { t.Qd[i] = o->Qd[i]; t.Qfd[i] = o->Qfd[i]; }
  for (int i = 0; i < 3; ++i) t.Rd[i] = o->Rd[i];
  std::vector<double> gen;
  const double* nz = nullptr;
  if (o->noise_mode == 1) nz = noise;
  if (o->noise_mode == 2) {
    gen.resize((size_t)N * 36);
    for (int64_t k = 0; k < N; ++k)
      for (int s = 0; s < 4; ++s) tvlqr_noise(o->seed, trial, (uint32_t)k, (uint32_t)s, &gen[(size_t)k * 36 + s * 9]);
    nz = gen.data();
  }
  return attitude_simulation(c, t, N, X_lqr, U_lqr, x0, nz, X_sim, U_sim, dX, K);
}
double orc_mc_slew_time(const double* X_sim, int64_t N_sim, const double* q_final, double t_final, double time_step,
                        double w_limit, double ang_limit, int literal, int64_t trial_index_1based) {
  return mc_slew_time(X_sim, N_sim, q_final, t_final, time_step, w_limit, ang_limit, literal, trial_index_1based);
}

// ---------------------------------------------------------------- Philox
void orc_philox4x32_10(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { philox4x32_10(ctr4, key2, out4); }
void orc_tvlqr_noise(uint64_t seed, uint32_t trial, uint32_t step, uint32_t stage, double* out9) {
  tvlqr_noise(seed, trial, step, stage, out9);
}

int orc_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
