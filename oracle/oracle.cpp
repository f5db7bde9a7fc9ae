// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI exports of the CPU restatement (orc_*.hpp) of the reference's hot
// path.  Loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs -- never by the product library.
//
// PARITY STATUS: the reference has no tests, fixtures or golden vectors
//  and cannot run here (no Julia).  Everything whose source
// IS in /root/reference (IGRF, orbit, field table, gramian cutoff, eigen-axis
// slew, Bryson weights, dynamics, rk3/rk4, TVLQR replay, MC post-processing) is
// pinned to tests/golden/ref_fixtures.json, the output of a line-by-line numpy
// transliteration of those sources (tests/golden/gen_ref_fixtures.py), and IGRF
// additionally to the reference's second implementation + table (igrf12syn) and
// to mpmath goldens.  AL-iLQR: TrajectoryOptimization.jl v0.1.2 is not vendored
// => PARITY UNPINNED against the real package; the algorithm is the frozen spec
// of SURVEY.md App. C, cross-checked by a second independent implementation
// (same fixture file) and with every unpinned choice a flip-tested switch.
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
#include "orc_comparison.hpp"

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
  int32_t k3_pair, k3_wide_occ;
  int32_t quat_error, k3_generic_inertia;
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
  o.quat_error = s->quat_error;
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
  s->k3_suspend_after = 150; s->k3_tail_share = 1; s->k3_early_factor = 2.0; s->k3_pair = 2; s->k3_wide_occ = 0;
  s->quat_error = o.quat_error; s->k3_generic_inertia = 0;
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

// ---------------------------------------------------------------- whole Monte-Carlo pipeline (CPU baseline)
// The loop body of reference src/monte_carlo.jl:118-262 with the solver block of src/TortoiseSat.jl:178-199, one trial
// per OpenMP task (schedule(dynamic,1): trials differ by 10x in work): scoping field pass -> gramian cutoff -> fine
// field table -> eigen-axis guess + Bryson weights -> AL-iLQR -> TVLQR replay with Philox noise -> slew-time rule.
// Mirrors ts_monte_carlo_run of the product argument for argument (a shared orbit is scoped once, as the fused GPU
// path does).  cpu_seconds (nullable, n_trials): per-trial CPU time, so that a bounded sample can be scaled by SUMS.
struct orc_mc_config {
  int64_t n_trials;
  int32_t shared_orbit, run_tvlqr;
  double t0, tf;
  int64_t N_scope;
  double cutoff, dt, alpha, beta;
  int32_t eigen_axis_fix, keep_trajectories;
  orc_ilqr_opts ilqr;
  orc_tvlqr_opts tvlqr;   // NB: product layout has more fields behind `seed`; only this prefix is read
};
int orc_mc_run(const orc_mc_config* cfg, const double* kep6, const orc_field_opts* fopts, const double* x0, const double* xf,
               const double* Jmat, const double* q_noise0, const uint32_t* stream_id, IlqrOutcome* out, double* cpu_seconds,
               int nthreads) {
  const int64_t n = cfg->n_trials;
  const IlqrOpts o = make_opts(&cfg->ilqr);
  if (nthreads < 1) nthreads = 1;
  struct Orbit {
    int64_t idx = 0, N = 0;
    double t_final = 0;
    std::vector<double> B;
  };
  auto scope = [&](int64_t f, Orbit& ob) {
    FieldOpts fo{fopts[f].GM, fopts[f].mjd, fopts[f].igrf_date, fopts[f].field_radius_m, cfg->t0, cfg->tf, cfg->N_scope};
    std::vector<double> B0((size_t)2 * cfg->N_scope * 3), G((size_t)2 * cfg->N_scope * 9);
    magnetic_simulation(kep6 + f * 6, fo, B0.data(), nullptr, nullptr);
    magnetic_gramian(B0.data(), 2 * cfg->N_scope, (cfg->tf - cfg->t0) / (double)cfg->N_scope, G.data());
    ob.idx = condition_based_time(G.data(), 2 * cfg->N_scope, cfg->cutoff);
    if (ob.idx <= 0) return;
    ob.t_final = (double)ob.idx * (cfg->tf - cfg->t0) / (double)cfg->N_scope;
    ob.N = (int64_t)std::floor((ob.t_final - cfg->t0) / cfg->dt);
    if (ob.N < 2) { ob.N = 0; return; }
    FieldOpts f2 = fo;
    f2.tf = ob.t_final;
    f2.N = ob.N;
    ob.B.assign((size_t)2 * ob.N * 3, 0.0);
    magnetic_simulation(kep6 + f * 6, f2, ob.B.data(), nullptr, nullptr);
  };
  Orbit shared;
  if (cfg->shared_orbit) scope(0, shared);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int64_t t = 0; t < n; ++t) {
    const double t_start = omp_get_wtime();
    Orbit own;
    if (!cfg->shared_orbit) scope(t, own);
    const Orbit& ob = cfg->shared_orbit ? shared : own;
    IlqrOutcome& r = out[t];
    std::memset(&r, 0, sizeof(r));
    if (ob.N < 2) {
      r.status = 5;  // NO_CUTOFF (magnetic_toolbox.jl:23: tf_index stays 0)
      if (cpu_seconds) cpu_seconds[t] = omp_get_wtime() - t_start;
      continue;
    }
    const int64_t N = ob.N;
    const int64_t nt = range_len(cfg->t0, cfg->dt, ob.t_final);
    std::vector<double> tt((size_t)nt), wg((size_t)nt * 3), qg((size_t)nt * 4);
    for (int64_t i = 0; i < nt; ++i) tt[(size_t)i] = cfg->t0 + cfg->dt * (double)i;
    eigen_axis_slew(x0 + t * 8, xf + t * 8, tt.data(), nt, wg.data(), qg.data(), cfg->eigen_axis_fix);
    IlqrProblem p;
    p.N = N;
    p.dt = cfg->dt;
    bryson_weights(wg.data(), nt, Jmat + t * 9, cfg->dt, cfg->alpha, cfg->beta, p.Qd, p.Qfd, p.Rd);
    for (int i = 0; i < 8; ++i) {
      p.x0[i] = x0[t * 8 + i];
      p.xf[i] = xf[t * 8 + i];
    }
    p.dyn.B_eci = ob.B.data();
    p.dyn.B_rows = 2 * N;
    p.dyn.index_scale = (double)N;
    p.dyn.clock_rate = 1.0 / (cfg->tf - cfg->t0);
    for (int i = 0; i < 9; ++i) p.dyn.J[i] = Jmat[t * 9 + i];
    inv3(p.dyn.J, p.dyn.Jinv);
    std::vector<double> X((size_t)N * 8), U((size_t)N * 3);
    alilqr_solve(p, o, nullptr, X.data(), U.data(), nullptr, &r);
    r.t_final = ob.t_final;
    if (cfg->run_tvlqr) {
      double x0l[8];
      for (int i = 0; i < 3; ++i) x0l[i] = x0[t * 8 + i];
      const double* q0 = x0 + t * 8 + 3;
      if (q_noise0) {  // TortoiseSat.jl:231-234
        const double* qn = q_noise0 + t * 3;
        const double th = std::sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2]);
        const double qp[4] = {std::cos(th / 2), qn[0] / th * std::sin(th / 2), qn[1] / th * std::sin(th / 2), qn[2] / th * std::sin(th / 2)};
        qmult(q0, qp, x0l + 3);
      } else {
        for (int i = 0; i < 4; ++i) x0l[3 + i] = q0[i];
      }
      x0l[7] = 0.0;
      TvlqrOpts tv;
      tv.dt = cfg->dt; tv.t0 = cfg->t0; tv.tf = ob.t_final; tv.dt_squared = cfg->tvlqr.dt_squared;
      for (int i = 0; i < 6; ++i) { tv.Qd[i] = cfg->tvlqr.Qd[i]; tv.Qfd[i] = cfg->tvlqr.Qfd[i]; }
      for (int i = 0; i < 3; ++i) tv.Rd[i] = cfg->tvlqr.Rd[i];
      const uint32_t sid = stream_id ? stream_id[t] : (uint32_t)t;
      std::vector<double> gen;
      if (cfg->tvlqr.noise_mode != 0) {
        gen.resize((size_t)N * 36);
        for (int64_t k = 0; k < N; ++k)
          for (int s4 = 0; s4 < 4; ++s4) tvlqr_noise(cfg->tvlqr.seed, sid, (uint32_t)k, (uint32_t)s4, &gen[(size_t)k * 36 + s4 * 9]);
      }
      std::vector<double> Xs((size_t)N * 8), Us((size_t)N * 3), dX((size_t)N * 6), K((size_t)N * 18);
      const int64_t ns = attitude_simulation(p.dyn, tv, N, X.data(), U.data(), x0l, gen.empty() ? nullptr : gen.data(), Xs.data(),
                                             Us.data(), dX.data(), K.data());
      r.slew_time = mc_slew_time(Xs.data(), ns, xf + t * 8 + 3, ob.t_final, cfg->dt, 0.05, 0.08727, 0, t + 1);
    }
    if (cpu_seconds) cpu_seconds[t] = omp_get_wtime() - t_start;
  }
  return 0;
}

// ---------------------------------------------------------------- comparison controller
void orc_psiaki_pd_simulation(int64_t N, const double* x0, const double* w_guess, const double* q_guess, const double* B, const double* J9,
                              double dt, double C1, double C2, double* X, double* M, double* Qe) {
  psiaki_pd_simulation(N, x0, w_guess, q_guess, B, J9, dt, C1, C2, X, M, Qe);
}
void orc_attitude_dynamics_linear(const double* x7, const double* u3, const double* xl7, const double* BB, const double* J9, double* dx7) {
  double Jinv[9];
  inv3(J9, Jinv);
  attitude_dynamics_linear(x7, u3, xl7, BB, J9, Jinv, dx7);
}

// ---------------------------------------------------------------- Philox
void orc_philox4x32_10(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { philox4x32_10(ctr4, key2, out4); }
void orc_tvlqr_noise(uint64_t seed, uint32_t trial, uint32_t step, uint32_t stage, double* out9) {
  tvlqr_noise(seed, trial, step, stage, out9);
}

int orc_max_threads(void) { return omp_get_max_threads(); }

}  // extern "C"
