// K4 -- batched TVLQR replay, one thread per trial (tvlqr_solver.cuh), plus the
// eigen-axis-slew / Bryson-weight preparation kernel.
//   attitude_simulation(...)      reference src/attitude_controller.jl:1-48 (+ :50-145)
//   eigen_axis_slew(x0,xf,t)      src/eigen_axis_slew.jl:1-38
//   Bryson weights                src/TortoiseSat.jl:157-168, src/monte_carlo.jl:165-176
// A single backward sweep and a single forward sweep per trial: <1% of a trial's
// FLOPs (~12 MFLOP vs ~2 GFLOP for the solve), latency-bound sequential
// scans -- kept deliberately simple.
#pragma once
#include "common.cuh"
#include "tvlqr_solver.cuh"

namespace ts {

struct K4Args {
  int64_t n_trials;
  const int64_t* N_i;
  const int64_t* offs;
  const double* X_lqr;   // ragged N x 8 at offs*8
  const double* U_lqr;   // ragged (N-1) x 3 at offs*3
  const double* x0_lqr;  // 8 per trial
  const double* Jmat;    // 9
  const double* B_eci;
  const int64_t* B_offs;
  const int64_t* B_rows;
  const double* index_scale;
  const double* clock_rate;
  const double* t_final;      // per trial (tf of t_sim and the "fail" sentinel)
  const double* q_final;      // 4 per trial
  const uint32_t* stream_id;  // Philox stream per trial (global trial id), nullable -> t
  ts_tvlqr_opts_dev opts;
  const double* noise;   // explicit: ragged (N x 36) at offs*36, nullable
  double* X_sim;         // nullable, ragged N x 8
  double* U_sim;         // nullable, ragged N x 3
  double* dX;            // nullable, ragged N x 6
  double* K;             // ragged N x 18 (scratch if the caller passes none)
  int64_t* N_sim;        // nullable
  double* slew_time;     // nullable
};

__global__ void __launch_bounds__(64) k4_tvlqr_kernel(const K4Args a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_trials) return;
  TvlqrIn in;
  in.N = (int)a.N_i[t];
  in.X_lqr = a.X_lqr + a.offs[t] * 8;
  in.U_lqr = a.U_lqr + a.offs[t] * 3;
  for (int i = 0; i < 8; ++i) in.x0[i] = a.x0_lqr[t * 8 + i];
  for (int i = 0; i < 9; ++i) in.I.J[i] = a.Jmat[t * 9 + i];
  inv3_gj(in.I.J, in.I.Jinv);
  in.Bt = a.B_eci + a.B_offs[t] * 3;
  in.B_rows = a.B_rows[t];
  in.index_scale = a.index_scale[t];
  in.clock_rate = a.clock_rate[t];
  in.noise = a.noise ? a.noise + a.offs[t] * 36 : nullptr;
  in.trial = a.stream_id ? a.stream_id[t] : (uint32_t)t;
  for (int i = 0; i < 4; ++i) in.q_final[i] = a.q_final[t * 4 + i];
  in.t_final = a.t_final[t];
  in.time_step = a.opts.dt;
  in.trial_index_1based = (long long)in.trial + 1;
  ts_tvlqr_opts_dev o = a.opts;
  o.tf = a.t_final[t];
  double* K = a.K + a.offs[t] * 18;
  tvlqr_gains(in, o, K);
  double slew = 0.0;
  const long long ns = tvlqr_replay(in, o, K, a.X_sim ? a.X_sim + a.offs[t] * 8 : nullptr, a.U_sim ? a.U_sim + a.offs[t] * 3 : nullptr,
                                    a.dX ? a.dX + a.offs[t] * 6 : nullptr, &slew);
  if (a.N_sim) a.N_sim[t] = ns;
  if (a.slew_time) a.slew_time[t] = slew;
}

// eigen_axis_slew + Bryson weights, one thread per trial.  t_k = t0 + k*dt, k = 0..nt-1 with
// nt = length(t0:dt:t_final).  Optional guess outputs (ragged nt x 3 / nt x 4 at goffs).
struct PrepArgs {
  int64_t n_trials;
  const double* x0;       // 8 per trial (omega, q, clock)
  const double* xf;       // 8
  const double* Jmat;     // 9
  const double* t_final;  // per trial
  double t0, dt, alpha, beta;
  int conj_fix;           // 0: the reference's literal qmult(q_f, q_0) (eigen_axis_slew.jl:16); 1: conj(q_f) (x) q_0
  double* Qd;             // 8 per trial
  double* Qfd;
  double* Rd;             // 3
  const int64_t* goffs;   // nullable
  double* w_guess;        // nullable
  double* q_guess;        // nullable
};

__global__ void __launch_bounds__(128) k_slew_prep(const PrepArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_trials) return;
  const double PI = 3.141592653589793;
  const double* x0 = a.x0 + t * 8;
  const double* xf = a.xf + t * 8;
  const double* J = a.Jmat + t * 9;
  const double q1[4] = {x0[3], x0[4], x0[5], x0[6]};
  // eigen_axis_slew.jl:16 passes the 7-vector [q2; -q2[2:4]] to qmult, which reads entries 1 and 2:4 only: the literal
  // error quaternion is qmult(q2, q1), NOT conj(q2) (x) q1.  Reproduced by default; conj_fix = 1 is the evident intent.
  const double sg = a.conj_fix ? -1.0 : 1.0;
  const double q2c[4] = {xf[3], sg * xf[4], sg * xf[5], sg * xf[6]};
  double qe[4];
  qmult(q2c, q1, qe);
  const double theta_f = 2 * acos(qe[0]);
  const double sh = sin(theta_f / 2);
  const double axis[3] = {-qe[1] / sh, -qe[2] / sh, -qe[3] / sh};
  const long long nt = range_len(a.t0, a.dt, a.t_final[t]);
  const double t_end = a.t0 + a.dt * (double)(nt - 1);
  const double al = PI / t_end;
  const double tstep = (a.t0 + a.dt * 1.0) - (a.t0 + a.dt * 0.0);  // t[2]-t[1]
  double w_max = 0.0, tau_max = -INFINITY;
  double th_prev = theta_f * 1 / 2 * (1.0 - cos(al * (a.t0 + a.dt * 0.0)));
  double dth_prev = 0.0, w_prev[3] = {0, 0, 0};
  double* wg = (a.w_guess && a.goffs) ? a.w_guess + a.goffs[t] * 3 : nullptr;
  double* qg = (a.q_guess && a.goffs) ? a.q_guess + a.goffs[t] * 4 : nullptr;
  for (long long i = 0; i < nt; ++i) {
    double dth;
    double th_next = 0.0;
    if (i + 1 < nt) {
      th_next = theta_f * 1 / 2 * (1.0 - cos(al * (a.t0 + a.dt * (double)(i + 1))));
      dth = (th_next - th_prev) / tstep;
    } else {
      dth = dth_prev;  // push!(d_theta, d_theta[end])
    }
    double w[3];
    for (int c = 0; c < 3; ++c) {
      w[c] = dth * axis[c];
      w_max = fmax(w_max, fabs(w[c]));
    }
    if (i > 0) {
      const double dw[3] = {w[0] - w_prev[0], w[1] - w_prev[1], w[2] - w_prev[2]};
      for (int r = 0; r < 3; ++r) {
        double s = J[r * 3 + 0] * dw[0];
        s += J[r * 3 + 1] * dw[1];
        s += J[r * 3 + 2] * dw[2];
        tau_max = fmax(tau_max, s / a.dt);
      }
    }
    if (wg)
      for (int c = 0; c < 3; ++c) wg[i * 3 + c] = w[c];
    if (qg) {
      const double hs = sin(th_prev / 2);
      const double qa[4] = {cos(th_prev / 2), axis[0] * hs, axis[1] * hs, axis[2] * hs};
      qmult(q1, qa, qg + i * 4);
    }
    for (int c = 0; c < 3; ++c) w_prev[c] = w[c];
    dth_prev = dth;
    th_prev = th_next;
  }
  const double m_max = tau_max / 1.e-5 * 1.e2;
  double* Qd = a.Qd + t * 8;
  double* Qfd = a.Qfd + t * 8;
  for (int i = 0; i < 3; ++i) {
    Qd[i] = a.alpha / (w_max * w_max);
    Qfd[i] = a.alpha / (w_max * w_max) * 10;
  }
  for (int i = 3; i < 7; ++i) {
    Qd[i] = a.alpha * a.beta;
    Qfd[i] = a.alpha * a.beta * 10;
  }
  Qd[7] = 0.0;
  Qfd[7] = 0.0;
  for (int i = 0; i < 3; ++i) a.Rd[t * 3 + i] = 1 / (m_max * m_max);
}

}  // namespace ts
