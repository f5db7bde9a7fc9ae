// K7 -- the paper's comparison controller on the same rollout machinery: the Psiaki-style
// PD magnetic controller closed loop of src/comparison/psiaki2005.jl:116-164, batched one thread per trial.
//   psiaki_controller(C_1,C_2,J,q,w,B_meas,m_limit)   src/comparison/psiaki_dynamics.jl:1-26
//   rk4_psiaki(f,x,dt,u,B_B,J)                        src/comparison/psiaki_dynamics.jl:63-73
//   attitude_dynamics(x,u,B_B,J)                      src/attitude_dynamics.jl:2-24
//   attitude_dynamics_linear(x,u,x_linear,B_B,J)      src/attitude_dynamics.jl:26-48
// The periodic-LQR baseline (psiaki2001_Period_LQR.jl) needs ControlSystems.care, which is not in the reference's
// Manifest ("unrunnable as pinned") and is not built.
#pragma once
#include "common.cuh"
#include "ilqr_math.cuh"

namespace ts {

// attitude_dynamics.jl:2-24: 7-state, body-frame field and raw moment (tau = u x B_B, no 1e-2 scaling)
__device__ __forceinline__ void attitude_dynamics7(const Inertia& I, const double x[7], const double u[3], const double BB[3], double dx[7]) {
  const double nq = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  const double q[4] = {x[3] / nq, x[4] / nq, x[5] / nq, x[6] / nq};
  const double w4[4] = {0.0, x[0], x[1], x[2]};
  double qd[4], tau[3], Jw[3], wJw[3];
  qmult(q, w4, qd);
  cross3(u, BB, tau);
  for (int c = 0; c < 3; ++c) Jw[c] = I.J[c * 3 + 0] * x[0] + I.J[c * 3 + 1] * x[1] + I.J[c * 3 + 2] * x[2];
  cross3(x, Jw, wJw);
  const double r0 = tau[0] - wJw[0], r1 = tau[1] - wJw[1], r2 = tau[2] - wJw[2];
  for (int c = 0; c < 3; ++c) dx[c] = I.Jinv[c * 3 + 0] * r0 + I.Jinv[c * 3 + 1] * r1 + I.Jinv[c * 3 + 2] * r2;
  for (int c = 0; c < 4; ++c) dx[3 + c] = 0.5 * qd[c];
}
// attitude_dynamics.jl:26-48: the same with q_dot driven by x_linear[4:6] instead of omega
__device__ __forceinline__ void attitude_dynamics_linear7(const Inertia& I, const double x[7], const double u[3], const double xl[7],
                                                          const double BB[3], double dx[7]) {
  const double nq = sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  const double q[4] = {x[3] / nq, x[4] / nq, x[5] / nq, x[6] / nq};
  const double w4[4] = {0.0, xl[3], xl[4], xl[5]};
  double qd[4], tau[3], Jw[3], wJw[3];
  qmult(q, w4, qd);
  cross3(u, BB, tau);
  for (int c = 0; c < 3; ++c) Jw[c] = I.J[c * 3 + 0] * x[0] + I.J[c * 3 + 1] * x[1] + I.J[c * 3 + 2] * x[2];
  cross3(x, Jw, wJw);
  const double r0 = tau[0] - wJw[0], r1 = tau[1] - wJw[1], r2 = tau[2] - wJw[2];
  for (int c = 0; c < 3; ++c) dx[c] = I.Jinv[c * 3 + 0] * r0 + I.Jinv[c * 3 + 1] * r1 + I.Jinv[c * 3 + 2] * r2;
  for (int c = 0; c < 4; ++c) dx[3 + c] = 0.5 * qd[c];
}

// psiaki_dynamics.jl:1-26: m = (B x T_req) / |B|^2 with T_req = -(C_1 w + C_2 inv(J) q[2:4])
__device__ __forceinline__ void psiaki_controller_dev(double C1, double C2, const Inertia& I, const double q[4], const double w[3],
                                                      const double Bm[3], double m[3]) {
  double T[3];
  for (int i = 0; i < 3; ++i) {
    const double jq = I.Jinv[i * 3 + 0] * q[1] + I.Jinv[i * 3 + 1] * q[2] + I.Jinv[i * 3 + 2] * q[3];
    T[i] = -(C1 * w[i] + C2 * jq);
  }
  double c[3];
  cross3(Bm, T, c);
  const double nb = sqrt(Bm[0] * Bm[0] + Bm[1] * Bm[1] + Bm[2] * Bm[2]);
  const double n2 = nb * nb;
  for (int i = 0; i < 3; ++i) m[i] = c[i] / n2;
}

// psiaki_dynamics.jl:63-73 (u and B_B held over the step)
__device__ __forceinline__ void rk4_psiaki_dev(const Inertia& I, const double x[7], double dt, const double u[3], const double BB[3], double xn[7]) {
  double f1[7], f2[7], f3[7], f4[7], xs[7];
  attitude_dynamics7(I, x, u, BB, f1);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + .5 * f1[i] * dt;
  attitude_dynamics7(I, xs, u, BB, f2);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + .5 * f2[i] * dt;
  attitude_dynamics7(I, xs, u, BB, f3);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + f3[i] * dt;
  attitude_dynamics7(I, xs, u, BB, f4);
  for (int i = 0; i < 7; ++i) xn[i] = x[i] + 1.0 / 6 * (f1[i] + 2 * f2[i] + 2 * f3[i] + f4[i]) * dt;
}

struct K7Args {
  int64_t n_trials;
  const int64_t* N_i;
  const int64_t* offs;
  const double* x0;        // n x 7
  const double* w_guess;   // ragged N x 3 at offs
  const double* q_guess;   // ragged N x 4
  const double* B_eci;     // ragged N x 3: field at every step (psiaki2005.jl:73-74)
  const double* Jmat;      // n x 9
  double dt, C1, C2;
  double* X;               // ragged N x 7
  double* M;               // ragged N x 3 (nullable)
  double* Qe;              // ragged N x 4 (nullable)
};

// psiaki2005.jl:116-164, one thread per trial (1-based i of the script -> 0-based k = i-1)
__global__ void __launch_bounds__(64) k7_psiaki_pd_kernel(const K7Args a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n_trials) return;
  const int64_t N = a.N_i[t], o = a.offs[t];
  Inertia I;
  for (int i = 0; i < 9; ++i) I.J[i] = a.Jmat[t * 9 + i];
  inv3_gj(I.J, I.Jinv);
  const double* B = a.B_eci + o * 3;
  const double* wg = a.w_guess + o * 3;
  const double* qg = a.q_guess + o * 4;
  double* X = a.X + o * 7;
  double x[7];
  for (int i = 0; i < 7; ++i) X[i] = x[i] = a.x0[t * 7 + i];
  if (a.M)
    for (int i = 0; i < 3; ++i) a.M[o * 3 + i] = 0.0;
  if (a.Qe)
    for (int i = 0; i < 4; ++i) a.Qe[o * 4 + i] = 0.0;
  if (N < 2) return;
  {  // :124-125 one explicit Euler step with zero moment
    const double z[3] = {0.0, 0.0, 0.0};
    double dx[7];
    attitude_dynamics7(I, x, z, B, dx);
    for (int i = 0; i < 7; ++i) X[7 + i] = x[i] = x[i] + a.dt * dx[i];
    if (a.M)
      for (int i = 0; i < 3; ++i) a.M[(o + 1) * 3 + i] = 0.0;   // overwritten below when the loop visits step 2
    if (a.Qe)
      for (int i = 0; i < 4; ++i) a.Qe[(o + 1) * 4 + i] = 0.0;
  }
  for (int64_t k = 1; k < N - 1; ++k) {   // i = 2 : length(t)-1
    const double qi[4] = {x[3], -x[4], -x[5], -x[6]};
    double Bm[3], wbar[3], qbar[4], m[3], xn[7];
    qrot(qi, B + k * 3, Bm);
    for (int i = 0; i < 3; ++i) wbar[i] = wg[k * 3 + i] - x[i];
    qmult(x + 3, qg + k * 4, qbar);
    psiaki_controller_dev(a.C1, a.C2, I, qbar, wbar, Bm, m);
    rk4_psiaki_dev(I, x, a.dt, m, Bm, xn);
    const double nq = sqrt(xn[3] * xn[3] + xn[4] * xn[4] + xn[5] * xn[5] + xn[6] * xn[6]);   // normalize(x[4:7,i+1])
    for (int i = 3; i < 7; ++i) xn[i] = xn[i] / nq;
    for (int i = 0; i < 7; ++i) X[(k + 1) * 7 + i] = x[i] = xn[i];
    if (a.M)
      for (int i = 0; i < 3; ++i) a.M[(o + k) * 3 + i] = m[i];
    if (a.Qe)
      for (int i = 0; i < 4; ++i) a.Qe[(o + k) * 4 + i] = qbar[i];
  }
  if (a.M)
    for (int i = 0; i < 3; ++i) a.M[(o + N - 1) * 3 + i] = 0.0;
  if (a.Qe)
    for (int i = 0; i < 4; ++i) a.Qe[(o + N - 1) * 4 + i] = 0.0;
}

__global__ void k7_attitude_dynamics_linear(int64_t n, const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ xl,
                                            const double* __restrict__ BB, const double* __restrict__ Jmat, double* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Inertia I;
  for (int c = 0; c < 9; ++c) I.J[c] = Jmat[c];
  inv3_gj(I.J, I.Jinv);
  double d[7];
  attitude_dynamics_linear7(I, x + i * 7, u + i * 3, xl + i * 7, BB + i * 3, d);
  for (int c = 0; c < 7; ++c) dx[i * 7 + c] = d[c];
}

}  // namespace ts
