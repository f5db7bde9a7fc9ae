// Lane-local math of the AL-iLQR / TVLQR kernels: 7-state attitude dynamics, its
// analytic Jacobians, and the rk3 / rk4 ZOH steps with their Jacobians.
//
// Replaces, per knot,
//   DerivFunction(dx,x,u)              reference src/DerivFunction.jl:1-48
//   gain_simulator / simulator         src/gain_simulator.jl:1-53, src/simulator.jl:1-42
//   rk3 / rk4 discretisers             src/attitude_controller.jl:122-132,178-187
//   ForwardDiff.jacobian! through them (inside TrajectoryOptimization and
//                                       attitude_controller.jl:103)
// The 8th ("clock") state of the reference is dynamically decoupled
// (d f/d x8 = 0 through floor(), quirk Q1/Q2), so the kernels carry 7 states and
// take the field row of each Runge-Kutta stage as an input.  Derivatives are
// analytic (chain rule through the stages) instead of dual numbers; they agree
// with the oracle's forward-mode duals to round-off (tests/test_hostsim.py).
//
// Everything here is plain per-thread code: TS_HD makes it compile for the GPU
// (nvcc) and for the host-side lane emulator used by the CPU tests.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TS_HD __host__ __device__ __forceinline__
#else
#define TS_HD inline
#endif

namespace ts {

struct Inertia {
  double J[9];
  double Jinv[9];
};

// 1/sqrt(s) instead of sqrt followed by four divisions.  On the GPU: the hardware seed (MUFU.RSQ64H, ~2^-22) and ONE
// third-order step y0 + y0 e (1/2 + 3/8 e), e = 1 - s y0^2 -- the arithmetic of the library's rsqrt() fast path (5 FP64
// operations) without its range checks and slow-path call (12 more instructions, three times per rollout knot).  The
// argument is the squared norm of an attitude quaternion (or a Cholesky pivot > 0); zero, denormal, infinite or NaN
// arguments give inf / NaN here, which the rollout's |x| < max_state_value test rejects like any diverged candidate.
TS_HD double rnorm(double s) {
#ifdef __CUDA_ARCH__
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s));
  const double e = fma(-s, y0 * y0, 1.0);
  return fma(fma(e, 0.375, 0.5), e * y0, y0);
#else
  return 1.0 / sqrt(s);
#endif
}
#define TS_SIXTH (1.0 / 6.0)

TS_HD void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
// J*w and Jinv*r.  DJ = true: the inertia matrix is diagonal (principal axes: every preset of input_parameters.jl:4-66) and
// the products with its exact zeros are left out: the same values (x + 0*y == x), 3 instead of 9 operations per product.
template <bool DJ>
TS_HD void mul_J(const Inertia& I, const double w[3], double o[3]) {
  if (DJ) {
    o[0] = I.J[0] * w[0]; o[1] = I.J[4] * w[1]; o[2] = I.J[8] * w[2];
  } else {
    for (int i = 0; i < 3; ++i) o[i] = I.J[i * 3 + 0] * w[0] + I.J[i * 3 + 1] * w[1] + I.J[i * 3 + 2] * w[2];
  }
}
template <bool DJ>
TS_HD void mul_Jinv(const Inertia& I, const double r[3], double o[3]) {
  if (DJ) {
    o[0] = I.Jinv[0] * r[0]; o[1] = I.Jinv[4] * r[1]; o[2] = I.Jinv[8] * r[2];
  } else {
    for (int i = 0; i < 3; ++i) o[i] = I.Jinv[i * 3 + 0] * r[0] + I.Jinv[i * 3 + 1] * r[1] + I.Jinv[i * 3 + 2] * r[2];
  }
}
// Hamilton product, scalar first (qmult.jl:1-3)
TS_HD void qmult(const double a[4], const double b[4], double o[4]) {
  o[0] = a[0] * b[0] - (a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
  o[1] = a[0] * b[1] + b[0] * a[1] + (a[2] * b[3] - a[3] * b[2]);
  o[2] = a[0] * b[2] + b[0] * a[2] + (a[3] * b[1] - a[1] * b[3]);
  o[3] = a[0] * b[3] + b[0] * a[3] + (a[1] * b[2] - a[2] * b[1]);
}
// a (x) (0, v): the Hamilton product with a pure-vector right factor, i.e. qmult without its products by the zero scalar
// part (the same values for finite operands; 12 instead of 19 operations, three times per rollout knot)
TS_HD void qmult_pure(const double a[4], const double v[3], double o[4]) {
  o[0] = -(a[1] * v[0] + a[2] * v[1] + a[3] * v[2]);
  o[1] = a[0] * v[0] + (a[2] * v[2] - a[3] * v[1]);
  o[2] = a[0] * v[1] + (a[3] * v[0] - a[1] * v[2]);
  o[3] = a[0] * v[2] + (a[1] * v[1] - a[2] * v[0]);
}
// qrot.jl:1-3
TS_HD void qrot(const double q[4], const double r[3], double o[3]) {
  double c1[3], w[3], c2[3];
  cross3(q + 1, r, c1);
  w[0] = c1[0] + q[0] * r[0];
  w[1] = c1[1] + q[0] * r[1];
  w[2] = c1[2] + q[0] * r[2];
  cross3(q + 1, w, c2);
  o[0] = r[0] + 2.0 * c2[0];
  o[1] = r[1] + 2.0 * c2[1];
  o[2] = r[2] + 2.0 * c2[2];
}

// xdot = f(x,u) for x = [omega(3); q(4)], field row Bn (ECI, Tesla).
// u_scale_mode 0: u*1e-2 (DerivFunction.jl:37); 1: u/100 (simulator/gain_simulator) -- quirk Q8.
template <int UMODE, bool DJ = false>
TS_HD void dyn_f(const Inertia& I, const double x[7], const double u[3], const double Bn[3], double dx[7]) {
  const double inq = rnorm(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  const double q[4] = {x[3] * inq, x[4] * inq, x[5] * inq, x[6] * inq};
  double qd[4];
  qmult_pure(q, x, qd);   // qmult(q, [0; omega]), DerivFunction.jl:33
  double BB[3];
  qrot(q, Bn, BB);
  double us[3];
  if (UMODE == 0) {
    us[0] = u[0] * 1.e-2; us[1] = u[1] * 1.e-2; us[2] = u[2] * 1.e-2;
  } else {
    us[0] = u[0] * 0.01; us[1] = u[1] * 0.01; us[2] = u[2] * 0.01;
  }
  double tau[3], Jw[3], wJw[3];
  cross3(us, BB, tau);
  mul_J<DJ>(I, x, Jw);
  cross3(x, Jw, wJw);
  const double r[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  mul_Jinv<DJ>(I, r, dx);
  for (int i = 0; i < 4; ++i) dx[3 + i] = 0.5 * qd[i];
}

// f and its Jacobians fx (7x7, row-major) and fu (only rows 0..2 are non-zero: 3x3 row-major).
template <int UMODE>
TS_HD void dyn_f_jac(const Inertia& I, const double x[7], const double u[3], const double Bn[3], double dx[7], double fx[49],
                     double fu[9]) {
  const double inq = rnorm(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  const double q[4] = {x[3] * inq, x[4] * inq, x[5] * inq, x[6] * inq};
  const double s = q[0];
  const double v[3] = {q[1], q[2], q[3]};
  const double w[3] = {x[0], x[1], x[2]};
  const double us_k = 1.e-2;
  const double us[3] = {u[0] * us_k, u[1] * us_k, u[2] * us_k};
  // ---- value
  double vxB[3], t1[3], c2[3], BB[3];
  cross3(v, Bn, vxB);
  t1[0] = vxB[0] + s * Bn[0];
  t1[1] = vxB[1] + s * Bn[1];
  t1[2] = vxB[2] + s * Bn[2];
  cross3(v, t1, c2);
  BB[0] = Bn[0] + 2.0 * c2[0];
  BB[1] = Bn[1] + 2.0 * c2[1];
  BB[2] = Bn[2] + 2.0 * c2[2];
  double tau[3], Jw[3], wJw[3];
  cross3(us, BB, tau);
  for (int i = 0; i < 3; ++i) Jw[i] = I.J[i * 3 + 0] * w[0] + I.J[i * 3 + 1] * w[1] + I.J[i * 3 + 2] * w[2];
  cross3(w, Jw, wJw);
  const double r[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  for (int i = 0; i < 3; ++i) dx[i] = I.Jinv[i * 3 + 0] * r[0] + I.Jinv[i * 3 + 1] * r[1] + I.Jinv[i * 3 + 2] * r[2];
  dx[3] = 0.5 * (-(v[0] * w[0] + v[1] * w[1] + v[2] * w[2]));
  dx[4] = 0.5 * (s * w[0] + (v[1] * w[2] - v[2] * w[1]));
  dx[5] = 0.5 * (s * w[1] + (v[2] * w[0] - v[0] * w[2]));
  dx[6] = 0.5 * (s * w[2] + (v[0] * w[1] - v[1] * w[0]));

  // ---- d(omega_dot)/d(omega) = -Jinv * ( hat(w) J - hat(J w) )
  double M[9];
  {
    // hat(w) J : row i = w x (column-wise) -> (hat(w) J)[i][j] = sum_k hat(w)[i][k] J[k][j]
    const double hw[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    const double hJw[9] = {0, -Jw[2], Jw[1], Jw[2], 0, -Jw[0], -Jw[1], Jw[0], 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        M[i * 3 + j] = hw[i * 3 + 0] * I.J[0 * 3 + j] + hw[i * 3 + 1] * I.J[1 * 3 + j] + hw[i * 3 + 2] * I.J[2 * 3 + j] - hJw[i * 3 + j];
  }
  for (int i = 0; i < 49; ++i) fx[i] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      fx[i * 7 + j] = -(I.Jinv[i * 3 + 0] * M[0 * 3 + j] + I.Jinv[i * 3 + 1] * M[1 * 3 + j] + I.Jinv[i * 3 + 2] * M[2 * 3 + j]);
  // ---- d(omega_dot)/du = Jinv * (-hat(BB)) * us_k
  {
    const double nhB[9] = {0, BB[2], -BB[1], -BB[2], 0, BB[0], BB[1], -BB[0], 0};  // -hat(BB)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        fu[i * 3 + j] = (I.Jinv[i * 3 + 0] * nhB[0 * 3 + j] + I.Jinv[i * 3 + 1] * nhB[1 * 3 + j] + I.Jinv[i * 3 + 2] * nhB[2 * 3 + j]) * us_k;
  }
  // ---- dBB/d(qhat) (3x4): column 0 = 2 (v x Bn); columns 1..3 = -2 hat(v x Bn) - 2 hat(v) hat(Bn) - 2 s hat(Bn)
  double dBq[12];
  {
    dBq[0 * 4 + 0] = 2.0 * vxB[0];
    dBq[1 * 4 + 0] = 2.0 * vxB[1];
    dBq[2 * 4 + 0] = 2.0 * vxB[2];
    const double hc[9] = {0, -vxB[2], vxB[1], vxB[2], 0, -vxB[0], -vxB[1], vxB[0], 0};
    const double hv[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    const double hB[9] = {0, -Bn[2], Bn[1], Bn[2], 0, -Bn[0], -Bn[1], Bn[0], 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double hvhB = hv[i * 3 + 0] * hB[0 * 3 + j] + hv[i * 3 + 1] * hB[1 * 3 + j] + hv[i * 3 + 2] * hB[2 * 3 + j];
        dBq[i * 4 + 1 + j] = -2.0 * hc[i * 3 + j] - 2.0 * hvhB - 2.0 * s * hB[i * 3 + j];
      }
  }
  // d(qhat)/dq = (I - qhat qhat') / |q|   (4x4, symmetric)
  double P[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) P[i * 4 + j] = ((i == j ? 1.0 : 0.0) - q[i] * q[j]) * inq;
  // d(omega_dot)/dq = Jinv * hat(us) * dBq * P
  {
    const double hu[9] = {0, -us[2], us[1], us[2], 0, -us[0], -us[1], us[0], 0};
    double T1[12], T2[12];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        T1[i * 4 + j] = hu[i * 3 + 0] * dBq[0 * 4 + j] + hu[i * 3 + 1] * dBq[1 * 4 + j] + hu[i * 3 + 2] * dBq[2 * 4 + j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        T2[i * 4 + j] = I.Jinv[i * 3 + 0] * T1[0 * 4 + j] + I.Jinv[i * 3 + 1] * T1[1 * 4 + j] + I.Jinv[i * 3 + 2] * T1[2 * 4 + j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        fx[i * 7 + 3 + j] = T2[i * 4 + 0] * P[0 * 4 + j] + T2[i * 4 + 1] * P[1 * 4 + j] + T2[i * 4 + 2] * P[2 * 4 + j] + T2[i * 4 + 3] * P[3 * 4 + j];
  }
  // ---- d(qdot)/d(omega): row0 = -v'/2 ; rows 1..3 = (s I + hat(v))/2
  fx[3 * 7 + 0] = -0.5 * v[0];
  fx[3 * 7 + 1] = -0.5 * v[1];
  fx[3 * 7 + 2] = -0.5 * v[2];
  fx[4 * 7 + 0] = 0.5 * s;      fx[4 * 7 + 1] = -0.5 * v[2];  fx[4 * 7 + 2] = 0.5 * v[1];
  fx[5 * 7 + 0] = 0.5 * v[2];   fx[5 * 7 + 1] = 0.5 * s;      fx[5 * 7 + 2] = -0.5 * v[0];
  fx[6 * 7 + 0] = -0.5 * v[1];  fx[6 * 7 + 1] = 0.5 * v[0];   fx[6 * 7 + 2] = 0.5 * s;
  // ---- d(qdot)/d(qhat) (4x4): row0 = [0, -w'/2]; rows 1..3 = [w/2, -hat(w)/2]; then * P
  {
    const double D[16] = {0,          -0.5 * w[0], -0.5 * w[1], -0.5 * w[2],
                          0.5 * w[0], 0,            0.5 * w[2], -0.5 * w[1],
                          0.5 * w[1], -0.5 * w[2],  0,           0.5 * w[0],
                          0.5 * w[2], 0.5 * w[1],  -0.5 * w[0],  0};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        fx[(3 + i) * 7 + 3 + j] = D[i * 4 + 0] * P[0 * 4 + j] + D[i * 4 + 1] * P[1 * 4 + j] + D[i * 4 + 2] * P[2 * 4 + j] + D[i * 4 + 3] * P[3 * 4 + j];
  }
}

// rk3 ZOH step (attitude_controller.jl:178-187 == TrajOpt rk3); Bs = field rows of the 3 stages.
template <int UMODE, bool DJ = false>
TS_HD void rk3_step7(const Inertia& I, const double x[7], const double u[3], const double* B1, const double* B2, const double* B3,
                     double dt, double xn[7]) {
  double k1[7], k2[7], k3[7], xs[7];
  dyn_f<UMODE, DJ>(I, x, u, B1, k1);
  for (int i = 0; i < 7; ++i) k1[i] = k1[i] * dt;
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * k1[i];
  dyn_f<UMODE, DJ>(I, xs, u, B2, k2);
  for (int i = 0; i < 7; ++i) k2[i] = k2[i] * dt;
  for (int i = 0; i < 7; ++i) xs[i] = x[i] - k1[i] + 2.0 * k2[i];
  dyn_f<UMODE, DJ>(I, xs, u, B3, k3);
  for (int i = 0; i < 7; ++i) k3[i] = k3[i] * dt;
  for (int i = 0; i < 7; ++i) xn[i] = x[i] + (k1[i] + 4.0 * k2[i] + k3[i]) * TS_SIXTH;
}

// y (7x10) = fx (7x7) * M (7x10) [+ fu in columns 7..9 of rows 0..2], all scaled by dt.
TS_HD void stage_chain(const double fx[49], const double fu[9], const double M[70], double dt, double out[70]) {
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) {
      double a = 0.0;
      for (int l = 0; l < 7; ++l) a += fx[i * 7 + l] * M[l * 10 + j];
      if (i < 3 && j >= 7) a += fu[i * 3 + (j - 7)];
      out[i * 10 + j] = a * dt;
    }
}

// Jacobian of the rk3 step: AB (7x10 row-major) = [A | B], A = d xn/d x, B = d xn/d u.  Also returns xn.
template <int UMODE>
TS_HD void rk3_jac7(const Inertia& I, const double x[7], const double u[3], const double* B1, const double* B2, const double* B3,
                    double dt, double xn[7], double AB[70]) {
  double k1[7], k2[7], k3[7], xs[7], fx[49], fu[9];
  double K1[70], K2[70], M[70];
  // stage 1: d k1 = dt [fx | fu]
  dyn_f_jac<UMODE>(I, x, u, B1, k1, fx, fu);
  for (int i = 0; i < 7; ++i) k1[i] = k1[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) {
      double a = (j < 7) ? fx[i * 7 + j] : ((i < 3) ? fu[i * 3 + (j - 7)] : 0.0);
      K1[i * 10 + j] = a * dt;
    }
  // stage 2 at x2 = x + k1/2 : d x2 = [I|0] + K1/2
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * k1[i];
  dyn_f_jac<UMODE>(I, xs, u, B2, k2, fx, fu);
  for (int i = 0; i < 7; ++i) k2[i] = k2[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) M[i * 10 + j] = ((i == j) ? 1.0 : 0.0) + 0.5 * K1[i * 10 + j];
  stage_chain(fx, fu, M, dt, K2);
  // stage 3 at x3 = x - k1 + 2 k2 : d x3 = [I|0] - K1 + 2 K2
  for (int i = 0; i < 7; ++i) xs[i] = x[i] - k1[i] + 2.0 * k2[i];
  dyn_f_jac<UMODE>(I, xs, u, B3, k3, fx, fu);
  for (int i = 0; i < 7; ++i) k3[i] = k3[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) M[i * 10 + j] = ((i == j) ? 1.0 : 0.0) - K1[i * 10 + j] + 2.0 * K2[i * 10 + j];
  double K3[70];
  stage_chain(fx, fu, M, dt, K3);
  for (int i = 0; i < 7; ++i) xn[i] = x[i] + (k1[i] + 4.0 * k2[i] + k3[i]) * TS_SIXTH;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j)
      AB[i * 10 + j] = ((i == j) ? 1.0 : 0.0) + (K1[i * 10 + j] + 4.0 * K2[i * 10 + j] + K3[i * 10 + j]) * TS_SIXTH;
}

// ---- register-resident linearisation: Jacobian-vector products ------------------------------
// The matrix form above keeps K1, K2, K3 and a stage Jacobian alive (~260 doubles) and spills;
// the kernel instead pushes the 10 unit directions through the three stages one at a time with
// analytic JVPs.  Live state: three StagePt (3 x 23 doubles) + one direction (~30 doubles).
struct StagePt {           // 11 doubles per stage: small enough for three of them to stay in registers
  double inq, s, v[3], w[3], Bn[3];
};
// f(x,u) at a stage point + the intermediates its JVP needs
template <bool DJ = false>
TS_HD void stage_eval(const Inertia& I, const double x[7], const double us[3], const double* Bn, StagePt& sp, double dx[7]) {
  sp.inq = rnorm(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  sp.s = x[3] * sp.inq;
  for (int i = 0; i < 3; ++i) {
    sp.v[i] = x[4 + i] * sp.inq;
    sp.w[i] = x[i];
    sp.Bn[i] = Bn[i];
  }
  double vxB[3], t1[3], c2[3], BB[3], tau[3], Jw[3], wJw[3];
  cross3(sp.v, sp.Bn, vxB);
  for (int i = 0; i < 3; ++i) t1[i] = vxB[i] + sp.s * sp.Bn[i];
  cross3(sp.v, t1, c2);
  for (int i = 0; i < 3; ++i) BB[i] = sp.Bn[i] + 2.0 * c2[i];
  cross3(us, BB, tau);
  mul_J<DJ>(I, sp.w, Jw);
  cross3(sp.w, Jw, wJw);
  const double r[3] = {tau[0] - wJw[0], tau[1] - wJw[1], tau[2] - wJw[2]};
  mul_Jinv<DJ>(I, r, dx);
  dx[3] = 0.5 * (-(sp.v[0] * sp.w[0] + sp.v[1] * sp.w[1] + sp.v[2] * sp.w[2]));
  dx[4] = 0.5 * (sp.s * sp.w[0] + (sp.v[1] * sp.w[2] - sp.v[2] * sp.w[1]));
  dx[5] = 0.5 * (sp.s * sp.w[1] + (sp.v[2] * sp.w[0] - sp.v[0] * sp.w[2]));
  dx[6] = 0.5 * (sp.s * sp.w[2] + (sp.v[0] * sp.w[1] - sp.v[1] * sp.w[0]));
}
// out = fx*vx + fu*vu at the stage point (t1, BB and J*w are recomputed: 36 FLOP instead of 12 live doubles)
template <bool DJ = false>
TS_HD void stage_jvp(const Inertia& I, const StagePt& sp, const double us[3], const double vx[7], const double vu[3], double out[7]) {
  const double dot = sp.s * vx[3] + sp.v[0] * vx[4] + sp.v[1] * vx[5] + sp.v[2] * vx[6];
  const double ds = (vx[3] - sp.s * dot) * sp.inq;
  const double dv[3] = {(vx[4] - sp.v[0] * dot) * sp.inq, (vx[5] - sp.v[1] * dot) * sp.inq, (vx[6] - sp.v[2] * dot) * sp.inq};
  const double dw[3] = {vx[0], vx[1], vx[2]};
  double a[3], b[3];
  cross3(dv, sp.w, a);
  cross3(sp.v, dw, b);
  out[3] = 0.5 * (-((dv[0] * sp.w[0] + dv[1] * sp.w[1] + dv[2] * sp.w[2]) + (sp.v[0] * dw[0] + sp.v[1] * dw[1] + sp.v[2] * dw[2])));
  for (int i = 0; i < 3; ++i) out[4 + i] = 0.5 * (ds * sp.w[i] + sp.s * dw[i] + a[i] + b[i]);
  double vxB[3], t1[3], cc[3], BB[3];
  cross3(sp.v, sp.Bn, vxB);
  for (int i = 0; i < 3; ++i) t1[i] = vxB[i] + sp.s * sp.Bn[i];
  cross3(sp.v, t1, cc);
  for (int i = 0; i < 3; ++i) BB[i] = sp.Bn[i] + 2.0 * cc[i];
  double dt1[3], c1[3], c2[3];
  cross3(dv, sp.Bn, dt1);
  for (int i = 0; i < 3; ++i) dt1[i] += ds * sp.Bn[i];
  cross3(dv, t1, c1);
  cross3(sp.v, dt1, c2);
  const double dBB[3] = {2.0 * (c1[0] + c2[0]), 2.0 * (c1[1] + c2[1]), 2.0 * (c1[2] + c2[2])};
  const double dus[3] = {vu[0] * 1.e-2, vu[1] * 1.e-2, vu[2] * 1.e-2};
  double ta[3], tb[3], g1[3], g2[3], Jw[3], Jdw[3];
  cross3(dus, BB, ta);
  cross3(us, dBB, tb);
  mul_J<DJ>(I, sp.w, Jw);
  mul_J<DJ>(I, dw, Jdw);
  cross3(dw, Jw, g1);
  cross3(sp.w, Jdw, g2);
  const double r[3] = {ta[0] + tb[0] - g1[0] - g2[0], ta[1] + tb[1] - g1[1] - g2[1], ta[2] + tb[2] - g1[2] - g2[2]};
  mul_Jinv<DJ>(I, r, out);
}
// Stage-1 JVPs of the rk3 step for the ten UNIT directions.  In the first stage the direction is e_c itself, so nine
// tenths of the generic stage_jvp are products with exact zeros; written out per direction type they cost ~25 (omega),
// ~70 (quaternion) and ~12 (control) operations instead of ~165, and the stage point's BB, t1 and J*w are formed once for
// all ten.  Same values as stage_jvp on e_c (sums with exact zeros dropped).  out = (fx e_c or fu e_c) * dt.
struct Stage1Pt {
  double t1[3], BB[3], Jw[3];
};
template <bool DJ>
TS_HD void stage1_prepare(const Inertia& I, const StagePt& sp, Stage1Pt& p1) {
  double vxB[3], cc[3];
  cross3(sp.v, sp.Bn, vxB);
  for (int i = 0; i < 3; ++i) p1.t1[i] = vxB[i] + sp.s * sp.Bn[i];
  cross3(sp.v, p1.t1, cc);
  for (int i = 0; i < 3; ++i) p1.BB[i] = sp.Bn[i] + 2.0 * cc[i];
  mul_J<DJ>(I, sp.w, p1.Jw);
}
template <bool DJ, int C>   // C = 0..2: direction e_C in omega
TS_HD void stage1_jvp_w(const Inertia& I, const StagePt& sp, const Stage1Pt& p1, double dt, double* out) {
  constexpr int C1 = (C + 1) % 3, C2 = (C + 2) % 3;
  // b = v x e_C, g1 = e_C x (J w), g2 = w x (J e_C)
  double b[3], g1[3], g2[3], Jc[3], r[3], o3[3];
  b[C] = 0.0; b[C1] = sp.v[C2]; b[C2] = -sp.v[C1];
  g1[C] = 0.0; g1[C1] = -p1.Jw[C2]; g1[C2] = p1.Jw[C1];
  if (DJ) {
    Jc[C] = I.J[C * 4]; Jc[C1] = 0.0; Jc[C2] = 0.0;
    g2[C] = 0.0; g2[C1] = sp.w[C2] * Jc[C]; g2[C2] = -(sp.w[C1] * Jc[C]);
  } else {
    for (int i = 0; i < 3; ++i) Jc[i] = I.J[i * 3 + C];
    cross3(sp.w, Jc, g2);
  }
  for (int i = 0; i < 3; ++i) r[i] = -g1[i] - g2[i];
  mul_Jinv<DJ>(I, r, o3);
  for (int i = 0; i < 3; ++i) out[i] = o3[i] * dt;
  out[3] = (0.5 * (-sp.v[C])) * dt;
  for (int i = 0; i < 3; ++i) out[4 + i] = (0.5 * (((i == C) ? sp.s : 0.0) + b[i])) * dt;
}
template <bool DJ, int C>   // C = 0..3: direction e_C in the raw quaternion (s, v0, v1, v2)
TS_HD void stage1_jvp_q(const Inertia& I, const StagePt& sp, const Stage1Pt& p1, const double us[3], double dt, double* out) {
  const double dot = (C == 0) ? sp.s : sp.v[C > 0 ? C - 1 : 0];   // q_hat[C]
  const double ds = (((C == 0) ? 1.0 : 0.0) - sp.s * dot) * sp.inq;
  double dv[3];
  for (int i = 0; i < 3; ++i) dv[i] = (((C == i + 1) ? 1.0 : 0.0) - sp.v[i] * dot) * sp.inq;
  double a[3];
  cross3(dv, sp.w, a);
  out[3] = (0.5 * (-(dv[0] * sp.w[0] + dv[1] * sp.w[1] + dv[2] * sp.w[2]))) * dt;
  for (int i = 0; i < 3; ++i) out[4 + i] = (0.5 * (ds * sp.w[i] + a[i])) * dt;
  double dt1[3], c1[3], c2[3];
  cross3(dv, sp.Bn, dt1);
  for (int i = 0; i < 3; ++i) dt1[i] += ds * sp.Bn[i];
  cross3(dv, p1.t1, c1);
  cross3(sp.v, dt1, c2);
  const double dBB[3] = {2.0 * (c1[0] + c2[0]), 2.0 * (c1[1] + c2[1]), 2.0 * (c1[2] + c2[2])};
  double tb[3], o3[3];
  cross3(us, dBB, tb);
  mul_Jinv<DJ>(I, tb, o3);
  for (int i = 0; i < 3; ++i) out[i] = o3[i] * dt;
}
template <bool DJ, int C>   // C = 0..2: direction e_C in the control
TS_HD void stage1_jvp_u(const Inertia& I, const Stage1Pt& p1, double dt, double* out) {
  constexpr int C1 = (C + 1) % 3, C2 = (C + 2) % 3;
  double ta[3], o3[3];   // (1e-2 e_C) x BB
  ta[C] = 0.0; ta[C1] = -(1.e-2 * p1.BB[C2]); ta[C2] = 1.e-2 * p1.BB[C1];
  mul_Jinv<DJ>(I, ta, o3);
  for (int i = 0; i < 3; ++i) out[i] = o3[i] * dt;
  for (int i = 3; i < 7; ++i) out[i] = 0.0;
}
// Jacobian of the rk3 step by JVPs; column c of [A|B] is written to colmajor[c*7 .. c*7+6] (the lane's knot record in
// shared memory, which also holds the ten stage-1 products between the two passes).
template <bool DJ = false>
TS_HD void rk3_jac7_jvp(const Inertia& I, const double x[7], const double u[3], const double* B1, const double* B2, const double* B3,
                        double dt, double* colmajor) {
  StagePt s1, s2, s3;
  double k1[7], k2[7], k3[7], xs[7];
  const double us[3] = {u[0] * 1.e-2, u[1] * 1.e-2, u[2] * 1.e-2};
  stage_eval<DJ>(I, x, us, B1, s1, k1);
  {  // pass 1: t1_c = (stage-1 JVP of e_c) * dt, unrolled over the ten directions
    Stage1Pt p1;
    stage1_prepare<DJ>(I, s1, p1);
    stage1_jvp_w<DJ, 0>(I, s1, p1, dt, colmajor + 0);
    stage1_jvp_w<DJ, 1>(I, s1, p1, dt, colmajor + 7);
    stage1_jvp_w<DJ, 2>(I, s1, p1, dt, colmajor + 14);
    stage1_jvp_q<DJ, 0>(I, s1, p1, us, dt, colmajor + 21);
    stage1_jvp_q<DJ, 1>(I, s1, p1, us, dt, colmajor + 28);
    stage1_jvp_q<DJ, 2>(I, s1, p1, us, dt, colmajor + 35);
    stage1_jvp_q<DJ, 3>(I, s1, p1, us, dt, colmajor + 42);
    stage1_jvp_u<DJ, 0>(I, p1, dt, colmajor + 49);
    stage1_jvp_u<DJ, 1>(I, p1, dt, colmajor + 56);
    stage1_jvp_u<DJ, 2>(I, p1, dt, colmajor + 63);
  }
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * (k1[i] * dt);
  stage_eval<DJ>(I, xs, us, B2, s2, k2);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] - k1[i] * dt + 2.0 * (k2[i] * dt);
  stage_eval<DJ>(I, xs, us, B3, s3, k3);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int c = 0; c < 10; ++c) {   // pass 2: stages 2 and 3 (generic JVPs)
    double vx[7], vu[3], t1[7], t2[7], t3[7], y[7];
    for (int i = 0; i < 7; ++i) vx[i] = (i == c) ? 1.0 : 0.0;
    for (int i = 0; i < 3; ++i) vu[i] = (7 + i == c) ? 1.0 : 0.0;
    for (int i = 0; i < 7; ++i) {
      t1[i] = colmajor[c * 7 + i];
      y[i] = vx[i] + 0.5 * t1[i];
    }
    stage_jvp<DJ>(I, s2, us, y, vu, t2);
    for (int i = 0; i < 7; ++i) {
      t2[i] *= dt;
      y[i] = vx[i] - t1[i] + 2.0 * t2[i];
    }
    stage_jvp<DJ>(I, s3, us, y, vu, t3);
    for (int i = 0; i < 7; ++i) colmajor[c * 7 + i] = vx[i] + (t1[i] + 4.0 * t2[i] + t3[i] * dt) * TS_SIXTH;
  }
}

// Directional derivatives of the rk4 ZOH step (attitude_controller.jl:122-132,134-145) by JVPs through the four stages:
// out[d*7 .. d*7+6] = d xn / d (x,u) . (vx_d, vu_d) for nd directions (dirs: nd x 10 = [vx(7) | vu(3)]).
// Same register-resident scheme as rk3_jac7_jvp: four 11-double stage records + one direction in flight.
TS_HD void rk4_jvp7(const Inertia& I, const double x[7], const double u[3], const double* B1, const double* B2, const double* B3,
                    const double* B4, double h, int nd, const double* dirs, double* out) {
  StagePt s1, s2, s3, s4;
  double k1[7], k2[7], k3[7], k4[7], xs[7];
  const double us[3] = {u[0] * 1.e-2, u[1] * 1.e-2, u[2] * 1.e-2};
  stage_eval(I, x, us, B1, s1, k1);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * (k1[i] * h);
  stage_eval(I, xs, us, B2, s2, k2);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * (k2[i] * h);
  stage_eval(I, xs, us, B3, s3, k3);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + k3[i] * h;
  stage_eval(I, xs, us, B4, s4, k4);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int d = 0; d < nd; ++d) {
    double vx[7], vu[3], t1[7], t2[7], t3[7], t4[7], y[7];
    for (int i = 0; i < 7; ++i) vx[i] = dirs[d * 10 + i];
    for (int i = 0; i < 3; ++i) vu[i] = dirs[d * 10 + 7 + i];
    stage_jvp(I, s1, us, vx, vu, t1);
    for (int i = 0; i < 7; ++i) {
      t1[i] *= h;
      y[i] = vx[i] + 0.5 * t1[i];
    }
    stage_jvp(I, s2, us, y, vu, t2);
    for (int i = 0; i < 7; ++i) {
      t2[i] *= h;
      y[i] = vx[i] + 0.5 * t2[i];
    }
    stage_jvp(I, s3, us, y, vu, t3);
    for (int i = 0; i < 7; ++i) {
      t3[i] *= h;
      y[i] = vx[i] + t3[i];
    }
    stage_jvp(I, s4, us, y, vu, t4);
    for (int i = 0; i < 7; ++i) out[d * 7 + i] = vx[i] + (t1[i] + 2.0 * t2[i] + 2.0 * t3[i] + t4[i] * h) * TS_SIXTH;
  }
}

// rk4 ZOH step (attitude_controller.jl:122-132) with one field row per stage.
template <int UMODE>
TS_HD void rk4_jac7(const Inertia& I, const double x[7], const double u[3], const double* B1, const double* B2, const double* B3,
                    const double* B4, double dt, double xn[7], double AB[70]) {
  double k1[7], k2[7], k3[7], k4[7], xs[7], fx[49], fu[9];
  double K1[70], K2[70], K3[70], K4[70], M[70];
  dyn_f_jac<UMODE>(I, x, u, B1, k1, fx, fu);
  for (int i = 0; i < 7; ++i) k1[i] = k1[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) {
      double a = (j < 7) ? fx[i * 7 + j] : ((i < 3) ? fu[i * 3 + (j - 7)] : 0.0);
      K1[i * 10 + j] = a * dt;
    }
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * k1[i];
  dyn_f_jac<UMODE>(I, xs, u, B2, k2, fx, fu);
  for (int i = 0; i < 7; ++i) k2[i] = k2[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) M[i * 10 + j] = ((i == j) ? 1.0 : 0.0) + 0.5 * K1[i * 10 + j];
  stage_chain(fx, fu, M, dt, K2);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + 0.5 * k2[i];
  dyn_f_jac<UMODE>(I, xs, u, B3, k3, fx, fu);
  for (int i = 0; i < 7; ++i) k3[i] = k3[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) M[i * 10 + j] = ((i == j) ? 1.0 : 0.0) + 0.5 * K2[i * 10 + j];
  stage_chain(fx, fu, M, dt, K3);
  for (int i = 0; i < 7; ++i) xs[i] = x[i] + k3[i];
  dyn_f_jac<UMODE>(I, xs, u, B4, k4, fx, fu);
  for (int i = 0; i < 7; ++i) k4[i] = k4[i] * dt;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j) M[i * 10 + j] = ((i == j) ? 1.0 : 0.0) + K3[i * 10 + j];
  stage_chain(fx, fu, M, dt, K4);
  for (int i = 0; i < 7; ++i) xn[i] = x[i] + (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) * TS_SIXTH;
  for (int i = 0; i < 7; ++i)
    for (int j = 0; j < 10; ++j)
      AB[i * 10 + j] = ((i == j) ? 1.0 : 0.0) + (K1[i * 10 + j] + 2.0 * K2[i * 10 + j] + 2.0 * K3[i * 10 + j] + K4[i * 10 + j]) * TS_SIXTH;
}

// Sequentially accumulated clock state of the reference (x8), replicated exactly:
// c = rate*dt ; stages at x8, x8 + c/2, x8 - c + 2c ; x8 += (c + 4c + c)/6.   No FMA.
struct ClockStep {
  double t1, t2, t3, next;
};
TS_HD ClockStep clock_rk3(double x8, double rate, double dt) {
#ifdef __CUDA_ARCH__
  const double c = __dmul_rn(rate, dt);
  ClockStep s;
  s.t1 = x8;
  s.t2 = __dadd_rn(x8, __ddiv_rn(c, 2.0));
  s.t3 = __dadd_rn(__dsub_rn(x8, c), __dmul_rn(2.0, c));
  s.next = __dadd_rn(x8, __ddiv_rn(__dadd_rn(__dadd_rn(c, __dmul_rn(4.0, c)), c), 6.0));
  return s;
#else
  const volatile double c = rate * dt;
  ClockStep s;
  s.t1 = x8;
  volatile double h = c / 2.0;
  s.t2 = x8 + h;
  volatile double a = x8 - c;
  volatile double b = 2.0 * c;
  s.t3 = a + b;
  volatile double c4 = 4.0 * c;
  volatile double sum = c + c4;
  sum = sum + c;
  volatile double inc = sum / 6.0;
  s.next = x8 + inc;
  return s;
#endif
}
// 0-based field row for clock value t: floor(t*index_scale + 1) - 1, clamped to the table.
TS_HD int field_row(double t, double index_scale, long long rows) {
#ifdef __CUDA_ARCH__
  const double v = __dadd_rn(__dmul_rn(t, index_scale), 1.0);
#else
  volatile double p = t * index_scale;
  const double v = p + 1.0;
#endif
  long long idx = (long long)floor(v);
  if (idx < 1) idx = 1;
  if (idx > rows) idx = rows;
  return (int)(idx - 1);
}

}  // namespace ts
