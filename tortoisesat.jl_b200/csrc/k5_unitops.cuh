// K5 -- element-wise batch versions of the reference's small building blocks, so that every
// function on the hot path can be called -- and parity-checked -- on its own:
//   kep_ECI(kep,t0,GM)                              reference src/kep_ECI.jl:1-49
//   OrbitPlotter(x,p,t)                             src/OrbitPlotter.jl:1-52
//   legendre(Val{:schmidt},phi,n_max,false)         src/legendre.jl:254-292
//   dlegendre(Val{:schmidt},phi,P,false)            src/dlegendre.jl:221-309
//   DerivFunction / gain_simulator (8 state)        src/DerivFunction.jl:1-48, src/gain_simulator.jl:1-53
//   attitude_dynamics(x,u,B_B,J) (7 state)          src/attitude_dynamics.jl:2-24
//   rk3 ZOH step                                    src/attitude_controller.jl:178-187
// One thread per element; these are correctness/drop-in paths, not performance paths (the fused
// kernels K1..K4 inline the same device functions).
#pragma once
#include "common.cuh"
#include "ilqr_math.cuh"
#include "k2_field.cuh"

namespace ts {

__global__ void k5_kep_eci(int64_t n, const double* __restrict__ kep6, const double* __restrict__ t0, double GM, double* __restrict__ rv6) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u[6];
  kep_eci_dev(kep6 + i * 6, t0 ? t0[i] : 0.0, GM, u);
  for (int c = 0; c < 6; ++c) rv6[i * 6 + c] = u[c];
}

__global__ void k5_orbit_rhs(int64_t n, const double* __restrict__ x6, double* __restrict__ dx6) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[6], dx[6];
  for (int c = 0; c < 6; ++c) x[c] = x6[i * 6 + c];
  orbit_rhs_dev(x, dx);
  for (int c = 0; c < 6; ++c) dx6[i * 6 + c] = dx[c];
}

// P, dP: n x (nmax+1)^2 row-major, entries above the diagonal 0 (like the reference's matrices).
__global__ void k5_legendre(int64_t n, const double* __restrict__ theta, int nmax, double* __restrict__ P, double* __restrict__ dP) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = nmax + 1;
  double* p = P + i * d * d;
  for (int k = 0; k < d * d; ++k) p[k] = 0.0;
  const double th = theta[i];
  const double c = cos(th);
  const double s = sqrt(__dsub_rn(1.0, __dmul_rn(c, c)));
  p[0] = 1.0;
  p[1 * d + 0] = c;
  p[1 * d + 1] = s;
  for (int nn = 2; nn <= nmax; ++nn) {
    for (int m = 0; m <= nn - 1; ++m) p[nn * d + m] = c_igrf.leg_a[nn][m] * c * p[(nn - 1) * d + m] - c_igrf.leg_b[nn][m] * p[(nn - 2) * d + m];
    p[nn * d + nn] = s * c_igrf.leg_d[nn] * p[(nn - 1) * d + (nn - 1)];
  }
  if (!dP) return;
  double* q = dP + i * d * d;
  for (int k = 0; k < d * d; ++k) q[k] = 0.0;
  const double PI = 3.141592653589793;
  double ph = fmod(th, 2 * PI);
  if (ph < 0) ph += 2 * PI;
  const double fact = (ph > PI) ? -1.0 : 1.0;
  for (int nn = 1; nn <= nmax; ++nn)
    for (int m = 0; m <= nn; ++m) {
      double v;
      if (m == 0)
        v = -c_igrf.dl_a[nn][0] * p[nn * d + 1] + c_igrf.dl_b[nn][0] * p[nn * d + 1];
      else if (m == nn && m != 1)
        v = c_igrf.dl_a[nn][m] * p[nn * d + m - 1];
      else
        v = c_igrf.dl_a[nn][m] * p[nn * d + m - 1] + c_igrf.dl_b[nn][m] * ((m + 1 <= nmax) ? p[nn * d + m + 1] : 0.0);
      q[nn * d + m] = v * fact;
    }
}

// mode 0: DerivFunction, 1: gain_simulator (8-state x, field row looked up from the clock state);
// mode 2: attitude_dynamics (7-state x, B = body-frame field per element, no 1e-2 scaling).
__global__ void k5_dynamics(int mode, int64_t n, const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ B,
                            int64_t B_rows, double index_scale, double clock_rate, const double* __restrict__ Jmat,
                            double* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Inertia I;
  for (int c = 0; c < 9; ++c) I.J[c] = Jmat[c];
  inv3_gj(I.J, I.Jinv);
  if (mode == 2) {
    const double* xi = x + i * 7;
    // attitude_dynamics takes the BODY-frame field and the raw moment: tau = u x B_B
    const double nq = sqrt(xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5] + xi[6] * xi[6]);
    const double q[4] = {xi[3] / nq, xi[4] / nq, xi[5] / nq, xi[6] / nq};
    const double w4[4] = {0.0, xi[0], xi[1], xi[2]};
    double qd[4], tau[3], Jw[3], wJw[3];
    qmult(q, w4, qd);
    cross3(u + i * 3, B + i * 3, tau);
    for (int c = 0; c < 3; ++c) Jw[c] = I.J[c * 3 + 0] * xi[0] + I.J[c * 3 + 1] * xi[1] + I.J[c * 3 + 2] * xi[2];
    cross3(xi, Jw, wJw);
    const double r0 = tau[0] - wJw[0], r1 = tau[1] - wJw[1], r2 = tau[2] - wJw[2];
    for (int c = 0; c < 3; ++c) dx[i * 7 + c] = I.Jinv[c * 3 + 0] * r0 + I.Jinv[c * 3 + 1] * r1 + I.Jinv[c * 3 + 2] * r2;
    for (int c = 0; c < 4; ++c) dx[i * 7 + 3 + c] = 0.5 * qd[c];
    return;
  }
  const double* xi = x + i * 8;
  const double* Bn = B + (int64_t)field_row(xi[7], index_scale, B_rows) * 3;
  double d7[7];
  if (mode == 0)
    dyn_f<0>(I, xi, u + i * 3, Bn, d7);
  else
    dyn_f<1>(I, xi, u + i * 3, Bn, d7);
  for (int c = 0; c < 7; ++c) dx[i * 8 + c] = d7[c];
  dx[i * 8 + 7] = clock_rate;
}

// rk3 ZOH step of the 8-state model (TrajOpt rk3 o DerivFunction), field rows looked up per stage from the clock.
__global__ void k5_rk3_step(int64_t n, const double* __restrict__ x, const double* __restrict__ u, const double* __restrict__ B,
                            int64_t B_rows, double index_scale, double clock_rate, const double* __restrict__ Jmat, double dt,
                            double* __restrict__ xn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Inertia I;
  for (int c = 0; c < 9; ++c) I.J[c] = Jmat[c];
  inv3_gj(I.J, I.Jinv);
  const double* xi = x + i * 8;
  const ClockStep cs = clock_rk3(xi[7], clock_rate, dt);
  double o7[7];
  rk3_step7<0>(I, xi, u + i * 3, B + (int64_t)field_row(cs.t1, index_scale, B_rows) * 3, B + (int64_t)field_row(cs.t2, index_scale, B_rows) * 3,
               B + (int64_t)field_row(cs.t3, index_scale, B_rows) * 3, dt, o7);
  for (int c = 0; c < 7; ++c) xn[i * 8 + c] = o7[c];
  xn[i * 8 + 7] = cs.next;
}

}  // namespace ts
