// IGRF-12 geocentric field synthesis -- device code shared by K1 (batched
// IGRF) and K2 (per-trial field tables).
//
// Replaces, on the GPU, the per-point work of the reference's
//   igrf12(date, r, lat, lon)            src/igrf.jl:70-274
//   legendre(Val{:schmidt}, theta, 13)   src/legendre.jl:254-292
//   dlegendre(Val{:schmidt}, theta, P)   src/dlegendre.jl:221-309
// One thread evaluates one point.  The Legendre rows n, n-1, n-2 and the
// cos/sin(m*phi) tables live in registers (fully unrolled n,m loops, every
// index a compile-time constant); the date-interpolated Gauss coefficients are
// staged per block in shared memory (uniform-address LDS.128 broadcasts of a
// (g,h) pair); the position-independent recursion constants come from
// __constant__ memory and fold into the FP64 instructions as c[][] operands.
//
// Arithmetic notes (parity with the CPU oracle is 1e-10 relative, measured
// ~1e-14):
//  * s = sqrt(1 - c*c) is computed with two roundings (__dmul_rn, __dsub_rn),
//    never fma(-c,c,1): the reference's cancellation near the poles, including
//    s == 0 exactly within ~1.5e-8 rad of a pole, must be reproduced.
//  * theta == 0 (lat == pi/2 exactly) takes the reference's pole branch.
//  * everything else may be FMA-contracted by nvcc.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ts {

constexpr int IGRF_NCOEF = 104;  // n=1..13, m=0..n

// Recursion constants, filled once per context by igrf_upload_constants().
//  leg_a/leg_b[n][m] (m<n): a_nm, b_nm of legendre.jl:278-281; leg_d[n]: sqrt((2n-1)/(2n))
//  dl_a/dl_b[n][m]: +-0.5*sqrt(..) of dlegendre.jl:274-299
struct IgrfConsts {
  double leg_a[14][14];
  double leg_b[14][14];
  double leg_d[14];
  double dl_a[14][14];
  double dl_b[14][14];
  // Rescaled recursion of the field kernels (igrf12_point): R[n][m] = P[n][m] / kap[n][m] with kap chosen so that the
  // three-term recurrence loses one multiplication per (n,m):  R[n][m] = (rl_a[n][m] * c) * R[n-1][m] - R[n-2][m]
  // (kap[m][m] = kap[m+1][m] = 1, kap[n][m] = leg_b[n][m] * kap[n-2][m]); the factor kap is folded into the staged Gauss
  // coefficients, and the derivative constants are rescaled to give dP[n][m] / kap[n][m] directly:
  //   dP'[n][0] = rd_0[n] * R[n][1];   dP'[n][m] = rd_a[n][m] * R[n][m-1] + rd_b[n][m] * R[n][m+1]
  double kap[14][14];
  double rl_a[14][14];
  double rd_0[14];
  double rd_a[14][14];
  double rd_b[14][14];
};
// The library is built as ONE translation unit (csrc/tortoise_b200.cu), so the
// definition lives here.
__constant__ IgrfConsts c_igrf;

// Interpolates the Gauss coefficients to `date` into shared memory:
// s_gh[k] = (g_k, h_k), k = n(n+1)/2 - 1 + m.  tabG/tabH are the 104x25 / 91x25
// device tables.  Mirrors igrf.jl:112-120,170-179,210-226.  Call with all
// threads of the block, followed by __syncthreads().
__device__ __forceinline__ void igrf_stage_coeffs(double2* s_gh, const double* __restrict__ tabG,
                                                  const double* __restrict__ tabH, double date) {
  const int idx = (date < 2020.0) ? (int)floor((date - 1900.0) * 0.2 + 1.0) : 24;
  const int epoch = 1900 + (idx - 1) * 5;
  const double dt = date - (double)epoch;
  const int c0 = idx - 1;
  const bool interp = date < 2015.0;
  for (int k = threadIdx.x; k < IGRF_NCOEF; k += blockDim.x) {
    // recover (n,m) from k
    int n = 1;
    while ((n + 1) * (n + 2) / 2 - 1 <= k) ++n;
    const int m = k - (n * (n + 1) / 2 - 1);
    const double g0 = tabG[k * 25 + c0];
    const double dg = interp ? __ddiv_rn(__dsub_rn(tabG[k * 25 + c0 + 1], g0), 5.0) : tabG[k * 25 + 24];
    double2 v;
    v.x = __dadd_rn(g0, __dmul_rn(dg, dt));
    v.y = 0.0;
    if (m > 0) {
      const int kh = k - n;  // H has no m=0 rows
      const double h0 = tabH[kh * 25 + c0];
      const double dh = interp ? __ddiv_rn(__dsub_rn(tabH[kh * 25 + c0 + 1], h0), 5.0) : tabH[kh * 25 + 24];
      v.y = __dadd_rn(h0, __dmul_rn(dh, dt));
    }
    // the rescaling factor of the Legendre recursion (IgrfConsts::kap) rides on the coefficients
    const double kp = c_igrf.kap[n][m];
    v.x *= kp;
    v.y *= kp;
    s_gh[k] = v;
    // (-m g, +m h): folds the factor -m of the d/dphi sum into the coefficients (igrf.jl:235)
    double2 vm;
    vm.x = -(double)m * v.x;
    vm.y = (double)m * v.y;
    s_gh[IGRF_NCOEF + k] = vm;
  }
}

__host__ __device__ inline int igrf_nmax_for_date(double date) {
  const int idx = (date < 2020.0) ? (int)floor((date - 1900.0) * 0.2 + 1.0) : 24;
  const int epoch = 1900 + (idx - 1) * 5;
  return (epoch < 1995) ? 10 : 13;
}

// Field at one point.  r in metres; lat, lon in rad; out (north, east, down) in nT.  POLE: theta == 0 exactly (the
// reference's pole branch, igrf.jl:235,270) -- a separate instantiation, so the 91 per-(n,m) selects between P and dP
// are not paid by every other point.
template <int NMAX, bool POLE>
__device__ __forceinline__ void igrf12_point_t(const double2* __restrict__ s_gh, double r_m, double lat, double lon,
                                               double& bn, double& be, double& bd) {
  const double PI = 3.141592653589793;
  const double theta = PI / 2 - lat;
  const double phi = (lon >= 0.0) ? lon : 2 * PI + lon;
  const double r = r_m / 1000.0;

  const double c = cos(theta);
  const double s = sqrt(__dsub_rn(1.0, __dmul_rn(c, c)));
  double sin_phi, cos_phi;
  sincos(phi, &sin_phi, &cos_phi);

  // cos(m phi), sin(m phi), m = 0..NMAX (igrf.jl:190-203 recurrence, identical for every n)
  double cm[NMAX + 1], sm[NMAX + 1];
  cm[0] = 1.0;
  sm[0] = 0.0;
  {
    const double two_c = 2 * cos_phi;
    double s2 = -sin_phi, c2 = cos_phi;  // sin(-phi), cos(-phi)
#pragma unroll
    for (int m = 1; m <= NMAX; ++m) {
      sm[m] = two_c * sm[m - 1] - s2;
      cm[m] = two_c * cm[m - 1] - c2;
      s2 = sm[m - 1];
      c2 = cm[m - 1];
    }
  }

  const double a = 6371.2;
  const double ratio = a / r;
  const double inv_r = 1.0 / r;
  double fact = ratio;

  double dVr = 0.0, dVt = 0.0, dVp = 0.0;
  double Pm1[NMAX + 2], Pm2[NMAX + 2], Pn[NMAX + 2];
#pragma unroll
  for (int i = 0; i < NMAX + 2; ++i) Pm1[i] = Pm2[i] = Pn[i] = 0.0;
  Pm1[0] = 1.0;  // row n = 0

#pragma unroll
  for (int n = 1; n <= NMAX; ++n) {
    // ---- Legendre row n (legendre.jl:271-290)
    if (n == 1) {
      Pn[0] = c;
      Pn[1] = s;
    } else {
#pragma unroll
      for (int m = 0; m <= n - 1; ++m) Pn[m] = fma(c_igrf.rl_a[n][m] * c, Pm1[m], -Pm2[m]);   // rescaled rows: R = P / kap
      Pn[n] = s * c_igrf.leg_d[n] * Pm1[n - 1];
    }
    Pn[n + 1] = 0.0;
    // -(n+1)/r : shared-reciprocal division with one correction step
    const double num = -(double)(n + 1);
    double cr = num * inv_r;
    cr = fma(fma(-cr, r, num), inv_r, cr);

    const int k0 = n * (n + 1) / 2 - 1;
    // ---- m = 0
    double aux_r, aux_t, aux_p = 0.0;
    {
      const double g = s_gh[k0].x;
      const double dP0 = c_igrf.rd_0[n] * Pn[1];
      aux_r = g * Pn[0];  // -(n+1)/r is factored out of the sum over m (one DMUL per n instead of per (n,m))
      aux_t = g * dP0;
    }
    // ---- m = 1..n
#pragma unroll
    for (int m = 1; m <= n; ++m) {
      const double2 gh = s_gh[k0 + m];
      double dPm;
      if (m == n)
        dPm = c_igrf.rd_a[n][m] * Pn[m - 1];                                   // (n = m = 1 reads the zero P[1][2]: same value)
      else
        dPm = c_igrf.rd_a[n][m] * Pn[m - 1] + c_igrf.rd_b[n][m] * Pn[m + 1];
      const double2 ghm = s_gh[IGRF_NCOEF + k0 + m];
      const double GcHs = gh.x * cm[m] + gh.y * sm[m];
      const double mGsHc = ghm.x * sm[m] + ghm.y * cm[m];  // = -m (g sin - h cos)
      aux_r += GcHs * Pn[m];
      aux_t += GcHs * dPm;
      aux_p += mGsHc * (POLE ? dPm : Pn[m]);
    }
    fact *= ratio;
    dVr += (cr * aux_r) * fact;
    dVp += aux_p * fact;
    dVt += aux_t * fact;
    // rotate rows
#pragma unroll
    for (int i = 0; i < NMAX + 2; ++i) {
      Pm2[i] = Pm1[i];
      Pm1[i] = Pn[i];
    }
  }
  dVr *= a;
  dVp *= a;
  dVt *= a;
  bn = inv_r * dVt;
  be = POLE ? (-1.0 / r) * dVp : (-1.0 / (r * sin(theta))) * dVp;
  bd = dVr;
}
// the pole instantiation is kept out of line: it is taken by measure-zero inputs (the lat = +pi/2 row of a grid map) and
// must not cost the common path registers
template <int NMAX>
__device__ __noinline__ void igrf12_point_pole(const double2* s_gh, double r_m, double lat, double lon, double* out3) {
  igrf12_point_t<NMAX, true>(s_gh, r_m, lat, lon, out3[0], out3[1], out3[2]);
}
template <int NMAX>
__device__ __forceinline__ void igrf12_point(const double2* __restrict__ s_gh, double r_m, double lat, double lon,
                                             double& bn, double& be, double& bd) {
  if (3.141592653589793 / 2 - lat == 0.0) {
    double o[3];
    igrf12_point_pole<NMAX>(s_gh, r_m, lat, lon, o);
    bn = o[0];
    be = o[1];
    bd = o[2];
  } else {
    igrf12_point_t<NMAX, false>(s_gh, r_m, lat, lon, bn, be, bd);
  }
}

}  // namespace ts
