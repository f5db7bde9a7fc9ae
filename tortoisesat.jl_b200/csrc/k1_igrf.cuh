// K1 -- batched IGRF-12 (geocentric), one point per thread, SoA in / SoA out.
// Replaces n scalar calls of igrf12() [reference src/igrf.jl:67-274], e.g. the
// 10^6-point map of igrf_data() [src/magnetic_toolbox.jl:108-121] and
// BASELINE config "10^8 random LEO points".
//
// Roofline: FP64 CUDA-core pipe.  Algorithmic work 2243 FLOP/point,
// 48 B/point of HBM traffic (3 doubles in, 3 out) -> AI ~ 47 FLOP/B, far right of
// the ridge; HBM is not binding.  Loads/stores are fully coalesced (consecutive
// threads -> consecutive doubles of each SoA array).
#pragma once
#include "common.cuh"
#include "igrf_device.cuh"

namespace ts {

constexpr int K1_THREADS = 128;
constexpr int K1_MIN_BLOCKS = 4;  // <= 128 registers: 16 warps per SM keep the FP64 pipe busier than 168 regs / 12 warps

// The date-interpolated, rescaled Gauss coefficients of one call (2 x 104 (g,h) pairs, igrf_stage_coeffs), computed ONCE
// per call by this one-block kernel; every block of K1 then copies the 3.3 KB table from L2 into its shared memory.
// (Round 1 re-did the interpolation -- table look-ups, two FP64 divisions, the (n,m) search -- in every block of 128
// points: ~17 % of a warp's lifetime spent before its first FP64 instruction.)
__global__ void __launch_bounds__(128) k1_stage_kernel(const double* __restrict__ tabG, const double* __restrict__ tabH, double date,
                                                       double2* __restrict__ gh) {
  igrf_stage_coeffs(gh, tabG, tabH, date);
}

template <int NMAX>
__global__ void __launch_bounds__(K1_THREADS, K1_MIN_BLOCKS)
k1_igrf12_batch(const double2* __restrict__ gh, int64_t n,
                const double* __restrict__ r_m, const double* __restrict__ lat, const double* __restrict__ lon,
                double* __restrict__ Bn, double* __restrict__ Be, double* __restrict__ Bd, int* __restrict__ bad_flag) {
  __shared__ double2 s_gh[2 * IGRF_NCOEF];
  // the point's own inputs are requested first: their DRAM latency overlaps the table copy and the barrier
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  const double la = live ? lat[i] : 0.0, lo = live ? lon[i] : 0.0, rr = live ? r_m[i] : 6371200.0;
  for (int k = threadIdx.x; k < 2 * IGRF_NCOEF; k += K1_THREADS) s_gh[k] = gh[k];
  __syncthreads();
  const double PI = 3.141592653589793;
  // one point per thread: a grid-stride loop lets the compiler hoist the c[][] recursion
  // constants into registers (255 regs + spills); without it the body needs 168, no spills.
  if (!live) return;
  double bn, be, bd;
  // reference validation, igrf.jl:84-88 (NaN inputs fail it too)
  if (!(la >= -PI / 2 && la <= PI / 2 && lo >= -PI && lo <= PI)) {
    bn = be = bd = nan("");
    *bad_flag = 1;
  } else {
    igrf12_point<NMAX>(s_gh, rr, la, lo, bn, be, bd);
  }
  Bn[i] = bn;
  Be[i] = be;
  Bd[i] = bd;
}

// Register-resident DFMA micro-benchmark: 8 independent chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

// Latency micro-benchmark (one warp): cycles per DEPENDENT operation for DFMA, DADD/DMUL pairs, rsqrt, and a shared-memory
// store -> __syncwarp -> load round trip: the numbers that bound the strictly sequential Riccati / rollout chains of K3.
__global__ void __launch_bounds__(32) fp64_latency_kernel(double* out, int iters, double a, double b) {
  __shared__ double sm[64];
  double x = threadIdx.x * 1e-9 + 1.0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x = fma(x, a, b);
  }
  long long t1 = clock64();
  double y = x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      y = __dadd_rn(y, b);
      y = __dmul_rn(y, a);
    }
  }
  long long t2 = clock64();
  double z = fabs(y) + 2.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) z = rsqrt(z) + 2.0;
  }
  long long t3 = clock64();
  double w = z;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[threadIdx.x] = w;
      __syncwarp();
      w = sm[(threadIdx.x + 1) & 31] + b;
      __syncwarp();
    }
  }
  long long t4 = clock64();
  if (threadIdx.x == 0) {
    out[0] = (double)(t1 - t0) / ((double)iters * 32);   // DFMA
    out[1] = (double)(t2 - t1) / ((double)iters * 32);   // DADD / DMUL
    out[2] = (double)(t3 - t2) / ((double)iters * 8);    // rsqrt + DADD
    out[3] = (double)(t4 - t3) / ((double)iters * 8);    // STS + syncwarp + LDS + DADD + syncwarp
    out[4] = w;
  }
}

inline void igrf_host_constants(IgrfConsts& h) {
  memset(&h, 0, sizeof(h));
  for (int n = 1; n <= 13; ++n) {
    for (int m = 0; m <= n - 1; ++m) {
      if (n >= 2) {
        const long aux = (long)(n - m) * (n + m);
        h.leg_a[n][m] = sqrt((double)((2 * n - 1) * (2 * n - 1)) / (double)aux);
        h.leg_b[n][m] = sqrt((double)((n + m - 1) * (n - m - 1)) / (double)aux);
      }
    }
    h.leg_d[n] = sqrt((double)(2 * n - 1) / (double)(2 * n));
    for (int m = 0; m <= n; ++m) {
      if (m == 0) {
        const double aux = sqrt((double)(n * (n + 1)) / 2.0);
        h.dl_a[n][0] = +0.5 * aux;
        h.dl_b[n][0] = -0.5 * aux;
      } else if (m == 1) {
        h.dl_a[n][1] = +0.5 * sqrt((double)(2 * n * (n + 1)));
        h.dl_b[n][1] = -0.5 * sqrt((double)((n + 2) * (n - 1)));
      } else {
        h.dl_a[n][m] = +0.5 * sqrt((double)((n + m) * (n - m + 1)));
        h.dl_b[n][m] = -0.5 * sqrt((double)((n + m + 1) * (n - m)));
      }
    }
  }
  // rescaled recursion (see IgrfConsts): kap, then the constants of R = P / kap and dP' = dP / kap
  for (int n = 0; n <= 13; ++n)
    for (int m = 0; m <= 13; ++m) h.kap[n][m] = 1.0;
  for (int m = 0; m <= 13; ++m)
    for (int n = m + 2; n <= 13; ++n) h.kap[n][m] = h.leg_b[n][m] * h.kap[n - 2][m];
  for (int n = 2; n <= 13; ++n)
    for (int m = 0; m <= n - 1; ++m) h.rl_a[n][m] = h.leg_a[n][m] * h.kap[n - 1][m] / h.kap[n][m];
  for (int n = 1; n <= 13; ++n) {
    h.rd_0[n] = (-h.dl_a[n][0] + h.dl_b[n][0]) * h.kap[n][1] / h.kap[n][0];
    for (int m = 1; m <= n; ++m) {
      h.rd_a[n][m] = h.dl_a[n][m] * h.kap[n][m - 1] / h.kap[n][m];
      h.rd_b[n][m] = (m + 1 <= n) ? h.dl_b[n][m] * h.kap[n][m + 1] / h.kap[n][m] : 0.0;
    }
  }
}

}  // namespace ts
