/* tortoise_b200.h -- C ABI of libtortoise_b200.so
 *
 * B200-native (sm_100a, FP64 CUDA cores) batched Monte-Carlo engine for the
 * data-parallel hot path of RoboticExplorationLab/TortoiseSat.jl: thousands of
 * independent magnetorquer slew trials.  The reference is a set of Julia
 * scripts with no FFI; every entry point below is what a Julia `ccall` (or the
 * Python ctypes host in tortoisesat.jl_b200/host.py) binds in place of the
 * reference function cited next to it.  See INTEGRATION.md for the bindings.
 *
 * Conventions
 *  - extern "C", plain pointers + sizes, all reals are FP64, all sizes int64_t.
 *  - return 0 = TS_OK, negative = error (message: ts_last_error()).  A failure of
 *    ONE trial is never a call failure: it is reported in
 *    ts_trial_outcome.status.
 *  - every data pointer is a HOST pointer unless the call has a
 *    `pointers_are_device` argument set to 1 (then all array arguments of that
 *    call are device pointers on the context's GPU; option structs stay host).
 *  - the caller owns its buffers for the duration of the (blocking) call; the
 *    library owns all device memory behind ts_ctx.  A ts_ctx is bound to one GPU
 *    and is not re-entrant.  There is NO CPU fallback: without a usable CUDA
 *    device ts_create() fails.
 *  - matrices are row-major "by sample / by knot": e.g. a field table is
 *    rows x 3, a state trajectory is N x 8 (Julia: declare them 3 x rows / 8 x N
 *    column-major and pass the array as is).
 */
#ifndef TORTOISE_B200_H
#define TORTOISE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TS_OK 0
#define TS_ERR_CUDA (-1)      /* CUDA runtime / launch failure */
#define TS_ERR_ARG (-2)       /* bad argument (null pointer, negative size, ...) */
#define TS_ERR_DATE (-3)      /* igrf12: date outside [1900, 2025]        (igrf.jl:80-81) */
#define TS_ERR_DOMAIN (-4)    /* igrf12: |lat| > pi/2 or |lon| > pi at >=1 point (igrf.jl:84-88);
                                 the offending outputs are NaN, the rest are valid */
#define TS_ERR_NOMEM (-5)

typedef struct ts_ctx ts_ctx;

/* ---- context -------------------------------------------------------------- */
int ts_create(ts_ctx** out, int device_id);
void ts_destroy(ts_ctx* ctx);
const char* ts_last_error(const ts_ctx* ctx); /* valid until the next call on ctx */
int ts_version(void);
/* multiprocessor count and name of the bound device */
int ts_device_info(ts_ctx* ctx, int* sm_count, char* name, int name_len);
/* number of kernels this context has launched since creation (bench bookkeeping) */
int64_t ts_launch_count(const ts_ctx* ctx);
/* wait for all work queued by this context */
int ts_synchronize(ts_ctx* ctx);
/* device time (ms, CUDA events on the context's stream) of the kernels launched by the
 * most recent call on ctx, excluding host<->device copies */
double ts_last_kernel_ms(const ts_ctx* ctx);

/* Measures the FP64 FMA peak of the bound GPU with a register-resident DFMA
 * micro-benchmark (the roofline denominator; MEASURED_PEAKS.json has no FP64 row). */
int ts_fp64_peak_probe(ts_ctx* ctx, double* tflops_out);

/* ---- K1: batched IGRF-12 --------------------------------------------------- *
 * Replaces igrf12(date, r, lat, lon) [src/igrf.jl:67-274] (+ legendre.jl:254-292,
 * dlegendre.jl:221-309) evaluated at n points -- e.g. the 10^6-point map of
 * igrf_data() [src/magnetic_toolbox.jl:108-121].  Geocentric: r in metres,
 * lat in [-pi/2, pi/2], lon in [-pi, pi] (rad).  Output north/east/down in nT.   */
int ts_igrf12_batch(ts_ctx* ctx, double date, int64_t n, const double* r_m, const double* lat, const double* lon,
                    double* Bn, double* Be, double* Bd, int pointers_are_device);

/* ---- K2: orbit + ECI field table + gramian cutoff --------------------------- *
 * One entry of ts_field_opts per trial (the reference keeps these in the `params`
 * struct / globals: p.GM, p.MJD, the hard-coded igrf date 2019 and the constant
 * field radius (alt+R_E)*1000 of src/magnetic_toolbox.jl:44,81).                   */
typedef struct ts_field_opts {
  double GM;             /* km^3/s^2                        (input_parameters.jl:26) */
  double mjd;            /* modified Julian day of t = 0    (input_parameters.jl:63) */
  double igrf_date;      /* year A.D. passed to igrf12      (magnetic_toolbox.jl:81: 2019) */
  double field_radius_m; /* radius passed to igrf12 [m]     (magnetic_toolbox.jl:81: (alt+R_E)*1000) */
  double t0, tf;         /* s; the orbit is propagated over [t0, 2 tf] with dt = (tf-t0)/N */
  int64_t N;             /* knot count; the table has 2N rows, pos/vel 2N+1 rows */
} ts_field_opts;

/* magnetic_simulation(p,t0,tf,N,mag_field) [src/magnetic_toolbox.jl:33-106] (which calls
 * kep_ECI [src/kep_ECI.jl:1-49] and Euler-integrates OrbitPlotter [src/OrbitPlotter.jl:1-52])
 * for n_trials trials.  kep6: n_trials x 6 = [e, a km, i deg, RAAN deg, argp deg, nu deg]
 * (NOT mutated, unlike kep_ECI.jl:7-8).  Ragged layout: trial t owns rows
 * [B_offs[t], B_offs[t]+2N_t) of B_eci (rows x 3, Tesla; last row 0 like the reference) and
 * rows [B_offs[t]+t, B_offs[t]+t+2N_t+1) of pos / vel (km, km/s; nullable).  B_offs has
 * n_trials+1 entries (host pointer).  rows_limit (host, nullable): if rows_limit[t] > 0 only
 * the first rows_limit[t] table rows are computed (the rest are 0) -- the solver only ever
 * reads rows <= floor(x8*N+1) (DerivFunction.jl:28, quirk Q1).                            */
int ts_magnetic_simulation_batch(ts_ctx* ctx, int64_t n_trials, const double* kep6, const ts_field_opts* opts,
                                 const int64_t* B_offs, const int64_t* rows_limit, double* B_eci, double* pos, double* vel,
                                 int pointers_are_device);

/* magnetic_gramian(B_N, dt) [src/magnetic_toolbox.jl:1-12]: G (rows x 9 per trial, row-major 3x3,
 * same row offsets as B_eci).  rows/dt: host arrays of n_trials.                           */
int ts_magnetic_gramian_batch(ts_ctx* ctx, int64_t n_trials, const double* B_eci, const int64_t* B_offs, const int64_t* rows,
                              const double* dt, double* G, int pointers_are_device);

/* condition_based_time(B_gram, cutoff) [src/magnetic_toolbox.jl:14-31]: first 1-based sample
 * whose gramian has cond_2 < cutoff, else 0.  tf_index: host array of n_trials.             */
int ts_condition_based_time_batch(ts_ctx* ctx, int64_t n_trials, const double* G, const int64_t* offs, const int64_t* rows,
                                  const double* cutoff, int64_t* tf_index, int pointers_are_device);

/* Fused magnetic_gramian + condition_based_time (no rows x 9 intermediate).               */
int ts_condition_cutoff_batch(ts_ctx* ctx, int64_t n_trials, const double* B_eci, const int64_t* B_offs, const int64_t* rows,
                              const double* dt, const double* cutoff, int64_t* tf_index, int pointers_are_device);

#ifdef __cplusplus
}
#endif
#endif /* TORTOISE_B200_H */
