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

#ifdef __cplusplus
}
#endif
#endif /* TORTOISE_B200_H */
