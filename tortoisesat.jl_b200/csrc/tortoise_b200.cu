// libtortoise_b200.so -- single translation unit: context + C-ABI entry points.
// Kernels live in the k*_*.cuh files included below.  sm_100a only; there is no
// CPU fallback anywhere in this library.
#include "common.cuh"
#include "igrf_device.cuh"
#include "k1_igrf.cuh"
#include "k2_field.cuh"
#include "k3_alilqr.cuh"

#include <algorithm>
#include <numeric>

namespace ts {
#include "igrf12_tables.inc"
}

using namespace ts;

extern "C" {

int ts_version(void) { return 100; }

int ts_create(ts_ctx** out, int device_id) {
  if (!out) return TS_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return TS_ERR_CUDA;  // no GPU -> no library (no CPU fallback)
  if (device_id < 0 || device_id >= ndev) return TS_ERR_ARG;
  ts_ctx* c = new ts_ctx();
  c->device = device_id;
  if (cudaSetDevice(device_id) != cudaSuccess) {
    delete c;
    return TS_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device_id);
  c->sm_count = prop.multiProcessorCount;
  snprintf(c->name, sizeof(c->name), "%s", prop.name);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaMalloc(&c->d_tabG, sizeof(TS_IGRF12_G)) != cudaSuccess || cudaMalloc(&c->d_tabH, sizeof(TS_IGRF12_H)) != cudaSuccess ||
      cudaMalloc(&c->d_flag, 64) != cudaSuccess) {
    ts_destroy(c);
    return TS_ERR_CUDA;
  }
  cudaMemcpy(c->d_tabG, TS_IGRF12_G, sizeof(TS_IGRF12_G), cudaMemcpyHostToDevice);
  cudaMemcpy(c->d_tabH, TS_IGRF12_H, sizeof(TS_IGRF12_H), cudaMemcpyHostToDevice);
  IgrfConsts h;
  igrf_host_constants(h);
  if (cudaMemcpyToSymbol(c_igrf, &h, sizeof(h)) != cudaSuccess) {
    ts_destroy(c);
    return TS_ERR_CUDA;
  }
  cudaMemset(c->d_flag, 0, 64);
  *out = c;
  return TS_OK;
}

void ts_destroy(ts_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (int i = 0; i < 8; ++i)
    if (c->scratch[i]) cudaFree(c->scratch[i]);
  if (c->d_tabG) cudaFree(c->d_tabG);
  if (c->d_tabH) cudaFree(c->d_tabH);
  if (c->d_flag) cudaFree(c->d_flag);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* ts_last_error(const ts_ctx* c) { return c ? c->err : "null context"; }
int64_t ts_launch_count(const ts_ctx* c) { return c ? c->launches : 0; }
double ts_last_kernel_ms(const ts_ctx* c) { return c ? c->last_kernel_ms : 0.0; }

int ts_device_info(ts_ctx* c, int* sm_count, char* name, int name_len) {
  if (!c) return TS_ERR_ARG;
  if (sm_count) *sm_count = c->sm_count;
  if (name && name_len > 0) snprintf(name, (size_t)name_len, "%s", c->name);
  return TS_OK;
}

int ts_synchronize(ts_ctx* c) {
  if (!c) return TS_ERR_ARG;
  TS_CUDA(c, cudaSetDevice(c->device));
  TS_CUDA(c, cudaStreamSynchronize(c->stream));
  return TS_OK;
}

int ts_fp64_peak_probe(ts_ctx* c, double* tflops_out) {
  if (!c || !tflops_out) return TS_ERR_ARG;
  TS_CUDA(c, cudaSetDevice(c->device));
  const int iters = 4096, blocks = c->sm_count * 8, threads = 256;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    KernelTimer t(c);
    fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>((double*)c->d_flag + 1, iters, 1.0000001, 1e-9);
    t.stop();
    c->launches++;
    TS_CUDA(c, cudaGetLastError());
    TS_CUDA(c, cudaStreamSynchronize(c->stream));
    t.read();
    const double flops = 2.0 * 8 * 16 * (double)iters * (double)blocks * threads;
    const double tf = flops / (c->last_kernel_ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  *tflops_out = best;
  return TS_OK;
}

int ts_igrf12_batch(ts_ctx* c, double date, int64_t n, const double* r_m, const double* lat, const double* lon, double* Bn,
                    double* Be, double* Bd, int pointers_are_device) {
  if (!c) return TS_ERR_ARG;
  if (n < 0 || (n > 0 && (!r_m || !lat || !lon || !Bn || !Be || !Bd))) return fail(c, TS_ERR_ARG, "ts_igrf12_batch: null array or n<0");
  if (!(date >= 1900.0 && date <= 2025.0))
    return fail(c, TS_ERR_DATE, "This IGRF version will not work for years outside the interval [1900, 2025).");
  if (n == 0) return TS_OK;
  TS_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = (size_t)n * sizeof(double);
  DevBuf dr, dla, dlo, dbn, dbe, dbd;
  int rc;
  if ((rc = dev_in(c, dr, r_m, bytes, pointers_are_device))) return rc;
  if ((rc = dev_in(c, dla, lat, bytes, pointers_are_device))) return rc;
  if ((rc = dev_in(c, dlo, lon, bytes, pointers_are_device))) return rc;
  if ((rc = dev_out(c, dbn, Bn, bytes, pointers_are_device))) return rc;
  if ((rc = dev_out(c, dbe, Be, bytes, pointers_are_device))) return rc;
  if ((rc = dev_out(c, dbd, Bd, bytes, pointers_are_device))) return rc;
  TS_CUDA(c, cudaMemsetAsync(c->d_flag, 0, sizeof(int), c->stream));
  const unsigned blocks = (unsigned)((n + K1_THREADS - 1) / K1_THREADS);
  KernelTimer t(c);
  if (igrf_nmax_for_date(date) == 13)
    k1_igrf12_batch<13><<<blocks, K1_THREADS, 0, c->stream>>>(c->d_tabG, c->d_tabH, date, n, (const double*)dr.d,
                                                               (const double*)dla.d, (const double*)dlo.d, (double*)dbn.d,
                                                               (double*)dbe.d, (double*)dbd.d, c->d_flag);
  else
    k1_igrf12_batch<10><<<blocks, K1_THREADS, 0, c->stream>>>(c->d_tabG, c->d_tabH, date, n, (const double*)dr.d,
                                                               (const double*)dla.d, (const double*)dlo.d, (double*)dbn.d,
                                                               (double*)dbe.d, (double*)dbd.d, c->d_flag);
  t.stop();
  c->launches++;
  TS_CUDA(c, cudaGetLastError());
  if ((rc = dev_back(c, dbn, Bn, bytes))) return rc;
  if ((rc = dev_back(c, dbe, Be, bytes))) return rc;
  if ((rc = dev_back(c, dbd, Bd, bytes))) return rc;
  int bad = 0;
  TS_CUDA(c, cudaMemcpyAsync(&bad, c->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  TS_CUDA(c, cudaStreamSynchronize(c->stream));
  t.read();
  if (bad) return fail(c, TS_ERR_DOMAIN, "The latitude must be between -pi/2 and +pi/2 rad and the longitude between -pi and +pi rad.");
  return TS_OK;
}

// ---------------------------------------------------------------------------- K2
static_assert(sizeof(ts_field_opts) == sizeof(ts_field_opts_dev), "field opts layout");


int ts_magnetic_simulation_batch(ts_ctx* c, int64_t n_trials, const double* kep6, const ts_field_opts* opts,
                                 const int64_t* B_offs, const int64_t* rows_limit, double* B_eci, double* pos, double* vel,
                                 int pointers_are_device) {
  if (!c) return TS_ERR_ARG;
  if (n_trials < 0 || (n_trials > 0 && (!kep6 || !opts || !B_offs || !B_eci)))
    return fail(c, TS_ERR_ARG, "ts_magnetic_simulation_batch: null argument");
  if (n_trials == 0) return TS_OK;
  bool any13 = false, any10 = false;
  int64_t maxN = 0;
  for (int64_t t = 0; t < n_trials; ++t) {
    if (!(opts[t].igrf_date >= 1900.0 && opts[t].igrf_date <= 2025.0))
      return fail(c, TS_ERR_DATE, "trial %lld: IGRF date outside [1900, 2025]", (long long)t);
    if (opts[t].N < 1 || B_offs[t + 1] - B_offs[t] < 2 * opts[t].N)
      return fail(c, TS_ERR_ARG, "trial %lld: N < 1 or B_offs does not leave 2N rows", (long long)t);
    (igrf_nmax_for_date(opts[t].igrf_date) == 13 ? any13 : any10) = true;
    if (opts[t].N > maxN) maxN = opts[t].N;
  }
  TS_CUDA(c, cudaSetDevice(c->device));
  const int64_t total_rows = B_offs[n_trials];
  int rc;
  DevBuf dkep, dB, dpos, dvel;
  if ((rc = dev_in(c, dkep, kep6, (size_t)n_trials * 6 * sizeof(double), pointers_are_device))) return rc;
  if ((rc = dev_out(c, dB, B_eci, (size_t)total_rows * 3 * sizeof(double), pointers_are_device))) return rc;
  const size_t pos_bytes = (size_t)(total_rows + n_trials) * 3 * sizeof(double);
  double* d_pos = nullptr;
  if (pos) {
    if ((rc = dev_out(c, dpos, pos, pos_bytes, pointers_are_device))) return rc;
    d_pos = (double*)dpos.d;
  } else {
    void* p = nullptr;
    if ((rc = scratch_reserve(c, 0, pos_bytes, &p))) return rc;
    d_pos = (double*)p;
  }
  if (vel && (rc = dev_out(c, dvel, vel, pos_bytes, pointers_are_device))) return rc;
  ts_field_opts_dev* d_opts = nullptr;
  int64_t *d_offs = nullptr, *d_lim = nullptr;
  if ((rc = upload(c, 1, (const ts_field_opts_dev*)opts, (size_t)n_trials, &d_opts))) return rc;
  if ((rc = upload(c, 2, B_offs, (size_t)n_trials + 1, &d_offs))) return rc;
  if (rows_limit && (rc = upload(c, 3, rows_limit, (size_t)n_trials, &d_lim))) return rc;
  KernelTimer tm(c);
  k2a_orbit_euler<<<(unsigned)((n_trials + 127) / 128), 128, 0, c->stream>>>(n_trials, (const double*)dkep.d, d_opts, d_offs, d_lim,
                                                                             d_pos, vel ? (double*)dvel.d : nullptr);
  c->launches++;
  dim3 grid((unsigned)n_trials, (unsigned)((2 * maxN + K2B_THREADS - 1) / K2B_THREADS));
  if (any13) {
    k2b_field_rows<13><<<grid, K2B_THREADS, 0, c->stream>>>(c->d_tabG, c->d_tabH, d_opts, d_offs, d_lim, d_pos, (double*)dB.d, 13);
    c->launches++;
  }
  if (any10) {
    k2b_field_rows<10><<<grid, K2B_THREADS, 0, c->stream>>>(c->d_tabG, c->d_tabH, d_opts, d_offs, d_lim, d_pos, (double*)dB.d, 10);
    c->launches++;
  }
  tm.stop();
  TS_CUDA(c, cudaGetLastError());
  if ((rc = dev_back(c, dB, B_eci, (size_t)total_rows * 3 * sizeof(double)))) return rc;
  if (pos && (rc = dev_back(c, dpos, pos, pos_bytes))) return rc;
  if (vel && (rc = dev_back(c, dvel, vel, pos_bytes))) return rc;
  TS_CUDA(c, cudaStreamSynchronize(c->stream));
  tm.read();
  return TS_OK;
}

static int gramian_common(ts_ctx* c, int64_t n_trials, const double* B_eci, const int64_t* B_offs, const int64_t* rows,
                          const double* dt, const double* cutoff, double* G, int64_t* tf_index, int pad) {
  if (!c) return TS_ERR_ARG;
  if (n_trials < 0 || (n_trials > 0 && (!B_eci || !B_offs || !rows || !dt))) return fail(c, TS_ERR_ARG, "gramian: null argument");
  if (n_trials == 0) return TS_OK;
  TS_CUDA(c, cudaSetDevice(c->device));
  int64_t total = 0;
  for (int64_t t = 0; t < n_trials; ++t) total = (B_offs[t] + rows[t] > total) ? B_offs[t] + rows[t] : total;
  int rc;
  DevBuf dB, dG;
  if ((rc = dev_in(c, dB, B_eci, (size_t)total * 3 * sizeof(double), pad))) return rc;
  if (G && (rc = dev_out(c, dG, G, (size_t)total * 9 * sizeof(double), pad))) return rc;
  int64_t *d_offs, *d_rows, *d_idx = nullptr;
  double *d_dt, *d_cut = nullptr;
  if ((rc = upload(c, 1, B_offs, (size_t)n_trials, &d_offs))) return rc;
  if ((rc = upload(c, 2, rows, (size_t)n_trials, &d_rows))) return rc;
  if ((rc = upload(c, 3, dt, (size_t)n_trials, &d_dt))) return rc;
  if (cutoff && (rc = upload(c, 4, cutoff, (size_t)n_trials, &d_cut))) return rc;
  if (tf_index) {
    void* p;
    if ((rc = scratch_reserve(c, 5, (size_t)n_trials * sizeof(int64_t), &p))) return rc;
    d_idx = (int64_t*)p;
  }
  KernelTimer tm(c);
  k2c_gramian_cutoff<<<(unsigned)((n_trials + 127) / 128), 128, 0, c->stream>>>(n_trials, (const double*)dB.d, d_offs, d_rows, d_dt,
                                                                                d_cut, G ? (double*)dG.d : nullptr, d_idx);
  tm.stop();
  c->launches++;
  TS_CUDA(c, cudaGetLastError());
  if (G && (rc = dev_back(c, dG, G, (size_t)total * 9 * sizeof(double)))) return rc;
  if (tf_index) TS_CUDA(c, cudaMemcpyAsync(tf_index, d_idx, (size_t)n_trials * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  TS_CUDA(c, cudaStreamSynchronize(c->stream));
  tm.read();
  return TS_OK;
}

int ts_magnetic_gramian_batch(ts_ctx* c, int64_t n_trials, const double* B_eci, const int64_t* B_offs, const int64_t* rows,
                              const double* dt, double* G, int pad) {
  if (c && !G) return fail(c, TS_ERR_ARG, "ts_magnetic_gramian_batch: G is null");
  return gramian_common(c, n_trials, B_eci, B_offs, rows, dt, nullptr, G, nullptr, pad);
}

int ts_condition_cutoff_batch(ts_ctx* c, int64_t n_trials, const double* B_eci, const int64_t* B_offs, const int64_t* rows,
                              const double* dt, const double* cutoff, int64_t* tf_index, int pad) {
  if (c && (!cutoff || !tf_index)) return fail(c, TS_ERR_ARG, "ts_condition_cutoff_batch: null argument");
  return gramian_common(c, n_trials, B_eci, B_offs, rows, dt, cutoff, nullptr, tf_index, pad);
}

int ts_condition_based_time_batch(ts_ctx* c, int64_t n_trials, const double* G, const int64_t* offs, const int64_t* rows,
                                  const double* cutoff, int64_t* tf_index, int pad) {
  if (!c) return TS_ERR_ARG;
  if (n_trials < 0 || (n_trials > 0 && (!G || !offs || !rows || !cutoff || !tf_index)))
    return fail(c, TS_ERR_ARG, "ts_condition_based_time_batch: null argument");
  if (n_trials == 0) return TS_OK;
  TS_CUDA(c, cudaSetDevice(c->device));
  int64_t total = 0;
  for (int64_t t = 0; t < n_trials; ++t) total = (offs[t] + rows[t] > total) ? offs[t] + rows[t] : total;
  int rc;
  DevBuf dG;
  if ((rc = dev_in(c, dG, G, (size_t)total * 9 * sizeof(double), pad))) return rc;
  int64_t *d_offs, *d_rows, *d_idx;
  double* d_cut;
  if ((rc = upload(c, 1, offs, (size_t)n_trials, &d_offs))) return rc;
  if ((rc = upload(c, 2, rows, (size_t)n_trials, &d_rows))) return rc;
  if ((rc = upload(c, 4, cutoff, (size_t)n_trials, &d_cut))) return rc;
  void* p;
  if ((rc = scratch_reserve(c, 5, (size_t)n_trials * sizeof(int64_t), &p))) return rc;
  d_idx = (int64_t*)p;
  KernelTimer tm(c);
  k2d_condition_time<<<(unsigned)((n_trials + 127) / 128), 128, 0, c->stream>>>(n_trials, (const double*)dG.d, d_offs, d_rows, d_cut, d_idx);
  tm.stop();
  c->launches++;
  TS_CUDA(c, cudaGetLastError());
  TS_CUDA(c, cudaMemcpyAsync(tf_index, d_idx, (size_t)n_trials * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  TS_CUDA(c, cudaStreamSynchronize(c->stream));
  tm.read();
  return TS_OK;
}

// ---------------------------------------------------------------------------- K3
static_assert(sizeof(ts_ilqr_opts) == sizeof(ts_ilqr_opts_dev), "ilqr opts layout");
static_assert(sizeof(ts_trial_outcome) == sizeof(ts_trial_outcome_dev) && sizeof(ts_trial_outcome) == 64, "outcome layout");

void ts_ilqr_default_opts(ts_ilqr_opts* o) {
  if (!o) return;
  o->max_outer = 20; o->max_inner = 50; o->max_linesearch = 20; o->dJ_counter_limit = 10;
  o->stage_cost_dt = 0; o->goal_mask = 0x7F;
  o->cost_tol = 1e-4; o->cost_tol_intermediate = 1e-3; o->grad_tol = 1e-5; o->grad_tol_intermediate = 1e-5;
  o->constraint_tol = 1e-3; o->penalty_initial = 1.0; o->penalty_scaling = 10.0; o->penalty_max = 1e8; o->dual_max = 1e8;
  o->ls_lower = 1e-8; o->ls_upper = 10.0; o->bp_reg_increase = 1.6; o->bp_reg_max = 1e8; o->bp_reg_min = 1e-8;
  o->bp_reg_fp = 10.0; o->max_cost_value = 1e8; o->max_state_value = 1e8; o->max_control_value = 1e8;
  o->u_max = 1.0; o->u_min = -1.0;
}

int ts_alilqr_solve_batch(ts_ctx* c, int64_t n_trials, const int64_t* N_i, const int64_t* offs, const double* x0,
                          const double* xf, const double* Jmat, const double* Qd, const double* Qfd, const double* Rd,
                          const double* B_eci, const int64_t* B_offs, const int64_t* B_rows, const double* index_scale,
                          const double* clock_rate, double dt, const double* U0, const ts_ilqr_opts* opts, double* X,
                          double* U, double* K, ts_trial_outcome* out, int pad) {
  if (!c) return TS_ERR_ARG;
  if (n_trials < 0 || (n_trials > 0 && (!N_i || !offs || !x0 || !xf || !Jmat || !Qd || !Qfd || !Rd || !B_eci || !B_offs || !B_rows ||
                                        !index_scale || !clock_rate || !X || !U || !out)))
    return fail(c, TS_ERR_ARG, "ts_alilqr_solve_batch: null argument");
  if (n_trials == 0) return TS_OK;
  if (!(dt > 0)) return fail(c, TS_ERR_ARG, "ts_alilqr_solve_batch: dt must be > 0");
  ts_ilqr_opts o;
  if (opts) o = *opts; else ts_ilqr_default_opts(&o);
  if (o.max_linesearch < 0 || o.max_linesearch > 63) return fail(c, TS_ERR_ARG, "max_linesearch must be in [0,63]");
  int64_t Nmax = 0, total_knots = 0, total_rows = 0;
  for (int64_t t = 0; t < n_trials; ++t) {
    if (N_i[t] < 2) return fail(c, TS_ERR_ARG, "trial %lld: N < 2", (long long)t);
    Nmax = std::max(Nmax, N_i[t]);
    total_knots = std::max(total_knots, offs[t] + N_i[t]);
    total_rows = std::max(total_rows, B_offs[t] + B_rows[t]);
    if (B_rows[t] < 1) return fail(c, TS_ERR_ARG, "trial %lld: empty field table", (long long)t);
  }
  TS_CUDA(c, cudaSetDevice(c->device));
  // launch geometry: persistent warps, 4 trials per warp
  int occ = 0;
  TS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_alilqr_kernel, K3_WARPS_PER_BLOCK * 32, K3_SMEM_BYTES));
  if (occ < 1) return fail(c, TS_ERR_CUDA, "k3 kernel does not fit on an SM");
  const int64_t groups = (n_trials + 3) / 4;
  const int64_t max_warps = (int64_t)c->sm_count * occ * K3_WARPS_PER_BLOCK;
  const int64_t warps = std::min(groups, max_warps);
  const int blocks = (int)((warps + K3_WARPS_PER_BLOCK - 1) / K3_WARPS_PER_BLOCK);
  const int64_t slots = (int64_t)blocks * K3_WARPS_PER_BLOCK * 4;
  // trials sorted by horizon (descending) so the four teams of a warp have similar trip counts
  std::vector<int64_t> order((size_t)n_trials);
  std::iota(order.begin(), order.end(), (int64_t)0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return N_i[a] > N_i[b]; });

  int rc;
  DevBuf dB, dU0, dX, dU, dK;
  if ((rc = dev_in(c, dB, B_eci, (size_t)total_rows * 3 * sizeof(double), pad))) return rc;
  if (U0 && (rc = dev_in(c, dU0, U0, (size_t)total_knots * 3 * sizeof(double), pad))) return rc;
  if ((rc = dev_out(c, dX, X, (size_t)total_knots * 8 * sizeof(double), pad))) return rc;
  if ((rc = dev_out(c, dU, U, (size_t)total_knots * 3 * sizeof(double), pad))) return rc;
  if (K && (rc = dev_out(c, dK, K, (size_t)total_knots * 24 * sizeof(double), pad))) return rc;
  // per-trial small arrays -> one packed upload
  const size_t T = (size_t)n_trials;
  const size_t n_i64 = 5 * T, n_f64 = (8 + 8 + 9 + 8 + 8 + 3 + 1 + 1) * T;
  std::vector<int64_t> hi(n_i64);
  std::vector<double> hf(n_f64);
  memcpy(&hi[0 * T], order.data(), T * 8);
  memcpy(&hi[1 * T], N_i, T * 8);
  memcpy(&hi[2 * T], offs, T * 8);
  memcpy(&hi[3 * T], B_offs, T * 8);
  memcpy(&hi[4 * T], B_rows, T * 8);
  size_t fo = 0;
  auto put = [&](const double* src, size_t per) { memcpy(&hf[fo], src, per * T * 8); fo += per * T; return fo - per * T; };
  const size_t o_x0 = put(x0, 8), o_xf = put(xf, 8), o_J = put(Jmat, 9), o_Qd = put(Qd, 8), o_Qfd = put(Qfd, 8), o_Rd = put(Rd, 3),
               o_is = put(index_scale, 1), o_cr = put(clock_rate, 1);
  int64_t* d_i = nullptr;
  double* d_f = nullptr;
  if ((rc = upload(c, 1, hi.data(), n_i64, &d_i))) return rc;
  if ((rc = upload(c, 2, hf.data(), n_f64, &d_f))) return rc;
  void *p_out, *p_work, *p_q;
  if ((rc = scratch_reserve(c, 3, T * sizeof(ts_trial_outcome), &p_out))) return rc;
  const size_t per_slot = (size_t)Nmax * (90 + 24 + 6 + 1) * sizeof(double) + (size_t)Nmax * 3 * sizeof(int) + 64;
  if ((rc = scratch_reserve(c, 6, per_slot * (size_t)slots + 256, &p_work))) return rc;
  if ((rc = scratch_reserve(c, 4, 64, &p_q))) return rc;
  TS_CUDA(c, cudaMemsetAsync(p_q, 0, 64, c->stream));
  K3Args a;
  a.n_trials = n_trials;
  a.order = d_i + 0 * T; a.N_i = d_i + 1 * T; a.offs = d_i + 2 * T; a.B_offs = d_i + 3 * T; a.B_rows = d_i + 4 * T;
  a.x0 = d_f + o_x0; a.xf = d_f + o_xf; a.Jmat = d_f + o_J; a.Qd = d_f + o_Qd; a.Qfd = d_f + o_Qfd; a.Rd = d_f + o_Rd;
  a.index_scale = d_f + o_is; a.clock_rate = d_f + o_cr;
  a.B_eci = (const double*)dB.d;
  a.dt = dt;
  a.U0 = U0 ? (const double*)dU0.d : nullptr;
  memcpy(&a.opts, &o, sizeof(o));
  a.X = (double*)dX.d; a.U = (double*)dU.d; a.K = K ? (double*)dK.d : nullptr;
  a.out = (ts_trial_outcome_dev*)p_out;
  a.Nmax = Nmax;
  {
    char* w = (char*)p_work;
    a.w_xu = (double*)w;   w += (size_t)slots * Nmax * 90 * sizeof(double);
    a.w_kd = (double*)w;   w += (size_t)slots * Nmax * 24 * sizeof(double);
    a.w_lam = (double*)w;  w += (size_t)slots * Nmax * 6 * sizeof(double);
    a.w_clk = (double*)w;  w += (size_t)slots * Nmax * sizeof(double);
    a.w_rows = (int*)w;
  }
  a.queue = (unsigned long long*)p_q;
  KernelTimer tm(c);
  k3_alilqr_kernel<<<blocks, K3_WARPS_PER_BLOCK * 32, K3_SMEM_BYTES, c->stream>>>(a);
  tm.stop();
  c->launches++;
  TS_CUDA(c, cudaGetLastError());
  if ((rc = dev_back(c, dX, X, (size_t)total_knots * 8 * sizeof(double)))) return rc;
  if ((rc = dev_back(c, dU, U, (size_t)total_knots * 3 * sizeof(double)))) return rc;
  if (K && (rc = dev_back(c, dK, K, (size_t)total_knots * 24 * sizeof(double)))) return rc;
  TS_CUDA(c, cudaMemcpyAsync(out, p_out, T * sizeof(ts_trial_outcome), cudaMemcpyDeviceToHost, c->stream));
  TS_CUDA(c, cudaStreamSynchronize(c->stream));
  tm.read();
  return TS_OK;
}

}  // extern "C"
