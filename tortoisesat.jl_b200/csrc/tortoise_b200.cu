// libtortoise_b200.so -- single translation unit: context + C-ABI entry points.
// Kernels live in the k*_*.cuh files included below.  sm_100a only; there is no
// CPU fallback anywhere in this library.
#include "common.cuh"
#include "igrf_device.cuh"
#include "k1_igrf.cuh"

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
  int64_t want = (n + K1_THREADS - 1) / K1_THREADS;
  const int64_t cap = (int64_t)c->sm_count * 16;  // persistent-style grid: a multiple of the SM count
  const int blocks = (int)(want < cap ? want : cap);
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

}  // extern "C"
