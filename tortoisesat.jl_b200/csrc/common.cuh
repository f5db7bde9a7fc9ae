// Context, error plumbing and small helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tortoise_b200.h"

struct ts_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_k3[3] = {nullptr, nullptr, nullptr};  // K3: start | persistent kernel done | straggler kernel done
  unsigned* d_k3_parked = nullptr;                     // device word holding the number of parked trials of the last K3 run
  bool k3_timed = false;
  int k3_park_cap = 0;                                 // parking places of the last K3 run
  int64_t k3_diag_n = 0;                               // trials covered by the K3 cycle diagnostics (scratch slot 19)
  cudaStream_t pipe[2] = {nullptr, nullptr};           // K1 host-pointer path: double-buffered copy/compute pipeline
  cudaEvent_t pipe_ev = nullptr;
  char err[512] = {0};
  char name[128] = {0};
  int64_t launches = 0;
  double last_kernel_ms = 0.0;
  double* d_tabG = nullptr;  // 104 x 25
  double* d_tabH = nullptr;  // 91 x 25
  double* d_tabGH = nullptr; // 3450 (igrf12syn)
  int* d_flag = nullptr;     // generic device error/flag word
  void* d_gh_stage = nullptr; // K1: the call's interpolated coefficient table (2 x 104 double2)
  // trajectories of the last ts_monte_carlo_run with keep_trajectories (device pointers into the scratch arenas)
  struct McLast {
    bool valid = false;
    int64_t n = 0, knots = 0, rows = 0;
    std::vector<int64_t> knot_offs, row_offs;
    double *d_X = nullptr, *d_U = nullptr, *d_Xs = nullptr, *d_Us = nullptr, *d_B = nullptr;
  } mc_last;
  // grow-only scratch arenas
  void* scratch[32] = {};
  size_t scratch_bytes[32] = {};
};

namespace ts {

inline int fail(ts_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define TS_CUDA(ctx, call)                                                                              \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return ts::fail((ctx), TS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// grow-only device scratch slot
inline int scratch_reserve(ts_ctx* ctx, int slot, size_t bytes, void** out) {
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], bytes);
    if (e != cudaSuccess) return fail(ctx, TS_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    ctx->scratch_bytes[slot] = bytes;
  }
  *out = ctx->scratch[slot];
  return TS_OK;
}

// RAII-less helper for "host or device pointer" arguments: stages host arrays
// into device memory owned by the call and copies results back.
struct DevBuf {
  void* d = nullptr;
  bool owned = false;
  ~DevBuf() {
    if (owned && d) cudaFree(d);
  }
};
inline int dev_in(ts_ctx* ctx, DevBuf& b, const void* p, size_t bytes, int is_device) {
  if (is_device || bytes == 0) {
    b.d = const_cast<void*>(p);
    return TS_OK;
  }
  cudaError_t e = cudaMalloc(&b.d, bytes);
  if (e != cudaSuccess) return fail(ctx, TS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  b.owned = true;
  e = cudaMemcpyAsync(b.d, p, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return fail(ctx, TS_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  return TS_OK;
}
inline int dev_out(ts_ctx* ctx, DevBuf& b, void* p, size_t bytes, int is_device) {
  if (is_device || bytes == 0 || p == nullptr) {
    b.d = p;
    return TS_OK;
  }
  cudaError_t e = cudaMalloc(&b.d, bytes);
  if (e != cudaSuccess) return fail(ctx, TS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  b.owned = true;
  return TS_OK;
}
inline int dev_back(ts_ctx* ctx, DevBuf& b, void* p, size_t bytes) {
  if (!b.owned || bytes == 0) return TS_OK;
  cudaError_t e = cudaMemcpyAsync(p, b.d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return fail(ctx, TS_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  return TS_OK;
}

// small host array -> device scratch slot
template <class T>
inline int upload(ts_ctx* c, int slot, const T* host, size_t count, T** dev) {
  void* d = nullptr;
  int rc = scratch_reserve(c, slot, count * sizeof(T) + 16, &d);
  if (rc) return rc;
  TS_CUDA(c, cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  *dev = (T*)d;
  return TS_OK;
}

struct KernelTimer {
  ts_ctx* c;
  explicit KernelTimer(ts_ctx* ctx) : c(ctx) { cudaEventRecord(c->ev0, c->stream); }
  void stop() { cudaEventRecord(c->ev1, c->stream); }
  // call after the stream has been synchronised
  void read() {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->last_kernel_ms = ms;
  }
};

}  // namespace ts
