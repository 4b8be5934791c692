// Shared plumbing of the sofima_b200 C-ABI library: context, error reporting,
// scratch memory, launch accounting.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sofima_b200.h"

struct sofima_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  int64_t launches = 0;
  char err[512] = {0};
  // Named scratch buffers that only ever grow; freed with the context.
  struct Buf {
    void* ptr = nullptr;
    size_t bytes = 0;
  };
  std::map<std::string, Buf> scratch;
  void* pinned = nullptr;  // small pinned host block for scalar read-back
  size_t pinned_bytes = 0;
  // Optional per-kernel timing with CUDA events on the launching stream (bench).
  bool timing = false;
  struct Rec {
    const char* name;
    cudaEvent_t e0, e1;
  };
  std::vector<Rec> recs;
  // Row-spectra cache of the flow path (flow.cu: sofima_xcorr_rowcache): forward row
  // transforms of every image row per distinct patch x-start, shared by all patches.
  struct RowCache {
    bool valid = false;
    const void* img[2] = {nullptr, nullptr};
    int dtype = 0, h[2] = {0, 0}, w[2] = {0, 0}, pw[2] = {0, 0}, L = 0;
    int pitch = 0;                           // float2 per cached row (L / 2 + 1 rounded up to 8)
    int nslots[2] = {0, 0};                  // distinct x starts per image
    float2* spec[2] = {nullptr, nullptr};    // [nxs][h][pitch]
    const int* xindex[2] = {nullptr, nullptr};  // [w]: x start -> slot, -1 = not cached
    const float2* fix = nullptr;             // [3][L / 2 + 1]: rect(pre), rect(post), W(post)
  } rowcache;
  // (L, pw, K, delta mask) of the twiddle digit tables "flow.tc_tab*" (flow_rowspec_tc.cuh)
  int tc_tab_key[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  int rowfix_key[3] = {0, 0, 0};  // (L, pw pre, pw post) of the table in scratch "flow.rc_fix"
};

namespace sofima {

extern thread_local char g_last_error[512];

inline int fail(sofima_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) strncpy(ctx->err, buf, sizeof(ctx->err) - 1);
  strncpy(g_last_error, buf, sizeof(g_last_error) - 1);
  return code;
}

#define SOFIMA_CUDA(ctx, expr)                                                   \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess)                                                       \
      return sofima::fail((ctx), SOFIMA_ECUDA, "%s:%d: %s -> %s", __FILE__,     \
                          __LINE__, #expr, cudaGetErrorString(_e));              \
  } while (0)

#define SOFIMA_CHECK_LAUNCH(ctx)                                                 \
  do {                                                                           \
    (ctx)->launches++;                                                           \
    SOFIMA_CUDA((ctx), cudaGetLastError());                                      \
  } while (0)

// Returns a device scratch buffer of at least `bytes` under `name` (grow-only).
inline int scratch(sofima_ctx* ctx, const char* name, size_t bytes, void** out) {
  auto& b = ctx->scratch[name];
  if (b.bytes < bytes) {
    if (b.ptr) {
      SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      SOFIMA_CUDA(ctx, cudaFree(b.ptr));
      b.ptr = nullptr;
      b.bytes = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
      b.ptr = nullptr;
      return fail(ctx, SOFIMA_ENOMEM, "cudaMalloc(%zu) for scratch '%s': %s", want,
                  name, cudaGetErrorString(e));
    }
    b.bytes = want;
  }
  *out = b.ptr;
  return SOFIMA_OK;
}

// RAII: brackets one kernel launch with events when ctx->timing is on.
struct LaunchTimer {
  sofima_ctx* ctx;
  sofima_ctx::Rec rec;
  bool on;
  LaunchTimer(sofima_ctx* c, const char* name) : ctx(c), on(c->timing) {
    if (!on) return;
    rec.name = name;
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    cudaEventRecord(rec.e0, ctx->stream);
  }
  ~LaunchTimer() {
    if (!on) return;
    cudaEventRecord(rec.e1, ctx->stream);
    ctx->recs.push_back(rec);
  }
};

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

}  // namespace sofima
