// Context management for the sofima_b200 C-ABI library.
#include "common.cuh"

namespace sofima {
thread_local char g_last_error[512] = {0};
}

extern "C" {

int sofima_abi_version(void) { return SOFIMA_B200_ABI_VERSION; }

int sofima_ctx_create(int device, void* stream, sofima_ctx** out) {
  if (!out) return sofima::fail(nullptr, SOFIMA_EINVAL, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return sofima::fail(nullptr, SOFIMA_ECUDA,
                        "no CUDA device available (%s); sofima_b200 has no CPU "
                        "fallback",
                        cudaGetErrorString(e));
  if (device < 0 || device >= count)
    return sofima::fail(nullptr, SOFIMA_EINVAL, "device %d out of range [0,%d)",
                        device, count);
  sofima::DeviceGuard guard(device);
  cudaDeviceProp prop;
  SOFIMA_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return sofima::fail(nullptr, SOFIMA_EUNSUPPORTED,
                        "device %d is sm_%d%d; this library is built for sm_100a "
                        "only",
                        device, prop.major, prop.minor);
  sofima_ctx* ctx = new sofima_ctx();
  ctx->device = device;
  ctx->stream = static_cast<cudaStream_t>(stream);
  ctx->num_sms = prop.multiProcessorCount;
  ctx->pinned_bytes = 4096;
  e = cudaHostAlloc(&ctx->pinned, ctx->pinned_bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    delete ctx;
    return sofima::fail(nullptr, SOFIMA_ECUDA, "cudaHostAlloc: %s",
                        cudaGetErrorString(e));
  }
  *out = ctx;
  return SOFIMA_OK;
}

int sofima_ctx_destroy(sofima_ctx* ctx) {
  if (!ctx) return SOFIMA_OK;
  sofima::DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->scratch)
    if (kv.second.ptr) cudaFree(kv.second.ptr);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  delete ctx;
  return SOFIMA_OK;
}

int sofima_ctx_set_stream(sofima_ctx* ctx, void* stream) {
  if (!ctx) return sofima::fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  ctx->stream = static_cast<cudaStream_t>(stream);
  return SOFIMA_OK;
}

const char* sofima_last_error(const sofima_ctx* ctx) {
  return ctx ? ctx->err : sofima::g_last_error;
}

int sofima_ctx_set_timing(sofima_ctx* ctx, int on) {
  if (!ctx) return sofima::fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  ctx->timing = on != 0;
  return SOFIMA_OK;
}

int sofima_ctx_timing_report(sofima_ctx* ctx, char* buf, int64_t buf_len) {
  if (!ctx || !buf || buf_len < 2) return sofima::fail(ctx, SOFIMA_EINVAL, "bad arguments");
  sofima::DeviceGuard guard(ctx->device);
  SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::map<std::string, std::pair<double, long long>> tot;
  for (auto& r : ctx->recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto& t = tot[r.name];
    t.first += ms;
    t.second += 1;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : tot) {
    char item[256];
    snprintf(item, sizeof(item), "%s\"%s\": {\"ms\": %.6f, \"n\": %lld}", first ? "" : ", ",
             kv.first.c_str(), kv.second.first, kv.second.second);
    out += item;
    first = false;
  }
  out += "}";
  if ((int64_t)out.size() + 1 > buf_len)
    return sofima::fail(ctx, SOFIMA_EINVAL, "timing report buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return SOFIMA_OK;
}

int sofima_ctx_trim(sofima_ctx* ctx, int64_t keep_bytes) {
  if (!ctx) return sofima::fail(nullptr, SOFIMA_EINVAL, "ctx is NULL");
  sofima::DeviceGuard guard(ctx->device);
  SOFIMA_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->rowcache.valid = false;  // its spectra live in scratch buffers
  ctx->rowfix_key[0] = 0;
  ctx->tc_tab_key[0][0] = ctx->tc_tab_key[1][0] = 0;
  for (auto it = ctx->scratch.begin(); it != ctx->scratch.end();) {
    if (it->second.ptr && (int64_t)it->second.bytes > keep_bytes) {
      SOFIMA_CUDA(ctx, cudaFree(it->second.ptr));
      it = ctx->scratch.erase(it);
    } else {
      ++it;
    }
  }
  return SOFIMA_OK;
}

int64_t sofima_ctx_launch_count(const sofima_ctx* ctx) {
  return ctx ? ctx->launches : 0;
}

}  // extern "C"
