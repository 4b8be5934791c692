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

int64_t sofima_ctx_launch_count(const sofima_ctx* ctx) {
  return ctx ? ctx->launches : 0;
}

}  // extern "C"
