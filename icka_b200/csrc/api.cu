// Handle lifecycle, error channel, version.  Part of libicka_b200.so (see include/icka_b200.h).
#include "common.cuh"

#include <string.h>

static thread_local char g_last_error[512] = "";

void icka_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

extern "C" int icka_version(void) { return 100; }

extern "C" const char* icka_last_error(void) { return g_last_error; }

extern "C" int icka_create(int device, icka_handle** out) {
  ICKA_REQUIRE(out != nullptr, "icka_create: out is null");
  *out = nullptr;
  int count = 0;
  ICKA_CUDA(cudaGetDeviceCount(&count));
  ICKA_REQUIRE(device >= 0 && device < count, "icka_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  ICKA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    ICKA_FAIL(ICKA_ERR_UNSUPPORTED,
              "icka_create: device %d is sm_%d%d; libicka_b200 is built for sm_100a only and has no fallback",
              device, prop.major, prop.minor);
  icka_handle* h = new icka_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->cc_minor = prop.minor;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->launches.store(0);
  h->encode_tiled = nullptr;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete h;
    ICKA_FAIL(ICKA_ERR_CUDA, "icka_create: cuTensorMapEncodeTiled not available from the driver");
  }
  h->encode_tiled = fn;
  h->workspace = nullptr;
  h->seed_base = nullptr;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaError_t ea = cudaSetDevice(device);
  if (ea == cudaSuccess) ea = cudaMalloc(&h->workspace, ICKA_WORKSPACE_BYTES);
  cudaSetDevice(prev);
  if (ea != cudaSuccess) {
    delete h;
    ICKA_FAIL(ICKA_ERR_CUDA, "icka_create: cannot allocate the handle's device scratch: %s", cudaGetErrorString(ea));
  }
  *out = h;
  return ICKA_OK;
}

extern "C" int icka_destroy(icka_handle* h) {
  if (h && h->workspace) cudaFree(h->workspace);
  delete h;
  return ICKA_OK;
}

extern "C" int icka_set_seed_base(icka_handle* h, const uint64_t* seed_base_dev) {
  ICKA_REQUIRE(h != nullptr, "null handle");
  ICKA_REQUIRE(icka_aligned(seed_base_dev, 8), "seed base must be 8-byte aligned");
  h->seed_base = reinterpret_cast<const unsigned long long*>(seed_base_dev);
  return ICKA_OK;
}

extern "C" int64_t icka_launch_count(const icka_handle* h) { return h ? (int64_t)h->launches.load() : 0; }
