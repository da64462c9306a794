// Library-level entry points of the C ABI: version, error reporting, device initialisation.
#include "common.cuh"
#include <stdlib.h>

namespace masr {

static thread_local std::string g_last_error;
static int g_sm_count = 148;

void set_error(const std::string& msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", int(e), cudaGetErrorString(e), file, line, what);
  g_last_error = buf;
  return MASR_E_CUDA;
}

int sm_count() { return g_sm_count; }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MASR_PDL"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

const uint64_t* g_seed_dev_ptr = nullptr;

__global__ void seed_bump_kernel(uint64_t* p, uint64_t inc) { *p += inc; }

}  // namespace masr

using namespace masr;

extern "C" int masr_abi_version(void) { return MASR_ABI_VERSION; }

extern "C" const char* masr_last_error(void) { return g_last_error.c_str(); }

extern "C" int masr_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("metaasr_b200 needs a CUDA device (sm_100); none is visible and there is no CPU fallback");
    return MASR_E_NOGPU;
  }
  MASR_REQUIRE(device >= 0 && device < n, "device index out of range");
  cudaDeviceProp prop;
  MASR_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("metaasr_b200 is built for sm_100a only; device is ") + prop.name);
    return MASR_E_NOGPU;
  }
  g_sm_count = prop.multiProcessorCount;
  return MASR_OK;
}

extern "C" int masr_set_seed_ptr(const uint64_t* dev_ptr) {
  g_seed_dev_ptr = dev_ptr;
  return MASR_OK;
}

extern "C" int masr_seed_bump(uint64_t* dev_ptr, uint64_t inc, void* stream) {
  MASR_REQUIRE(dev_ptr != nullptr, "seed_bump: null pointer");
  seed_bump_kernel<<<1, 1, 0, as_stream(stream)>>>(dev_ptr, inc);
  MASR_LAUNCH_CHECK();
  return MASR_OK;
}
