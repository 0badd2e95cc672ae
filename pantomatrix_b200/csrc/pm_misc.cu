// ABI version / device probe.
#include "pm_common.cuh"
#include "../../include/pm_emage.h"

extern "C" int pm_abi_version(void) { return PM_ABI_VERSION; }

extern "C" int pm_memset_async(void* ptr, int value, long long bytes, void* stream) {
  PM_REQUIRE(ptr && bytes >= 0);
  if (bytes == 0) return PM_OK;
  const cudaError_t e = cudaMemsetAsync(ptr, value, (size_t)bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? PM_OK : (int)e;
}

extern "C" int pm_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
