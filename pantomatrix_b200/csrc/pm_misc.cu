// ABI version / device probe.
#include "pm_common.cuh"
#include "../../include/pm_emage.h"

extern "C" int pm_abi_version(void) { return PM_ABI_VERSION; }

extern "C" int pm_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
