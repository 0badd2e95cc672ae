// tcgen05 tap-GEMM (placeholder until the kernel lands): exported so the ABI is complete.
#include "pm_common.cuh"
#include "../../include/pm_emage.h"

extern "C" int pm_tapgemm_tc(const uint16_t*, long long, long long, int, int, int, int, const uint16_t*, long long,
                             int, int, int, const float*, int, int, const float*, long long, int, int, float,
                             float*, long long, int, uint16_t*, long long, long long, int, int, void*) {
  return PM_EUNSUPPORTED;
}

extern "C" int pm_split_bf16(const float*, long long, int, int, int, int, uint16_t*, long long, long long, int, int,
                             void*) {
  return PM_EUNSUPPORTED;
}
