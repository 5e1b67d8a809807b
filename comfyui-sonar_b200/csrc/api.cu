// ABI version probe for the ctypes loader.
#include <cuda_runtime.h>

#include "../../include/sonar_b200.h"
#include "common.cuh"

extern "C" int sonar_abi_version(void) { return SONAR_B200_ABI_VERSION; }

extern "C" int sonar_set_device(int device) { return (int)cudaSetDevice(device); }

namespace sonar {
int& grid_limit_ctas_per_sm() {
  static thread_local int limit = 0;
  return limit;
}
}  // namespace sonar

// Co-scheduling hint for the calling thread's next launches (0 = off): see common.cuh streaming_grid_shared.
extern "C" int sonar_set_grid_limit(int ctas_per_sm) {
  sonar::grid_limit_ctas_per_sm() = ctas_per_sm > 0 ? ctas_per_sm : 0;
  return 0;
}
