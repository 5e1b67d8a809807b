// ABI version probe for the ctypes loader.
#include <cuda_runtime.h>

#include "../../include/sonar_b200.h"

extern "C" int sonar_abi_version(void) { return SONAR_B200_ABI_VERSION; }

extern "C" int sonar_set_device(int device) { return (int)cudaSetDevice(device); }
