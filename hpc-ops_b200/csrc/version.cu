// Library identity (reference: src/C/version.cc, src/C/built_json.cu).
#include "host_utils.h"

#ifndef HPC_B200_VERSION
#define HPC_B200_VERSION "0.1.0+b200"
#endif
#ifndef HPC_B200_BUILT_JSON
#define HPC_B200_BUILT_JSON "{}"
#endif

extern "C" const char* hpc_version() { return HPC_B200_VERSION; }
extern "C" const char* hpc_built_json() { return HPC_B200_BUILT_JSON; }
