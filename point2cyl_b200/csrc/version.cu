#include "common.cuh"
extern "C" int p2c_version(void) { return 1; }
extern "C" const char* p2c_arch(void) { return "sm_100a"; }
