#include "common.cuh"
extern "C" int p2c_version(void) { return 2; }
extern "C" const char* p2c_arch(void) { return "sm_100a"; }

int g_p2c_sm_budget = 0;   // 0 = every SM of the device
extern "C" int p2c_set_sm_budget(int sms) {
  const int prev = g_p2c_sm_budget;
  g_p2c_sm_budget = sms > 0 ? sms : 0;
  return prev;
}

// programmatic dependent launch for the kernels that support it (common.cuh: p2c_launch); env P2C_PDL=0 disables
int g_p2c_pdl = 1;
extern "C" int p2c_set_pdl(int on) {
  const int prev = g_p2c_pdl;
  g_p2c_pdl = on ? 1 : 0;
  return prev;
}
