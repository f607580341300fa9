// Tensor-core (tcgen05) paths of the per-point MLP layer.  Placeholder until the UMMA kernel lands:
// reports "unsupported" so p2c_linear routes every shape to the fp32 SIMT kernel.
#include "common.cuh"

int p2c_linear_tc(const float*, int64_t, const float*, const float*, const float*, const float*,
                  const float*, int64_t, float*, int64_t, int, int, int, double*, int, float*, float*,
                  int, cudaStream_t) {
  return P2C_EUNSUPPORTED;
}
