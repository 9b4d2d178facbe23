// Phase retrieval operator — placeholder until the FFT kernels land (see below in this round).
#include "operator.cuh"

struct PhaseTables {
  int pad = 0;
};

int phase_create(dps_operator* op, int pad) {
  (void)op; (void)pad;
  dps_set_error("phase retrieval: not built yet");
  return DPS_ERR_UNSUPPORTED;
}
void phase_destroy(dps_operator* op) {
  delete op->phase;
  op->phase = nullptr;
}
int phase_forward(const dps_operator*, const FwdArgs&, cudaStream_t) {
  dps_set_error("phase retrieval: not built yet");
  return DPS_ERR_UNSUPPORTED;
}
int phase_adjoint(const dps_operator*, const AdjArgs&, cudaStream_t) {
  dps_set_error("phase retrieval: not built yet");
  return DPS_ERR_UNSUPPORTED;
}
