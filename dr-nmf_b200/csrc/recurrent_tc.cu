#include "internal.h"
namespace drnmf {
int launch_recurrent_tc(drnmf_handle* h, FwdWorkspace& w, int B, int T, float* H_user, cudaStream_t st) {
  set_error("tcgen05 recurrent kernel not built yet");
  return DRNMF_ERR_INVALID;
}
}
