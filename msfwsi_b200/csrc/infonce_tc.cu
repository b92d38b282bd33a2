// placeholder until the tcgen05 main loop lands (next commit)
#include "infonce_plan.cuh"
namespace msf {
int launch_infonce_tc(const void*, const void*, int64_t, int64_t, int, float, const NcePlan&, float*, float*, cudaStream_t) {
  set_error("tcgen05 InfoNCE kernel not built");
  return MSF_ERR_UNSUPPORTED;
}
}  // namespace msf
