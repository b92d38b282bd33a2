// Error plumbing and device check behind include/msfwsi_b200.h.
#include <stdarg.h>

#include "common.cuh"

namespace msf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace msf

extern "C" int msf_abi_version(void) { return MSF_ABI_VERSION; }
extern "C" const char* msf_last_error(void) { return msf::g_err; }

extern "C" int msf_device_check(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MSF_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MSF_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  MSF_REQUIRE(prop.major == 10, MSF_ERR_UNSUPPORTED,
              "msfwsi_b200 is built for sm_100a only; current device is sm_%d%d", prop.major, prop.minor);
  return MSF_OK;
}

// ---- opt-in launch profiler (bench.py): CUDA events recorded on the launching stream right around the main kernel(s)
// of each C-ABI call.  Off by default; the only mutable global state of the library, guarded by a mutex.
#include <mutex>
#include <vector>

namespace msf {
namespace {
struct ProfSlot {
  cudaEvent_t a = nullptr, b = nullptr;
  int kernel = -1;
  double work = 0.0;
};
std::mutex g_prof_mu;
std::vector<ProfSlot> g_prof_slots;
int g_prof_used = 0, g_prof_dropped = 0;
bool g_prof_on = false;
const char* const kKernelNames[MSF_K_COUNT] = {
    "gather_concat_fwd", "gather_concat_bwd", "cosine_loss_fwd", "cosine_loss_bwd", "rownorm", "infonce_flash_fwd",
    "infonce_twopass_fwd", "infonce_simt_fwd", "infonce_bwd", "gemm_bf16", "crop_resample_fwd", "crop_resample_bwd", "ema_multi",
    "bn2d_stats", "bn2d_apply", "bn2d_apply_res", "bn2d_bwd_reduce", "bn2d_bwd_elemt", "bn2d_apply_pool", "bn2d_pool_bwd_elemt",
    "adam_multi", "grad_check_multi", "stem_s2d", "peer_allreduce_f64", "jigsaw_tiles", "gemm_grouped", "head_bn_finalize",
    "head_bn_elementwise", "gemm_grouped_f32", "infonce_key_grad"};
const char kKernelBound[MSF_K_COUNT] = {'h', 'h', 'h', 'h', 'h', 't', 't', 'f', 'h', 't', 'h', 'h', 'h',
                                        'h', 'h', 'h', 'h', 'h', 'h', 'h', 'h', 'h', 'h', 'l', 'h', 't', 'l', 'h', 'f', 't'};
}  // namespace

ProfScope::ProfScope(void* stream, int kernel, double work) : slot_(-1), stream_(stream) {
  if (!g_prof_on) return;  // unsynchronised read: profiling is switched while no call is in flight
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on) return;
  if (g_prof_used >= static_cast<int>(g_prof_slots.size())) { ++g_prof_dropped; return; }
  slot_ = g_prof_used++;
  g_prof_slots[slot_].kernel = kernel;
  g_prof_slots[slot_].work = work;
  cudaEventRecord(g_prof_slots[slot_].a, static_cast<cudaStream_t>(stream_));
}
ProfScope::~ProfScope() {
  if (slot_ < 0) return;
  cudaEventRecord(g_prof_slots[slot_].b, static_cast<cudaStream_t>(stream_));
}
}  // namespace msf

extern "C" int msf_prof_begin(int capacity) {
  using namespace msf;
  MSF_REQUIRE(capacity > 0 && capacity <= (1 << 22), MSF_ERR_INVALID, "capacity %d out of range", capacity);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  MSF_REQUIRE(!g_prof_on, MSF_ERR_INVALID, "profiler already running");
  g_prof_slots.resize(capacity);
  for (auto& s : g_prof_slots) {
    if (!s.a) MSF_CUDA_OK(cudaEventCreate(&s.a));
    if (!s.b) MSF_CUDA_OK(cudaEventCreate(&s.b));
  }
  g_prof_used = g_prof_dropped = 0;
  g_prof_on = true;
  return MSF_OK;
}

extern "C" int msf_prof_end(msf_prof_record* out, int* dropped) {
  using namespace msf;
  MSF_REQUIRE(out, MSF_ERR_INVALID, "NULL output");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  MSF_REQUIRE(g_prof_on, MSF_ERR_INVALID, "profiler not running");
  g_prof_on = false;
  for (int k = 0; k < MSF_K_COUNT; ++k) {
    out[k].kernel = k;
    out[k].launches = 0;
    out[k].work = 0.0;
    out[k].ms = 0.0;
  }
  int rc = MSF_OK;
  for (int i = 0; i < g_prof_used; ++i) {
    ProfSlot& s = g_prof_slots[i];
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(s.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, s.a, s.b);
    if (e != cudaSuccess) { set_error("profiler event failed: %s", cudaGetErrorString(e)); rc = MSF_ERR_CUDA; continue; }
    out[s.kernel].launches += 1;
    out[s.kernel].work += s.work;
    out[s.kernel].ms += ms;
  }
  if (dropped) *dropped = g_prof_dropped;
  for (auto& s : g_prof_slots) {
    if (s.a) cudaEventDestroy(s.a);
    if (s.b) cudaEventDestroy(s.b);
  }
  g_prof_slots.clear();
  return rc;
}

extern "C" const char* msf_prof_kernel_name(int kernel) {
  return (kernel >= 0 && kernel < MSF_K_COUNT) ? msf::kKernelNames[kernel] : "?";
}
extern "C" int msf_prof_kernel_bound(int kernel) {
  return (kernel >= 0 && kernel < MSF_K_COUNT) ? msf::kKernelBound[kernel] : 0;
}
