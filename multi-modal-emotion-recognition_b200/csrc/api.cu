// Library-level entry points: version, error reporting, debug knobs.
#include <stdarg.h>

#include "common.cuh"

namespace mmer {

extern int g_debug[16];
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return MMER_ERR_CUDA;
}

static long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }
long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  // SMs left to a communication kernel that runs beside the step (data-parallel overlap): every persistent grid of
  // this library is sized from this number, and a persistent grid that does not fit in one wave runs twice as long
  const int reserve = g_debug[MMER_DEBUG_RESERVE_SMS];
  return reserve > 0 && reserve < cached ? cached - reserve : cached;
}

}  // namespace mmer

extern "C" {

int mmer_version(void) { return MMER_VERSION; }
const char* mmer_last_error(void) { return mmer::g_err; }
int mmer_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return MMER_ERR_ARG;
  mmer::g_debug[key] = value;
  return 0;
}
int64_t mmer_launch_count(void) { return (int64_t)mmer::launch_count(); }
int mmer_event_create(void** event_out) {
  MMER_CHECK_ARG(event_out != nullptr, "event_create: null pointer");
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e != cudaSuccess) return mmer::cuda_fail(e, "cudaEventCreateWithFlags");
  *event_out = (void*)ev;
  return 0;
}
int mmer_event_destroy(void* event) {
  if (event == nullptr) return 0;
  cudaError_t e = cudaEventDestroy((cudaEvent_t)event);
  return e == cudaSuccess ? 0 : mmer::cuda_fail(e, "cudaEventDestroy");
}
int mmer_stream_wait_event(void* stream, void* event) {
  MMER_CHECK_ARG(event != nullptr, "stream_wait_event: null event");
  cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0);
  return e == cudaSuccess ? 0 : mmer::cuda_fail(e, "cudaStreamWaitEvent");
}
int mmer_debug_get(int key) { return (key < 0 || key >= 16) ? 0 : mmer::g_debug[key]; }

}  // extern "C"
