// Library-wide state: last-error string, kernel-launch counter, version / device probe.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int* g_f16_overflow = nullptr;      // device flag registered by the host side (tsr_set_f16_overflow_flag)

int* tsr_f16_overflow_ptr() { return g_f16_overflow; }

void tsr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char* tsr_last_error(void) { return g_err; }

int tsr_version(void) { return 100; }

long long tsr_launch_count_inc(int n) { return g_launches.fetch_add(n) + n; }
long long tsr_launch_count(void) { return g_launches.load(); }
void tsr_launch_count_reset(void) { g_launches.store(0); }
// kernels launched on the library's behalf without passing through its entry points: a CUDA-graph replay of n captured launches
void tsr_launch_count_add(long long n) { g_launches.fetch_add(n); }

// "fp16" precision mode: a sticky device int that the kernels storing fp16 activations (tsr_conv2d_tc2 with TSR_TC2_F16,
// tsr_bn_apply with fp16 output) set to 1 when a value they store is not finite (|x| > 65504 or NaN).  The caller owns
// the memory (the library never allocates); NULL switches the check off.
void tsr_set_f16_overflow_flag(int* flag_dev) { g_f16_overflow = flag_dev; }

// 0 when the current device can run this library (compute capability 10.x), an error code otherwise.
int tsr_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    tsr_set_error("check_device: no usable CUDA device");
    return TSR_ERR_CUDA;
  }
  if (prop.major != 10) {
    tsr_set_error("check_device: built for sm_100a only, device is sm_%d%d", prop.major, prop.minor);
    return TSR_ERR_UNSUPPORTED;
  }
  return TSR_OK;
}

}  // extern "C"
