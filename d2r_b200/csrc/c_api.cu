// C ABI glue: probes, error reporting, dtype dispatch of d2r_gemm.
#include <string.h>

#include "common.cuh"

namespace d2r {

std::atomic<long long> g_launches{0};

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int gemm_tc(const d2r_gemm_args& a, cudaStream_t stream);
int gemm_tc_prof_slots();
extern std::atomic<long long*> g_prof_buf;
int gemm_simt(const d2r_gemm_args& a, cudaStream_t stream);

}  // namespace d2r

extern "C" {

int d2r_abi_version(void) { return D2R_B200_ABI_VERSION; }
const char* d2r_build_arch(void) { return "sm_100a"; }
const char* d2r_last_error(void) { return d2r::last_error_buf(); }
int64_t d2r_launch_count(void) { return d2r::g_launches.load(); }

int d2r_gemm_set_profile(int64_t* records) {
  d2r::g_prof_buf.store(reinterpret_cast<long long*>(records));
  return d2r::gemm_tc_prof_slots();
}

int d2r_gemm(const d2r_gemm_args* args, void* stream) {
  if (!args) return d2r::set_error(D2R_ERR_ARG, "gemm: null args");
  if (!args->a || !args->b || !args->c) return d2r::set_error(D2R_ERR_ARG, "gemm: null operand");
  auto st = static_cast<cudaStream_t>(stream);
  if (args->dtype == D2R_BF16) return d2r::gemm_tc(*args, st);
  if (args->dtype == D2R_F32) return d2r::gemm_simt(*args, st);
  return d2r::set_error(D2R_ERR_ARG, "gemm: bad dtype %d", args->dtype);
}

}  // extern "C"
