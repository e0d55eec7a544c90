// Shared helpers for the debvader_b200 C-ABI library (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <atomic>

#include "../../include/debvader_b200.h"

namespace dbv {

// thread-local message behind dbv_last_error()
inline std::string& last_error() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

extern std::atomic<long long> g_launches;  // every kernel this library launches

#define DBV_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return dbv::fail(DBV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DBV_LAUNCH_CHECK()                                                                      \
  do {                                                                                          \
    dbv::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess)                                                                      \
      return dbv::fail(DBV_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DBV_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return dbv::fail(DBV_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Environment switches (kernel A/B selection, ablations, tuning knobs) exist only in the ablation build
// (-DDBV_ABLATE: libdebvader_b200_ablate.so, used by tools/ and a few cross-check tests).  The product library reads
// no environment variable that changes what it computes or which kernel runs (DBV_VERBOSE only prints the plans).
#ifdef DBV_ABLATE
static inline const char* dbv_env(const char* name) { return getenv(name); }
#define DBV_DBG(x) (x)
#else
static inline const char* dbv_env(const char*) { return nullptr; }
#define DBV_DBG(x) 0
#endif

constexpr int kNumSMs = 148;

// ---- DC2 architecture constants (reference model/model.py:61-161, train.py:104-107) -----------
constexpr int S_ = 59;       // stamp size
constexpr int CB_ = 6;       // bands
constexpr int LAT = 32;      // latent dim
constexpr int NPAR = 560;    // 32 + 32*33/2
constexpr int STAMP_ELTS = S_ * S_ * CB_;

}  // namespace dbv
