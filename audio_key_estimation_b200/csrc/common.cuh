// common.cuh -- shared host/device helpers for libake_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/ake_b200.h"

namespace ake {

// ---- error plumbing -------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);
int64_t& launch_counter();

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define AKE_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      ::ake::fail(AKE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// Call after every kernel launch: counts the launch and surfaces launch-configuration errors.
#define AKE_LAUNCHED()                                                                          \
  do {                                                                                          \
    ++::ake::launch_counter();                                                                  \
    AKE_CUDA(cudaGetLastError());                                                               \
  } while (0)

// Wrap every extern "C" body: exceptions never cross the ABI.
template <class F>
inline int guarded(F&& f) {
  try {
    f();
    return AKE_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return AKE_ERR_INVALID;
  }
}

inline int current_device() {
  int dev = 0;
  AKE_CUDA(cudaGetDevice(&dev));
  return dev;
}

// SM count of the current device (persistent kernels launch one CTA, or a fixed few, per SM); cached per device
inline int sm_count() {
  static int n[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) fail(AKE_ERR_UNSUPPORTED, "device index %d", dev);
  if (!n[dev]) AKE_CUDA(cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev));
  return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE (and per-function) attribute: remember the largest size
// opted into for (current device, kernel) so a second GPU driven from the same process gets its own opt-in.
template <class K>
inline void ensure_dyn_smem(K kern, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> seen;
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = seen[{dev, reinterpret_cast<const void*>(kern)}];
  if (bytes > cur) {
    AKE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
  }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Arena {
  char* base;
  size_t cap, off = 0;
  Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c) {}
  template <class T>
  T* take(size_t n) {
    size_t bytes = align_up(n * sizeof(T), 256);
    if (base && off + bytes > cap) fail(AKE_ERR_WORKSPACE, "workspace too small: need > %zu, have %zu", off + bytes, cap);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
};

// ---- optional per-section device timing (ake_profile_enable / ake_profile_collect) -----------
// When enabled, a ProfScope records a CUDA event pair on the launch stream around the kernels of one
// tagged section; the events are resolved and summed per tag by ake_profile_collect().  Disabled (the
// default) it costs one relaxed atomic load.
bool profile_enabled();
void profile_record(const char* tag, cudaStream_t st, bool begin, int n_launches);
struct ProfScope {
  const char* tag;
  cudaStream_t st;
  bool on;
  int64_t launches0;
  ProfScope(const char* t, cudaStream_t s) : tag(t), st(s), on(profile_enabled()), launches0(launch_counter()) {
    if (on) profile_record(tag, st, true, 0);
  }
  ~ProfScope() {
    if (on) profile_record(tag, st, false, (int)(launch_counter() - launches0));
  }
};

// cqt.cu: the front-end with the per-clip lengths either on the host (validated + staged) or already on the device
void run_cqt(::ake_cqt* p, const float* audio, long long stride, const int64_t* lengths_host, const long long* lengths_dev, int B,
             long long n_max, int mode, float* out, int T_max, int* seq_len_out, void* ws, size_t ws_bytes, cudaStream_t st);

constexpr float kLeakySlope = 0.01f;  // nn.LeakyReLU() default (models.py:197,234,315)
constexpr float kBnEps = 1e-5f;       // nn.BatchNorm2d default

}  // namespace ake
