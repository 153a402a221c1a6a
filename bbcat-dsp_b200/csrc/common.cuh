// common.cuh -- error plumbing and small helpers shared by the libbbx translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/bbx.h"

namespace bbx {

// last error message of the calling thread (bbx_last_error)
void set_error(const char* fmt, ...);
const char* get_error();

#define BBX_CUDA_TRY(expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      bbx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return BBX_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define BBX_REQUIRE(cond, ...)     \
  do {                             \
    if (!(cond)) {                 \
      bbx::set_error(__VA_ARGS__); \
      return BBX_ERR_INVALID;      \
    }                              \
  } while (0)

inline uint32_t ceil_div(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// growable device scratch used by the host-pointer entry points (one per thread)
struct DeviceScratch {
  void* ptr = nullptr;
  size_t cap = 0;
  int device = -1;
  int ensure(size_t bytes);
  ~DeviceScratch();
};
DeviceScratch& scratch(int which);  // which = 0..3, thread-local, one set per device (the current device's)

// owner of a half-built object inside a *_create function: BBX_CUDA_TRY / BBX_REQUIRE return early, and the destructor then
// releases what was allocated so far through the object's own destroy entry point (which accepts partial objects)
template <typename T>
struct CreateGuard {
  T* p;
  int (*destroy)(T*);
  CreateGuard(T* obj, int (*d)(T*)) : p(obj), destroy(d) {}
  ~CreateGuard() {
    if (p) destroy(p);
  }
  T* release() {
    T* q = p;
    p = nullptr;
    return q;
  }
  CreateGuard(const CreateGuard&) = delete;
  CreateGuard& operator=(const CreateGuard&) = delete;
};

// make sure a CUDA device is usable; sets the error and returns BBX_ERR_CUDA otherwise
int require_device();

// Every handle-based entry point runs with the device its object was created on and gives the caller's current device back
// on return (a single process may hold engines and delay / filter-bank objects on several GPUs).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace bbx
