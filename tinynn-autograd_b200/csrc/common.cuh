// Shared internals of libtnn_b200: error plumbing, the per-process context and launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/tnn_b200.h"

namespace tnn {

struct Context {
  bool inited = false;
  int device = 0;
  int sm_count = 148;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;       // compute stream: every kernel goes here
  cudaStream_t copy_stream = nullptr;  // H2D prefetch
  cudaStream_t comm_stream = nullptr;  // gradient all-reduce chunks running beside the optimiser (comm.cu)
  cudaStream_t d2h_stream = nullptr;   // asynchronous read-backs (tnn_d2h_async): a host read of step i's
                                       // loss must not wait for step i+1, which is already queued
  cudaEvent_t ev_comm = nullptr;       // last comm-stream fence
  cudaEvent_t ev_copy = nullptr;       // last copy-stream fence
  cudaEvent_t ev_compute = nullptr;    // last compute-stream fence
  uint64_t launches = 0;               // kernels launched by this library
  int prof_family = 0;                 // tnn_prof_enable
  void* scratch = nullptr;             // reduction partials (grown on demand)
  size_t scratch_bytes = 0;
  void* l2_flush_buf = nullptr;
  size_t l2_flush_bytes = 0;
};

Context& ctx();
void set_error(const std::string& msg);
int fail(const char* file, int line, const std::string& msg);
// reduction scratch on the compute stream (valid until the next call that asks for scratch)
int get_scratch(size_t nbytes, void** out);

// bracket a launch with events when profiling of `family` is on
void prof_begin(int family);
void prof_end(int family);

}  // namespace tnn

#define TNN_FAIL(msg) return ::tnn::fail(__FILE__, __LINE__, (msg))

#define TNN_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess)                                                          \
      return ::tnn::fail(__FILE__, __LINE__,                                        \
                         std::string(#call) + ": " + cudaGetErrorString(_e));       \
  } while (0)

#define TNN_REQUIRE_INIT()                                                          \
  do {                                                                              \
    if (!::tnn::ctx().inited) TNN_FAIL("tnn_init() has not been called");           \
  } while (0)

// after a <<<>>> launch: count it and surface launch-configuration errors immediately
#define TNN_POST_LAUNCH()                                                           \
  do {                                                                              \
    ::tnn::ctx().launches++;                                                        \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess)                                                          \
      return ::tnn::fail(__FILE__, __LINE__,                                        \
                         std::string("kernel launch: ") + cudaGetErrorString(_e));  \
  } while (0)

namespace tnn {

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
};
template <>
struct Vec4<double> {
  using type = double4;
};

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride elementwise kernel: enough CTAs to fill the chip a few times over,
// in multiples of the SM count
inline int ew_grid(int64_t work_items, int threads, int per_sm = 8) {
  int64_t blocks = ceil_div(work_items, threads);
  int64_t cap = (int64_t)ctx().sm_count * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace tnn
