// sfb_internal.hpp -- host-side internals shared by the translation units of libsfb.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sfb.h"

namespace sfbi {

constexpr int kNumSlots = 3;  // staging slots of the host-buffer pipeline (H2D / compute / D2H overlap)
constexpr int kCountersPerStream = 16;  // work-queue heads per stream, recycled in stream order

struct Slot
{
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  void* dev = nullptr;
  size_t bytes = 0;
};

// device workspace owned by a handle (polish scratch, sparse working set, staging of host buffers)
struct Scratch
{
  void* dev = nullptr;
  size_t bytes = 0;
};

std::string& create_error();

}  // namespace sfbi

struct sfb_context
{
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  unsigned long long* counters = nullptr;  // work-queue heads, one per launch in flight
  int next_counter[sfbi::kNumSlots + 1] = {};
  uint64_t launches = 0;
  std::string last_error;
  sfbi::Slot slots[sfbi::kNumSlots];
  sfbi::Scratch scratch[sfbi::kNumSlots + 1];  // [kNumSlots] belongs to the handle's own stream
  cudaEvent_t ev_start = nullptr;
  cudaEvent_t ev_order = nullptr;  // orders the handle's workspaces across a change of stream (sfb_set_stream)
  int dinf_guard = 1;  // SFB_OPT_DUAL_INF_DX_GUARD
  int force_polish_scratch = 0;  // SFB_OPT_FORCE_POLISH_SCRATCH
  int polish_form = 0;           // SFB_OPT_POLISH_FORM
  bool ekf_force_generic = false;
  bool dense_force_generic = false;  // SFB_DENSE_FORCE_GENERIC=1: bypass the tall-skinny register kernel (A/B measurements)
  int sparse_tw = 0;     // SFB_SPARSE_TW=4|8|32 overrides the tile-width heuristic of the sparse QP path (A/B measurements)
  int sparse_kernel = 0; // SFB_SPARSE_KERNEL=tiled (1) forces the HBM-tiled kernel, =cta (2) / unset: the on-chip kernel when the instance fits in shared memory
  sfbi::Scratch sparse_cta_ws; // per-CTA global vectors of the on-chip sparse kernel
  sfbi::Scratch sparse_scale_ws; // equilibration of an fp32 on-chip solve, handed to its fp64 polish pass
  sfbi::Scratch sparse_ws;     // tiled working set of the sparse QP path
  sfbi::Scratch sparse_stage;  // device copies of host buffers (sparse path)
  sfbi::Scratch csc_tmp;       // CSC -> CSR permuted A values (sfb_qp_solve_sparse_batch_csc_f64)
  sfbi::Scratch act_tmp;       // active sets of an fp32 solve when the caller did not ask for them (mixed-precision polish)
};

namespace sfbi {

int fail(sfb_context* h, int code, const char* fmt, ...);
#define SFB_CUDA(h, call)                                                                              \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return fail(h, SFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,  \
                  __LINE__);                                                                           \
  } while (0)

// 0 = host, 1 = device
int mem_space(const void* p);
// classify a set of pointers (nullptr entries ignored): 0 all host, 1 all device, -1 mixed
int classify(std::initializer_list<const void*> ps);
unsigned long long* next_counter(sfb_context* h, int slot);
int ensure_slot(sfb_context* h, Slot& s, size_t bytes);
int ensure_scratch(sfb_context* h, Scratch& s, size_t bytes, cudaStream_t st);
int check_params(sfb_context* h, const sfb_qp_params* prm, int64_t batch, int n, int m);
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace sfbi
