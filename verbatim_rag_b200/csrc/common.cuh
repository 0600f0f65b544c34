// Shared host-side plumbing of libvrag_b200: context, error reporting, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vrag_b200.h"

namespace vrag {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define VRAG_CUDA(expr)                                                                              \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      throw ::vrag::Error(VRAG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                             __FILE__ + ":" + std::to_string(__LINE__) + ")");       \
  } while (0)

#define VRAG_CHECK(cond, code, msg)                       \
  do {                                                    \
    if (!(cond)) throw ::vrag::Error((code), (msg));      \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Device buffer owned by the library (workspaces, weights, corpora).
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void reserve(size_t n) {
    if (n <= bytes) return;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    VRAG_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace vrag

// Optional per-launch CUDA-event profiler (bench.py's roofline leg): start/stop events on the launching stream
// around every kernel, summed per class on read.  Off by default (no events recorded).
namespace vrag {
enum ProfClass { PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_ROWOPS = 2, PROF_SCAN = 3, PROF_SELECT = 4, PROF_OTHER = 5,
                 PROF_NCLASS = 6 };
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cls; size_t e0, e1; };
  std::vector<Rec> recs;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
};
}  // namespace vrag

// The context behind the opaque C handle.
struct vrag_ctx {
  vrag::Profiler prof;
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  vrag::PFN_encodeTiled encode_tiled = nullptr;
  std::string last_error;
  std::mutex mu;           // one call at a time per context (plugin objects are shared across threads)
  int gemm_stages = 5;     // operand ring depth of the tcgen05 GEMM (VRAG_GEMM_STAGES = 3 | 4 | 5)
  uint64_t launches = 0;   // kernels launched through this context (bench.py's gpu_launches)
  // pinned staging for the *_host entry points
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  void* pinned_reserve(size_t n);
};

namespace vrag {

// 2D row-major tensor map, 128-byte swizzle, box = {box_cols (128 bytes worth), box_rows}.
// (CU_TENSOR_MAP_SWIZZLE_64B: inner box of 64 bytes.)
CUtensorMap make_tmap_2d(vrag_ctx* ctx, const void* base, CUtensorMapDataType dt, size_t elem_bytes, uint64_t rows,
                         uint64_t cols, uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols,
                         CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);

struct ProfScope {
  vrag_ctx* c;
  int cls;
  cudaStream_t st;   // the stream the bracketed kernels are launched on (default: the context's main stream)
  size_t e0 = 0;
  ProfScope(vrag_ctx* ctx, int k, cudaStream_t stream = nullptr) : c(ctx), cls(k), st(stream ? stream : ctx->stream) {
    if (c->prof.on) {
      e0 = c->prof.used;
      cudaEventRecord(c->prof.get(), st);
    }
  }
  ~ProfScope() {
    if (c->prof.on) {
      size_t e1 = c->prof.used;
      cudaEventRecord(c->prof.get(), st);
      c->prof.recs.push_back({cls, e0, e1});
    }
  }
};

struct StreamGuard {
  vrag_ctx* c;
  explicit StreamGuard(vrag_ctx* ctx) : c(ctx) { VRAG_CUDA(cudaSetDevice(ctx->device)); }
};

}  // namespace vrag
