// Internal helpers shared by the translation units of libdrb.so (not part of the ABI).
#ifndef DRB_INTERNAL_H
#define DRB_INTERNAL_H

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <new>

#include "drb.h"

// Records the message for drb_last_error() (thread-local) and returns `code`.
int drb_fail(int code, const char* fmt, ...);

static inline int64_t drb_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

#ifdef __CUDACC__
#include <cuda_runtime.h>

#include <vector>

struct drb_prof_rec { const char* name; cudaEvent_t beg, end; };

struct drb_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  int64_t launches;
  int sticky;  // first CUDA error seen (sticky)
  int profile; // when set, every kernel launch is bracketed by CUDA events on `stream` (bench.py roofline leg)
  std::vector<drb_prof_rec> recs;
};

// RAII bracket around one kernel launch: records events on the launching stream when profiling is enabled
struct drb_prof_scope {
  drb_ctx* ctx; cudaEvent_t end = nullptr;
  drb_prof_scope(drb_ctx* c, const char* name) : ctx(c) {
    if (!c->profile) return;
    cudaEvent_t beg;
    cudaEventCreate(&beg);
    cudaEventCreate(&end);
    cudaEventRecord(beg, c->stream);
    c->recs.push_back({name, beg, end});
  }
  ~drb_prof_scope() {
    if (end) cudaEventRecord(end, ctx->stream);
  }
};

#define DRB_CUDA_TRY(ctx, expr)                                                                        \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) {                                                                           \
      (ctx)->sticky = (int)e_;                                                                         \
      return drb_fail(DRB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__,    \
                      __LINE__);                                                                       \
    }                                                                                                  \
  } while (0)

// after a <<<>>> launch: count it and surface launch-configuration errors immediately
#define DRB_LAUNCH_CHECK(ctx, name)                                                                    \
  do {                                                                                                 \
    (ctx)->launches++;                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                                               \
    if (e_ != cudaSuccess) {                                                                           \
      (ctx)->sticky = (int)e_;                                                                         \
      return drb_fail(DRB_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_));            \
    }                                                                                                  \
  } while (0)
#endif  // __CUDACC__

#endif
