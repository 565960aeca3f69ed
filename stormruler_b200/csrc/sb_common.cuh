// sb_common.cuh -- context, error handling and layout constants shared by the library TUs.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/stormb200.h"

namespace sb {

// ---- layout constants ("SB_TREE v1") ---------------------------------------------
// A CTA tile is 2048 consecutive elements / rows: 8 warps x 4 sub-iterations x 32 lanes x 2.
// Vector capacities and the ELL leading dimension are padded to a multiple of kTile so every
// lane can use unguarded 128-bit accesses; only reductions mask by the logical length.
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSub = 4;
constexpr int kTile = kThreads * 2 * kSub; // 2048
constexpr int kMaxDots = 3;
constexpr int32_t kColPad = INT32_MIN;

inline int64_t pad_up(int64_t n) { return ((n + kTile - 1) / kTile) * kTile; }
inline int64_t num_tiles(int64_t n) { return (n + kTile - 1) / kTile; }

// ---- errors ----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define SB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::sb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SB_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define SB_REQUIRE(cond, msg)                                                   \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ::sb::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg);  \
      return SB_ERR_INVALID;                                                    \
    }                                                                           \
  } while (0)

#define SB_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc != SB_OK) return _rc; \
  } while (0)

// Device-resident reduction scratch: per-tile partial sums and the output slots.
struct RedScratch {
  double* partials = nullptr; // [kMaxDots][cap_tiles]
  int64_t cap_tiles = 0;
  double* result = nullptr;   // [64] result slots of the stand-alone dots
};

} // namespace sb

struct sb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  sb::RedScratch red;
  double* h_pinned = nullptr; // pinned staging (results, flags)
  size_t pinned_doubles = 0;
  int64_t launches = 0;
  int sm_count = 0;
  // solver workspaces, grown on demand and reused across solves
  std::vector<double*> work;
  size_t work_n = 0;
  struct SolverState* d_state = nullptr; // device
  double* d_hist = nullptr;
  int64_t hist_cap = 0;
  double* d_trace = nullptr;
  int64_t trace_cap = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace sb {
int ensure_red_scratch(sb_ctx* ctx, int64_t n);
} // namespace sb
