// TEST INFRASTRUCTURE -- not product code.
//
// Compiles the product's OWN device code of sb_eval_group -- struct GroupBody, stormruler_b200/csrc/sb_group_body.cuh,
// included verbatim -- for the host, and drives it the way ew_kernel<ND, GroupBody> + block_reduce_partials +
// final_reduce_kernel do (csrc/sb_kernels.cuh): 256 threads per 2048-row tile, thread t owns elements
// tile*2048 + (t/32)*256 + j*64 + 2*(t%32) + {0,1} for j = 0..3; per-lane masked accumulation; xor butterfly per warp;
// 8 warp sums left to right per tile; tile partials strided over 256 threads in batches of 8, butterfly, 8 warp sums.
// The host stand-ins of the CUDA primitives round every operation separately (this file is built with
// -ffp-contract=off). tests/test_dropin_emulated.py checks the result against the emulator of the C ABI (statements)
// and the oracle's restatement of the reduction tree (dots): the kernel body's logic is pinned without a GPU; what is
// left for the device run is the launch plumbing, which is shared with kernels already proven there.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/stormb200.h"

#define __device__
#define __forceinline__ inline

struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }

namespace sb {
constexpr int kThreads = 256, kWarps = 8, kSub = 4, kTile = 2048, kMaxDots = 3;
static inline double2 ld2(const double* p, int64_t e) { return double2{p[e], p[e + 1]}; }
static inline void st2(double* p, int64_t e, double2 v) { p[e] = v.x, p[e + 1] = v.y; }
// csrc/sb_kernels.cuh: acc_pair
static inline void acc_pair(double& acc, int64_t e0, int64_t n, double p0, double p1) {
  acc = __dadd_rn(acc, (e0 < n) ? p0 : 0.0);
  acc = __dadd_rn(acc, (e0 + 1 < n) ? p1 : 0.0);
}
#include "../../stormruler_b200/csrc/sb_group_body.cuh"

static double butterfly_lane0(double* v) { // v[32]: every lane ends with the same bits; lane 0 is returned
  for (int m = 16; m >= 1; m >>= 1) {
    double w[32];
    for (int l = 0; l < 32; ++l) w[l] = __dadd_rn(v[l], v[l ^ m]);
    std::memcpy(v, w, sizeof w);
  }
  return v[0];
}
} // namespace sb

// Vectors must have a capacity padded to a multiple of 2048 doubles (like sb_vec_alloc's): the kernel stores whole tiles.
extern "C" __attribute__((visibility("default"))) int group_body_host_run(size_t n, int n_stmt, const sb_chain* stmts,
                                                                           int n_dots, const double* const* da,
                                                                           const double* const* db, double* out) {
  using namespace sb;
  if (n_stmt < 0 || n_stmt > SB_GROUP_MAX_STMT || n_dots < 0 || n_dots > kMaxDots) return -1;
  GroupBody body{};
  for (int s = 0; s < n_stmt; ++s) body.st[s] = stmts[s];
  body.n_stmt = n_stmt, body.n_dots = n_dots;
  for (int d = 0; d < n_dots; ++d) body.da[d] = da[d], body.db[d] = db[d];
  const int64_t tiles = ((int64_t) n + kTile - 1) / kTile;
  std::vector<double> partial((size_t) kMaxDots * (size_t) tiles, 0.0);
  for (int64_t tile = 0; tile < tiles; ++tile) {
    double lane_acc[kMaxDots][kThreads];
    for (int t = 0; t < kThreads; ++t) {
      const int warp = t >> 5, lane = t & 31;
      double acc[kMaxDots] = {0.0, 0.0, 0.0};
      GroupBody::Regs regs;
      for (int j = 0; j < kSub; ++j)
        body.run(tile * kTile + warp * (kTile / kWarps) + j * 64 + 2 * lane, (int64_t) n, regs, acc);
      for (int d = 0; d < kMaxDots; ++d) lane_acc[d][t] = acc[d];
    }
    for (int d = 0; d < kMaxDots; ++d) { // block_reduce_partials
      double s = 0.0;
      for (int w = 0; w < kWarps; ++w) {
        const double v = butterfly_lane0(&lane_acc[d][32 * w]);
        s = w == 0 ? v : __dadd_rn(s, v);
      }
      partial[(size_t) d * tiles + tile] = s;
    }
  }
  for (int d = 0; d < n_dots; ++d) { // final_reduce_kernel
    double tsum[kThreads];
    for (int t = 0; t < kThreads; ++t) {
      double s = 0.0;
      for (int64_t q0 = t; q0 < tiles; q0 += 8 * kThreads)
        for (int b = 0; b < 8; ++b) {
          const int64_t q = q0 + (int64_t) b * kThreads;
          if (q < tiles) s = __dadd_rn(s, partial[(size_t) d * tiles + q]);
        }
      tsum[t] = s;
    }
    double total = 0.0;
    for (int w = 0; w < kWarps; ++w) {
      const double v = butterfly_lane0(&tsum[32 * w]);
      total = w == 0 ? v : __dadd_rn(total, v);
    }
    out[d] = total;
  }
  return 0;
}
