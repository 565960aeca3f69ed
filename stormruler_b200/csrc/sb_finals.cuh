// sb_finals.cuh -- what happens behind every reduction of the fused CG / BiCGStab solvers: the solver state recorder,
// the scalar updates ("Final" functors; reference: SolverCg.hpp:73-124, SolverBiCgStab.hpp:83-164, Solver.hpp:116-147),
// and the in-kernel reducer that runs them as the LAST CTA of the kernel that produced the partial sums.
#pragma once

// Included by sb_kernels.cuh behind RedPtrs / warp_butterfly.

namespace sb {

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ double safe_divide(double x, double y) {
  // Crow/MathUtils.hpp:49-52
  return (y == 0.0) ? 0.0 : __ddiv_rn(x, y);
}

struct Recorder {
  SolverState* st;
  double* hist;
  double* trace;
  __device__ void push_trace(double v) const {
    if (trace != nullptr && st->n_trace < st->trace_cap) trace[st->n_trace] = v;
    st->n_trace++;
  }
  __device__ void push_hist(double v) const {
    if (hist != nullptr && st->n_hist < st->hist_cap) hist[st->n_hist] = v;
    st->n_hist++;
  }
  // Solver.hpp:124-128: early exit when the initial residual is already below abs_tol.
  __device__ void init_error(double err) const {
    st->initial_err = err, st->abs_err = err, st->rel_err = 0.0;
    st->iteration = 0;
    push_hist(err);
    if (st->abs_tol > 0.0 && err < st->abs_tol) st->converged = 1, st->done = 1;
    if (st->max_iter <= 0) st->done = 1;
  }
  // Solver.hpp:132-140: one pass of the iteration loop after iterate() returned `err`.
  __device__ void iteration_error(double err) const {
    st->abs_err = err;
    st->rel_err = __ddiv_rn(err, st->initial_err); // no zero guard
    push_hist(err);
    bool conv = (st->abs_tol > 0.0) && (err < st->abs_tol);
    conv |= (st->rel_tol > 0.0) && (st->rel_err < st->rel_tol);
    st->iteration++;
    if (conv) st->converged = 1;
    if (conv || st->iteration >= st->max_iter) st->done = 1;
  }
};

// One-CTA final stage (no folding: initialisation, NCCL mode): run the scalar update in place, then publish the stop.
template<class Inner>
struct PublishFinal {
  Inner inner;
  SolveBlock* blk;
  __device__ void operator()(const double* s) const {
    inner(s);
    if (inner.rec.st->done) blk->final_() = *inner.rec.st, blk->done = 1;
  }
};

// ---- CG -------------------------------------------------------------------------------------------
struct CgInitFinal { // after r = b - A x fused with <r,r>   (SolverCg.hpp:73,80,83)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.st->gamma = s[0];
    rec.push_trace(s[0]);
    rec.init_error(sqrt(s[0]));
  }
};
struct CgAlphaFinal { // after z = A p fused with <p,z>       (:95-96)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.push_trace(s[0]);
    rec.st->alpha = safe_divide(rec.st->gamma, s[0]);
  }
};
struct CgBetaFinal { // after the x/r update fused with <r,r>  (:109,114,121,124)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    const double gamma_bar = rec.st->gamma;
    rec.st->gamma = s[0];
    rec.push_trace(s[0]);
    rec.st->beta = safe_divide(s[0], gamma_bar);
    rec.iteration_error(sqrt(s[0]));
  }
};

// ---- BiCGStab -------------------------------------------------------------------------------------
struct BiInitFinal { // r = b - A x, r~ = r, rho = <r~,r>      (SolverBiCgStab.hpp:83,88-91)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.st->rho = s[0];
    rec.push_trace(s[0]);
    rec.init_error(sqrt(s[0]));
  }
};
struct BiAlphaFinal { // after v = A p fused with <r~,v>        (:137,139)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    rec.push_trace(s[0]);
    rec.st->alpha = safe_divide(rec.st->rho, s[0]);
  }
};
struct BiOmegaFinal { // after t = A r fused with <t,t>, <t,r>  (:158-160)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    // g++ evaluates safe_divide's arguments right to left: <t,t> is traced before <t,r>.
    rec.push_trace(s[0]);
    rec.push_trace(s[1]);
    rec.st->omega = safe_divide(s[1], s[0]);
  }
};
struct BiEndFinal { // after the final update fused with <r,r> and <r~,r>   (:164 and next :115-117)
  Recorder rec;
  __device__ void operator()(const double* s) const {
    const double nrm = sqrt(s[0]);
    rec.push_trace(nrm);
    rec.iteration_error(nrm);
    if (!rec.st->done) {
      // head of the next iteration: rho_bar <- rho, rho <- <r~,r>, beta <- (alpha*rho)/(omega*rho_bar)
      const double rho_bar = rec.st->rho;
      rec.st->rho = s[1];
      rec.push_trace(s[1]);
      rec.st->beta = safe_divide(__dmul_rn(rec.st->alpha, s[1]), __dmul_rn(rec.st->omega, rho_bar));
    }
  }
};


// ---- the reduction finished INSIDE the kernel that produced the partial sums ("in-kernel reducer") --------------------
// A reducing kernel of the fused solvers is launched with one extra CTA, the last of its grid. While the tile CTAs
// work, that CTA runs the final stage of SB_TREE over their partial sums -- thread t adds partial[t], partial[t + 256],
// ... in that order, exactly like final_stage() -- taking each partial as soon as it exists: the slots live in a set of
// their own that always holds a NaN sentinel between kernels, a tile CTA deposits its sum with one relaxed store (the
// value is the flag: no fence, no ticket, nothing that makes a tile CTA wait for its own y stores), the reducer polls
// the slot, takes the value and puts the sentinel back. Behind the last partial: warp butterflies, the 8 warp sums left
// to right, the rank-ordered all-reduce over NVLink peer memory, the solver's scalar update, the stop flag, the count
// of the distributed apply. What the one-CTA kernel behind the producer did (final_reduce_kernel), minus its launch,
// minus the producer's kernel boundary in front of it, and with the partials consumed as they arrive instead of after
// the grid has drained. Nobody waits for the reducer inside the kernel: the consumer is the NEXT kernel.
enum FinalKind : int32_t { kFinalNone = 0, kFinalCgAlpha, kFinalCgBeta, kFinalBiAlpha, kFinalBiOmega, kFinalBiEnd };

struct ReducerArgs {
  int32_t kind = kFinalNone; // kFinalNone: no reducer CTA in this launch (the tile CTAs store to RedPtrs as always)
  int32_t nd = 0;            // number of sums
  int64_t n_tiles = 0;
  double* slots = nullptr;   // [nd][cap_tiles], sentinel-filled between kernels
  int64_t cap_tiles = 0;
  Recorder rec{};
  SolveBlock* blk = nullptr;
  CommDev comm{};            // mode SB_COMM_P2P and world > 1: all-reduce over the ranks
  CommCtrl* bump = nullptr;  // the kernel is a distributed apply: count it
  unsigned long long* wait_ns = nullptr; // optional timeline: the wait for the other ranks' sums
};

// What a tile CTA does with its sums when the launch has a reducer.
__device__ __forceinline__ void deposit_partial(double* slot, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(slot), "d"(v) : "memory");
}

static __device__ __noinline__ void reducer_role(const ReducerArgs& ra) {
  __shared__ double s_w[kMaxDots][kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long t_begin = globaltimer_ns();
  if (threadIdx.x < 2) prefetch_l1(reinterpret_cast<const unsigned char*>(ra.rec.st) + 128 * threadIdx.x); // the scalar update's state
  bool gave_up = false;
  double sums[kMaxDots];
  // Thread t takes partial[t], partial[t + 256], ... of every sum, in that order (final_stage's order). The loads of a
  // batch -- all sums, four slots each -- are issued together: one L2 round trip for everything that has already been
  // deposited; only a slot that still holds the sentinel is polled. (The first version polled slot by slot: six
  // dependent round trips per thread for two sums at 617 tiles, 4-9 us behind the last tile CTA, more than the one-CTA
  // kernel it replaced: profiles/r02_ab_n8_10M.txt, variant red+stream.)
  constexpr int kBatch = 4;
  double s[kMaxDots];
#pragma unroll
  for (int d = 0; d < kMaxDots; ++d) s[d] = 0.0;
  for (int64_t q0 = threadIdx.x; q0 < ra.n_tiles; q0 += (int64_t) kBatch * kThreads) {
    unsigned long long v[kMaxDots][kBatch];
#pragma unroll
    for (int d = 0; d < kMaxDots; ++d)
#pragma unroll
      for (int b = 0; b < kBatch; ++b)
        v[d][b] = (d < ra.nd && q0 + (int64_t) b * kThreads < ra.n_tiles) ? kArSentinel : 0ull;
    unsigned spins = 0;
    for (;;) { // every pass re-reads ALL slots of the batch that are still missing, independently of each other
      bool missing = false;
#pragma unroll
      for (int d = 0; d < kMaxDots; ++d) {
        const unsigned long long* part = reinterpret_cast<const unsigned long long*>(ra.slots + (int64_t) d * ra.cap_tiles);
#pragma unroll
        for (int b = 0; b < kBatch; ++b)
          if (v[d][b] == kArSentinel) v[d][b] = ld_relaxed_gpu(part + q0 + (int64_t) b * kThreads);
      }
#pragma unroll
      for (int d = 0; d < kMaxDots; ++d)
#pragma unroll
        for (int b = 0; b < kBatch; ++b) missing |= v[d][b] == kArSentinel;
      if (!missing || gave_up) break;
      __nanosleep(20);
      // a tile CTA that never deposits (it cannot happen short of a device fault) must not hang the device
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t_begin > ra.comm.timeout_ns + 20000000000ull) gave_up = true;
    }
#pragma unroll
    for (int d = 0; d < kMaxDots; ++d) {
      unsigned long long* part = reinterpret_cast<unsigned long long*>(ra.slots + (int64_t) d * ra.cap_tiles);
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int64_t q = q0 + (int64_t) b * kThreads;
        if (d < ra.nd && q < ra.n_tiles) {
          part[q] = kArSentinel;
          s[d] = __dadd_rn(s[d], __longlong_as_double((long long) v[d][b]));
        }
      }
    }
  }
#pragma unroll
  for (int d = 0; d < kMaxDots; ++d) {
    if (d < ra.nd) {
      const double w = warp_butterfly(s[d]);
      if (lane == 0) s_w[d][warp] = w;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int d = 0; d < ra.nd; ++d) {
      double t = s_w[d][0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) t = __dadd_rn(t, s_w[d][w]);
      sums[d] = t;
    }
  }
  if (ra.comm.mode == SB_COMM_P2P && ra.comm.world > 1) {
    const unsigned long long t0 = (ra.wait_ns != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0;
    if (ra.nd == 1) {
      double one[1] = {sums[0]};
      allreduce_p2p<1>(ra.comm, one);
      sums[0] = one[0];
    } else if (ra.nd == 2) {
      double two[2] = {sums[0], sums[1]};
      allreduce_p2p<2>(ra.comm, two);
      sums[0] = two[0], sums[1] = two[1];
    } else {
      allreduce_p2p<kMaxDots>(ra.comm, sums);
    }
    if (ra.wait_ns != nullptr && threadIdx.x == 0) *ra.wait_ns = globaltimer_ns() - t0;
  }
  if (threadIdx.x == 0) {
    switch (ra.kind) {
      case kFinalCgAlpha: PublishFinal<CgAlphaFinal>{CgAlphaFinal{ra.rec}, ra.blk}(sums); break;
      case kFinalCgBeta: PublishFinal<CgBetaFinal>{CgBetaFinal{ra.rec}, ra.blk}(sums); break;
      case kFinalBiAlpha: PublishFinal<BiAlphaFinal>{BiAlphaFinal{ra.rec}, ra.blk}(sums); break;
      case kFinalBiOmega: PublishFinal<BiOmegaFinal>{BiOmegaFinal{ra.rec}, ra.blk}(sums); break;
      case kFinalBiEnd: PublishFinal<BiEndFinal>{BiEndFinal{ra.rec}, ra.blk}(sums); break;
      default: break;
    }
    if (ra.bump != nullptr) ra.bump->apply_seq = ra.bump->apply_seq + 1; // the distributed apply I belong to is complete
  }
}

} // namespace sb
