// sb_kernels.cuh -- kernel framework: tiling, the fixed reduction tree and the expression evaluator.
//
// Everything on the Krylov path is HBM-bound fp64 streaming work, so the framework
// is built around three rules: every lane moves 128-bit words (double2 / int2+double2), all loads of
// a tile are issued before the first store (the target may alias an operand, which would otherwise
// stop the compiler from hoisting them), and every reduction is fused into the kernel that
// produces its operand, with a fixed-shape tree so results do not depend on the grid or the run.
//
// Floating-point contract: element-wise arithmetic uses __dadd_rn/__dmul_rn/__ddiv_rn/__dsub_rn so
// nvcc can never contract a*b+c into an FMA -- the reference's canonical build (g++ -O2, baseline
// x86-64) rounds every operation separately.
#pragma once

#include "sb_comm.cuh"

namespace sb {

struct RedPtrs {
  double* partials;      // [kMaxDots][cap_tiles]
  int64_t cap_tiles;
};

__device__ __forceinline__ double2 ld2(const double* p, int64_t e) {
  return *reinterpret_cast<const double2*>(p + e);
}
__device__ __forceinline__ void st2(double* p, int64_t e, double2 v) {
  *reinterpret_cast<double2*>(p + e) = v;
}

// xor butterfly: after the loop every lane holds the same bits (a+b == b+a in IEEE-754).
__device__ __forceinline__ double warp_butterfly(double v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}

} // namespace sb
#include "sb_finals.cuh" // Recorder, the solvers' scalar updates, the in-kernel reducer (needs warp_butterfly)
namespace sb {

// First element handled by this lane in sub-iteration j of tile `tile` (= blockIdx.x, except in the
// distributed apply kernel whose first CTAs are halo-pack CTAs).
__device__ __forceinline__ int64_t lane_elem(int64_t tile, int j) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  return tile * kTile + warp * (kTile / kWarps) + j * 64 + 2 * lane;
}

// CTA-level combine of SB_TREE v1: warp butterflies, then the 8 warp sums are added left to right and
// stored as partial[tile]. No fence, no atomics: the totals are produced by final_reduce_kernel, a
// one-CTA launch that follows in stream order. (An earlier version let the last CTA finish the
// reduction behind a __threadfence + ticket atomic; ncu/CUDA-event timing showed that the fence, which
// must wait for the CTA's outstanding y stores, stretched every CTA's lifetime and cost ~25 us per
// reducing kernel at 10 M cells, against ~2 us for the extra launch.)
template<int ND>
__device__ __forceinline__ void block_reduce_partials(double (&acc)[ND], const RedPtrs& red, int64_t tile) {
  __shared__ double s_w[ND][kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double v = warp_butterfly(acc[d]);
    if (lane == 0) s_w[d][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < ND) {
    const int d = threadIdx.x;
    double s = s_w[d][0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) s = __dadd_rn(s, s_w[d][w]);
    // one relaxed store: when the launch has an in-kernel reducer (sb_finals.cuh) the slot is being polled
    deposit_partial(red.partials + (int64_t) d * red.cap_tiles + tile, s);
  }
}

// Final stage of SB_TREE v1, run by one CTA of 256 threads. Thread t adds partial[t], partial[t+256], ...
// in that order (loads of a batch are independent and issued together; only the additions are
// sequential), butterfly per warp, the 8 warp sums added left to right. The totals are valid in thread 0.
// `s_w` is caller-provided shared scratch; the function ends behind a __syncthreads.
template<int ND>
__device__ __forceinline__ void final_stage(int64_t n_tiles, const RedPtrs& red, double (&s_w)[kMaxDots][kWarps],
                                            double (&sums)[ND]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kBatch = 8;
  double s[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) s[d] = 0.0;
  for (int64_t q0 = threadIdx.x; q0 < n_tiles; q0 += (int64_t) kBatch * kThreads) {
    double v[ND][kBatch];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      const double* part = red.partials + (int64_t) d * red.cap_tiles;
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int64_t q = q0 + (int64_t) b * kThreads;
        v[d][b] = (q < n_tiles) ? __ldcg(part + q) : 0.0;
      }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int64_t q = q0 + (int64_t) b * kThreads;
        s[d] = (q < n_tiles) ? __dadd_rn(s[d], v[d][b]) : s[d];
      }
    }
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double v = warp_butterfly(s[d]);
    if (lane == 0) s_w[d][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double t = s_w[d][0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) t = __dadd_rn(t, s_w[d][w]);
      sums[d] = t;
    }
  }
}

// The solver state a Final functor updates (PublishFinal<...>), or null (StoreFinal, NoFinal).
template<class F>
__device__ __forceinline__ auto final_state_of(const F& f, int) -> decltype((const void*) f.inner.rec.st) {
  return f.inner.rec.st;
}
template<class F>
__device__ __forceinline__ const void* final_state_of(const F&, long) {
  return nullptr;
}

// The one-CTA kernel behind every reducing kernel of the one-kernel-per-step schedule. `fin(sums)` runs on one
// thread: that is where the solver scalars (alpha, beta, residual, stop flag) are updated.
template<int ND, class Final>
__global__ void __launch_bounds__(kThreads) final_reduce_kernel(int64_t n_tiles, RedPtrs red, Final fin, CommDev comm,
                                                                CommCtrl* bump, const int* __restrict__ done,
                                                                unsigned long long* wait_ns) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  // the scalar update's state: on its way into L1 while the partial sums are read and the ranks exchange theirs
  if (const void* st = final_state_of(fin, 0); st != nullptr && threadIdx.x < 2)
    prefetch_l1(reinterpret_cast<const unsigned char*>(st) + 128 * threadIdx.x);
  __shared__ double s_w[kMaxDots][kWarps];
  double sums[ND];
  final_stage<ND>(n_tiles, red, s_w, sums);
  // multi-GPU, P2P mode: exchange the rank sums with every peer inside this kernel (rank-ordered total)
  if (comm.mode == SB_COMM_P2P && comm.world > 1) {
    const unsigned long long t0 = (wait_ns != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0;
    allreduce_p2p<ND>(comm, sums);
    if (wait_ns != nullptr && threadIdx.x == 0) *wait_ns = globaltimer_ns() - t0;
  }
  if (threadIdx.x == 0) {
    fin(sums);
    if (bump != nullptr) bump->apply_seq = bump->apply_seq + 1; // the distributed apply in front of me is complete
  }
}

template<int M>
struct StoreFinal {
  double* out;
  __device__ void operator()(const double* sums) const {
#pragma unroll
    for (int k = 0; k < M; ++k) out[k] = sums[k];
  }
};

// NCCL mode: the solver's scalar update runs after ncclAllReduce has combined the rank sums.
template<int ND, class Final>
__global__ void scalar_final_kernel(const double* __restrict__ sums, Final fin, const int* __restrict__ done) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  double s[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) s[d] = sums[d];
  fin(s);
}

int nccl_allreduce_sum(sb_ctx* ctx, double* d_buf, int count); // sb_comm.cu

// Launch helper shared by all reducing kernels: the one-CTA final stage, right behind the producer.
template<int ND, class Final>
inline int launch_final(sb_ctx* ctx, int64_t n, const Final& fin, const int* done, CommCtrl* bump = nullptr,
                        unsigned long long* ar_wait_ns = nullptr, bool pdl = false) {
  const RedPtrs red{ctx->red.partials, ctx->red.cap_tiles};
  unsigned long long* const no_wait = nullptr;
  if (ctx->debug & 4) { // experiment: rank-local sums only
    SB_CUDA(launch_kernel(ctx, final_reduce_kernel<ND, Final>, 1, kThreads, 0, num_tiles(n), red, fin, CommDev{}, bump, done, no_wait));
    ctx->launches++;
    return SB_OK;
  }
  if (ctx->comm.mode == SB_COMM_NCCL && ctx->comm.world > 1) {
    SB_CUDA(launch_kernel(ctx, final_reduce_kernel<ND, StoreFinal<ND>>, 1, kThreads, 0, num_tiles(n), red,
                          StoreFinal<ND>{ctx->d_ar}, CommDev{}, (CommCtrl*) nullptr, done, no_wait));
    SB_TRY(nccl_allreduce_sum(ctx, ctx->d_ar, ND));
    SB_CUDA(launch_kernel(ctx, scalar_final_kernel<ND, Final>, 1, 1, 0, (const double*) ctx->d_ar, fin, done));
    ctx->launches += 2;
    return SB_OK;
  }
  if (ctx->prof_mid != nullptr) { // profiled solve: where the reducing kernel ends and its final stage begins
    cudaEvent_t e;
    SB_CUDA(cudaEventCreate(&e));
    ctx->prof_mid->push_back(e); // owned by the solve's guard from here on
    SB_CUDA(cudaEventRecord(e, ctx->stream));
  }
  PdlScope pdl_scope(ctx, pdl);
  SB_CUDA(launch_kernel(ctx, final_reduce_kernel<ND, Final>, 1, kThreads, 0, num_tiles(n), red, fin, ctx->comm, bump, done, ar_wait_ns));
  ctx->launches++;
  return SB_OK;
}

struct NoFinal {
  __device__ void operator()(const double*) const {}
};

// ---- folding a reduction into the kernel that consumes it ----------------------------------------------------------
// The one-CTA final stage costs a launch per reduction (~2.7 us of boundary + ~3 us of kernel inside a replayed graph:
// a third of a BiCGStab iteration at 1.26 M cells per rank, profiles/r02_stepwise_unfolded_1M.json). A kernel that
// needs the reduced scalars can finish the reduction itself instead: its CTA 0 does what the one-CTA kernel did --
// SB_TREE's final stage over the producer's tile partials, the rank-ordered all-reduce over NVLink peer memory, the
// solver's scalar update -- on a private copy of state version `in`, stores the result as version `in ^ 1` and raises
// that version's ready flag; every other CTA has its tile's loads in flight by then, acquires the flag (one L2 round
// trip once it is up) and reads the scalars from the new version. Nobody ever reads a field that is being written:
// the version a kernel's CTAs read is complete before the flag goes up. CTA 0 also lowers the flag of version `in`
// (nobody waits for it in this kernel; the next folding kernel will raise it again), publishes the stop (`final_`,
// `done`) and counts the producer when it was a distributed apply.
// (A first version let EVERY CTA run the scalar update on a copy in shared memory behind a mailbox: two CTA barriers
// and three dependent memory round trips per CTA stretched every CTA's lifetime, and the element-wise kernels lost
// 15-25 % of their bandwidth -- profiles/r02_stepwise_folded_v1_every_cta_waits_*.json.)
template<int ND, class Final>
struct Fold {
  int64_t n_tiles = -1;             // tiles of the producer; < 0: nothing to fold (the kernel reads version `in` as it is)
  RedPtrs red{};
  Final fin{};                      // fin.rec: history / trace buffers
  SolveBlock* blk = nullptr;
  int32_t in = 0;                   // state version to start from; the result goes to in ^ 1
  CommDev comm{};                   // mode SB_COMM_P2P and world > 1: all-reduce over the ranks
  CommCtrl* bump = nullptr;         // the producer was a distributed apply: count it
  unsigned long long* wait_ns = nullptr; // optional: how long CTA 0 waited for the other ranks' sums
};

// CTA 0 of a folding kernel, all threads. Deliberately NOT inlined: inlined, its register needs became those of the
// whole kernel -- the half update went from 32 to 85 registers, i.e. from eight resident CTAs per SM to two
// (profiles/r02_stepwise_folded_v2_acquire_poll_10M.json). As a call made by a CTA that owns no tile it lives within
// the register budget the kernel's __launch_bounds__ gives the element-wise body, nothing is live across it.
// Every resident CTA of the kernel waits for this chain, so it is kept to one memory round trip per stage: the tile
// partials arrive by bulk copy in the CTA's (otherwise unused) dynamic shared memory -- in one piece up to
// `smem_bytes / (8 ND)` tiles, in pieces of a multiple of 256 tiles beyond, so that thread t still adds
// partial[t], partial[t + 256], ... in that order -- while thread 0 already has the old state version in flight; one
// fence, one flag store. (With the register-staged final stage of the one-CTA kernel, compiled under the 32-register
// budget of the half update, the chain took 4 us at 617 tiles and 10 us at 4 937: profiles/r02_fold_reducer_chain.txt.)
template<int ND, class Final>
__device__ __noinline__ void fold_reduce(const Fold<ND, Final>& f, unsigned char* smem, int smem_bytes) {
  __shared__ double s_fin[kMaxDots][kWarps];
  __shared__ __align__(8) uint64_t bar;
  const unsigned long long t_start = (f.wait_ns != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SolverState st; // thread 0: version `in`, fetched around L1 (its lines must not linger in this SM's L1)
  if (threadIdx.x == 0) {
    f.blk->ready[f.in] = 0; // nobody waits for it during this kernel
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&f.blk->ver(f.in));
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(&st);
#pragma unroll
    for (int k = 0; k < (int) (sizeof(SolverState) / 8); ++k) dst[k] = __ldcg(src + k);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t) __cvta_generic_to_shared(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  double* buf = reinterpret_cast<double*>(smem);
  const int64_t per_piece = ((int64_t) (smem_bytes / 8) / ND) & ~(int64_t) 255; // tiles per piece, a multiple of 256
  double s[ND];
#pragma unroll
  for (int d = 0; d < ND; ++d) s[d] = 0.0;
  uint32_t phase = 0;
  for (int64_t base = 0; base < f.n_tiles; base += per_piece) {
    const int64_t cnt = f.n_tiles - base < per_piece ? f.n_tiles - base : per_piece;
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t) (((cnt * 8) + 15) & ~(int64_t) 15); // partial sets are 16-byte aligned and padded
      const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes * ND) : "memory");
#pragma unroll
      for (int d = 0; d < ND; ++d)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t) __cvta_generic_to_shared(buf + d * per_piece)),
                     "l"(f.red.partials + (int64_t) d * f.red.cap_tiles + base), "r"(bytes), "r"(bar_s)
                     : "memory");
    }
    {
      uint32_t ok = 0;
      const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar);
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar_s), "r"(phase)
                     : "memory");
      phase ^= 1u;
    }
#pragma unroll
    for (int d = 0; d < ND; ++d)
      for (int64_t q = threadIdx.x; q < cnt; q += kThreads) s[d] = __dadd_rn(s[d], buf[d * per_piece + q]);
    __syncthreads(); // the next piece overwrites the buffer
  }
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    const double v = warp_butterfly(s[d]);
    if (lane == 0) s_fin[d][warp] = v;
  }
  __syncthreads();
  double sums[ND];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double t = s_fin[d][0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) t = __dadd_rn(t, s_fin[d][w]);
      sums[d] = t;
    }
  }
  if (f.comm.mode == SB_COMM_P2P && f.comm.world > 1) {
    const unsigned long long t0 = (f.wait_ns != nullptr && threadIdx.x == 0) ? globaltimer_ns() : 0;
    allreduce_p2p<ND>(f.comm, sums);
    if (f.wait_ns != nullptr && threadIdx.x == 0) *f.wait_ns = globaltimer_ns() - t0;
  }
  if (threadIdx.x == 0) {
    Final fin = f.fin;
    fin.rec.st = &st;
    fin(sums);
    f.blk->ver(f.in ^ 1) = st;
    if (st.done) f.blk->final_() = st, f.blk->done = 1;
    if (f.bump != nullptr) f.bump->apply_seq = f.bump->apply_seq + 1; // the distributed apply in front of me is complete
    __threadfence();
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(&f.blk->ready[f.in ^ 1]), "r"(1) : "memory");
    // one GPU, profiled run: the timeline word holds how long this chain took, first instruction to flag
    if (f.wait_ns != nullptr && !(f.comm.mode == SB_COMM_P2P && f.comm.world > 1)) *f.wait_ns = globaltimer_ns() - t_start;
  }
}

// Every other CTA: wait until the new version is complete. ONE thread per CTA polls, around L1 with a plain volatile
// load and a short sleep between polls; the CTA follows through a barrier. The scalars are then read by ordinary loads
// that miss L1 (nothing of this version's lines can be there: nobody reads them before the flag) and find in L2 what
// CTA 0 wrote before its fence.
//  * An acquire load would be the textbook form, but it drags a fence behind it that waits for whatever the thread has
//    in flight (profiles/r02_stepwise_folded_v2_acquire_poll_10M.json).
//  * One poller per warp without a pause turned the first wave of CTAs into 8 000 threads hammering one L2 line: the
//    reducer's own loads and its flag store queued behind them, and every folding kernel at 10 M cells took 15-20 us
//    longer than with the one-CTA final stage in front (profiles/r02_stepwise_folded_v7_tma_staged_poll_storm_10M.json).
__device__ __forceinline__ void fold_wait(const int* ready) {
  if (threadIdx.x == 0) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(ready) : "memory");
    while (v == 0) {
      __nanosleep(100);
      asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(ready) : "memory");
    }
  }
  __syncthreads();
}

// Generic tiled element-wise kernel. Body provides
//   struct Regs; void load(int64_t e0, Regs&) const; void run(int64_t e0, int64_t n, Regs&, double (&acc)[max(ND,1)]) const;
// `done` (may be null) is the solver's device-side stop flag: once set, later launches are no-ops, so
// the iterate stays exactly what the reference returns at that iteration (DESIGN.md "stopping").
template<int ND, class Body>
__global__ void __launch_bounds__(kThreads) ew_kernel(int64_t n, Body body, RedPtrs red,
                                                      const int* __restrict__ done) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  typename Body::Regs r[kSub];
#pragma unroll
  for (int j = 0; j < kSub; ++j) body.load(lane_elem(blockIdx.x, j), r[j]);
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
#pragma unroll
  for (int j = 0; j < kSub; ++j) body.run(lane_elem(blockIdx.x, j), n, r[j], acc);
  if constexpr (ND > 0) block_reduce_partials<ND>(acc, red, blockIdx.x);
}

// masked accumulation: out-of-range elements contribute +0.0 (SB_TREE v1)
__device__ __forceinline__ void acc_pair(double& acc, int64_t e0, int64_t n, double p0, double p1) {
  acc = __dadd_rn(acc, (e0 < n) ? p0 : 0.0);
  acc = __dadd_rn(acc, (e0 + 1 < n) ? p1 : 0.0);
}

// ---- postfix expression evaluator --------------------------------------------------------------
struct RuntimeProg {
  int32_t n;
  uint8_t code[SB_EXPR_MAX_OPS];
  __device__ __forceinline__ int size() const { return n; }
  __device__ __forceinline__ int op(int k) const { return code[k]; }
  static constexpr int kMax = SB_EXPR_MAX_OPS;
};

template<uint8_t... Ops>
struct StaticProg {
  __device__ __forceinline__ constexpr int size() const { return (int) sizeof...(Ops); }
  __device__ __forceinline__ constexpr int op(int k) const {
    constexpr uint8_t code[sizeof...(Ops)] = {Ops...};
    return code[k];
  }
  static constexpr int kMax = (int) sizeof...(Ops);
};

// A 6-deep shifting register stack: no dynamic register indexing, so it never spills to local memory;
// with a StaticProg the whole loop unrolls and folds to the straight-line expression.
template<int NV, class Prog>
__device__ __forceinline__ double eval_prog(const Prog& P, const double (&in)[NV], const double (&sc)[SB_EXPR_MAX_SCAL]) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
#pragma unroll
  for (int k = 0; k < Prog::kMax; ++k) {
    if (k < P.size()) {
      const int op = P.op(k);
      if (op < SB_OP_ADD) {
        double v;
        if (op < SB_OP_SCAL0) {
          v = in[0];
          if constexpr (NV > 1) v = (op == 1) ? in[1] : v;
          if constexpr (NV > 2) v = (op == 2) ? in[2] : v;
          if constexpr (NV > 3) v = (op == 3) ? in[3] : v;
        } else {
          const int q = op - SB_OP_SCAL0;
          v = (q == 0) ? sc[0] : (q == 1) ? sc[1] : (q == 2) ? sc[2] : sc[3];
        }
        s5 = s4, s4 = s3, s3 = s2, s2 = s1, s1 = s0, s0 = v;
      } else if (op == SB_OP_NEG) {
        s0 = -s0;
      } else {
        double v;
        if (op == SB_OP_ADD) v = __dadd_rn(s1, s0);
        else if (op == SB_OP_SUB) v = __dsub_rn(s1, s0);
        else if (op == SB_OP_MUL) v = __dmul_rn(s1, s0);
        else v = __ddiv_rn(s1, s0);
        s0 = v, s1 = s2, s2 = s3, s3 = s4, s4 = s5;
      }
    }
  }
  return s0;
}

// The same stack machine for a RUN-TIME program, on both elements of a 128-bit pair at once: a rolled loop over the
// program with real branches (the opcode sits in the kernel parameters, so every branch is warp-uniform). The
// unrolled form above folds away with a StaticProg, but with a run-time program ptxas predicates all 24 steps --
// every step then issues both divisions' slow paths whatever the opcode: 760 us for a 3-vector statement at 10.1 M
// cells, where the static kernels need 50 (profiles/r02_launches_cgs_10M_runtime_program_summary.txt).
template<int NV>
__device__ __forceinline__ double2 eval_prog_pair(const RuntimeProg& P, const double2 (&in)[NV],
                                                  const double (&sc)[SB_EXPR_MAX_SCAL]) {
  double2 s0 = {0.0, 0.0}, s1 = s0, s2 = s0, s3 = s0, s4 = s0, s5 = s0;
#pragma unroll 1
  for (int k = 0; k < P.n; ++k) {
    const int op = P.code[k];
    if (op < SB_OP_ADD) {
      double2 v;
      if (op < SB_OP_SCAL0) {
        v = in[0];
        if constexpr (NV > 1) v = (op == 1) ? in[1] : v;
        if constexpr (NV > 2) v = (op == 2) ? in[2] : v;
        if constexpr (NV > 3) v = (op == 3) ? in[3] : v;
      } else {
        const int q = op - SB_OP_SCAL0;
        const double c = (q == 0) ? sc[0] : (q == 1) ? sc[1] : (q == 2) ? sc[2] : sc[3];
        v = {c, c};
      }
      s5 = s4, s4 = s3, s3 = s2, s2 = s1, s1 = s0, s0 = v;
    } else if (op == SB_OP_NEG) {
      s0 = {-s0.x, -s0.y};
    } else {
      double2 v;
      if (op == SB_OP_ADD) v = {__dadd_rn(s1.x, s0.x), __dadd_rn(s1.y, s0.y)};
      else if (op == SB_OP_SUB) v = {__dsub_rn(s1.x, s0.x), __dsub_rn(s1.y, s0.y)};
      else if (op == SB_OP_MUL) v = {__dmul_rn(s1.x, s0.x), __dmul_rn(s1.y, s0.y)};
      else v = {__ddiv_rn(s1.x, s0.x), __ddiv_rn(s1.y, s0.y)};
      s0 = v, s1 = s2, s2 = s3, s3 = s4, s4 = s5;
    }
  }
  return s0;
}

template<int AOP>
__device__ __forceinline__ double apply_assign(double y, double v) {
  if constexpr (AOP == SB_ASSIGN) return v;
  else if constexpr (AOP == SB_ADD_ASSIGN) return __dadd_rn(y, v);
  else if constexpr (AOP == SB_SUB_ASSIGN) return __dsub_rn(y, v);
  else if constexpr (AOP == SB_MUL_ASSIGN) return __dmul_rn(y, v);
  else return __ddiv_rn(y, v);
}

template<int NV, int AOP, class Prog>
struct EvalBody {
  double* y;
  const double* v[NV];
  double sc[SB_EXPR_MAX_SCAL];
  Prog prog;
  struct Regs {
    double2 in[NV];
    double2 y;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const {
#pragma unroll
    for (int k = 0; k < NV; ++k) r.in[k] = ld2(v[k], e0);
    if constexpr (AOP != SB_ASSIGN) r.y = ld2(y, e0);
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs& r, double (&)[1]) const {
    double2 out;
    if constexpr (std::is_same_v<Prog, RuntimeProg>) {
      const double2 v = eval_prog_pair<NV>(prog, r.in, sc);
      out.x = apply_assign<AOP>(r.y.x, v.x);
      out.y = apply_assign<AOP>(r.y.y, v.y);
    } else {
      double a[NV], b[NV];
#pragma unroll
      for (int k = 0; k < NV; ++k) a[k] = r.in[k].x, b[k] = r.in[k].y;
      out.x = apply_assign<AOP>(r.y.x, eval_prog<NV>(prog, a, sc));
      out.y = apply_assign<AOP>(r.y.y, eval_prog<NV>(prog, b, sc));
    }
    st2(y, e0, out);
  }
};

// ---- stand-alone dot products ------------------------------------------------------------------
template<int M>
struct DotBody {
  const double* a[M];
  const double* b[M];
  struct Regs {
    double2 a[M], b[M];
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const {
#pragma unroll
    for (int k = 0; k < M; ++k) r.a[k] = ld2(a[k], e0), r.b[k] = ld2(b[k], e0);
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, Regs& r, double (&acc)[M]) const {
#pragma unroll
    for (int k = 0; k < M; ++k)
      acc_pair(acc[k], e0, n, __dmul_rn(r.a[k].x, r.b[k].x), __dmul_rn(r.a[k].y, r.b[k].y));
  }
};

#include "sb_group_body.cuh" // struct GroupBody (sb_eval_group); a file of its own so that a host harness can compile it

struct FillBody {
  double* y;
  double value;
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t e0, int64_t, Regs&, double (&)[1]) const {
    st2(y, e0, make_double2(value, value));
  }
};

} // namespace sb
