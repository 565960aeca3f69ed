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
    red.partials[(int64_t) d * red.cap_tiles + tile] = s;
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

// The one-CTA kernel behind every reducing kernel of the one-kernel-per-step schedule. `fin(sums)` runs on one
// thread: that is where the solver scalars (alpha, beta, residual, stop flag) are updated.
template<int ND, class Final>
__global__ void __launch_bounds__(kThreads) final_reduce_kernel(int64_t n_tiles, RedPtrs red, Final fin, CommDev comm,
                                                                CommCtrl* bump, const int* __restrict__ done) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  __shared__ double s_w[kMaxDots][kWarps];
  double sums[ND];
  final_stage<ND>(n_tiles, red, s_w, sums);
  // multi-GPU, P2P mode: exchange the rank sums with every peer inside this kernel (rank-ordered total)
  if (comm.mode == SB_COMM_P2P && comm.world > 1) allreduce_p2p<ND>(comm, sums);
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
inline int launch_final(sb_ctx* ctx, int64_t n, const Final& fin, const int* done, CommCtrl* bump = nullptr) {
  const RedPtrs red{ctx->red.partials, ctx->red.cap_tiles};
  if (ctx->debug & 4) { // experiment: rank-local sums only
    SB_CUDA(launch_kernel(ctx, final_reduce_kernel<ND, Final>, 1, kThreads, 0, num_tiles(n), red, fin, CommDev{}, bump, done));
    ctx->launches++;
    return SB_OK;
  }
  if (ctx->comm.mode == SB_COMM_NCCL && ctx->comm.world > 1) {
    SB_CUDA(launch_kernel(ctx, final_reduce_kernel<ND, StoreFinal<ND>>, 1, kThreads, 0, num_tiles(n), red,
                          StoreFinal<ND>{ctx->d_ar}, CommDev{}, (CommCtrl*) nullptr, done));
    SB_TRY(nccl_allreduce_sum(ctx, ctx->d_ar, ND));
    SB_CUDA(launch_kernel(ctx, scalar_final_kernel<ND, Final>, 1, 1, 0, (const double*) ctx->d_ar, fin, done));
    ctx->launches += 2;
    return SB_OK;
  }
  SB_CUDA(launch_kernel(ctx, final_reduce_kernel<ND, Final>, 1, kThreads, 0, num_tiles(n), red, fin, ctx->comm, bump, done));
  ctx->launches++;
  return SB_OK;
}

struct NoFinal {
  __device__ void operator()(const double*) const {}
};

// ---- folding a reduction into the kernel that consumes it ----------------------------------------------------------
// The one-CTA final stage costs a launch per reduction (~2.7 us of boundary + ~3 us of kernel inside a replayed graph:
// a third of a BiCGStab iteration at 1.26 M cells per rank, profiles/r02_stepwise_unfolded_1M.json). A kernel that
// needs the reduced scalars can finish the reduction itself instead: CTA 0 runs SB_TREE's final stage over the
// producer's tile partials and posts the sums into the all-reduce mailbox of every rank (its own included; the single
// mailbox of a one-GPU context is local); EVERY CTA polls its rank's mailbox, adds the rank sums in rank order and
// runs the solver's scalar update on a copy of the solver state in shared memory -- identical inputs, identical code,
// identical bits in all CTAs of all ranks. The copy comes from state version `in`; CTA 0 writes the updated state to
// version `out` (never the version the other CTAs are reading), publishes the stop (`final_`, `done`) and bumps the
// apply sequence number when the producer was a distributed apply. The mailbox of fold #a is emptied by CTA 0 of the
// kernel that handles fold #a+1 (every CTA of the earlier kernel has finished by then), before it posts #a+1.
struct FoldShared {
  SolverState st;
  double s_fin[kMaxDots][kWarps];
  double s_all[kMaxRanks][4];
  double s_local[4];
};

template<int ND, class Final>
struct Fold {
  int64_t n_tiles = -1;             // tiles of the producer; < 0: nothing to fold (the kernel reads version `in` as it is)
  RedPtrs red{};
  Final fin{};                      // fin.rec: history / trace buffers (CTA 0 records)
  SolveBlock* blk = nullptr;
  int32_t in = 0;                   // state version to start from; the result goes to in ^ 1
  CommDev comm{};                   // world <= 1: one GPU
  unsigned long long* box = nullptr;             // one GPU: mailbox [2][4]
  const unsigned long long* ar_base = nullptr;   // all-reduces completed before this solve's folds (mailbox parity)
  CommCtrl* bump = nullptr;         // the producer was a distributed apply: count it
  unsigned long long* wait_ns = nullptr; // optional: longest wait of CTA 0 for the other ranks' sums (atomicMax)
};

__device__ __forceinline__ unsigned long long* fold_box(const CommDev& comm, unsigned long long* box, int rank, unsigned long long par,
                                                        int src, int d) {
  if (comm.world > 1) return &comm.ctrl(rank)->ar_slot[par][src][d];
  return box + par * 4 + d;
}

// All threads of a folding kernel. On return sh.st is the updated state (valid for every thread).
template<int ND, class Final>
__device__ __forceinline__ void fold_prologue(const Fold<ND, Final>& f, FoldShared& sh) {
  const int world = f.comm.world > 1 ? f.comm.world : 1, me = f.comm.world > 1 ? f.comm.rank : 0;
  const SolverState* in = &f.blk->ver[f.in];
  const unsigned long long par = (*f.ar_base + (unsigned long long) in->folds) & 1ull;
  if (threadIdx.x == 0) sh.st = *in;
  if (blockIdx.x == 0) {
    if (threadIdx.x < world * 4) { // empty the mailbox of the previous fold BEFORE my sums go out
      const int r = threadIdx.x >> 2, d = threadIdx.x & 3;
      st_relaxed_sys(fold_box(f.comm, f.box, me, par ^ 1ull, r, d), kArSentinel);
    }
    double sums[ND];
    final_stage<ND>(f.n_tiles, f.red, sh.s_fin, sums);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int d = 0; d < ND; ++d) sh.s_local[d] = sums[d];
      if (world > 1) __threadfence_system(); // the reset above is performed before any peer can answer what goes out below
      else __threadfence();
      if (f.bump != nullptr) f.bump->apply_seq = f.bump->apply_seq + 1; // the distributed apply in front of me is complete
    }
    __syncthreads();
    if (threadIdx.x < world * ND) {
      const int r = threadIdx.x / ND, d = threadIdx.x % ND;
      st_relaxed_sys(fold_box(f.comm, f.box, r, par, me, d), (unsigned long long) __double_as_longlong(sh.s_local[d]));
    }
  }
  if (threadIdx.x < world * ND) {
    const int r = threadIdx.x / ND, d = threadIdx.x % ND;
    const unsigned long long* box = fold_box(f.comm, f.box, me, par, r, d);
    unsigned long long v = ld_relaxed_sys(box);
    if (v == kArSentinel) {
      CommCtrl* ctl = f.comm.world > 1 ? f.comm.ctrl(me) : nullptr;
      const unsigned long long t0 = globaltimer_ns(), limit = f.comm.world > 1 ? f.comm.timeout_ns : kDefaultSpinTimeoutNs;
      unsigned spins = 0;
      while ((v = ld_relaxed_sys(box)) == kArSentinel) {
        if ((++spins & 63u) == 0 && (globaltimer_ns() - t0 > limit || (ctl != nullptr && comm_failed(ctl)))) {
          if (ctl != nullptr) comm_fail(ctl, 0xC000 + r);
          break;
        }
      }
      if (f.wait_ns != nullptr && blockIdx.x == 0) atomicMax(f.wait_ns, globaltimer_ns() - t0);
    }
    sh.s_all[r][d] = __longlong_as_double((long long) v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double t = sh.s_all[0][d];
      for (int r = 1; r < world; ++r) t = __dadd_rn(t, sh.s_all[r][d]);
      tot[d] = t;
    }
    Final fin = f.fin;
    fin.rec.st = &sh.st;
    if (blockIdx.x != 0) fin.rec.hist = nullptr, fin.rec.trace = nullptr;
    fin(tot);
    sh.st.folds++;
    if (blockIdx.x == 0) {
      f.blk->ver[f.in ^ 1] = sh.st;
      if (sh.st.done) {
        f.blk->final_ = sh.st;
        f.blk->done = 1;
      }
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

// The same kernel in front of which a reduction is folded (fold_prologue): the tile's loads are issued first, so the
// final stage, the mailbox round trip and the scalar update run under their latency; the body then reads the solver
// scalars from the CTA's own copy of the state (Body::st is redirected to it). One extra CTA beyond the tiles is never
// needed: a fold without tiles (the flush at the end of a solve) is launched with n = 0 and one CTA.
template<int ND, class Body, int FND, class Final>
__global__ void __launch_bounds__(kThreads) ew_fold_kernel(int64_t n, Body body, RedPtrs red, const int* __restrict__ done,
                                                           Fold<FND, Final> fold) {
  __shared__ FoldShared sh;
  if (is_done(done)) return;
  const bool tile = (int64_t) blockIdx.x * kTile < n;
  typename Body::Regs r[kSub];
  if (tile) {
#pragma unroll
    for (int j = 0; j < kSub; ++j) body.load(lane_elem(blockIdx.x, j), r[j]);
  }
  if (fold.n_tiles >= 0) {
    fold_prologue(fold, sh);
  } else {
    if (threadIdx.x == 0) sh.st = fold.blk->ver[fold.in];
    __syncthreads();
  }
  if (sh.st.done || !tile) return; // the stopping rule has just fired: the iterate stays what it is
  body.st = &sh.st;
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
    double a[NV], b[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) a[k] = r.in[k].x, b[k] = r.in[k].y;
    double2 out;
    out.x = apply_assign<AOP>(r.y.x, eval_prog<NV>(prog, a, sc));
    out.y = apply_assign<AOP>(r.y.y, eval_prog<NV>(prog, b, sc));
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
