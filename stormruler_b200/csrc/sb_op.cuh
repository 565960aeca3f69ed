// sb_op.cuh -- device layout of the FVM operator and the operator-apply kernels.
//
// Reference semantics: the face loop of source_apps/playground/Playground.cpp:115-131 (interior
// faces) plus the boundary-ghost pattern of Feathers/ConvectionScheme.hpp:95-106, restated as a
// cell-row gather so no atomics are needed and every cell sums its face contributions in ascending
// face index, the order the CPU loop produces.
//
// HBM layout (column-major ELL, leading dimension ld = n padded to the 2048-row CTA tile):
//   col  [W][ld] int32   neighbour cell | ~i (Dirichlet mirror ghost of cell i) | kColPad
//   v0   [W][ld] fp64    coef form: a = ((area/vol_i)*dt)/dist ; faithful form: g = area/vol_i
//   v1   [W][ld] fp64    faithful form only: dist
//   diag [ld]    fp64    coef form only: prefill - sum(a) - sum_ghost(a+a)
// A lane owns two consecutive rows, so entry k of both rows is one int2 + one double2 load and a warp
// reads 256 B + 512 B contiguous per k: fully coalesced 64/128-bit streams. Algorithmic bytes per
// apply (coef form): 24*N + 12*entries = 72 B/cell on tets.
#pragma once

#include "sb_kernels.cuh"

#include <atomic>
#include <cstdlib>

namespace sb {

#include "sb_apply_rows.cuh" // OpDev, gather(), apply_rows<FORM, W>: a file of its own so that a host harness can compile it

// Multi-GPU (P2P mode): what the apply kernel needs for the fused halo exchange. n_pack == 0: none.
struct ApplyDist {
  CommDev comm;
  HaloDev halo;
  int64_t x_off = 0;           // byte offset of x inside the slab (the same on every rank)
  int32_t n_pack = 0;          // the first n_pack CTAs of the grid pack + push my boundary values
  int32_t coherent_gather = 0; // 1: BOUNDARY tiles gather x with ld.global.ca instead of the read-only path
                               // (their halo tail is written by peers while the kernel runs); interior tiles
                               // never touch the halo and keep ld.global.nc
  int32_t no_ack = 0;          // 1: skip the ack round before the halo push. Valid inside the fused solvers: between two
                               // applies that write the same halo tail there is always an all-reduce, and a rank
                               // contributes to it only after its earlier apply has completed
  int32_t pre_pushed = 0;      // 1: the kernel that produced x has already pushed the boundary values (halo_push_tile):
                               // no pack CTAs (n_pack == 0), the boundary tiles only acquire the flags
  int32_t post_flags = 0;      // 1 (with pre_pushed): the push was lazy -- this kernel's first CTA raises the flags
  unsigned long long* wait_ns = nullptr; // optional: longest halo-flag wait of any boundary CTA (atomicMax)
};

// Boundary tiles (they read the halo tail) wait until every neighbour's values of THIS apply have landed.
__device__ __forceinline__ void apply_halo_wait(const ApplyDist& ad, int64_t tile) {
  if ((ad.n_pack > 0 || ad.pre_pushed) && tile >= ad.halo.first_boundary_tile) {
    CommCtrl* me = ad.comm.ctrl(ad.comm.rank);
    if (threadIdx.x < ad.halo.n_nbr) {
      const unsigned long long t0 = ad.wait_ns != nullptr ? globaltimer_ns() : 0;
      wait_flag_ge(&me->halo_flag[ad.halo.nbr_rank[threadIdx.x]], ld_acquire_sys(&me->apply_seq) + 1, me,
                   0xB000 + ad.halo.nbr_rank[threadIdx.x], ad.comm.timeout_ns);
      if (ad.wait_ns != nullptr) atomicMax(ad.wait_ns, globaltimer_ns() - t0);
    }
    __syncthreads();
  }
}

// Epilogues fuse dot products into the apply: they see the input pair x[e0..e0+1] and the freshly
// computed output pair, so <x,Ax>-style reductions cost no extra vector pass.
struct NoEpi {
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t, int64_t, double2, double2, Regs&, double (&)[1]) const {}
};

// acc[0] += x.y
struct EpiXY {
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t e0, int64_t n, double2 x, double2 y, Regs&, double (&acc)[1]) const {
    acc_pair(acc[0], e0, n, __dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y));
  }
};

// acc[0] += u.y  (u: a third vector, e.g. r~ in BiCGStab)
struct EpiUY {
  const double* u;
  struct Regs {
    double2 u;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.u = ld2(u, e0); }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, double2, double2 y, Regs& r, double (&acc)[1]) const {
    acc_pair(acc[0], e0, n, __dmul_rn(r.u.x, y.x), __dmul_rn(r.u.y, y.y));
  }
};

// acc[0] += y.y ; acc[1] += y.x   (BiCGStab: <t,t>, <t,r>)
struct EpiYYandYX {
  struct Regs {};
  __device__ __forceinline__ void load(int64_t, Regs&) const {}
  __device__ __forceinline__ void run(int64_t e0, int64_t n, double2 x, double2 y, Regs&, double (&acc)[2]) const {
    acc_pair(acc[0], e0, n, __dmul_rn(y.x, y.x), __dmul_rn(y.y, y.y));
    acc_pair(acc[1], e0, n, __dmul_rn(y.x, x.x), __dmul_rn(y.y, x.y));
  }
};

// r = b - A x fused into the apply (Operator::Residual, Operator.hpp:95-99), with acc[0] += r.r.
// The kernel stores the value returned through `out`.
struct EpiResidual {
  const double* b;
  struct Regs {
    double2 b;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const { r.b = ld2(b, e0); }
};

// y <- A x with a fused reduction epilogue. RESID: store b - A x instead (and reduce <r,r>).
template<int FORM, int W, int ND, bool RESID, class Epi>
__global__ void __launch_bounds__(kThreads) apply_kernel(OpDev op, const double* __restrict__ x, double* __restrict__ y,
                                                         Epi epi, RedPtrs red, ApplyDist ad,
                                                         const int* __restrict__ done) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  if (ad.n_pack > 0 && (int) blockIdx.x < ad.n_pack) {
    halo_pack_role(ad.comm, ad.halo, x, ad.x_off, ad.n_pack, ad.no_ack != 0);
    return;
  }
  const int64_t tile = (int64_t) blockIdx.x - ad.n_pack;
  const int coh = ad.coherent_gather && tile >= ad.halo.first_boundary_tile;
  if (ad.post_flags && blockIdx.x == 0) halo_post_flags(ad.comm, ad.halo); // lazy push: the producer has completed
  apply_halo_wait(ad, tile);
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
#pragma unroll 1
  for (int j = 0; j < kSub; ++j) {
    const int64_t e0 = lane_elem(tile, j);
    const double2 xo = ld2(x, e0);
    typename Epi::Regs er;
    epi.load(e0, er);
    double2 yo = make_double2(0.0, 0.0);
    if constexpr (FORM == SB_FORM_FAITHFUL) {
      if (op.prefill == 2) yo = ld2(y, e0); // read by the lane that overwrites it below
    }
    double2 out = apply_rows<FORM, W>(op, x, e0, xo, yo, coh);
    if constexpr (RESID) {
      out.x = __dsub_rn(er.b.x, out.x);
      out.y = __dsub_rn(er.b.y, out.y);
      acc_pair(acc[0], e0, op.n, __dmul_rn(out.x, out.x), __dmul_rn(out.y, out.y));
    }
    st2(y, e0, out);
    if constexpr (!RESID) epi.run(e0, op.n, xo, out, er, acc);
  }
  if constexpr (ND > 0) block_reduce_partials<ND>(acc, red, tile);
}


// ---- apply kernel v2: warp-private TMA (bulk-copy) staging of the streamed operands ----------------
// ncu on v1 (profiles/r01_apply_v1_ncu_details.txt): DRAM traffic equals the algorithmic bytes but only
// ~60 % of peak bandwidth is reached; warps sit on long-scoreboard stalls because the column indices
// must arrive before the dependent gathers can even be issued. Here the streamed operands of a warp's
// 64-row stage (W index slices, W coefficient slices, diagonal, own x, optional epilogue vector: all
// contiguous runs) are fetched by cp.async.bulk into a per-warp shared-memory ring, completion tracked
// by a per-warp mbarrier. The operator itself is stored blocked (one contiguous record per 64-row
// slice, see OpDev::blk), so a stage is 2-3 bulk copies issued by one elected lane and the kernel
// reads 3-4 sequential HBM streams instead of 12 (measured: more concurrent streams = lower DRAM
// efficiency). No registers are held by loads in flight, so two stages per warp and three
// CTAs per SM keep ~200 KB in flight per SM, and the L1 is left entirely to the gathered x.
// The row->lane mapping, operation order and reduction tree are identical to v1 (bit-identical output).
constexpr int kStages = 2;

template<int W>
struct StageLayout {
  static constexpr int col = 0;              // W slices of 64 int32   } one slice record of the
  static constexpr int coef = W * 256;       // W slices of 64 fp64    } blocked operator layout,
  static constexpr int diag = coef + W * 512; //                        } fetched by ONE bulk copy
  static constexpr int slice = diag + 512;
  static constexpr int xown = slice;
  static constexpr int bytes = xown + 512;
  static constexpr int cta_bytes = bytes * kStages * kWarps;
  // resident CTAs per SM as the ring allows (227 KB per SM, ~1 KB static + 1 KB reserved per CTA), at most 3: told to
  // ptxas through __launch_bounds__ so that it uses the registers that occupancy leaves free anyway (85 at three CTAs)
  // instead of spilling down to 64
  static constexpr int fit = (227 * 1024) / (cta_bytes + 2048);
  // (64 against 72 registers at W = 4 -- __launch_bounds__(256, 4) against (256, 3) -- measured equal: 2 233-2 238 against
  // 2 252-2 260 it/s, profiles/r02_ab_apply_registers.txt)
  static constexpr int min_ctas = fit >= 3 ? 3 : (fit >= 1 ? fit : 1);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// The same with an L2 evict-first policy: data that is read once per kernel (the operator's slice records) should not
// push the vectors out of L2 when a rank is small enough for them to live there.
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap(); // a lost bulk copy must fail loudly, never hang the device
}

template<int W, int ND, bool RESID, class Epi>
__global__ void __launch_bounds__(kThreads, StageLayout<W>::min_ctas) apply_kernel_tma(OpDev op, const double* __restrict__ x,
                                                            double* __restrict__ y, Epi epi, RedPtrs red,
                                                            ApplyDist ad, const int* __restrict__ done) {
  using L = StageLayout<W>;
  extern __shared__ __align__(128) unsigned char sb_smem[];
  __shared__ __align__(8) uint64_t bars[kWarps][kStages];
  pdl_trigger();
  // (No in-kernel reducer role here, unlike ew_solver_kernel: merely CARRYING the out-of-line call -- a stack frame and
  // 640 B more static shared memory, never executed without SB_TUNE_IN_KERNEL_REDUCER -- cost this kernel 6 % at 10.1 M
  // cells: 153.9 against 143.8 us per apply + dot slot, BiCGStab 2 198 against 2 323 it/s on the same box,
  // profiles/r02_ab_apply_without_reducer_call.txt. An apply's reduction always ends in the one-CTA final stage.)
  if (ad.n_pack > 0 && (int) blockIdx.x < ad.n_pack) { // halo-pack CTAs: scheduled first, overlap the interior tiles
    pdl_wait();
    if (!is_done(done)) halo_pack_role(ad.comm, ad.halo, x, ad.x_off, ad.n_pack, ad.no_ack != 0);
    return;
  }
  const int64_t tile = (int64_t) blockIdx.x - ad.n_pack;
  const int coh = ad.coherent_gather && tile >= ad.halo.first_boundary_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = sb_smem + (size_t) warp * kStages * L::bytes;
  const int64_t row0 = tile * kTile + warp * (kTile / kWarps);

  // A stage = the operator's slice record (never written by a kernel: may be fetched BEFORE
  // griddepcontrol.wait, i.e. while the previous kernel is still draining) + this warp's own run of x,
  // which previous kernels produce (fetched after the wait).
  // `dep` (always 0 at run time) is added to the source addresses: see the hazard note in the main loop.
  auto issue_op = [&](int j, uint32_t dep) {
    const int s = j % kStages;
    uint64_t* bar = &bars[warp][s];
    mbar_expect_tx(bar, (uint32_t) L::bytes);
    const unsigned char* src = op.blk + ((row0 + j * 64) >> 6) * (int64_t) L::slice + dep;
    if (op.stream_hint) bulk_g2s_stream(wbase + s * L::bytes, src, L::slice, bar);
    else bulk_g2s(wbase + s * L::bytes, src, L::slice, bar);
  };
  auto issue_vec = [&](int j, uint32_t dep) {
    const int s = j % kStages;
    unsigned char* dst = wbase + s * L::bytes;
    uint64_t* bar = &bars[warp][s];
    const int64_t r = row0 + j * 64;
    bulk_g2s(dst + L::xown, reinterpret_cast<const unsigned char*>(x + r) + dep, 512, bar);
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int j = 0; j < kStages; ++j) issue_op(j, 0u);
  }
  pdl_wait();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kStages; ++j) issue_vec(j, 0u);
  }
  __syncwarp();
  if (is_done(done)) { // drain the copies in flight: a CTA must not exit with bulk copies landing in its smem
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_wait(&bars[warp][s], 0);
    return;
  }
  // The epilogue's third vector (r~ or b) is read with plain coalesced loads, all four sub-iterations up
  // front so they are in flight under the pipeline: measured 10 % faster than staging it as a third bulk
  // copy per stage (apply + <r~,v> at 10.1 M cells: 138 us against 152 us), and 512 B less ring per stage.
  if (ad.post_flags && blockIdx.x == 0) halo_post_flags(ad.comm, ad.halo); // lazy push: the producer has completed
  typename Epi::Regs er[kSub];
#pragma unroll
  for (int j = 0; j < kSub; ++j) epi.load(row0 + j * 64 + 2 * lane, er[j]);
  apply_halo_wait(ad, tile); // the streamed operands are already in flight while boundary tiles wait

  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;

#pragma unroll
  for (int j = 0; j < kSub; ++j) {
    const int s = j % kStages;
    mbar_wait(&bars[warp][s], (uint32_t) ((j / kStages) & 1));
    const unsigned char* src = wbase + s * L::bytes;
    int2 c[W];
    double2 a[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      c[k] = reinterpret_cast<const int2*>(src + L::col + k * 256)[lane];
      a[k] = reinterpret_cast<const double2*>(src + L::coef + k * 512)[lane];
    }
    const double2 dg = reinterpret_cast<const double2*>(src + L::diag)[lane];
    const double2 xo = reinterpret_cast<const double2*>(src + L::xown)[lane];
    // WAR hazard between the generic proxy and the async proxy: the ld.shared above have been ISSUED, but
    // their data may not have left shared memory yet (under DRAM-bound gather traffic the load/store unit
    // queues back up), and the bulk copy issued below overwrites this very slot. So every loaded register is
    // folded into one word and the refill's source addresses are made to depend on it (`& op.zero`: always
    // 0, but the compiler cannot know): the copy cannot issue before the fold, the fold cannot issue before
    // every load of the warp has returned. (Without this ~3 of 90 000 stages per launch read the NEXT
    // stage's bytes at 6 M cells -- none at 1 M, where everything sits in L2; tests/test_gpu_scale.py.)
    uint32_t fold = 0;
    {
      auto mix = [&](double v) {
        const long long b = __double_as_longlong(v);
        fold ^= (uint32_t) b ^ (uint32_t) (b >> 32);
      };
#pragma unroll
      for (int k = 0; k < W; ++k) {
        fold ^= (uint32_t) c[k].x ^ (uint32_t) c[k].y;
        mix(a[k].x), mix(a[k].y);
      }
      mix(dg.x), mix(dg.y), mix(xo.x), mix(xo.y);
    }
    const uint32_t dep = fold & (uint32_t) op.zero;
    __syncwarp(); // every lane has copied its slice out of the ring slot
    if (lane == 0 && j + kStages < kSub) issue_op(j + kStages, dep), issue_vec(j + kStages, dep);

    double g0[W], g1[W];
    if (op.debug & 1) {
#pragma unroll
      for (int k = 0; k < W; ++k) g0[k] = xo.x, g1[k] = xo.y;
    } else {
#pragma unroll
      for (int k = 0; k < W; ++k) {
        g0[k] = (c[k].x >= 0) ? gather(x + c[k].x, coh) : 0.0;
        g1[k] = (c[k].y >= 0) ? gather(x + c[k].y, coh) : 0.0;
      }
    }
    double u0 = __dmul_rn(dg.x, xo.x), u1 = __dmul_rn(dg.y, xo.y);
#pragma unroll
    for (int k = 0; k < W; ++k) {
      const double t0 = __dadd_rn(u0, __dmul_rn(a[k].x, g0[k]));
      const double t1 = __dadd_rn(u1, __dmul_rn(a[k].y, g1[k]));
      u0 = (c[k].x >= 0) ? t0 : u0;
      u1 = (c[k].y >= 0) ? t1 : u1;
    }
    double2 out = make_double2(u0, u1);
    const int64_t e0 = row0 + j * 64 + 2 * lane;
    if constexpr (RESID) {
      out.x = __dsub_rn(er[j].b.x, out.x);
      out.y = __dsub_rn(er[j].b.y, out.y);
      acc_pair(acc[0], e0, op.n, __dmul_rn(out.x, out.x), __dmul_rn(out.y, out.y));
    }
    st2(y, e0, out);
    if constexpr (!RESID) epi.run(e0, op.n, xo, out, er[j], acc);
  }
  if constexpr (ND > 0) block_reduce_partials<ND>(acc, red, tile);
}

// ---- element-wise kernel with a folded reduction in front, bulk-copy staged ---------------------------------------------
// (sb_kernels.cuh: Fold, fold_reduce, fold_wait.) CTA 0 is a dedicated reducer and owns no tile; CTA k > 0 owns tile k - 1.
// A tile's inputs -- the 512-byte runs of the body's NV input vectors for the four sub-iterations of each warp -- are
// fetched by bulk copies (cp.async.bulk) into shared memory, all sixteen NV per warp issued by one lane before
// anything else; THEN the CTA waits for the ready flag (one thread polls, the CTA follows: no fence, no call), reads the solver
// scalars from the new state version (Body::st is redirected to it) and consumes the stages. Nothing is held in
// registers across the wait, the flag's round trip and the scalar loads run under the latency of the copies, and the
// bytes in flight per SM are set by shared memory (NV x 16 KB per CTA: 7 / 4 / 3 / 2 resident CTAs for 2 / 3 / 4 / 5
// inputs = 160-224 KB per SM), not by the register file. What was tried before
// (profiles/r02_stepwise_folded_v{1,2,4,5,6}_*.json; half update at 10 M cells, 37 us with the one-CTA final stage in
// front): every CTA running the scalar update behind a mailbox (50 us); CTA 0 = reducer AND tile owner, inlined (its
// registers became everybody's: 85 instead of 32, 50 us) or as a call (the tile's registers spill around it: 104 us);
// the fold in front of the loads (two dependent L2 round trips per CTA before its first load: 63 us); register-staged
// loads in front of the wait (32 data registers live across it: five resident CTAs instead of eight, 50 us).
// A fold without tiles (the flush at the end of a BiCGStab solve) is launched with n = 0: CTA 0 alone.
template<class Body>
struct EwStage {
  static constexpr int stage = Body::NV * 512;     // one sub-iteration of one warp
  static constexpr int warp = stage * kSub;
  static constexpr int cta = warp * kWarps;        // = NV x 16 KB
  static constexpr int min_ctas = Body::NV <= 2 ? 7 : (Body::NV == 3 ? 4 : (Body::NV == 4 ? 3 : 2));
};

template<int ND, class Body, int FND, class Final>
__global__ void __launch_bounds__(kThreads, EwStage<Body>::min_ctas) ew_fold_kernel(int64_t n, Body body, RedPtrs red, const int* __restrict__ done,
                                                                                   const __grid_constant__ Fold<FND, Final> fold) {
  // (__grid_constant__: CTA 0 passes `fold` to fold_reduce by reference; without it EVERY thread of every CTA copied
  // the 200-byte parameter to its local stack on entry -- as much store traffic as the half update itself produces)
  using E = EwStage<Body>;
  constexpr int NV = Body::NV;
  extern __shared__ __align__(128) unsigned char sb_smem[];
  __shared__ __align__(8) uint64_t bars[kWarps][kSub];
  if (is_done(done)) return;
  const bool active = fold.n_tiles >= 0;
  if (blockIdx.x == 0) {
    if (active) fold_reduce(fold, sb_smem, E::cta);
    return;
  }
  const int64_t tile = (int64_t) blockIdx.x - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = sb_smem + (size_t) warp * E::warp;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kSub; ++j) mbar_init(&bars[warp][j], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
    for (int j = 0; j < kSub; ++j) {
      const int64_t r0 = tile * kTile + warp * (kTile / kWarps) + j * 64;
      mbar_expect_tx(&bars[warp][j], (uint32_t) E::stage);
#pragma unroll
      for (int k = 0; k < NV; ++k) bulk_g2s(wbase + j * E::stage + k * 512, body.in(k) + r0, 512, &bars[warp][j]);
    }
  }
  __syncwarp(); // the mbarriers lane 0 initialised are waited on by the whole warp
  const SolverState* st = &fold.blk->ver(active ? (fold.in ^ 1) : fold.in);
  if (active) fold_wait(&fold.blk->ready[fold.in ^ 1]);
  const bool stopped = __ldcg(&st->done) != 0; // the stopping rule has just fired: the iterate stays what it is
  body.st = st;
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
#pragma unroll
  for (int j = 0; j < kSub; ++j) {
    mbar_wait(&bars[warp][j], 0); // also when stopped: a CTA must not exit with bulk copies landing in its smem
    if (stopped) continue;
    double2 in[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) in[k] = reinterpret_cast<const double2*>(wbase + j * E::stage + k * 512)[lane];
    typename Body::Regs g;
    body.fill(g, in);
    body.run(lane_elem(tile, j), n, g, acc);
  }
  if (stopped) return;
  if constexpr (ND > 0) block_reduce_partials<ND>(acc, red, tile);
}

// ---- element-wise kernel of the fused solvers' stepwise schedule -------------------------------------------------------
// ew_kernel with two optional roles:
//  * push (SB_TUNE_PUSH_ON_PRODUCE, pa.push != 0): the tiles are taken in REVERSE order (the boundary block is the tail
//    of the local order, so its tiles become the first CTAs of the grid), and a CTA that owns a boundary tile ends with
//    halo_push_tile (sb_comm.cuh): the part of `y` -- the vector the body writes, the input of the next apply -- that
//    the neighbours need is on its way over NVLink before the interior tiles of this kernel have even started;
//  * in-kernel reducer (ra.kind != kFinalNone): the last CTA of the grid finishes the kernel's reductions (sb_finals.cuh).
struct PushArgs {
  CommDev comm;
  HaloDev halo;
  const double* y = nullptr; // the produced vector
  int64_t y_off = 0;         // its byte offset inside the slab (the same on every rank)
  int64_t n_tiles = 0;
  int32_t push = 0;
  int32_t lazy = 0;          // 1: stores only; the consuming apply raises the flags (halo_post_flags)
};

// RED: the instantiation that carries the reducer role. The default path (RED = false) does not even contain the
// out-of-line call: carrying it cost the apply kernel 6 % (see apply_kernel_tma).
template<int ND, class Body, bool RED>
__global__ void __launch_bounds__(kThreads) ew_solver_kernel(int64_t n, Body body, RedPtrs red, const int* __restrict__ done,
                                                             const __grid_constant__ PushArgs pa,
                                                             const __grid_constant__ ReducerArgs ra) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  if constexpr (RED) {
    if (ra.kind != kFinalNone && blockIdx.x == gridDim.x - 1) {
      reducer_role(ra);
      return;
    }
  }
  const int64_t tile = pa.push ? pa.n_tiles - 1 - (int64_t) blockIdx.x : (int64_t) blockIdx.x;
  typename Body::Regs r[kSub];
#pragma unroll
  for (int j = 0; j < kSub; ++j) body.load(lane_elem(tile, j), r[j]);
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
#pragma unroll
  for (int j = 0; j < kSub; ++j) body.run(lane_elem(tile, j), n, r[j], acc);
  if constexpr (ND > 0) block_reduce_partials<ND>(acc, red, tile);
  if (pa.push && tile >= pa.halo.first_boundary_tile) halo_push_tile(pa.comm, pa.halo, pa.y, pa.y_off, tile, pa.lazy != 0);
}

// y = x / diag (Jacobi). The diagonal lives inside the blocked slice records (or in OpDev::diag for the v1 layout).
struct JacobiBody {
  OpDev op;
  const double* x;
  double* y;
  struct Regs {
    double2 x, d;
  };
  __device__ __forceinline__ void load(int64_t e0, Regs& r) const {
    r.x = ld2(x, e0);
    if (op.blk != nullptr) {
      const unsigned char* rec = op.blk + (e0 >> 6) * (int64_t) op.slice_bytes + (int64_t) op.width * 768;
      r.d = *reinterpret_cast<const double2*>(rec + (e0 & 63) * 8);
    } else {
      r.d = ld2(op.diag, e0);
    }
  }
  __device__ __forceinline__ void run(int64_t e0, int64_t n, Regs& r, double (&)[1]) const {
    // padding rows have a zero diagonal: keep them 0 instead of producing inf/NaN
    double2 out;
    out.x = (e0 < n) ? __ddiv_rn(r.x.x, r.d.x) : 0.0;
    out.y = (e0 + 1 < n) ? __ddiv_rn(r.x.y, r.d.y) : 0.0;
    st2(y, e0, out);
  }
};

} // namespace sb

struct sb_op {
  sb::OpDev d;
  int64_t n_entries = 0;
  int64_t device_bytes = 0;
  void* buffers[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // distributed operator (sb_dist_op_create): halo plan; halo.n_nbr == 0 otherwise
  sb::HaloDev halo;
  int64_t halo_base = 0, n_halo = 0;
  int64_t recv_ptr[sb::kMaxRanks + 1] = {};
  int32_t* d_send_idx = nullptr;
  int32_t* d_push_ptr = nullptr; // push-on-produce plan (HaloDev::push_ptr / push_entry); null: not available
  int2* d_push_entry = nullptr;
  bool distributed = false;
};

namespace sb {

int halo_exchange(sb_ctx* ctx, const sb_op* op, const double* x, const int* done, int64_t* x_off); // sb_comm.cu

// Launch y <- A x (+ epilogue) on the context's stream, dispatching on form and ELL width.
struct ApplyOpts {
  const OpDev* per_call = nullptr; // replaces the operator's kernel arguments for this launch (sb_apply_accumulate: other dt/prefill)
  bool fold_later = false;         // the reduction (and the count of this apply) is folded into the kernel that consumes
                                   // it (sb_kernels.cuh: Fold) instead of a one-CTA final stage behind this launch
  int halo_mode = 0;               // distributed operator, P2P: 0 = pack CTAs inside the apply + ack round (any caller);
                                   // 1 = pack CTAs, no ack round (fused solvers: ApplyDist::no_ack); 2 = the producer of
                                   // x has pushed the boundary values already (halo_push_tile), no pack CTAs; 3 = the
                                   // same, lazily: stores only, this kernel's first CTA raises the flags
  unsigned long long* halo_wait_ns = nullptr; // optional timeline: longest halo-flag wait of a boundary CTA
  unsigned long long* ar_wait_ns = nullptr;   // optional timeline: the final stage's wait for the other ranks' sums
  bool pdl = false, pdl_final = false;        // programmatic-serialization attribute on the apply / on its final stage
};

template<int ND, bool RESID, class Epi, class Final>
int launch_apply(sb_ctx* ctx, const sb_op* op, const double* x, double* y, const Epi& epi, const Final& fin,
                 const int* done, const ApplyOpts& ao = ApplyOpts{}) {
  OpDev d = ao.per_call != nullptr ? *ao.per_call : op->d;
  d.stream_hint = ctx->stream_operator;
  if constexpr (ND > 0) {
    SB_TRY(ensure_red_scratch(ctx, d.n));
  }
  const RedPtrs red{ctx->red.partials, ctx->red.cap_tiles};
  ApplyDist ad;
  CommCtrl* bump = nullptr;
  if (op->distributed && ctx->comm.world > 1) {
    const bool exchange = op->halo.n_nbr > 0 && !(ctx->debug & 2);
    int64_t x_off = 0;
    // NCCL: pack kernel + grouped send/recv in front of the apply. P2P: validated here, done inside the apply kernel.
    if (exchange) SB_TRY(halo_exchange(ctx, op, x, done, &x_off));
    if (ctx->comm.mode == SB_COMM_P2P && exchange) {
      ad.comm = ctx->comm, ad.halo = op->halo, ad.x_off = x_off;
      const int64_t total = op->halo.send_ptr[op->halo.n_nbr];
      if (ao.halo_mode >= 2) {
        ad.n_pack = 0, ad.pre_pushed = 1, ad.post_flags = ao.halo_mode == 3 ? 1 : 0;
      } else {
        ad.n_pack = (int32_t) std::max<int64_t>(1, std::min<int64_t>(64, (total + 2 * kThreads - 1) / (2 * kThreads)));
        ad.no_ack = (ao.fold_later || ao.halo_mode == 1) ? 1 : 0;
      }
      ad.coherent_gather = 1;
      ad.wait_ns = ao.halo_wait_ns;
    }
    // The apply sequence number counts EVERY apply of a distributed operator on every rank, whether this rank has
    // neighbours in this operator or not: the flags of different operators (other partitions, a rank without
    // neighbours) are compared against the same counter, which therefore has to advance in lockstep on all ranks.
    if (ctx->comm.mode == SB_COMM_P2P && !(ctx->debug & 2)) bump = ctx->comm.ctrl(ctx->comm.rank);
  }
  const unsigned grid = (unsigned) (num_tiles(d.n) + ad.n_pack);
  {
  PdlScope pdl_scope(ctx, ao.pdl);
#define SB_LAUNCH(FORM, W) \
  SB_CUDA(launch_kernel(ctx, apply_kernel<FORM, W, ND, RESID, Epi>, grid, kThreads, 0, d, x, y, epi, red, ad, done))
#define SB_WIDTHS(FORM)                   \
  switch (d.width) {                      \
    case 0: case 1: SB_LAUNCH(FORM, 1); break; \
    case 2: SB_LAUNCH(FORM, 2); break;    \
    case 3: SB_LAUNCH(FORM, 3); break;    \
    case 4: SB_LAUNCH(FORM, 4); break;    \
    case 5: SB_LAUNCH(FORM, 5); break;    \
    case 6: SB_LAUNCH(FORM, 6); break;    \
    case 7: SB_LAUNCH(FORM, 7); break;    \
    case 8: SB_LAUNCH(FORM, 8); break;    \
    case 10: SB_LAUNCH(FORM, 10); break;  \
    case 12: SB_LAUNCH(FORM, 12); break;  \
    case 14: SB_LAUNCH(FORM, 14); break;  \
    case 16: SB_LAUNCH(FORM, 16); break;  \
    default:                              \
      set_error("operator width %d not supported (1..8, 10, 12, 14, 16)", d.width); \
      return SB_ERR_INVALID;              \
  }
  if (d.form == SB_FORM_COEF && d.blk != nullptr) {
#define SB_LAUNCH_TMA(W)                                                                                        \
  {                                                                                                             \
    auto kern = apply_kernel_tma<W, ND, RESID, Epi>;                                                            \
    constexpr int smem = StageLayout<W>::cta_bytes;                                                \
    /* the attribute is per device: a process may hold contexts on several GPUs (one host thread each) */         \
    static std::atomic<uint64_t> configured{0};                                                                 \
    const uint64_t bit = 1ull << (ctx->device & 63);                                                            \
    if (!(configured.load(std::memory_order_acquire) & bit)) {                                                  \
      SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));                   \
      configured.fetch_or(bit, std::memory_order_release);                                                      \
    }                                                                                                           \
    SB_CUDA(launch_kernel(ctx, kern, grid, kThreads, smem, d, x, y, epi, red, ad, done));                       \
  }
    switch (d.width) {
      case 0: case 1: SB_LAUNCH_TMA(1) break;
      case 2: SB_LAUNCH_TMA(2) break;
      case 3: SB_LAUNCH_TMA(3) break;
      case 4: SB_LAUNCH_TMA(4) break;
      case 5: SB_LAUNCH_TMA(5) break;
      case 6: SB_LAUNCH_TMA(6) break;
      case 7: SB_LAUNCH_TMA(7) break;
      case 8: SB_LAUNCH_TMA(8) break;
      case 10: SB_LAUNCH_TMA(10) break;
      case 12: SB_LAUNCH_TMA(12) break;
      case 14: SB_LAUNCH_TMA(14) break;
      case 16: SB_LAUNCH_TMA(16) break;
      default:
        set_error("operator width %d not supported (1..8, 10, 12, 14, 16)", d.width);
        return SB_ERR_INVALID;
    }
#undef SB_LAUNCH_TMA
  } else if (d.form == SB_FORM_COEF) {
    SB_WIDTHS(SB_FORM_COEF)
  } else {
    SB_WIDTHS(SB_FORM_FAITHFUL)
  }
#undef SB_WIDTHS
#undef SB_LAUNCH
  }
  ctx->launches++;
  if (ao.fold_later) return SB_OK;
  if constexpr (ND > 0) return launch_final<ND>(ctx, d.n, fin, done, bump, ao.ar_wait_ns, ao.pdl_final);
  if (bump != nullptr) {
    SB_CUDA(launch_kernel(ctx, seq_bump_kernel, 1, 1, 0, bump, done));
    ctx->launches++;
  }
  return SB_OK;
}

} // namespace sb
