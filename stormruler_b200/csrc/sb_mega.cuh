// sb_mega.cuh -- the persistent whole-solve kernel of the fused CG / BiCGStab solvers.
//
// The one-kernel-per-step schedule (sb_solvers.cu) pays a kernel boundary per step: 8 launches per BiCGStab
// iteration (5 steps + 3 one-CTA reduction stages), each ~2.6 us of drain + launch + ramp-up, plus a second wave of
// CTAs per launch. At 10 M cells on one GPU that is 5 % of an iteration; at 1.26 M cells per rank (10 M cells on 8
// GPUs) it is a third of it, and the cross-GPU all-reduce sits behind yet another launch. Here ONE cooperative
// kernel runs the whole iteration loop: a grid of (SMs x resident CTAs) persistent CTAs, every CTA owns the row
// tiles {blockIdx.x, blockIdx.x + grid, ...} in every step, and the steps are separated by grid-wide barriers in
// global memory (~1 us) instead of kernel boundaries:
//
//   BiCGStab iteration (SolverBiCgStab.hpp:93-165)          CG iteration (SolverCg.hpp:86-126)
//     p = r + beta (p - omega v)        | barrier             z = A p, <p,z>          | barrier + all-reduce
//     v = A p, <r~,v>                   | barrier + all-reduce x += a p, r -= a z, <r,r> | barrier + all-reduce
//     r -= alpha v                      | barrier             p = r + beta p          | barrier
//     t = A r, <t,t>, <t,r>             | barrier + all-reduce
//     x = (x + alpha p) + omega r, r -= omega t, <r,r>, <r~,r> | barrier + all-reduce
//
// Same element-wise bodies, same scalar updates, same reduction tree (SB_TREE v1) as the one-kernel-per-step
// schedule, so the two are bit-identical to each other and to the oracle (tests/test_gpu_mega.py).
//
// Reducing barrier: every CTA stores its tile partials and arrives; CTA 0 sees the last arrival, runs the final stage
// of SB_TREE over the partials and "sends" the rank sums into the all-reduce mailbox of every rank (its own included;
// on one GPU the mailbox is local); EVERY CTA polls the mailbox, adds the rank sums in rank order and runs the
// solver's scalar update on its own copy of the solver state in shared memory -- identical inputs, identical code,
// identical bits, so all CTAs of all ranks take the same decisions (alpha, beta, omega, stop) without another
// broadcast. Only CTA 0 records the residual history and writes the state back. The mailbox of all-reduce #a is
// reset by CTA 0 while it handles #a+1 (every local CTA has read it by then), before it sends #a+1 -- a peer cannot
// post #a+2 before it has seen my #a+1.
//
// Tiles: every CTA owns the tiles {blockIdx.x + k grid} of the first three quarters of a step (no bookkeeping) and
// draws the remaining ones from a counter in global memory, one at a time, two tiles ahead of where it works: the
// CTAs that finish early take what is left (a static split leaves CTA 0 waiting 35 us of a 170 us apply step at
// 10 M cells for the slowest CTA: profiles/r02_persistent_v1_*).
//
// Every streamed operand of every step arrives by bulk copy (cp.async.bulk) in a per-warp shared-memory ring:
// the apply steps stage the operator's slice records and the warp's own run of x like apply_kernel_tma does, the
// element-wise steps stage the 512-byte runs of their 2-5 input vectors, so the bytes in flight per SM (~190 KB)
// do not depend on the register budget of three resident CTAs.
//
// Operator apply inside the loop: the TMA-staged tile pipeline of apply_kernel_tma, with the ring running across
// the tiles of a CTA (no pipeline ramp per tile). Gathers use the coherent path (ld.global.ca): x is written by
// this very kernel, and every grid barrier ends in a gpu-scope fence (which invalidates L1). Distributed operator:
// the boundary values are packed and pushed to the neighbours' halo tails by the first CTAs at the start of the
// apply step, boundary tiles (the last ones of every CTA) acquire the neighbours' flags. No ack round is needed:
// between two applies that write the same halo tail there is always an all-reduce, and a rank contributes to it
// only after all its CTAs have finished the earlier apply.
#pragma once

#include "sb_solver_bodies.cuh"

namespace sb {

// Per-context device block of the persistent kernel (grid barrier, single-GPU mailbox, abort flag).
struct MegaCtrl {
  unsigned long long arrive; // grid-barrier arrival counter (zeroed by the host before every launch)
  unsigned long long pad0[15];
  unsigned long long abort;  // a spin wait timed out: every CTA leaves (code of the first failure)
  unsigned long long release;// generation of the last completed plain grid barrier (written by its last arriver)
  unsigned long long ar_seq; // single-GPU: all-reduces completed (mailbox parity); multi-GPU uses CommCtrl::ar_seq
  unsigned long long pad1[13];
  unsigned long long box[2][4]; // single-GPU all-reduce mailbox
  unsigned long long pad2[8];
  unsigned long long dyn[16];   // dynamic-tile counters of three consecutive steps, 32 bytes apart (dyn_slot)
};
// The tile counter of step p is dyn_slot(p); CTA 0 zeroes the counter of step p+2 at the barrier that ends step p.
__device__ __forceinline__ unsigned long long* dyn_slot(MegaCtrl* mc, unsigned long long step) { return &mc->dyn[(step % 3ull) * 4]; }

constexpr int kMegaStamps = SB_TIMELINE_WORDS; // timeline record per iteration (include/stormb200.h: sb_solver_opts::h_timeline)

struct MegaArgs {
  OpDev op;
  ApplyDist ad;                 // ad.n_pack > 0: distributed operator with neighbours (x_off unused here)
  double *x, *r, *p, *v, *t, *rt; // CG: v = z; t, rt unused
  int64_t off_p, off_r;         // byte offsets of p and r inside the slab (halo pushes)
  SolveBlock* blk;
  double* hist;
  double* trace;
  RedPtrs red;
  MegaCtrl* mc;
  unsigned long long timeout_ns;
  // optional timeline (include/stormb200.h: sb_solver_opts::h_timeline), kMegaStamps words per iteration
  unsigned long long* timeline;
  int32_t timeline_iters;
};

// ---- grid barrier ------------------------------------------------------------------------------------------
struct MegaRun {
  unsigned long long gen = 0;  // barriers passed in this launch
  unsigned long long ar = 0;   // all-reduce sequence number of the next reducing barrier
  unsigned long long seq = 0;  // distributed applies completed (CommCtrl::apply_seq)
  uint32_t par = 0;            // bit s: phase parity of the next wait on this warp's mbarrier s
  int32_t pre = 0;             // stages of the NEXT step already in flight in ring slots 0..pre-1 (issued before the barrier)
  int32_t stamp_it = -1;       // timeline: iteration being stamped (-1: off)
  int32_t stamp_b = 0;
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Spin on `ready()`; gives up on timeout or when another CTA has given up. Returns false in that case.
template<class Pred>
__device__ __forceinline__ bool mega_spin(Pred ready, MegaCtrl* mc, unsigned long long timeout_ns, unsigned long long code) {
  if (ready()) return true;
  const unsigned long long t0 = globaltimer_ns();
  for (unsigned spins = 1;; ++spins) {
    if (ready()) return true;
    if ((spins & 127u) == 0) {
      if (ld_relaxed_gpu(&mc->abort) != 0) return false;
      if (globaltimer_ns() - t0 > timeout_ns) {
        atomicCAS(&mc->abort, 0ull, code);
        return false;
      }
    }
  }
}

// Shared-memory scratch of a CTA.
struct MegaShared {
  SolverState st;                       // this CTA's copy of the solver state
  double s_w[2][4][kMaxDots][kWarps];   // tile combine: warp sums of up to 4 tiles, double-buffered
  long long s_next[2];                  // tile scheduler: the tile after next, double-buffered by tile parity
  double s_fin[kMaxDots][kWarps];       // final stage (the reducing CTA)
  double s_all[kMaxRanks][4];           // mailbox values
  double s_local[4];
  int abort;
  int last;                             // this CTA was the last one to arrive at the current barrier
};

__device__ __forceinline__ void stamp(const MegaArgs& a, MegaRun& run, int slot, unsigned long long v) {
  if (run.stamp_it >= 0 && blockIdx.x == 0) a.timeline[(int64_t) run.stamp_it * kMegaStamps + slot] = v;
}

// All threads. Everything this CTA wrote is visible to every CTA that leaves the barrier. The CTA whose arrival
// completes the count learns it from its own atomic (sh.last): it releases the others (plain barrier) or runs the
// reduction (reducing barrier) at once -- nobody polls the counter the arrivals are hammering.
__device__ __forceinline__ void grid_arrive(const MegaArgs& a, MegaRun& run, MegaShared& sh) {
  fence_proxy_async_all(); // my generic-proxy stores vs. the bulk copies (async proxy) issued after the barrier
  __syncthreads();
  run.gen++;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(&a.mc->arrive, 1ull);
    sh.last = (old + 1 == run.gen * gridDim.x) ? 1 : 0;
    if (sh.last) {
      __threadfence();                        // acquire: what the other CTAs wrote before they arrived
      *dyn_slot(a.mc, run.gen + 1) = 0;       // tile counter of the step after next (nobody uses it now)
    }
  }
  __syncthreads();
}

// Plain grid barrier. Returns false when the kernel must be abandoned.
__device__ __forceinline__ bool grid_barrier(const MegaArgs& a, MegaRun& run, MegaShared& sh) {
  const unsigned long long t0 = run.stamp_it >= 0 ? globaltimer_ns() : 0;
  grid_arrive(a, run, sh);
  if (threadIdx.x == 0) {
    if (sh.last) {
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&a.mc->release), "l"(run.gen) : "memory");
    } else {
      const unsigned long long want = run.gen;
      const unsigned long long* rel = &a.mc->release;
      if (!mega_spin([&] { return ld_acquire_gpu(rel) >= want; }, a.mc, a.timeout_ns, 0xD000 + run.gen)) sh.abort = 1;
      __threadfence(); // as cooperative_groups' grid sync: the fence (it invalidates this SM's L1) orders every thread
                       // of the CTA, through the __syncthreads below, behind the arrivals the release stands for
    }
    if (run.stamp_it >= 0) {
      const unsigned long long t1 = globaltimer_ns();
      stamp(a, run, 6 + run.stamp_b, t1 - t0), stamp(a, run, 1 + run.stamp_b, t1);
    }
  }
  run.stamp_b++;
  __syncthreads();
  fence_proxy_async_all();
  return sh.abort == 0;
}

__device__ __forceinline__ unsigned long long* mega_box(const MegaArgs& a, int rank, unsigned long long par, int src, int d) {
  if (a.ad.comm.world > 1) return &a.ad.comm.ctrl(rank)->ar_slot[par][src][d];
  return &a.mc->box[par][d];
}

// Reducing barrier: grid barrier + final stage of SB_TREE + all-reduce over the ranks + the solver's scalar update
// (`fin`) on this CTA's state copy. n_tiles: tiles whose partials the producers of this step have written.
template<int ND, class Final>
__device__ __forceinline__ bool reduce_barrier(const MegaArgs& a, MegaRun& run, MegaShared& sh, int64_t n_tiles, const Final& fin) {
  const int world = a.ad.comm.world, me = a.ad.comm.rank;
  const unsigned long long par = run.ar & 1ull;
  const unsigned long long t0 = run.stamp_it >= 0 ? globaltimer_ns() : 0;
  grid_arrive(a, run, sh);
  if (sh.last) {
    // every local CTA has arrived, hence has read the mailbox of the previous all-reduce: make it empty again
    // BEFORE my sums go out (a peer posts into it only after it has seen those)
    if (threadIdx.x < world * 4) {
      const int r = threadIdx.x >> 2, d = threadIdx.x & 3;
      st_relaxed_sys(mega_box(a, me, par ^ 1ull, r, d), kArSentinel);
    }
    double sums[ND];
    final_stage<ND>(n_tiles, a.red, sh.s_fin, sums);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int d = 0; d < ND; ++d) sh.s_local[d] = sums[d];
      // release: the reset above (and everything the arrivals made visible) is performed before any rank / CTA can
      // act on the values sent below
      if (world > 1) __threadfence_system();
      else __threadfence();
    }
    __syncthreads();
    if (threadIdx.x < world * ND) {
      const int r = threadIdx.x / ND, d = threadIdx.x % ND;
      st_relaxed_sys(mega_box(a, r, par, me, d), (unsigned long long) __double_as_longlong(sh.s_local[d]));
    }
  }
  // every CTA: collect the rank sums
  if (threadIdx.x < world * ND) {
    const int r = threadIdx.x / ND, d = threadIdx.x % ND;
    const unsigned long long* box = mega_box(a, me, par, r, d);
    unsigned long long v = kArSentinel;
    const unsigned long long t1 = (run.stamp_it >= 0 && sh.last) ? globaltimer_ns() : 0;
    if (!mega_spin([&] { return (v = ld_relaxed_sys(box)) != kArSentinel; }, a.mc, a.timeout_ns, 0xC000 + r)) sh.abort = 1;
    sh.s_all[r][d] = __longlong_as_double((long long) v);
    // the reducing CTA: how long the other ranks' sums took to arrive after its own were posted
    if (run.stamp_it >= 0 && sh.last) atomicMax(&a.timeline[(int64_t) run.stamp_it * kMegaStamps + 11 + run.stamp_b], globaltimer_ns() - t1);
  }
  __syncthreads();
  if (threadIdx.x == 0 && sh.abort == 0) {
    double tot[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double s = sh.s_all[0][d];
      for (int r = 1; r < world; ++r) s = __dadd_rn(s, sh.s_all[r][d]);
      tot[d] = s;
    }
    fin(tot);
    if (run.stamp_it >= 0) {
      const unsigned long long t2 = globaltimer_ns();
      stamp(a, run, 6 + run.stamp_b, t2 - t0), stamp(a, run, 1 + run.stamp_b, t2);
    }
  }
  run.ar++;
  run.stamp_b++;
  __threadfence(); // acquire side for the vector data of the other CTAs (and drops stale L1 lines)
  __syncthreads();
  fence_proxy_async_all();
  return sh.abort == 0;
}

// ---- tile scheduler + tile combine ----------------------------------------------------------------------------------
// The sequence of tiles a CTA works on in one step. The host sizes the grid so that the tiles divide evenly:
// R = ceil(tiles / resident CTAs) rounds, grid = ceil(tiles / R) CTAs (601 tiles on 444 resident CTAs: 301 CTAs with
// two tiles each instead of 157 CTAs with two and 287 with one). Tiles blockIdx.x + k grid, k < Ks, are the CTA's
// own (no bookkeeping, one CTA barrier per four tiles for SB_TREE's tile combine, "the 8 warp sums added left to
// right"); with six rounds or more the last two rounds are handed out by a counter in global memory instead, one tile
// at a time, to whichever CTA gets there first (SMs do not run at exactly the same speed, and at twelve rounds the
// static split left a tail of 6-35 us per apply step). `cur` and `nxt` are known to every thread; from two tiles
// before the dynamic part thread 0 draws the tile after `nxt` while the CTA works on `cur` (the ticket is consumed
// only at the CTA barrier that ends the tile, which also publishes it).
template<int ND>
struct TileSched {
  static constexpr int NA = ND > 0 ? ND : 1;
  MegaShared& sh;
  const RedPtrs& red;
  unsigned long long* ctr;
  long long n_tiles, G, Ks, base, i = 0, cur, nxt;
  long long ticket = 0;      // thread 0: the tile drawn in begin()
  long long chunk_tile0 = 0; // first tile of the pending combine chunk
  int cnt = 0, buf = 0;      // pending tiles in the combine chunk, its buffer
  __device__ __forceinline__ TileSched(MegaShared& s, const RedPtrs& r, const MegaArgs& a, const MegaRun& run, long long tiles)
      : sh(s), red(r), ctr(dyn_slot(a.mc, run.gen)), n_tiles(tiles), G(gridDim.x) {
    const long long R = (n_tiles + G - 1) / G;
    Ks = R >= 6 ? R - 2 : R;
    base = Ks * G;
    cur = blockIdx.x; // the grid never exceeds the number of tiles
    nxt = cur + G;    // the second tile is always static (Ks >= 2 whenever there is one)
  }
  __device__ __forceinline__ bool more() const { return cur < n_tiles; }
  __device__ __forceinline__ bool dynamic_from_here() const { return i + 2 >= Ks && Ks * G < n_tiles; }
  // start of a tile: thread 0 looks two tiles ahead (only where the dynamic part begins)
  __device__ __forceinline__ void begin() {
    if (threadIdx.x == 0 && dynamic_from_here()) ticket = nxt < n_tiles ? base + (long long) atomicAdd(ctr, 1ull) : n_tiles;
  }
  __device__ __forceinline__ void flush() {
    if constexpr (ND > 0) {
      if (threadIdx.x < cnt * ND) {
        const int c = threadIdx.x / ND, d = threadIdx.x % ND;
        double t = sh.s_w[buf][c][d][0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) t = __dadd_rn(t, sh.s_w[buf][c][d][w]);
        red.partials[(int64_t) d * red.cap_tiles + chunk_tile0 + (long long) c * G] = t;
      }
      buf ^= 1, cnt = 0;
    }
  }
  // end of a tile: park the warp sums (reducing steps), advance; CTA barrier where something has to be published
  __device__ __forceinline__ void end(double (&acc)[NA]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if constexpr (ND > 0) {
      if (cnt == 0) chunk_tile0 = cur;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        const double v = warp_butterfly(acc[d]);
        if (lane == 0) sh.s_w[buf][cnt][d][warp] = v;
        acc[d] = 0.0;
      }
      ++cnt;
    }
    const bool dyn = dynamic_from_here();
    long long nxt2 = nxt + G; // static successor of nxt
    if (dyn) {
      if (threadIdx.x == 0) sh.s_next[i & 1] = ticket;
      __syncthreads();
      nxt2 = sh.s_next[i & 1];
      flush(); // dynamic tiles are not strided: one tile per chunk from here on
    } else if (ND > 0 && (cnt == 4 || nxt >= n_tiles)) {
      __syncthreads();
      flush();
    }
    cur = nxt, nxt = nxt2, ++i;
  }
};

// ---- per-warp staging ring ---------------------------------------------------------------------------------------
constexpr int kMaxSlots = 4;
template<int W>
struct MegaRing {
  static constexpr int apply_bytes = kStages * StageLayout<W>::bytes;
  static constexpr int warp_bytes = apply_bytes > 8192 ? apply_bytes : 8192;
  static constexpr int cta_bytes = warp_bytes * kWarps;
};

__device__ __forceinline__ void ring_wait(uint64_t* bars, MegaRun& run, int slot) {
  mbar_wait(&bars[slot], (run.par >> slot) & 1u);
  run.par ^= 1u << slot;
}

// ---- element-wise step --------------------------------------------------------------------------------------------
// A stage = the 512-byte runs of the body's NV input vectors for one 64-element sub-iteration of the warp, one
// mbarrier per ring slot; S = min(4, ring / stage) stages in flight per warp. Same element -> lane mapping, same
// per-element arithmetic (Body::run) and same accumulation order as ew_kernel.
template<int ND, int RING, class Body>
__device__ __forceinline__ void ew_phase(const MegaArgs& a, MegaRun& run, MegaShared& sh, const Body& body, unsigned char* smem,
                                         uint64_t (*bars_all)[kMaxSlots]) {
  constexpr int NV = Body::NV, kStage = NV * 512;
  constexpr int S = RING / kStage < kMaxSlots ? RING / kStage : kMaxSlots;
  static_assert(S >= 1, "ring too small for this body");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = a.op.n;
  unsigned char* wbase = smem + (size_t) warp * RING;
  uint64_t* bars = bars_all[warp];
  TileSched<ND> ts(sh, a.red, a, run, num_tiles(n));
  auto issue = [&](long long tile, int j, int slot, uint32_t dep) {
    const int64_t r0 = tile * kTile + warp * (kTile / kWarps) + j * 64;
    mbar_expect_tx(&bars[slot], (uint32_t) kStage);
#pragma unroll
    for (int k = 0; k < NV; ++k)
      bulk_g2s(wbase + slot * kStage + k * 512, reinterpret_cast<const unsigned char*>(body.in(k) + r0) + dep, 512, &bars[slot]);
  };
  if (lane == 0 && run.pre == 0) {
#pragma unroll
    for (int q = 0; q < S; ++q) issue(ts.cur, q, q, 0u); // S <= kSub: all in the first tile
  }
  run.pre = 0;
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
  int slot = 0;
  while (ts.more()) {
    ts.begin();
#pragma unroll
    for (int j = 0; j < kSub; ++j) {
      ring_wait(bars, run, slot);
      double2 in[NV];
      uint32_t fold = 0;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        in[k] = reinterpret_cast<const double2*>(wbase + slot * kStage + k * 512)[lane];
        const long long b0 = __double_as_longlong(in[k].x), b1 = __double_as_longlong(in[k].y);
        fold ^= (uint32_t) b0 ^ (uint32_t) (b0 >> 32) ^ (uint32_t) b1 ^ (uint32_t) (b1 >> 32);
      }
      // generic-proxy loads vs. the async-proxy refill of this slot: see apply_kernel_tma (sb_op.cuh)
      const uint32_t dep = fold & (uint32_t) a.op.zero;
      __syncwarp();
      if (lane == 0) {
        if (j + S < kSub) issue(ts.cur, j + S, slot, dep);
        else if (ts.nxt < ts.n_tiles) issue(ts.nxt, j + S - kSub, slot, dep);
      }
      typename Body::Regs g;
      body.fill(g, in);
      body.run(lane_elem(ts.cur, j), n, g, acc);
      slot = slot + 1 == S ? 0 : slot + 1;
    }
    ts.end(acc);
  }
}

// Issue the first stages of the NEXT step before the barrier that precedes it, so that their latency overlaps the
// barrier. Legal because a step's first tile is always the CTA's own tile blockIdx.x, whose elements are read and
// written by the same warp (same lanes) in every step: what the copies read is final, and this warp's own stores are
// ordered before them by the proxy fence + __syncwarp. If the step never runs (the solver stopped), the kernel drains
// the copies before it exits.
template<int RING, class Body>
__device__ __forceinline__ void ew_prefetch(MegaRun& run, const Body& body, unsigned char* smem, uint64_t (*bars_all)[kMaxSlots]) {
  constexpr int NV = Body::NV, kStage = NV * 512;
  constexpr int S = RING / kStage < kMaxSlots ? RING / kStage : kMaxSlots;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fence_proxy_async_all();
  __syncwarp();
  if (lane == 0) {
    unsigned char* wbase = smem + (size_t) warp * RING;
#pragma unroll
    for (int q = 0; q < S; ++q) {
      const int64_t r0 = (int64_t) blockIdx.x * kTile + warp * (kTile / kWarps) + q * 64;
      mbar_expect_tx(&bars_all[warp][q], (uint32_t) kStage);
#pragma unroll
      for (int k = 0; k < NV; ++k)
        bulk_g2s(wbase + q * kStage + k * 512, reinterpret_cast<const unsigned char*>(body.in(k) + r0), 512, &bars_all[warp][q]);
    }
  }
  run.pre = S;
}

template<int W>
__device__ __forceinline__ void apply_prefetch(const MegaArgs& a, MegaRun& run, const double* x, unsigned char* smem,
                                               uint64_t (*bars_all)[kMaxSlots]) {
  using L = StageLayout<W>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  fence_proxy_async_all();
  __syncwarp();
  if (lane == 0) {
    unsigned char* wbase = smem + (size_t) warp * MegaRing<W>::warp_bytes;
#pragma unroll
    for (int j = 0; j < kStages; ++j) {
      const int64_t r = (int64_t) blockIdx.x * kTile + warp * (kTile / kWarps) + j * 64;
      mbar_expect_tx(&bars_all[warp][j], (uint32_t) L::bytes);
      bulk_g2s(wbase + j * L::bytes, a.op.blk + (r >> 6) * (int64_t) L::slice, L::slice, &bars_all[warp][j]);
      bulk_g2s(wbase + j * L::bytes + L::xown, reinterpret_cast<const unsigned char*>(x + r), 512, &bars_all[warp][j]);
    }
  }
  run.pre = kStages;
}

// ---- operator-apply step -------------------------------------------------------------------------------------------
// apply_kernel_tma's pipeline with the ring running across this CTA's tiles. `x_off`: byte offset of x in the slab.
template<int W, int ND, bool RESID, class Epi>
__device__ __forceinline__ void apply_phase(const MegaArgs& a, MegaRun& run, MegaShared& sh, const double* __restrict__ x,
                                            double* __restrict__ y, int64_t x_off, const Epi& epi, unsigned char* smem,
                                            uint64_t (*bars_all)[kMaxSlots], int apply_ordinal) {
  using L = StageLayout<W>;
  const OpDev& op = a.op;
  const ApplyDist& ad = a.ad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbase = smem + (size_t) warp * MegaRing<W>::warp_bytes;
  uint64_t* bars = bars_all[warp];
  TileSched<ND> ts(sh, a.red, a, run, num_tiles(op.n));
  auto issue = [&](long long tile, int j, uint32_t dep) {
    const int s = j & 1; // kStages == 2, four stages per tile: the slot is the parity of the sub-iteration
    unsigned char* dst = wbase + s * L::bytes;
    const int64_t r = tile * kTile + warp * (kTile / kWarps) + j * 64;
    mbar_expect_tx(&bars[s], (uint32_t) L::bytes);
    bulk_g2s(dst, op.blk + (r >> 6) * (int64_t) L::slice + dep, L::slice, &bars[s]);
    bulk_g2s(dst + L::xown, reinterpret_cast<const unsigned char*>(x + r) + dep, 512, &bars[s]);
  };
  static_assert(kStages == 2 && kSub == 4, "slot arithmetic of the apply step");
  if (lane == 0 && run.pre == 0) issue(ts.cur, 0, 0u), issue(ts.cur, 1, 0u);
  run.pre = 0;
  const unsigned long long seq = run.seq + 1; // this apply's number (all CTAs of all ranks agree)
  if (ad.n_pack > 0) {
    // halo push: boundary values straight into the neighbours' halo tails over NVLink, one element per thread
    CommCtrl* me = ad.comm.ctrl(ad.comm.rank);
    const int64_t G = gridDim.x, total = ad.halo.send_ptr[ad.halo.n_nbr];
    const int64_t n_pack = (total + kThreads - 1) / kThreads < G ? (total + kThreads - 1) / kThreads : G;
    if ((int64_t) blockIdx.x < n_pack) {
      for (int64_t i = (int64_t) blockIdx.x * kThreads + threadIdx.x; i < total; i += n_pack * kThreads) {
        int k = 0;
#pragma unroll
        for (int q = 1; q < kMaxRanks; ++q) k += (q < ad.halo.n_nbr && i >= ad.halo.send_ptr[q]) ? 1 : 0;
        double* dst = reinterpret_cast<double*>(ad.comm.base[ad.halo.nbr_rank[k]] + x_off) + ad.halo.send_dst[k] + (i - ad.halo.send_ptr[k]);
        *dst = __ldcg(x + ad.halo.send_idx[i]);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence_system(); // this CTA's peer stores are performed before the ticket is taken
        const unsigned long long ticket = atomicAdd(&me->pack_ticket, 1ull);
        if (ticket == (unsigned long long) n_pack - 1) {
          me->pack_ticket = 0;
          __threadfence_system();
          for (int k = 0; k < ad.halo.n_nbr; ++k) st_relaxed_sys(&ad.comm.ctrl(ad.halo.nbr_rank[k])->halo_flag[ad.comm.rank], seq);
        }
      }
    }
  }
  double acc[ND > 0 ? ND : 1];
#pragma unroll
  for (int d = 0; d < (ND > 0 ? ND : 1); ++d) acc[d] = 0.0;
  bool halo_ready = ad.n_pack == 0;
  while (ts.more()) {
    ts.begin();
    const int64_t row0 = ts.cur * kTile + warp * (kTile / kWarps);
    typename Epi::Regs er[kSub];
#pragma unroll
    for (int j = 0; j < kSub; ++j) epi.load(row0 + j * 64 + 2 * lane, er[j]);
    if (!halo_ready && ts.cur >= ad.halo.first_boundary_tile) {
      // boundary rows gather from the halo tail: wait until every neighbour's values of THIS apply have landed
      CommCtrl* me = ad.comm.ctrl(ad.comm.rank);
      if (lane < ad.halo.n_nbr) {
        const unsigned long long* flag = &me->halo_flag[ad.halo.nbr_rank[lane]];
        const unsigned long long t0 = run.stamp_it >= 0 ? globaltimer_ns() : 0;
        if (!mega_spin([&] { return ld_acquire_sys(flag) >= seq; }, a.mc, a.timeout_ns, 0xB000 + ad.halo.nbr_rank[lane])) sh.abort = 1;
        if (run.stamp_it >= 0) atomicMax(&a.timeline[(int64_t) run.stamp_it * kMegaStamps + 16 + apply_ordinal], globaltimer_ns() - t0);
      }
      __syncwarp();
      halo_ready = true;
    }
#pragma unroll
    for (int j = 0; j < kSub; ++j) {
      const int s = j & 1;
      ring_wait(bars, run, s);
      const unsigned char* src = wbase + s * L::bytes;
      int2 c[W];
      double2 cf[W];
#pragma unroll
      for (int e = 0; e < W; ++e) {
        c[e] = reinterpret_cast<const int2*>(src + L::col + e * 256)[lane];
        cf[e] = reinterpret_cast<const double2*>(src + L::coef + e * 512)[lane];
      }
      const double2 dg = reinterpret_cast<const double2*>(src + L::diag)[lane];
      const double2 xo = reinterpret_cast<const double2*>(src + L::xown)[lane];
      // generic-proxy loads vs. the async-proxy refill of this slot: see apply_kernel_tma (sb_op.cuh)
      uint32_t fold = 0;
      {
        auto mix = [&](double v) {
          const long long b = __double_as_longlong(v);
          fold ^= (uint32_t) b ^ (uint32_t) (b >> 32);
        };
#pragma unroll
        for (int e = 0; e < W; ++e) {
          fold ^= (uint32_t) c[e].x ^ (uint32_t) c[e].y;
          mix(cf[e].x), mix(cf[e].y);
        }
        mix(dg.x), mix(dg.y), mix(xo.x), mix(xo.y);
      }
      const uint32_t dep = fold & (uint32_t) op.zero;
      __syncwarp();
      if (lane == 0) {
        if (j + kStages < kSub) issue(ts.cur, j + kStages, dep);
        else if (ts.nxt < ts.n_tiles) issue(ts.nxt, j + kStages - kSub, dep);
      }
      double g0[W], g1[W];
#pragma unroll
      for (int e = 0; e < W; ++e) {
        g0[e] = (c[e].x >= 0) ? __ldca(x + c[e].x) : 0.0;
        g1[e] = (c[e].y >= 0) ? __ldca(x + c[e].y) : 0.0;
      }
      double u0 = __dmul_rn(dg.x, xo.x), u1 = __dmul_rn(dg.y, xo.y);
#pragma unroll
      for (int e = 0; e < W; ++e) {
        const double t0 = __dadd_rn(u0, __dmul_rn(cf[e].x, g0[e]));
        const double t1 = __dadd_rn(u1, __dmul_rn(cf[e].y, g1[e]));
        u0 = (c[e].x >= 0) ? t0 : u0;
        u1 = (c[e].y >= 0) ? t1 : u1;
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
    ts.end(acc);
  }
  run.seq = seq;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------
// Resident CTAs per SM the kernel is compiled for: the TMA ring (16 stages of 768 W + 1024 bytes per CTA) allows
// three CTAs per SM up to W = 4, two up to W = 7, one beyond; the register budget follows.
constexpr int mega_ctas_per_sm(int W) { return W <= 4 ? 3 : (W <= 7 ? 2 : 1); }

template<int KIND, int W>
__global__ void __launch_bounds__(kThreads, mega_ctas_per_sm(W)) krylov_persistent_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) unsigned char sb_smem[];
  __shared__ __align__(8) uint64_t bars[kWarps][kMaxSlots];
  __shared__ MegaShared sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool dist = a.ad.comm.world > 1;
  CommCtrl* ctl = dist ? a.ad.comm.ctrl(a.ad.comm.rank) : nullptr;
  if (threadIdx.x == 0) {
    sh.st = a.blk->ver(0);
    sh.abort = 0;
  }
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) mbar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async_all();
  }
  MegaRun run;
  run.ar = dist ? ctl->ar_seq : a.mc->ar_seq;
  run.seq = dist ? ctl->apply_seq : 0;
  __syncthreads();
  SolverState* S = &sh.st;
  const Recorder rec{S, blockIdx.x == 0 ? a.hist : nullptr, blockIdx.x == 0 ? a.trace : nullptr};
  const int64_t n_tiles = num_tiles(a.op.n);
  constexpr int RING = MegaRing<W>::warp_bytes;
  bool ok = true;
  long long it = 0;
  while (ok && !S->done) {
    run.stamp_it = (a.timeline != nullptr && blockIdx.x == 0 && it < a.timeline_iters) ? (int32_t) it : -1;
    if (a.timeline != nullptr && blockIdx.x != 0 && it < a.timeline_iters) run.stamp_it = (int32_t) it; // halo waits of all CTAs
    run.stamp_b = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) stamp(a, run, 0, globaltimer_ns());
    if constexpr (KIND == (int) Kind::BiCgStab) {
      const BiDirectionBody dir{S, a.p, a.r, a.v};
      const BiHalfBody half{S, a.r, a.v};
      const BiEndBody fin{S, a.x, a.r, a.p, a.t, a.rt};
      ew_phase<0, RING>(a, run, sh, dir, sb_smem, bars);
      apply_prefetch<W>(a, run, a.p, sb_smem, bars);
      if (!(ok = grid_barrier(a, run, sh))) break;
      apply_phase<W, 1, false>(a, run, sh, a.p, a.v, a.off_p, EpiUY{a.rt}, sb_smem, bars, 0);
      ew_prefetch<RING>(run, half, sb_smem, bars);
      if (!(ok = reduce_barrier<1>(a, run, sh, n_tiles, BiAlphaFinal{rec}))) break;
      ew_phase<0, RING>(a, run, sh, half, sb_smem, bars);
      apply_prefetch<W>(a, run, a.r, sb_smem, bars);
      if (!(ok = grid_barrier(a, run, sh))) break;
      apply_phase<W, 2, false>(a, run, sh, a.r, a.t, a.off_r, EpiYYandYX{}, sb_smem, bars, 1);
      ew_prefetch<RING>(run, fin, sb_smem, bars);
      if (!(ok = reduce_barrier<2>(a, run, sh, n_tiles, BiOmegaFinal{rec}))) break;
      ew_phase<2, RING>(a, run, sh, fin, sb_smem, bars);
      ew_prefetch<RING>(run, dir, sb_smem, bars);
      if (!(ok = reduce_barrier<2>(a, run, sh, n_tiles, BiEndFinal{rec}))) break;
    } else {
      const CgUpdateBody upd{S, a.x, a.r, a.p, a.v};
      const CgDirectionBody dir{S, a.p, a.r};
      apply_phase<W, 1, false>(a, run, sh, a.p, a.v, a.off_p, EpiXY{}, sb_smem, bars, 0);
      ew_prefetch<RING>(run, upd, sb_smem, bars);
      if (!(ok = reduce_barrier<1>(a, run, sh, n_tiles, CgAlphaFinal{rec}))) break;
      ew_phase<1, RING>(a, run, sh, upd, sb_smem, bars);
      ew_prefetch<RING>(run, dir, sb_smem, bars);
      if (!(ok = reduce_barrier<1>(a, run, sh, n_tiles, CgBetaFinal{rec}))) break;
      ew_phase<0, RING>(a, run, sh, dir, sb_smem, bars);
      apply_prefetch<W>(a, run, a.p, sb_smem, bars);
      if (!(ok = grid_barrier(a, run, sh))) break;
    }
    ++it;
  }
  // copies issued for a step that did not run any more: a CTA must not exit with bulk copies landing in its smem
  for (int q = 0; q < run.pre; ++q) ring_wait(bars[warp], run, q);
  // Every CTA has read the last mailbox before CTA 0 empties it (the next kernel expects empty mailboxes).
  run.stamp_it = -1;
  if (ok) ok = grid_barrier(a, run, sh);
  if (blockIdx.x == 0 && ok) {
    if (run.ar > 0 && threadIdx.x < a.ad.comm.world * 4) {
      const int r = threadIdx.x >> 2, d = threadIdx.x & 3;
      st_relaxed_sys(mega_box(a, a.ad.comm.rank, (run.ar - 1) & 1ull, r, d), kArSentinel);
    }
    if (threadIdx.x == 0) {
      a.blk->ver(0) = *S, a.blk->final_() = *S, a.blk->done = S->done;
      if (dist) ctl->ar_seq = run.ar, ctl->apply_seq = run.seq;
      else a.mc->ar_seq = run.ar;
    }
  }
  if (!ok && blockIdx.x == 0 && threadIdx.x == 0 && dist) comm_fail(ctl, ld_relaxed_gpu(&a.mc->abort));
}

} // namespace sb
