// sb_comm.cuh -- device side of the multi-GPU path: NVLink peer stores, sequence flags, and the
// in-kernel all-reduce (DESIGN.md "Multi-GPU").
//
// P2P protocol (all state lives in each rank's CommCtrl, written by peers over NVLink):
//   halo exchange, apply #s = apply_seq + 1 on every rank -- all inside the apply kernel:
//     pack CTAs      (the first n_pack blocks of the grid, so they are scheduled first and overlap the
//                    interior tiles): block 0 tells each neighbour "I have finished reading the halos of
//                    apply #s-1" (ack_flag; true because griddepcontrol.wait has returned); every pack CTA
//                    waits until each neighbour has acked #s-1 (its halo tail may be overwritten), stores
//                    its share of boundary values straight into the neighbours' halo tails, fences; the
//                    last one (ticket) release-stores halo_flag[me] = s at every neighbour.
//     boundary tiles acquire halo_flag[q] >= s for every neighbour q before their first gather;
//                    interior tiles never wait.
//     apply_seq is bumped by the one-CTA kernel that follows the apply (final reduce, or seq_bump_kernel
//                    for a plain sb_apply), so all CTAs of an apply kernel read the same value.
//   all-reduce #a (inside the one-CTA final-reduce kernel): thread (r, d) stores local sum d into
//     rank r's mailbox ar_slot[a&1][me][d]; then spins on its own mailbox [a&1][r][d] until the value
//     differs from the NaN sentinel, resets it, and thread 0 adds the P values in rank order.
//     Two parities suffice: a rank can start all-reduce #a+2 only after every peer contributed to
//     #a+1, i.e. after every peer has read (and reset) its #a mailboxes.
// Every spin loop is bounded (CommDev::timeout_ns = sb_ctx::spin_timeout_ns: env SB_SPIN_TIMEOUT_S, default 120 s):
// on expiry the kernel records the reason in CommCtrl::error and gives up WITHOUT trapping (a trap would destroy the
// CUDA context of this rank and, one timeout later, of every peer); every later wait of this rank sees the error
// word and returns at once, so whatever is queued drains in microseconds with meaningless values, and the host turns
// the error word into SB_ERR_COMM at the end of the solve (sb_comm_status for the stand-alone entry points).
#pragma once

#include "sb_common.cuh"

namespace sb {

constexpr unsigned long long kDefaultSpinTimeoutNs = 120ull * 1000ull * 1000ull * 1000ull;

// Programmatic dependent launch hooks: every stepwise kernel calls pdl_trigger() (dependents may be scheduled) and
// then pdl_wait() (all prerequisite grids have completed and their writes are visible) before it touches anything a
// previous kernel wrote. By default the library launches without the programmatic-serialization attribute, so both
// are no-ops: measured on the B200 in both rounds, per kernel class (SB_TUNE_PDL_FINAL / _AFTER_FINAL / _APPLY), on
// 1, 2 and 8 GPUs, the attribute lost 1-8 % inside a replayed CUDA graph.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// The solver's stop flag, read around L1.
__device__ __forceinline__ bool is_done(const int* done) {
  if (done == nullptr) return false;
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(done) : "memory");
  return v != 0;
}

struct HaloDev {
  int32_t n_nbr = 0;               // 0: operator is not distributed
  int32_t first_boundary_tile = 0; // row tiles >= this one contain boundary cells
  int32_t nbr_rank[kMaxRanks] = {};
  int64_t send_ptr[kMaxRanks + 1] = {};
  int64_t send_dst[kMaxRanks] = {}; // element offset inside neighbour k's vector
  const int32_t* send_idx = nullptr; // device
  // push-on-produce plan (sb_op.cu: attach_halo): the send list once more, sorted by the 2048-row tile of its source
  // cell. push_ptr[q] .. push_ptr[q + 1]: the entries of boundary tile first_boundary_tile + q; an entry is
  // {source cell (local index) | neighbour slot << 28, element offset inside that neighbour's vector}.
  const int32_t* push_ptr = nullptr;
  const int2* push_entry = nullptr;
  int32_t n_push_tiles = 0;          // tiles first_boundary_tile .. last owned tile
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

static __device__ __noinline__ void comm_fail(CommCtrl* ctrl, unsigned long long code) {
  atomicCAS(&ctrl->error, 0ull, code); // the first failure is the one reported
  __threadfence_system();
}
__device__ __forceinline__ bool comm_failed(const CommCtrl* ctrl) { return ld_relaxed_gpu(&ctrl->error) != 0; }

// Spin until *flag >= want (acquire). False: gave up (timeout, or this rank has already failed).
__device__ __forceinline__ bool wait_flag_ge(const unsigned long long* flag, unsigned long long want, CommCtrl* ctrl,
                                             unsigned long long code, unsigned long long timeout_ns) {
  if (ld_acquire_sys(flag) >= want) return true;
  if (comm_failed(ctrl)) return false;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > timeout_ns || comm_failed(ctrl)) {
      comm_fail(ctrl, code);
      return false;
    }
  }
  return true;
}

// ---- halo pack, P2P: boundary values go straight into the neighbours' halo tails -------------------
// Run by the first `n_pack` CTAs of the apply kernel. x_off: byte offset of the vector inside the slab
// (identical on every rank).
__device__ __forceinline__ void halo_pack_role(const CommDev& comm, const HaloDev& halo, const double* __restrict__ x,
                                               int64_t x_off, int n_pack, bool no_ack) {
  CommCtrl* me = comm.ctrl(comm.rank);
  const unsigned long long seq = ld_acquire_sys(&me->apply_seq) + 1; // this apply's number
  if (!no_ack) {
    if (blockIdx.x == 0 && threadIdx.x < halo.n_nbr) {
      // every earlier kernel of this rank has completed (griddepcontrol.wait) -> the halos of apply
      // #seq-1 are free to overwrite
      st_release_sys(&comm.ctrl(halo.nbr_rank[threadIdx.x])->ack_flag[comm.rank], seq - 1);
    }
    if (threadIdx.x < halo.n_nbr)
      wait_flag_ge(&me->ack_flag[halo.nbr_rank[threadIdx.x]], seq - 1, me, 0xA000 + halo.nbr_rank[threadIdx.x], comm.timeout_ns);
    __syncthreads();
  }
  const int64_t total = halo.send_ptr[halo.n_nbr];
  for (int64_t i = (int64_t) blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t) n_pack * kThreads) {
    int k = 0;
#pragma unroll
    for (int q = 1; q < kMaxRanks; ++q) k += (q < halo.n_nbr && i >= halo.send_ptr[q]) ? 1 : 0;
    double* dst = reinterpret_cast<double*>(comm.base[halo.nbr_rank[k]] + x_off) + halo.send_dst[k] + (i - halo.send_ptr[k]);
    *dst = x[halo.send_idx[i]];
  }
  __threadfence_system(); // my peer stores are performed before the ticket below is taken
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long ticket = atomicAdd(&me->pack_ticket, 1ull);
    if (ticket == (unsigned long long) n_pack - 1) {
      me->pack_ticket = 0;
      // ONE system fence, then posted (relaxed) flag stores: fence + relaxed store is a release, and the
      // stores to the neighbours do not wait for each other. (A st.release.sys per neighbour serialises a
      // full fence round trip per flag: measured ~25 us per apply with 4-7 neighbours at 8 GPUs.)
      __threadfence_system();
      for (int k = 0; k < halo.n_nbr; ++k)
        st_relaxed_sys(&comm.ctrl(halo.nbr_rank[k])->halo_flag[comm.rank], seq);
    }
  }
}

// ---- halo push by the PRODUCER of an apply's input (SB_TUNE_PUSH_ON_PRODUCE) ----------------------------------------
// Called by every CTA of a pushing element-wise kernel that owns a boundary tile, right after the tile's values have
// been stored to y: the CTA forwards the values its neighbours need (this tile's share of the send list) into their
// halo tails, fences, takes a ticket; the CTA that takes the last ticket raises halo_flag[me] = #(next apply) at every
// neighbour. Boundary tiles are the first CTAs of the producer's grid, so the values travel while the rest of the
// kernel, the kernel boundary and the apply's interior tiles run; the apply itself has no pack CTAs and its boundary
// tiles normally find the flags up. No ack round: a producer runs behind an all-reduce to which every neighbour
// contributed only after its previous apply of this vector had completed (sb_op.cuh: ApplyDist::no_ack).
//
// `lazy` (SB_TUNE_PUSH_LAZY): the CTA only issues the peer stores -- no fence, no ticket, no flag. The stores are posted
// writes that drain while the rest of the producer runs (boundary tiles come first), the producer's completion is
// what makes them performed, and the flags are raised by the first CTA of the CONSUMING apply (halo_post_flags) as
// soon as that kernel starts. Measured at 8 GPUs the fence + ticket + fence + flag chain of the eager form doubled
// the producer kernels (10 -> 20 us), and the same chain inside the apply's pack CTAs arrived 10 us after the boundary
// tiles needed it (profiles/r02_ab_n8_10M.txt).
__device__ __forceinline__ void halo_push_tile(const CommDev& comm, const HaloDev& halo, const double* y, int64_t y_off,
                                               int64_t tile, bool lazy) {
  __syncthreads(); // the tile's stores to y are visible to the whole CTA
  CommCtrl* me = comm.ctrl(comm.rank);
  const int q = (int) (tile - halo.first_boundary_tile);
  const int32_t beg = halo.push_ptr[q], end = halo.push_ptr[q + 1];
  // four entries per thread and pass, every load of a pass issued before its first store: a boundary tile forwards
  // ~2 000 values, and one entry at a time (entry -> value -> store, ~1.5 us of dependent latency each) kept the
  // pushing CTAs alive 6 us beyond the tile's own work (direction kernel 10.8 -> 16.7 us at 8 GPUs,
  // profiles/r02_ab_n8_10M_lazy.txt)
  constexpr int kPush = 4;
  for (int32_t i0 = beg + (int32_t) threadIdx.x; i0 < end; i0 += kPush * kThreads) {
    int2 en[kPush];
    double v[kPush];
#pragma unroll
    for (int u = 0; u < kPush; ++u) {
      const int32_t i = i0 + u * kThreads;
      en[u] = i < end ? halo.push_entry[i] : make_int2(0, 0);
    }
#pragma unroll
    for (int u = 0; u < kPush; ++u) v[u] = (i0 + u * kThreads < end) ? __ldcg(y + (en[u].x & 0x0fffffff)) : 0.0;
#pragma unroll
    for (int u = 0; u < kPush; ++u) {
      if (i0 + u * kThreads < end) {
        const int k = (int) ((uint32_t) en[u].x >> 28);
        double* dst = reinterpret_cast<double*>(comm.base[halo.nbr_rank[k]] + y_off) + en[u].y;
        *dst = v[u];
      }
    }
  }
  if (lazy) return;
  __threadfence_system(); // my peer stores are performed before the ticket below is taken
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long ticket = atomicAdd(&me->push_ticket, 1ull);
    if (ticket == (unsigned long long) halo.n_push_tiles - 1) {
      me->push_ticket = 0;
      const unsigned long long seq = ld_acquire_sys(&me->apply_seq) + 1; // the apply that will consume these values
      __threadfence_system();
      for (int k = 0; k < halo.n_nbr; ++k)
        st_relaxed_sys(&comm.ctrl(halo.nbr_rank[k])->halo_flag[comm.rank], seq);
    }
  }
}

// Lazy push, consumer side: the kernel that produced x has completed (stream order; griddepcontrol.wait has returned),
// so its peer stores are performed; one system fence makes that cumulative, then the flags go out as posted stores.
// Called by the first CTA of the apply, before anything else.
__device__ __forceinline__ void halo_post_flags(const CommDev& comm, const HaloDev& halo) {
  if (threadIdx.x < halo.n_nbr) {
    CommCtrl* me = comm.ctrl(comm.rank);
    const unsigned long long seq = ld_acquire_sys(&me->apply_seq) + 1; // this apply's number
    __threadfence_system();
    st_relaxed_sys(&comm.ctrl(halo.nbr_rank[threadIdx.x])->halo_flag[comm.rank], seq);
  }
}

// Plain sb_apply on a distributed operator has no final-reduce kernel behind it: this bumps apply_seq.
static __global__ void seq_bump_kernel(CommCtrl* me, const int* __restrict__ done) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  me->apply_seq = me->apply_seq + 1;
}

// NCCL mode: gather into a contiguous local send buffer.
static __global__ void __launch_bounds__(kThreads) halo_pack_local_kernel(HaloDev halo, const double* __restrict__ x,
                                                                   double* __restrict__ sendbuf,
                                                                   const int* __restrict__ done) {
  pdl_trigger();
  pdl_wait();
  if (is_done(done)) return;
  const int64_t total = halo.send_ptr[halo.n_nbr];
  for (int64_t i = (int64_t) blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t) gridDim.x * kThreads)
    sendbuf[i] = x[halo.send_idx[i]];
}

// ---- in-kernel all-reduce (P2P), called by all threads of the one-CTA final-reduce kernel ----------
// `sums` valid in thread 0 on entry and on exit (rank-ordered total, bit-identical on every rank).
template<int ND>
__device__ __forceinline__ void allreduce_p2p(const CommDev& comm, double (&sums)[ND]) {
  __shared__ double s_local[ND];
  __shared__ double s_all[kMaxRanks][ND];
  CommCtrl* me = comm.ctrl(comm.rank);
  const unsigned long long par = me->ar_seq & 1ull;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < ND; ++d) s_local[d] = sums[d];
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < comm.world * ND) {
    const int r = t / ND, d = t % ND;
    st_relaxed_sys(&comm.ctrl(r)->ar_slot[par][comm.rank][d], (unsigned long long) __double_as_longlong(s_local[d]));
    unsigned long long* box = &me->ar_slot[par][r][d];
    unsigned long long v = ld_relaxed_sys(box);
    if (v == kArSentinel && !comm_failed(me)) {
      const unsigned long long t0 = globaltimer_ns();
      unsigned spins = 0;
      while ((v = ld_relaxed_sys(box)) == kArSentinel) {
        if ((++spins & 63u) == 0 && (globaltimer_ns() - t0 > comm.timeout_ns || comm_failed(me))) {
          comm_fail(me, 0xC000 + r);
          break;
        }
      }
    }
    st_relaxed_sys(box, kArSentinel);
    s_all[r][d] = __longlong_as_double((long long) v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      double tot = s_all[0][d];
      for (int r = 1; r < comm.world; ++r) tot = __dadd_rn(tot, s_all[r][d]);
      sums[d] = tot;
    }
    me->ar_seq = me->ar_seq + 1;
  }
}

} // namespace sb
