// sb_api.cu -- context, vectors and BLAS-1 entry points of the C ABI (include/stormb200.h).
#include "sb_kernels.cuh"

#include <cstdlib>
#include <mutex>

namespace sb {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

static __global__ void fill_sentinel_kernel(unsigned long long* p, int64_t n) {
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = kArSentinel;
}

int ensure_red_scratch(sb_ctx* ctx, int64_t n) {
  const int64_t tiles = num_tiles(n) > 0 ? num_tiles(n) : 1;
  if (tiles <= ctx->red.cap_tiles) return SB_OK;
  if (ctx->red.partials != nullptr) {
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    SB_CUDA(cudaFree(ctx->red.partials));
    ctx->red.partials = nullptr;
  }
  const int64_t cap = (tiles + tiles / 4 + 64 + 1) & ~(int64_t) 1; // even: every partial set starts 16-byte aligned (bulk copies)
  // two ordinary sets + the sentinel-managed set of the in-kernel reducer
  SB_CUDA(cudaMalloc(&ctx->red.partials, sizeof(double) * 3 * kMaxDots * cap));
  ctx->red.cap_tiles = cap;
  ctx->red.slots = ctx->red.partials + 2 * kMaxDots * cap;
  const int64_t words = kMaxDots * cap;
  fill_sentinel_kernel<<<(unsigned) ((words + 255) / 256), 256, 0, ctx->stream>>>(reinterpret_cast<unsigned long long*>(ctx->red.slots), words);
  SB_CUDA(cudaGetLastError());
  return SB_OK;
}

static RedPtrs red_ptrs(const sb_ctx* ctx) { return RedPtrs{ctx->red.partials, ctx->red.cap_tiles}; }

} // namespace sb

using namespace sb;

extern "C" {

const char* sb_last_error(void) { return sb::g_error; }
int sb_version(void) { return 100; }

int sb_ctx_create(int device, sb_ctx** out) {
  SB_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  int count = 0;
  SB_CUDA(cudaGetDeviceCount(&count));
  SB_REQUIRE(device >= 0 && device < count, "no such CUDA device");
  SB_CUDA(cudaSetDevice(device));
  sb_ctx* ctx = new sb_ctx();
  ctx->device = device;
  if (const char* dbg = std::getenv("SB_DEBUG")) ctx->debug = std::atoi(dbg);
  ctx->spin_timeout_ns = sb::kDefaultSpinTimeoutNs;
  if (const char* t = std::getenv("SB_SPIN_TIMEOUT_S")) {
    const double secs = std::atof(t);
    if (secs > 0.0) ctx->spin_timeout_ns = (unsigned long long) (secs * 1e9);
  }
  SB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  SB_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
  SB_CUDA(cudaMalloc(&ctx->red.result, sizeof(double) * 64));
  ctx->pinned_doubles = 4096;
  SB_CUDA(cudaMallocHost(&ctx->h_pinned, sizeof(double) * ctx->pinned_doubles));
  SB_CUDA(cudaEventCreate(&ctx->ev0));
  SB_CUDA(cudaEventCreate(&ctx->ev1));
  SB_TRY(ensure_red_scratch(ctx, 1 << 20));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = ctx;
  return SB_OK;
}

int sb_ctx_destroy(sb_ctx* ctx) {
  if (ctx == nullptr) return SB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (double* w : ctx->work) vec_free(ctx, w);
  ctx->work.clear();
  for (double* w : ctx->basis) vec_free(ctx, w);
  ctx->basis.clear();
  vec_cache_release(ctx);
  cudaFree(ctx->d_gmres_scal);
  cudaFree(ctx->d_gmres_ptrs);
  cudaFreeHost(ctx->h_gmres);
  comm_teardown(ctx);
  cudaFree(ctx->d_mega);
  cudaFree(ctx->d_timeline);
  cudaFree(ctx->d_solve);
  cudaFree(ctx->d_hist);
  cudaFree(ctx->d_trace);
  cudaFree(ctx->red.partials);
  cudaFree(ctx->red.result);
  cudaFreeHost(ctx->h_pinned);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SB_OK;
}

int sb_sync(sb_ctx* ctx) {
  SB_REQUIRE(ctx != nullptr, "ctx is null");
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

void* sb_ctx_stream(sb_ctx* ctx) { return ctx ? (void*) ctx->stream : nullptr; }
int64_t sb_ctx_launch_count(sb_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ---- vectors ------------------------------------------------------------------------------------
int sb_vec_alloc(sb_ctx* ctx, size_t n, double** d_out) {
  SB_REQUIRE(ctx != nullptr && d_out != nullptr, "null argument");
  return vec_alloc(ctx, n, d_out);
}

int sb_vec_free(sb_ctx* ctx, double* d) {
  SB_REQUIRE(ctx != nullptr, "ctx is null");
  return vec_free(ctx, d);
}

int sb_vec_upload(sb_ctx* ctx, double* d, const double* h_src, size_t n) {
  SB_REQUIRE(ctx != nullptr && d != nullptr && (h_src != nullptr || n == 0), "null argument");
  SB_CUDA(cudaMemcpyAsync(d, h_src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

int sb_vec_download(sb_ctx* ctx, const double* d, double* h_dst, size_t n) {
  SB_REQUIRE(ctx != nullptr && d != nullptr && (h_dst != nullptr || n == 0), "null argument");
  SB_CUDA(cudaMemcpyAsync(h_dst, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SB_OK;
}

} // extern "C"

// ---- expression evaluation ------------------------------------------------------------------------
namespace {

template<int NV, int AOP, class Prog>
int launch_eval(sb_ctx* ctx, double* y, size_t n, const sb_expr* e, const Prog& prog) {
  EvalBody<NV, AOP, Prog> body;
  body.y = y;
  for (int k = 0; k < NV; ++k) body.v[k] = e->vec[k] != nullptr ? e->vec[k] : y;
  for (int k = 0; k < SB_EXPR_MAX_SCAL; ++k) body.sc[k] = e->scal[k];
  body.prog = prog;
  const unsigned grid = (unsigned) num_tiles((int64_t) n);
  SB_CUDA(launch_kernel(ctx, ew_kernel<0, EvalBody<NV, AOP, Prog>>, grid, kThreads, 0, (int64_t) n, body, RedPtrs{},
                        (const int*) nullptr));
  ctx->launches++;
  return SB_OK;
}

template<int NV, class Prog>
int dispatch_aop(sb_ctx* ctx, double* y, size_t n, int aop, const sb_expr* e, const Prog& prog) {
  switch (aop) {
    case SB_ASSIGN: return launch_eval<NV, SB_ASSIGN>(ctx, y, n, e, prog);
    case SB_ADD_ASSIGN: return launch_eval<NV, SB_ADD_ASSIGN>(ctx, y, n, e, prog);
    case SB_SUB_ASSIGN: return launch_eval<NV, SB_SUB_ASSIGN>(ctx, y, n, e, prog);
    case SB_MUL_ASSIGN: return launch_eval<NV, SB_MUL_ASSIGN>(ctx, y, n, e, prog);
    case SB_DIV_ASSIGN: return launch_eval<NV, SB_DIV_ASSIGN>(ctx, y, n, e, prog);
  }
  set_error("sb_eval: unknown assign_op %d", aop);
  return SB_ERR_INVALID;
}

template<uint8_t... Ops>
bool prog_is(const sb_expr* e) {
  constexpr uint8_t code[] = {Ops...};
  if (e->n_ops != (int) sizeof...(Ops)) return false;
  return std::memcmp(e->ops, code, sizeof...(Ops)) == 0;
}

// The expression shapes the reference solvers actually build get fully unrolled
// kernels; anything else runs through the same evaluator with the program read at run time.
#define V0 SB_OP_VEC0
#define V1 SB_OP_VEC1
#define V2 SB_OP_VEC2
#define S0 SB_OP_SCAL0
#define S1 SB_OP_SCAL1
#define ADD SB_OP_ADD
#define SUB SB_OP_SUB
#define MUL SB_OP_MUL
#define DIV SB_OP_DIV
#define SB_STATIC_PROGRAMS(X)                       \
  X(1, V0)                                          \
  X(2, V0, S0, V1, MUL, ADD)                        \
  X(2, V0, S0, V1, MUL, SUB)                        \
  X(3, V0, S0, V1, S1, V2, MUL, SUB, MUL, ADD)      \
  X(3, V0, S0, V1, S1, V2, MUL, ADD, MUL, ADD)      \
  X(3, V0, S0, V1, S0, V2, MUL, ADD, MUL, ADD)      \
  X(3, V0, S0, V1, S0, V2, MUL, SUB, MUL, ADD)      \
  X(2, S0, V0, MUL, S1, V1, MUL, ADD)               \
  X(1, V0, S0, DIV)                                 \
  X(2, V0, V1, ADD)                                 \
  X(2, V0, V1, SUB)                                 \
  X(2, S0, V0, V1, SUB, MUL)                        \
  X(1, S0, V0, MUL)                                 \
  X(1, S0)

int eval_dispatch(sb_ctx* ctx, double* y, size_t n, int aop, const sb_expr* e, int nv) {
#define X(NV, ...) \
  if (nv <= NV && prog_is<__VA_ARGS__>(e)) return dispatch_aop<NV>(ctx, y, n, aop, e, StaticProg<__VA_ARGS__>{});
  SB_STATIC_PROGRAMS(X)
#undef X
  RuntimeProg prog;
  prog.n = e->n_ops;
  std::memcpy(prog.code, e->ops, SB_EXPR_MAX_OPS);
  switch (nv) {
    case 0:
    case 1: return dispatch_aop<1>(ctx, y, n, aop, e, prog);
    case 2: return dispatch_aop<2>(ctx, y, n, aop, e, prog);
    case 3: return dispatch_aop<3>(ctx, y, n, aop, e, prog);
    default: return dispatch_aop<4>(ctx, y, n, aop, e, prog);
  }
}
#undef V0
#undef V1
#undef V2
#undef S0
#undef S1
#undef ADD
#undef SUB
#undef MUL
#undef DIV

} // namespace

extern "C" {

int sb_eval(sb_ctx* ctx, double* y, size_t n, int assign_op, const sb_expr* e) {
  SB_REQUIRE(ctx != nullptr && y != nullptr && e != nullptr, "null argument");
  SB_REQUIRE(e->n_ops >= 1 && e->n_ops <= SB_EXPR_MAX_OPS, "expression length out of range");
  // validate: operands exist, stack never underflows / overflows, exactly one result
  int depth = 0, nv = 0;
  for (int k = 0; k < e->n_ops; ++k) {
    const int op = e->ops[k];
    if (op >= SB_OP_VEC0 && op <= SB_OP_VEC3) {
      SB_REQUIRE(e->vec[op] != nullptr, "expression references a null vector operand");
      if (op + 1 > nv) nv = op + 1;
      depth++;
    } else if (op >= SB_OP_SCAL0 && op <= SB_OP_SCAL3) {
      depth++;
    } else if (op == SB_OP_NEG) {
      SB_REQUIRE(depth >= 1, "expression stack underflow");
    } else if (op >= SB_OP_ADD && op <= SB_OP_DIV) {
      SB_REQUIRE(depth >= 2, "expression stack underflow");
      depth--;
    } else {
      set_error("sb_eval: unknown opcode %d", op);
      return SB_ERR_INVALID;
    }
    SB_REQUIRE(depth <= 6, "expression too deep (max stack depth 6)");
  }
  SB_REQUIRE(depth == 1, "expression must leave exactly one value");
  if (n == 0) return SB_OK;
  return eval_dispatch(ctx, y, n, assign_op, e, nv);
}

int sb_fill(sb_ctx* ctx, double* y, size_t n, double value) {
  SB_REQUIRE(ctx != nullptr && y != nullptr, "null argument");
  if (n == 0) return SB_OK;
  FillBody body{y, value};
  SB_CUDA(launch_kernel(ctx, ew_kernel<0, FillBody>, (unsigned) num_tiles((int64_t) n), kThreads, 0, (int64_t) n, body,
                        RedPtrs{}, (const int*) nullptr));
  ctx->launches++;
  return SB_OK;
}

int sb_copy(sb_ctx* ctx, double* y, const double* x, size_t n) {
  SB_REQUIRE(ctx != nullptr && y != nullptr && x != nullptr, "null argument");
  if (n == 0 || x == y) return SB_OK;
  SB_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  return SB_OK;
}

} // extern "C"

// ---- reductions -----------------------------------------------------------------------------------
namespace {
template<int M>
int launch_dots(sb_ctx* ctx, const double* const* a, const double* const* b, size_t n, double* d_out) {
  DotBody<M> body;
  for (int k = 0; k < M; ++k) body.a[k] = a[k], body.b[k] = b[k];
  SB_TRY(ensure_red_scratch(ctx, (int64_t) n));
  SB_CUDA(launch_kernel(ctx, ew_kernel<M, DotBody<M>>, (unsigned) num_tiles((int64_t) n), kThreads, 0, (int64_t) n, body,
                        red_ptrs(ctx), (const int*) nullptr));
  ctx->launches++;
  return launch_final<M>(ctx, (int64_t) n, StoreFinal<M>{d_out}, nullptr);
}
} // namespace

extern "C" {

int sb_dot_batch(sb_ctx* ctx, int m, const double* const* h_a, const double* const* h_b, size_t n,
                 double* h_out) {
  SB_REQUIRE(ctx != nullptr && h_a != nullptr && h_b != nullptr && h_out != nullptr, "null argument");
  SB_REQUIRE(m >= 1 && m <= 60, "batch size out of range (1..60)");
  for (int k = 0; k < m; ++k) SB_REQUIRE(h_a[k] != nullptr && h_b[k] != nullptr, "null vector");
  if (n == 0) {
    // multi-GPU: every reduction is a collective (in-kernel all-reduce / ncclAllReduce); a rank that skipped it would
    // leave its peers waiting, so empty ranks are refused here and in sb_dist_op_create
    if (ctx->comm.world > 1) {
      set_error("sb_dot_batch: n = 0 on a multi-GPU context (reductions are collective: every rank must own rows)");
      return SB_ERR_INVALID;
    }
    for (int k = 0; k < m; ++k) h_out[k] = 0.0;
    return SB_OK;
  }
  int k = 0;
  while (k < m) {
    const int left = m - k;
    if (left >= 3) {
      SB_TRY(launch_dots<3>(ctx, h_a + k, h_b + k, n, ctx->red.result + k));
      k += 3;
    } else if (left == 2) {
      SB_TRY(launch_dots<2>(ctx, h_a + k, h_b + k, n, ctx->red.result + k));
      k += 2;
    } else {
      SB_TRY(launch_dots<1>(ctx, h_a + k, h_b + k, n, ctx->red.result + k));
      k += 1;
    }
  }
  SB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->red.result, sizeof(double) * m, cudaMemcpyDeviceToHost,
                          ctx->stream));
  SB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int q = 0; q < m; ++q) h_out[q] = ctx->h_pinned[q];
  return SB_OK;
}

int sb_eval_group(sb_ctx* ctx, size_t n, int n_stmt, const sb_chain* h_stmts, int n_dots,
                  const double* const* h_dot_a, const double* const* h_dot_b, double* h_out) {
  SB_REQUIRE(ctx != nullptr, "null argument");
  SB_REQUIRE(n_stmt >= 0 && n_stmt <= SB_GROUP_MAX_STMT, "statement count out of range (0..8)");
  SB_REQUIRE(n_dots >= 0 && n_dots <= SB_GROUP_MAX_DOTS, "dot count out of range (0..8)");
  SB_REQUIRE(n_stmt + n_dots > 0, "empty group");
  SB_REQUIRE(n_stmt == 0 || h_stmts != nullptr, "null statement array");
  SB_REQUIRE(n_dots == 0 || (h_dot_a != nullptr && h_dot_b != nullptr && h_out != nullptr), "null dot arrays");
  GroupBody body{};
  for (int s = 0; s < n_stmt; ++s) {
    const sb_chain& ch = h_stmts[s];
    SB_REQUIRE(ch.y != nullptr, "statement without a target");
    SB_REQUIRE(ch.n_terms >= 1 && ch.n_terms <= SB_GROUP_MAX_TERMS, "term count out of range (1..8)");
    for (int t = 0; t < ch.n_terms; ++t) SB_REQUIRE(ch.x[t] != nullptr, "null term vector");
    SB_REQUIRE(ch.base != nullptr || ch.sub[0] == 0, "a chain without a base cannot start with a subtraction");
    body.st[s] = ch;
  }
  for (int d = 0; d < n_dots; ++d) SB_REQUIRE(h_dot_a[d] != nullptr && h_dot_b[d] != nullptr, "null vector");
  if (n == 0) {
    if (ctx->comm.world > 1 && n_dots > 0) {
      set_error("sb_eval_group: n = 0 on a multi-GPU context (reductions are collective: every rank must own rows)");
      return SB_ERR_INVALID;
    }
    for (int d = 0; d < n_dots; ++d) h_out[d] = 0.0;
    return SB_OK;
  }
  // the first kMaxDots dots ride on the statement kernel; the rest (long batches) are stand-alone dot kernels
  const int fused = n_stmt > 0 ? std::min(n_dots, kMaxDots) : 0;
  body.n_stmt = n_stmt, body.n_dots = fused;
  for (int d = 0; d < fused; ++d) body.da[d] = h_dot_a[d], body.db[d] = h_dot_b[d];
  const unsigned grid = (unsigned) num_tiles((int64_t) n);
  if (n_stmt > 0) {
    if (fused > 0) SB_TRY(ensure_red_scratch(ctx, (int64_t) n));
    const RedPtrs red = fused > 0 ? red_ptrs(ctx) : RedPtrs{};
    switch (fused) {
      case 0:
        SB_CUDA(launch_kernel(ctx, ew_kernel<0, GroupBody>, grid, kThreads, 0, (int64_t) n, body, red, (const int*) nullptr));
        break;
      case 1:
        SB_CUDA(launch_kernel(ctx, ew_kernel<1, GroupBody>, grid, kThreads, 0, (int64_t) n, body, red, (const int*) nullptr));
        break;
      case 2:
        SB_CUDA(launch_kernel(ctx, ew_kernel<2, GroupBody>, grid, kThreads, 0, (int64_t) n, body, red, (const int*) nullptr));
        break;
      default:
        SB_CUDA(launch_kernel(ctx, ew_kernel<3, GroupBody>, grid, kThreads, 0, (int64_t) n, body, red, (const int*) nullptr));
        break;
    }
    ctx->launches++;
    if (fused == 1) SB_TRY(launch_final<1>(ctx, (int64_t) n, StoreFinal<1>{ctx->red.result}, nullptr));
    if (fused == 2) SB_TRY(launch_final<2>(ctx, (int64_t) n, StoreFinal<2>{ctx->red.result}, nullptr));
    if (fused == 3) SB_TRY(launch_final<3>(ctx, (int64_t) n, StoreFinal<3>{ctx->red.result}, nullptr));
  }
  int k = fused;
  while (k < n_dots) {
    const int left = n_dots - k;
    if (left >= 3) {
      SB_TRY(launch_dots<3>(ctx, h_dot_a + k, h_dot_b + k, n, ctx->red.result + k));
      k += 3;
    } else if (left == 2) {
      SB_TRY(launch_dots<2>(ctx, h_dot_a + k, h_dot_b + k, n, ctx->red.result + k));
      k += 2;
    } else {
      SB_TRY(launch_dots<1>(ctx, h_dot_a + k, h_dot_b + k, n, ctx->red.result + k));
      k += 1;
    }
  }
  if (n_dots > 0) {
    SB_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->red.result, sizeof(double) * n_dots, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < n_dots; ++q) h_out[q] = ctx->h_pinned[q];
  }
  return SB_OK;
}

int sb_dot(sb_ctx* ctx, const double* a, const double* b, size_t n, double* h_out) {
  return sb_dot_batch(ctx, 1, &a, &b, n, h_out);
}

int sb_norm2(sb_ctx* ctx, const double* a, size_t n, double* h_out) {
  SB_REQUIRE(h_out != nullptr, "null argument");
  double s = 0.0;
  SB_TRY(sb_dot_batch(ctx, 1, &a, &a, n, &s));
  *h_out = sqrt(s); // norm_2 = sqrt(sum |a_i|^2), MatrixAlgorithms.hpp:262-270
  return SB_OK;
}

} // extern "C"
