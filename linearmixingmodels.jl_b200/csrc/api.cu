// liblmm C ABI (include/lmm.h): contexts, posterior handles, host-side orchestration of the
// batched blocked Cholesky and of the OILMM / IndependentMOGP / ILMM inference path.
// There is no CPU fallback anywhere in this file: every compute entry point needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lmm.h"
#include "kernels.h"

using namespace lmm;

namespace {
const double LOG2PI = 1.8378770664093453;  // log(2π)
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen: no link-time dependency; picks up the libnccl.so.2 already loaded by the
// host process (torch bundles one) or the system one.
// ------------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommSplit)(void*, int, int, void**, void*) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = (int (*)(NcclId*))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(api.handle, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
    api.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllGather");
    api.CommSplit = (int (*)(void*, int, int, void**, void*))dlsym(api.handle, "ncclCommSplit");
    api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
  });
  return api;
}
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct lmm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::string err;
  int distance_form = 0;
  int outer_block = 8;
  bool outer_block_user = false;
  int nranks = 1, rank = 0;
  void* comm = nullptr;
  void* comm_small = nullptr;  // few-CTA communicator for the small, latency-critical exchanges on the panel chain
  int nccl_small_ctas = 0;  // 0: NCCL's own choice
  int profile_partition = 0;  // option "profile_partition": per-phase CUDA-event times of the row-cyclic schedule on stderr
  void* comm2 = nullptr;  // second communicator (ncclCommSplit): the large exchanges of the partitioned factorisation, which
                          // overlap the panel chain's small ones on another stream
  cudaStream_t xchg_stream = nullptr;
  void* xbuf2 = nullptr;
  size_t xbuf2_bytes = 0;
  int64_t launches = 0, h2d = 0, d2h = 0;
  double timings[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t ev[8];
  // latent groups run their (latency-bound) panel steps on separate streams so that one group's
  // diagonal-tile factorisation overlaps the other groups' trailing updates
  static constexpr int MAX_GROUPS = 8;
  int ngroups = 4;
  cudaStream_t gstream[MAX_GROUPS];
  cudaEvent_t ev_fork, ev_join[MAX_GROUPS];
  // block-level look-ahead for small batches (ILMM: batch 1): panel stream (high priority) +
  // trailing-update stream, chained by per-block events
  cudaStream_t panel_stream = nullptr, update_stream = nullptr;
  std::vector<cudaEvent_t> blk_ev;
  int lookahead = 2;  // 0 off, 1 left-looking K-split, 2 right-looking (default)
  // one large factor (general ILMM, batch 1) partitioned row-cyclically over the ranks of the communicator
  int partition_ilmm = 0;
  int partition_now = 0;  // set by the callers whose factorisation is replicated on every rank (ILMM joint factor)
  int dist_error = 0;  // NCCL failure inside the partitioned schedule (reported by the caller)
  void* xbuf = nullptr;  // exchange buffers of the row-cyclic schedule (send | all-gathered), grown on demand
  size_t xbuf_bytes = 0;

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  int fail_cuda(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s at api.cu:%d (%s)", cudaGetErrorString(e), line, what);
    err = buf;
    cudaGetLastError();  // clear non-sticky error state
    return e == cudaErrorMemoryAllocation ? LMM_E_OOM : LMM_E_CUDA;
  }
};

#define CU(expr)                                                        \
  do {                                                                  \
    cudaError_t e__ = (expr);                                           \
    if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #expr, __LINE__); \
  } while (0)

namespace {

struct DevBuf {
  lmm_ctx* c = nullptr;
  void* p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, c->stream);
    p = nullptr;
  }
  cudaError_t alloc(lmm_ctx* ctx, size_t bytes) {
    release();
    c = ctx;
    if (bytes == 0) bytes = 8;
    return cudaMallocAsync(&p, bytes, ctx->stream);
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
  void* detach() {
    void* q = p;
    p = nullptr;
    return q;
  }
};

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Copy n doubles from a caller pointer (host or device) into device memory at dst.
cudaError_t copy_in(lmm_ctx* ctx, double* dst, const double* src, size_t n) {
  if (is_device_ptr(src)) return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
  ctx->h2d += (int64_t)(n * sizeof(double));
  return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
}
cudaError_t copy_out(lmm_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  ctx->d2h += (int64_t)bytes;
  return cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream);
}

inline int ntiles(int n) { return (n + TILE - 1) / TILE; }

void shard_range(const lmm_ctx* ctx, int m, int& lo, int& hi) {
  lo = (int)(((int64_t)m * ctx->rank) / ctx->nranks);
  hi = (int)(((int64_t)m * (ctx->rank + 1)) / ctx->nranks);
}

int check_descs(lmm_ctx* ctx, const lmm_gp_desc* d, int m) {
  for (int i = 0; i < m; ++i) {
    if (d[i].kind < 0 || d[i].kind > 2) return ctx->fail(LMM_E_UNSUPPORTED, "unsupported kernel kind (only SE, Matern32, Matern52)");
    if (!(d[i].variance > 0.0) || !(d[i].inv_lengthscale > 0.0)) return ctx->fail(LMM_E_ARG, "kernel variance and inv_lengthscale must be positive");
  }
  return LMM_OK;
}

// lml_i = -(N log2π + logdet_i + quad_i)/2 ; optional regulariser slot.
__global__ void lml_terms_kernel(double* terms, int slot0, int nb, const double* logdet, const double* quad, int n, double log2pi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) terms[slot0 + i] = -((double)n * log2pi + logdet[i] + quad[i]) / 2.0;
}
__global__ void regulariser_kernel(double* slot, double c0, const double* resid, double sigma2) {
  slot[0] = -(c0 + resid[0] / sigma2) / 2.0;
}
__global__ void add_scalar_kernel(double* v, size_t n, double s) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += s;
}
// out[i*stride_out + n] = a[i*stride_in + n] + s  (latent-major copy with offset)
__global__ void copy_add_kernel(double* out, size_t stride_out, const double* in, size_t stride_in, int n, double s) {
  const int i = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[(size_t)i * stride_out + k] = in[(size_t)i * stride_in + k] + s;
}
// v[i][k] = mean_i + v[i][k]   and   w = a*x + y helpers for rand
__global__ void add_mean_kernel(double* v, size_t stride, int n, const LatentParams* params) {
  const int i = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) v[(size_t)i * stride + k] += params[i].mean;
}
__global__ void axpy_kernel(double* y, const double* x, size_t n, double a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fma(a, x[i], y[i]);
}

// ---- batched blocked Cholesky (left-looking over blocks of `outer_block` tile columns) -------
// For every block column [s0, s1): one wide trailing update against all previous columns
// (K = s0 tiles, output written once), then per tile column: narrow update inside the block,
// diagonal-tile factor (+ inverse, logdet, info), panel TRSM as a GEMM with the inverse.
cudaError_t chol_factor_stream(lmm_ctx* ctx, cudaStream_t st, TiledSym L, double* W, size_t wstride, int batch, double* logdet,
                               int* info) {
  const int nt = L.nt, ob = ctx->outer_block;
  GemmArgs g{};
  g.A = operand(L);
  g.B = operand(L);
  g.C = operand(L);
  g.W = W;
  g.w_batch_stride = wstride;
  g.sym = 1;
  cudaError_t e;
  for (int s0 = 0; s0 < nt; s0 += ob) {
    const int s1 = (s0 + ob < nt) ? s0 + ob : nt;
    if (s0 > 0) {
      g.i0 = s0; g.j0 = s0; g.k0 = 0; g.k1 = s0;
      if ((e = launch_gemm(st, GEMM_UPDATE, g, s1 - s0, nt - s0, batch)) != cudaSuccess) return e;
      ++ctx->launches;
      ctx->timings[6] += 1;
    }
    for (int jj = s0; jj < s1; ++jj) {
      if (jj > s0) {
        g.i0 = jj; g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(st, GEMM_UPDATE, g, 1, nt - jj, batch)) != cudaSuccess) return e;
        ++ctx->launches;
        ctx->timings[6] += 1;
      }
      if ((e = launch_potrf_tile(st, L, W, wstride, jj, batch, logdet, info)) != cudaSuccess) return e;
      ++ctx->launches;
      if (jj + 1 < nt) {
        g.i0 = jj + 1; g.j0 = jj;
        if ((e = launch_gemm(st, GEMM_TRSM, g, 1, nt - jj - 1, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
    }
  }
  return cudaSuccess;
}

// Block-level look-ahead (small batches: nothing else can hide the panel latency).  The wide
// update of block column b is split along K: part A (all columns before block b-1) runs on the
// update stream concurrently with the latency-bound panel steps of block b-1 on the high-priority
// panel stream; part B (the columns of block b-1) follows on the panel stream.
cudaError_t chol_factor_lookahead(lmm_ctx* ctx, TiledSym L, double* W, size_t wstride, int batch, double* logdet, int* info) {
  const int nt = L.nt;
  // the panel chain is the critical path here: narrower blocks for smaller matrices (measured)
  const int ob = ctx->outer_block_user ? ctx->outer_block : (nt <= 40 ? 3 : nt <= 96 ? 6 : 8);
  const int nblk = (nt + ob - 1) / ob;
  cudaError_t e;
  while ((int)ctx->blk_ev.size() < 2 * nblk + 2) {
    cudaEvent_t ev;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    ctx->blk_ev.push_back(ev);
  }
  cudaStream_t X = ctx->panel_stream, Y = ctx->update_stream;
  cudaEvent_t* evI = ctx->blk_ev.data();          // inner(b) done on X
  cudaEvent_t* evA = ctx->blk_ev.data() + nblk;   // part A(b) done on Y
  GemmArgs g{};
  g.A = operand(L); g.B = operand(L); g.C = operand(L);
  g.W = W; g.w_batch_stride = wstride; g.sym = 1;
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(X, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(Y, ctx->ev_fork, 0)) != cudaSuccess) return e;
  bool y_used = false;
  for (int b = 0; b < nblk; ++b) {
    const int s0 = b * ob, s1 = (s0 + ob < nt) ? s0 + ob : nt;
    const int sp = (b >= 1) ? (b - 1) * ob : 0;  // first column of block b-1
    if (b >= 2) {  // part A on Y: k in [0, sp)
      if ((e = cudaStreamWaitEvent(Y, evI[b - 2], 0)) != cudaSuccess) return e;
      g.i0 = s0; g.j0 = s0; g.k0 = 0; g.k1 = sp;
      if ((e = launch_gemm(Y, GEMM_UPDATE, g, s1 - s0, nt - s0, batch)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(evA[b], Y)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(X, evA[b], 0)) != cudaSuccess) return e;
      ++ctx->launches;
      ctx->timings[6] += 1;
      y_used = true;
    }
    if (b >= 1) {  // part B on X: k in [sp, s0)
      g.i0 = s0; g.j0 = s0; g.k0 = sp; g.k1 = s0;
      if ((e = launch_gemm(X, GEMM_UPDATE, g, s1 - s0, nt - s0, batch)) != cudaSuccess) return e;
      ++ctx->launches;
      ctx->timings[6] += 1;
    }
    for (int jj = s0; jj < s1; ++jj) {
      if (jj > s0) {
        g.i0 = jj; g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, nt - jj, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      if ((e = launch_potrf_tile(X, L, W, wstride, jj, batch, logdet, info)) != cudaSuccess) return e;
      ++ctx->launches;
      if (jj + 1 < nt) {
        g.i0 = jj + 1; g.j0 = jj;
        if ((e = launch_gemm(X, GEMM_TRSM, g, 1, nt - jj - 1, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
    }
    if ((e = cudaEventRecord(evI[b], X)) != cudaSuccess) return e;
  }
  if ((e = cudaStreamWaitEvent(ctx->stream, evI[nblk - 1], 0)) != cudaSuccess) return e;
  if (y_used) {
    if ((e = cudaEventRecord(ctx->ev_join[0], Y)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Trailing update of the columns >= s2 (tile rows first_row, first_row + row_step, ... < nt) by the k-tiles [k0, k1): ONE
// launch.  (Chunking it into launches of <= 132 CTAs, to keep a few SMs free for the panel chain on the other stream, was
// measured and is much slower -- N=16384: 48 -> 68 ms, N=8192: 8.5 -> 10.3 ms: every launch boundary costs a pipeline
// fill and a tail, while a single launch keeps the block scheduler streaming CTAs.)
cudaError_t launch_trailing(lmm_ctx* ctx, cudaStream_t st, GemmArgs g, int s2, int nt, int first_row, int row_step, int k0, int k1,
                            int batch) {
  const int nrows = first_row >= nt ? 0 : (nt - 1 - first_row) / row_step + 1;
  if (nrows <= 0) return cudaSuccess;
  g.i0 = first_row; g.j0 = s2; g.k0 = k0; g.k1 = k1; g.row_step = row_step;
  cudaError_t e = launch_gemm(st, GEMM_UPDATE, g, nt - s2, nrows, batch);
  if (e != cudaSuccess) return e;
  ++ctx->launches;
  ctx->timings[6] += 1;
  return cudaSuccess;
}

// Right-looking block schedule with look-ahead (small batches).  After block column kb is factored on the
// high-priority panel stream X, its update of the NEXT block column runs on X (so the next panel can start at
// once) while its update of everything further right runs as one large GEMM on the low-priority stream Y:
//   X: [wait Y(kb-2)] update(kb-1 -> kb), panel(kb)            Y: [wait X(kb)] update(kb -> kb+2 .. end)
// The Y launches have thousands of tiles (no tail effect, unlike the wide left-looking update of one block
// column) and keep every SM busy while the latency-bound panel steps run beside them; each C tile is
// read-modify-written once per block column of L (K = `ob` tiles per launch).
cudaError_t chol_factor_rightlooking(lmm_ctx* ctx, TiledSym L, double* W, size_t wstride, int batch, double* logdet, int* info) {
  const int nt = L.nt;
  // the panel chain is the critical path: narrow blocks for small matrices, wider ones (fewer read-modify-write
  // passes over the trailing matrix) once the trailing GEMMs dominate (measured: tools/bench_batch1.py)
  const int ob = ctx->outer_block_user ? ctx->outer_block : (nt <= 32 ? 1 : nt <= 72 ? 2 : nt <= 112 ? 3 : 4);
  const int nblk = (nt + ob - 1) / ob;
  cudaError_t e;
  while ((int)ctx->blk_ev.size() < 2 * nblk + 2) {
    cudaEvent_t ev;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    ctx->blk_ev.push_back(ev);
  }
  cudaStream_t X = ctx->panel_stream, Y = ctx->update_stream;
  cudaEvent_t* evX = ctx->blk_ev.data();          // panel(b) done on X
  cudaEvent_t* evY = ctx->blk_ev.data() + nblk;   // trailing update from block b done on Y
  GemmArgs g{};
  g.A = operand(L); g.B = operand(L); g.C = operand(L);
  g.W = W; g.w_batch_stride = wstride; g.sym = 1;
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(X, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(Y, ctx->ev_fork, 0)) != cudaSuccess) return e;
  for (int b = 0; b < nblk; ++b) {
    const int s0 = b * ob, s1 = (s0 + ob < nt) ? s0 + ob : nt;
    if (b >= 1) {
      // every earlier update of this block column (Y launches up to b-2) must have landed
      if (b >= 2 && (e = cudaStreamWaitEvent(X, evY[b - 2], 0)) != cudaSuccess) return e;
      g.i0 = s0; g.j0 = s0; g.k0 = s0 - ob; g.k1 = s0;
      if ((e = launch_gemm(X, GEMM_UPDATE, g, s1 - s0, nt - s0, batch)) != cudaSuccess) return e;
      ++ctx->launches;
      ctx->timings[6] += 1;
    }
    for (int jj = s0; jj < s1; ++jj) {
      if (jj > s0) {
        g.i0 = jj; g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, nt - jj, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      if ((e = launch_potrf_tile(X, L, W, wstride, jj, batch, logdet, info)) != cudaSuccess) return e;
      ++ctx->launches;
      if (jj + 1 < nt) {
        g.i0 = jj + 1; g.j0 = jj;
        if ((e = launch_gemm(X, GEMM_TRSM, g, 1, nt - jj - 1, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
    }
    if ((e = cudaEventRecord(evX[b], X)) != cudaSuccess) return e;
    const int s2 = s1 + ob;  // first column of block b+2
    if (s2 < nt) {
      if ((e = cudaStreamWaitEvent(Y, evX[b], 0)) != cudaSuccess) return e;
      if ((e = launch_trailing(ctx, Y, g, s2, nt, s2, 1, s0, s1, batch)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(evY[b], Y)) != cudaSuccess) return e;
    } else if ((e = cudaEventRecord(evY[b], Y)) != cudaSuccess) {  // no trailing launch left: keep the event chain defined
      return e;
    }
  }
  if ((e = cudaStreamWaitEvent(ctx->stream, evX[nblk - 1], 0)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(ctx->ev_join[0], Y)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0)) != cudaSuccess) return e;
  return cudaSuccess;
}

// Row-cyclic multi-GPU factorisation of ONE large matrix (north star: "ILMM runs on one GPU unless its blocked
// Cholesky is explicitly row-cyclic partitioned").  Every rank holds the whole packed-lower matrix and runs the same
// right-looking schedule; rank r owns the tile rows I = r (mod G) of the TRAILING matrix and applies the updates to
// those rows only.  Before block column b is factored its tiles are exchanged (pack own rows -> ncclAllGather over
// NVLink -> unpack the others' rows); the latency-bound panel (diagonal-tile factor, TRSM-as-GEMM of all rows, 2-3 %
// of the flops) is then computed redundantly by every rank, so the finished columns of L are complete everywhere and
// nothing downstream (solves, predictions, logdet) needs a collective.  Per block: one all-gather of (nt - s0) * ob
// tiles; the trailing GEMMs -- 97 % of the work -- are split G ways.
cudaError_t chol_factor_rowcyclic(lmm_ctx* ctx, TiledSym L, double* W, size_t wstride, double* logdet, int* info) {
  const int nt = L.nt, G = ctx->nranks, me = ctx->rank;
  const int ob = ctx->outer_block_user ? ctx->outer_block : (nt <= 72 ? 2 : nt <= 112 ? 3 : 4);
  const int nblk = (nt + ob - 1) / ob;
  cudaError_t e;
  while ((int)ctx->blk_ev.size() < 2 * nblk + 2) {
    cudaEvent_t ev;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    ctx->blk_ev.push_back(ev);
  }
  // exchange buffers: [send: slots*ob tiles][recv: G*slots*ob tiles], sized for the first (largest) exchange
  const int max_slots = (nt + G - 1) / G;
  const size_t send_elems = (size_t)max_slots * ob * TT, need = (send_elems * (size_t)(G + 1)) * sizeof(double);
  if (ctx->xbuf_bytes < need) {
    if (ctx->xbuf) cudaFree(ctx->xbuf);
    ctx->xbuf = nullptr;
    ctx->xbuf_bytes = 0;
    if ((e = cudaMalloc(&ctx->xbuf, need)) != cudaSuccess) return e;
    ctx->xbuf_bytes = need;
  }
  double* sendb = (double*)ctx->xbuf;
  double* recvb = sendb + send_elems;
  cudaStream_t X = ctx->panel_stream, Y = ctx->update_stream;
  cudaEvent_t* evX = ctx->blk_ev.data();
  cudaEvent_t* evY = ctx->blk_ev.data() + nblk;
  GemmArgs g{};
  g.A = operand(L); g.B = operand(L); g.C = operand(L);
  g.W = W; g.w_batch_stride = wstride; g.sym = 1;
  auto first_own = [&](int s) { return s + (((me - s % G) % G) + G) % G; };
  auto own_count = [&](int s) { const int f = first_own(s); return f >= nt ? 0 : (nt - 1 - f) / G + 1; };
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(X, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(Y, ctx->ev_fork, 0)) != cudaSuccess) return e;
  for (int b = 0; b < nblk; ++b) {
    const int s0 = b * ob, s1 = (s0 + ob < nt) ? s0 + ob : nt;
    if (b >= 1) {
      if (b >= 2 && (e = cudaStreamWaitEvent(X, evY[b - 2], 0)) != cudaSuccess) return e;
      // own rows of block column b: update with block b-1 ...
      const int cnt = own_count(s0);
      if (cnt > 0) {
        g.row_step = G; g.i0 = first_own(s0); g.j0 = s0; g.k0 = s0 - ob; g.k1 = s0;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, s1 - s0, cnt, 1)) != cudaSuccess) return e;
        ++ctx->launches;
        ctx->timings[6] += 1;
      }
      // ... then exchange the block column so that every rank can factor it
      const int slots = (nt - s0 + G - 1) / G;
      if ((e = launch_rowcyclic_pack(X, L, s0, s1, s0, nt, G, me, slots, sendb)) != cudaSuccess) return e;
      const size_t cntel = (size_t)slots * (s1 - s0) * TT;
      if (nccl_api().AllGather(sendb, recvb, cntel, NCCL_DOUBLE, ctx->comm_small ? ctx->comm_small : ctx->comm, X) != 0) {
        ctx->dist_error = 1;
        return cudaErrorUnknown;
      }
      if ((e = launch_rowcyclic_unpack(X, L, s0, s1, s0, nt, G, me, slots, recvb)) != cudaSuccess) return e;
      ctx->launches += 2;
    }
    g.row_step = 1;
    for (int jj = s0; jj < s1; ++jj) {  // the panel: every rank, all rows
      if (jj > s0) {
        g.i0 = jj; g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, nt - jj, 1)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      if ((e = launch_potrf_tile(X, L, W, wstride, jj, 1, logdet, info)) != cudaSuccess) return e;
      ++ctx->launches;
      if (jj + 1 < nt) {
        g.i0 = jj + 1; g.j0 = jj;
        if ((e = launch_gemm(X, GEMM_TRSM, g, 1, nt - jj - 1, 1)) != cudaSuccess) return e;
        ++ctx->launches;
      }
    }
    if ((e = cudaEventRecord(evX[b], X)) != cudaSuccess) return e;
    const int s2 = s1 + ob;
    const int cnt2 = s2 < nt ? own_count(s2) : 0;
    if (cnt2 > 0) {
      if ((e = cudaStreamWaitEvent(Y, evX[b], 0)) != cudaSuccess) return e;
      if ((e = launch_trailing(ctx, Y, g, s2, nt, first_own(s2), G, s0, s1, 1)) != cudaSuccess) return e;
    }
    if ((e = cudaEventRecord(evY[b], Y)) != cudaSuccess) return e;
  }
  if ((e = cudaStreamWaitEvent(ctx->stream, evX[nblk - 1], 0)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(ctx->ev_join[0], Y)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0)) != cudaSuccess) return e;
  return cudaSuccess;
}

// Second row-cyclic schedule ("partition_ilmm" = 2): the panel's TRSM is distributed as well and the large exchange leaves
// the critical path.  Per block column b = [s0, s1), next block [s1, s2):
//   X (panel stream, communicator 1): update(b-1 -> b) on own rows; all-gather of the DIAGONAL block rows [s0, s1) (<= ob
//     tile rows); diagonal block factored redundantly; TRSM of the OWN rows >= s1; all-gather of the NEXT block's rows
//     [s1, s2) of the finished panel -- all the next update(b -> b+1) needs besides the own rows.
//   Z (exchange stream, communicator 2): all-gather of the rows >= s2 of the finished panel -- the bulk of the data --
//     concurrently with the next panel; it only gates
//   Y (update stream): update(b -> b+2..end) on own rows.
// Everything on the panel chain is small (<= 2 ob tile rows exchanged, 1/G of the TRSM and update waves).
cudaError_t chol_factor_rowcyclic2(lmm_ctx* ctx, TiledSym L, double* W, size_t wstride, double* logdet, int* info) {
  const int nt = L.nt, G = ctx->nranks, me = ctx->rank;
  const int ob = ctx->outer_block_user ? ctx->outer_block : (nt <= 72 ? 2 : nt <= 160 ? 3 : 4);
  const int nblk = (nt + ob - 1) / ob;
  cudaError_t e;
  while ((int)ctx->blk_ev.size() < 3 * nblk + 3) {
    cudaEvent_t ev;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    ctx->blk_ev.push_back(ev);
  }
  auto grow = [&](void*& buf, size_t& have, size_t need) -> cudaError_t {
    if (have >= need) return cudaSuccess;
    if (buf) cudaFree(buf);
    buf = nullptr;
    have = 0;
    cudaError_t ee = cudaMalloc(&buf, need);
    if (ee == cudaSuccess) have = need;
    return ee;
  };
  // small exchanges (<= ob rows): [send | recv]; large ones: sized for the first block
  const int small_slots = (2 * ob + G - 1) / G;
  const size_t small_send = (size_t)small_slots * ob * TT;
  if ((e = grow(ctx->xbuf, ctx->xbuf_bytes, small_send * (size_t)(G + 1) * sizeof(double))) != cudaSuccess) return e;
  const int big_slots0 = (nt + G - 1) / G;
  const size_t big_send = (size_t)big_slots0 * ob * TT;
  if ((e = grow(ctx->xbuf2, ctx->xbuf2_bytes, big_send * (size_t)(G + 1) * sizeof(double))) != cudaSuccess) return e;
  double* ssend = (double*)ctx->xbuf;
  double* srecv = ssend + small_send;
  double* bsend = (double*)ctx->xbuf2;
  double* brecv = bsend + big_send;
  cudaStream_t X = ctx->panel_stream, Y = ctx->update_stream, Z = ctx->xchg_stream;
  cudaEvent_t* evX = ctx->blk_ev.data();
  cudaEvent_t* evY = evX + nblk;
  cudaEvent_t* evZ = evY + nblk;
  GemmArgs g{};
  g.A = operand(L); g.B = operand(L); g.C = operand(L);
  g.W = W; g.w_batch_stride = wstride; g.sym = 1;
  auto first_own = [&](int s) { return s + (((me - s % G) % G) + G) % G; };
  auto own_count = [&](int s) { const int f = first_own(s); return f >= nt ? 0 : (nt - 1 - f) / G + 1; };
  auto gather = [&](cudaStream_t st, void* comm, int s0, int s1, int ra, int rb, double* sendb, double* recvb) -> cudaError_t {
    if (rb <= ra) return cudaSuccess;
    const int slots = (rb - ra + G - 1) / G;
    cudaError_t ee;
    if ((ee = launch_rowcyclic_pack(st, L, s0, s1, ra, rb, G, me, slots, sendb)) != cudaSuccess) return ee;
    if (nccl_api().AllGather(sendb, recvb, (size_t)slots * (s1 - s0) * TT, NCCL_DOUBLE, comm, st) != 0) {
      ctx->dist_error = 1;
      return cudaErrorUnknown;
    }
    if ((ee = launch_rowcyclic_unpack(st, L, s0, s1, ra, rb, G, me, slots, recvb)) != cudaSuccess) return ee;
    ctx->launches += 2;
    return cudaSuccess;
  };
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(X, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(Y, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(Z, ctx->ev_fork, 0)) != cudaSuccess) return e;
  // optional phase profile of the panel chain: [0] wait for the trailing update, [1] own-row update, [2] exchange,
  // [3] redundant diagonal / next-block rows, [4] own-row TRSM
  std::vector<cudaEvent_t> pev;
  const bool prof = ctx->profile_partition != 0;
  auto mark = [&]() {
    if (!prof) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, X);
    pev.push_back(ev);
  };
  for (int b = 0; b < nblk; ++b) {
    const int s0 = b * ob, s1 = (s0 + ob < nt) ? s0 + ob : nt, s2 = (s1 + ob < nt) ? s1 + ob : nt;
    mark();
    if (b >= 2 && (e = cudaStreamWaitEvent(X, evY[b - 2], 0)) != cudaSuccess) return e;
    mark();
    if (b >= 1) {
      const int cnt = own_count(s0);
      if (cnt > 0) {  // own rows of block column b <- block b-1 (B operand rows [s0, s1) arrived with the previous panel)
        g.row_step = G; g.i0 = first_own(s0); g.j0 = s0; g.k0 = s0 - ob; g.k1 = s0;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, s1 - s0, cnt, 1)) != cudaSuccess) return e;
        ++ctx->launches;
        ctx->timings[6] += 1;
      }
      mark();
      // the diagonal block AND the next block's rows, in one small exchange
      if ((e = gather(X, ctx->comm_small ? ctx->comm_small : ctx->comm, s0, s1, s0, s2, ssend, srecv)) != cudaSuccess) return e;
    }
    if (b == 0) mark();
    mark();
    g.row_step = 1;
    for (int jj = s0; jj < s1; ++jj) {  // diagonal block and the next block's rows [s1, s2): every rank
      if (jj > s0) {
        g.i0 = jj; g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, s2 - jj, 1)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      if ((e = launch_potrf_tile(X, L, W, wstride, jj, 1, logdet, info)) != cudaSuccess) return e;
      ++ctx->launches;
      if (jj + 1 < s2) {
        g.i0 = jj + 1; g.j0 = jj;
        if ((e = launch_gemm(X, GEMM_TRSM, g, 1, s2 - jj - 1, 1)) != cudaSuccess) return e;
        ++ctx->launches;
      }
    }
    mark();
    const int cnt1 = s2 < nt ? own_count(s2) : 0;
    if (cnt1 > 0) {  // own rows below: in-block updates + TRSM, column by column
      g.row_step = G; g.i0 = first_own(s2);
      for (int jj = s0; jj < s1; ++jj) {
        if (jj > s0) {
          g.j0 = jj; g.k0 = s0; g.k1 = jj;
          if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, cnt1, 1)) != cudaSuccess) return e;
          ++ctx->launches;
        }
        g.j0 = jj;
        if ((e = launch_gemm(X, GEMM_TRSM, g, 1, cnt1, 1)) != cudaSuccess) return e;
        ++ctx->launches;
      }
    }
    mark();
    if ((e = cudaEventRecord(evX[b], X)) != cudaSuccess) return e;
    if (s2 < nt) {
      // the bulk of the panel travels beside the next panel's work and only gates the trailing update
      if ((e = cudaStreamWaitEvent(Z, evX[b], 0)) != cudaSuccess) return e;
      if ((e = gather(Z, ctx->comm2, s0, s1, s2, nt, bsend, brecv)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(evZ[b], Z)) != cudaSuccess) return e;
      const int cnt2 = own_count(s2);
      if (cnt2 > 0) {
        if ((e = cudaStreamWaitEvent(Y, evZ[b], 0)) != cudaSuccess) return e;
        if ((e = launch_trailing(ctx, Y, g, s2, nt, first_own(s2), G, s0, s1, 1)) != cudaSuccess) return e;
      }
    }
    if ((e = cudaEventRecord(evY[b], Y)) != cudaSuccess) return e;
  }
  if ((e = cudaStreamWaitEvent(ctx->stream, evX[nblk - 1], 0)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(ctx->ev_join[0], Y)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[0], 0)) != cudaSuccess) return e;
  if ((e = cudaEventRecord(ctx->ev_join[1], Z)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[1], 0)) != cudaSuccess) return e;
  if (prof) {
    cudaStreamSynchronize(X);
    double acc[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < nblk; ++b)
      for (int k = 0; k < 5; ++k) {
        float ms = 0;
        cudaEventElapsedTime(&ms, pev[(size_t)b * 6 + k], pev[(size_t)b * 6 + k + 1]);
        acc[k] += ms;
      }
    fprintf(stderr, "[liblmm rank %d] row-cyclic chain, nt=%d ob=%d: wait_trailing %.2f ms, own_update %.2f, exchange %.2f, "
                    "diag+next rows %.2f, own TRSM %.2f\n", me, nt, ob, acc[0], acc[1], acc[2], acc[3], acc[4]);
    for (cudaEvent_t ev : pev) cudaEventDestroy(ev);
  }
  return cudaSuccess;
}

// Fork the batch into latent groups on separate streams (joined back into ctx->stream).
cudaError_t chol_factor(lmm_ctx* ctx, TiledSym L, double* W, size_t wstride, int batch, double* logdet, int* info) {
  const int G = ctx->ngroups < batch ? ctx->ngroups : batch;
  if (ctx->partition_ilmm && ctx->partition_now && batch == 1 && ctx->comm && ctx->nranks > 1 && L.nt >= 2 * ctx->nranks && nccl_api().AllGather)
    return (ctx->partition_ilmm == 2 && ctx->comm2) ? chol_factor_rowcyclic2(ctx, L, W, wstride, logdet, info)
                                                    : chol_factor_rowcyclic(ctx, L, W, wstride, logdet, info);
  if (ctx->lookahead == 2 && batch <= 2 && L.nt >= 12) return chol_factor_rightlooking(ctx, L, W, wstride, batch, logdet, info);
  if (ctx->lookahead && batch <= 2 && L.nt >= 12) return chol_factor_lookahead(ctx, L, W, wstride, batch, logdet, info);
  if (G <= 1 || L.nt <= 1) return chol_factor_stream(ctx, ctx->stream, L, W, wstride, batch, logdet, info);
  cudaError_t e;
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  for (int gi = 0; gi < G; ++gi) {
    const int b0 = (int)((int64_t)batch * gi / G), b1 = (int)((int64_t)batch * (gi + 1) / G);
    cudaStream_t st = ctx->gstream[gi];
    if ((e = cudaStreamWaitEvent(st, ctx->ev_fork, 0)) != cudaSuccess) return e;
    TiledSym Lg{L.base + (size_t)b0 * L.batch_stride, L.nt, L.batch_stride};
    if ((e = chol_factor_stream(ctx, st, Lg, W + (size_t)b0 * wstride, wstride, b1 - b0, logdet + b0, info + b0)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ctx->ev_join[gi], st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[gi], 0)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// X <- X L^{-T} for a rectangular tiled X (rows = e.g. test points): the same update/TRSM sweep
// with X's tile rows appended under the factor.
cudaError_t trsm_right_lt_stream(lmm_ctx* ctx, cudaStream_t st, TiledRect X, TiledSym L, const double* W, size_t wstride, int batch) {
  const int nt = L.nt, ob = ctx->outer_block;
  GemmArgs g{};
  g.A = operand(X);
  g.B = operand(L);
  g.C = operand(X);
  g.W = W;
  g.w_batch_stride = wstride;
  g.sym = 0;
  g.i0 = 0;
  cudaError_t e;
  for (int s0 = 0; s0 < nt; s0 += ob) {
    const int s1 = (s0 + ob < nt) ? s0 + ob : nt;
    if (s0 > 0) {
      g.j0 = s0; g.k0 = 0; g.k1 = s0;
      if ((e = launch_gemm(st, GEMM_UPDATE, g, s1 - s0, X.ntr, batch)) != cudaSuccess) return e;
      ++ctx->launches;
    }
    for (int jj = s0; jj < s1; ++jj) {
      if (jj > s0) {
        g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(st, GEMM_UPDATE, g, 1, X.ntr, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      g.j0 = jj;
      if ((e = launch_gemm(st, GEMM_TRSM, g, 1, X.ntr, batch)) != cudaSuccess) return e;
      ++ctx->launches;
    }
  }
  return cudaSuccess;
}

cudaError_t trsm_right_lt(lmm_ctx* ctx, TiledRect X, TiledSym L, const double* W, size_t wstride, int batch) {
  const int G = ctx->ngroups < batch ? ctx->ngroups : batch;
  if (G <= 1) return trsm_right_lt_stream(ctx, ctx->stream, X, L, W, wstride, batch);
  cudaError_t e;
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  for (int gi = 0; gi < G; ++gi) {
    const int b0 = (int)((int64_t)batch * gi / G), b1 = (int)((int64_t)batch * (gi + 1) / G);
    cudaStream_t st = ctx->gstream[gi];
    if ((e = cudaStreamWaitEvent(st, ctx->ev_fork, 0)) != cudaSuccess) return e;
    TiledRect Xg{X.base + (size_t)b0 * X.batch_stride, X.ntr, X.ntc, X.batch_stride};
    TiledSym Lg{L.base + (size_t)b0 * L.batch_stride, L.nt, L.batch_stride};
    if ((e = trsm_right_lt_stream(ctx, st, Xg, Lg, W + (size_t)b0 * wstride, wstride, b1 - b0)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(ctx->ev_join[gi], st)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[gi], 0)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// X <- X L^{-T} for an upper-triangular X given as full rectangular tiles (zero tiles skipped).
cudaError_t trsm_right_lt_upper(lmm_ctx* ctx, cudaStream_t st, TiledRect X, TiledSym L, const double* W, size_t wstride, int batch) {
  const int nt = L.nt, ob = ctx->outer_block;
  GemmArgs g{};
  g.A = operand(X); g.B = operand(L); g.C = operand(X);
  g.W = W; g.w_batch_stride = wstride;
  g.sym = 0; g.upper = 1; g.k_from_row = 1; g.i0 = 0;
  cudaError_t e;
  for (int s0 = 0; s0 < nt; s0 += ob) {
    const int s1 = (s0 + ob < nt) ? s0 + ob : nt;
    if (s0 > 0) {
      g.j0 = s0; g.k0 = 0; g.k1 = s0;
      if ((e = launch_gemm(st, GEMM_UPDATE, g, s1 - s0, s0, batch)) != cudaSuccess) return e;  // rows < s0 have k < s0 terms
      ++ctx->launches;
    }
    for (int jj = s0; jj < s1; ++jj) {
      if (jj > s0) {
        g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(st, GEMM_UPDATE, g, 1, jj, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      g.j0 = jj;
      if ((e = launch_gemm(st, GEMM_TRSM, g, 1, jj + 1, batch)) != cudaSuccess) return e;
      ++ctx->launches;
    }
  }
  return cudaSuccess;
}

// Marks a factorisation that every rank of the communicator performs on identical inputs (the joint ILMM factor,
// the batch-1 potrf primitive): with the "partition_ilmm" option such a call runs the row-cyclic multi-GPU schedule.
struct PartitionScope {
  lmm_ctx* c;
  explicit PartitionScope(lmm_ctx* ctx) : c(ctx) { c->partition_now = 1; }
  ~PartitionScope() { c->partition_now = 0; }
};

size_t factor_bytes_per_latent(int nt) { return (sym_tiles(nt) + (size_t)nt) * TT * sizeof(double); }

void fill_params(std::vector<LatentParams>& hp, const lmm_gp_desc* d, const double* noise, int lo, int hi, double ls_scale = 1.0) {
  hp.resize(hi - lo);
  for (int i = lo; i < hi; ++i) {
    LatentParams& q = hp[i - lo];
    q.kind = d[i].kind;
    q.pad = 0;
    q.variance = d[i].variance;
    q.inv_ls = d[i].inv_lengthscale * ls_scale;
    q.noise = noise[i];
    q.mean = d[i].mean_const;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// posterior handle
// ------------------------------------------------------------------------------------------------
// POST_JOINT: IndependentMOGP conditioned under a dense Σy (AbstractGPs generic path): one joint (mN) factor like
// POST_ILMM, identity mixing, no projection.
enum { POST_OILMM = 0, POST_IMOGP = 1, POST_ILMM = 2, POST_JOINT = 3 };

struct lmm_post {
  lmm_ctx* ctx = nullptr;
  int kind = POST_OILMM;
  int m = 0, p = 0, N = 0, D = 1, nt = 0;
  int lo = 0, hi = 0;  // resident latents [lo, hi)
  std::vector<lmm_gp_desc> descs;
  std::vector<double> noise;  // per latent (all m)
  std::vector<double> H;      // p x m column-major (U sqrt(S) for OILMM)
  std::vector<double> U, S;
  double sigma2 = 0.0;
  // device
  double* d_xpad = nullptr;  // [Npad][D]
  double* d_L = nullptr;     // TiledSym, batch = hi - lo (ILMM: batch 1 over mN)
  double* d_W = nullptr;     // [batch][nt] tiles
  double* d_alpha = nullptr; // [batch][Npad]
  double* d_delta = nullptr; // [batch][Npad]
  LatentParams* d_params = nullptr;
  double* d_H = nullptr;
  double* d_noise_vec = nullptr;  // [nloc][Npad] per-point training noise (sequentially conditioned posteriors), else null
  double* d_Ept = nullptr;        // POST_ILMM: [N][m*m] per-point projected noise blocks ΣT (extended by sequential conditioning)
  size_t bytes = 0;
  int big_n = 0, big_nt = 0;  // ILMM joint dimension mN and its tile count

  int nloc() const { return hi - lo; }
  size_t npad() const { return (size_t)nt * TILE; }
  bool joint() const { return kind == POST_ILMM || kind == POST_JOINT; }
  TiledSym Lsym() const { return TiledSym{d_L, joint() ? big_nt : nt, sym_tiles(joint() ? big_nt : nt) * TT}; }
  size_t wstride() const { return (size_t)(joint() ? big_nt : nt) * TT; }
};

// ------------------------------------------------------------------------------------------------
// C ABI: context
// ------------------------------------------------------------------------------------------------
extern "C" const char* lmm_version(void) { return "liblmm 0.1.0 (sm_100a)"; }

extern "C" int lmm_ctx_create(int device, lmm_ctx** out) {
  if (!out) return LMM_E_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return LMM_E_CUDA;  // no CPU fallback
  }
  if (device < 0 || device >= ndev) return LMM_E_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return LMM_E_CUDA;
  lmm_ctx* ctx = new lmm_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return LMM_E_CUDA;
  }
  for (auto& e : ctx->ev) cudaEventCreate(&e);
  for (auto& g : ctx->gstream) cudaStreamCreateWithFlags(&g, cudaStreamNonBlocking);
  {
    int lo_pri = 0, hi_pri = 0;
    cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
    cudaStreamCreateWithPriority(&ctx->panel_stream, cudaStreamNonBlocking, hi_pri);
    cudaStreamCreateWithPriority(&ctx->update_stream, cudaStreamNonBlocking, lo_pri);
    cudaStreamCreateWithPriority(&ctx->xchg_stream, cudaStreamNonBlocking, hi_pri);
  }
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  for (auto& e : ctx->ev_join) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return LMM_OK;
}

extern "C" int lmm_ctx_destroy(lmm_ctx* ctx) {
  if (!ctx) return LMM_E_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_small && nccl_api().ok) nccl_api().CommDestroy(ctx->comm_small);
  if (ctx->comm2 && nccl_api().ok) nccl_api().CommDestroy(ctx->comm2);
  if (ctx->comm && nccl_api().ok) nccl_api().CommDestroy(ctx->comm);
  if (ctx->xbuf2) cudaFree(ctx->xbuf2);
  cudaStreamDestroy(ctx->xchg_stream);
  for (auto& e : ctx->ev) cudaEventDestroy(e);
  for (auto& g : ctx->gstream) cudaStreamDestroy(g);
  if (ctx->xbuf) cudaFree(ctx->xbuf);
  cudaStreamDestroy(ctx->panel_stream);
  cudaStreamDestroy(ctx->update_stream);
  for (auto& e : ctx->blk_ev) cudaEventDestroy(e);
  cudaEventDestroy(ctx->ev_fork);
  for (auto& e : ctx->ev_join) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return LMM_OK;
}

extern "C" const char* lmm_last_error(lmm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int lmm_ctx_set_option(lmm_ctx* ctx, const char* key, double value) {
  if (!ctx || !key) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  std::string k(key);
  if (k == "distance_form") {
    if (value != 0.0 && value != 1.0) return ctx->fail(LMM_E_ARG, "distance_form must be 0 or 1");
    ctx->distance_form = (int)value;
  } else if (k == "outer_block") {
    if (value == 0.0) {  // back to the built-in choice
      ctx->outer_block = 8;
      ctx->outer_block_user = false;
      return LMM_OK;
    }
    if (value < 1 || value > 64) return ctx->fail(LMM_E_ARG, "outer_block must be in [1, 64] (0 = automatic)");
    ctx->outer_block = (int)value;
    ctx->outer_block_user = true;
  } else if (k == "streams") {
    if (value < 1 || value > lmm_ctx::MAX_GROUPS) return ctx->fail(LMM_E_ARG, "streams must be in [1, 8]");
    ctx->ngroups = (int)value;
  } else if (k == "lookahead") {
    if (value != 0.0 && value != 1.0 && value != 2.0) return ctx->fail(LMM_E_ARG, "lookahead must be 0, 1 (left-looking, K-split) or 2 (right-looking)");
    ctx->lookahead = (int)value;
  } else if (k == "nccl_small_ctas") {  // takes effect at lmm_comm_init
    if (value < 0 || value > 32) return ctx->fail(LMM_E_ARG, "nccl_small_ctas must be in [0, 32]");
    ctx->nccl_small_ctas = (int)value;
  } else if (k == "profile_partition") {
    ctx->profile_partition = value != 0.0;
  } else if (k == "partition_ilmm") {
    if (value != 0.0 && value != 1.0 && value != 2.0) return ctx->fail(LMM_E_ARG, "partition_ilmm must be 0, 1 or 2");
    ctx->partition_ilmm = (int)value;
  } else if (k == "gemm_small") {
    if (value < 0 || value > 4096) return ctx->fail(LMM_E_ARG, "gemm_small must be in [0, 4096]");
    set_gemm_small_threshold((int)value);
  } else if (k == "gemm_impl") {
    if (value != 0.0 && value != 1.0 && value != 2.0) return ctx->fail(LMM_E_UNSUPPORTED, "gemm_impl must be 0, 1 or 2");
    set_gemm_impl((int)value);
  } else {
    return ctx->fail(LMM_E_UNSUPPORTED, "unknown option " + k);
  }
  return LMM_OK;
}

extern "C" int lmm_ctx_counters(lmm_ctx* ctx, int64_t* kernel_launches, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (kernel_launches) *kernel_launches = ctx->launches;
  if (h2d_bytes) *h2d_bytes = ctx->h2d;
  if (d2h_bytes) *d2h_bytes = ctx->d2h;
  return LMM_OK;
}

extern "C" int lmm_ctx_last_timings(lmm_ctx* ctx, double out_ms[8]) {
  if (!ctx || !out_ms) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  for (int i = 0; i < 8; ++i) out_ms[i] = ctx->timings[i];
  return LMM_OK;
}

extern "C" int lmm_comm_unique_id(void* out_128_bytes) {
  if (!out_128_bytes) return LMM_E_ARG;
  NcclApi& api = nccl_api();
  if (!api.ok) return LMM_E_NCCL;
  NcclId id;
  if (api.GetUniqueId(&id) != 0) return LMM_E_NCCL;
  memcpy(out_128_bytes, &id, 128);
  return LMM_OK;
}

extern "C" int lmm_comm_set_shard(lmm_ctx* ctx, int nranks, int rank) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->nranks = nranks;
  ctx->rank = rank;
  return LMM_OK;
}

extern "C" int lmm_comm_init(lmm_ctx* ctx, const void* unique_id_128_bytes, int nranks, int rank) {
  if (!ctx || !unique_id_128_bytes || nranks < 1 || rank < 0 || rank >= nranks) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  NcclApi& api = nccl_api();
  if (!api.ok) return ctx->fail(LMM_E_NCCL, "libnccl.so.2 not found (dlopen)");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return LMM_E_CUDA;
  NcclId id;
  memcpy(&id, unique_id_128_bytes, 128);
  int r = api.CommInitRank(&ctx->comm, nranks, id, rank);
  if (r != 0) return ctx->fail(LMM_E_NCCL, std::string("ncclCommInitRank: ") + (api.GetErrorString ? api.GetErrorString(r) : "?"));
  ctx->nranks = nranks;
  ctx->rank = rank;
  // Communicators of the partitioned factorisation (optional: without them it uses the main one).  Their exchanges sit on
  // the panel chain while the trailing GEMMs hold every SM (one 192 KB CTA each), so an NCCL kernel waits for as many
  // SMs to drain as it has CTAs: cap them (ncclConfig_t minCTAs / maxCTAs; the prefix of the struct as of NCCL 2.18).
  if (api.CommSplit && nranks > 1) {
    struct { size_t size; unsigned magic; unsigned version; int blocking, cgaClusterSize, minCTAs, maxCTAs; const char* netName; int splitShare; } cfg;
    const int UNDEF = -2147483647 - 1;
    auto split = [&](void** out, int max_ctas) {
      cfg = {sizeof(cfg), 0xcafebeefu, 21800u, UNDEF, UNDEF, 1, max_ctas, nullptr, UNDEF};
      if (api.CommSplit(ctx->comm, 0, rank, out, &cfg) != 0) {
        *out = nullptr;
        if (api.CommSplit(ctx->comm, 0, rank, out, nullptr) != 0) *out = nullptr;
      }
    };
    if (ctx->nccl_small_ctas > 0) split(&ctx->comm_small, ctx->nccl_small_ctas);
    if (api.CommSplit(ctx->comm, 0, rank, &ctx->comm2, nullptr) != 0) ctx->comm2 = nullptr;
  }
  return LMM_OK;
}

// ------------------------------------------------------------------------------------------------
// Orthogonal validation (host): src/orthogonal_matrix.jl:21-23  isapprox(U'U, I)
// ------------------------------------------------------------------------------------------------
extern "C" int lmm_orthogonal_validate(const double* U, int p, int m) {
  if (!U || p <= 0 || m <= 0) return LMM_E_ARG;
  // |U'U - I|_F <= sqrt(eps) * max(|U'U|_F, |I|_F)
  double diff2 = 0.0, g2 = 0.0;
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      double s = 0.0;
      for (int j = 0; j < p; ++j) s += U[(size_t)a * p + j] * U[(size_t)b * p + j];
      g2 += s * s;
      const double d = s - (a == b ? 1.0 : 0.0);
      diff2 += d * d;
    }
  const double rtol = 1.4901161193847656e-08;  // sqrt(eps(Float64))
  const double nrm = std::max(std::sqrt(g2), std::sqrt((double)m));
  if (!(std::sqrt(diff2) <= rtol * nrm)) return LMM_E_NOT_ORTHOGONAL;
  return LMM_OK;
}

extern "C" int lmm_reorder_indices(int N, int p, int direction, int64_t* out) {
  if (N <= 0 || p <= 0 || !out || (direction != 0 && direction != 1)) return LMM_E_ARG;
  // direction 0: vec(reshape(1:pN, N, p)')  (src/independent_mogp.jl:138)
  // direction 1: vec(reshape(1:pN, p, N)')  (src/independent_mogp.jl:144)
  if (direction == 0) {
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < p; ++j) out[(size_t)i * p + j] = (int64_t)j * N + i;
  } else {
    for (int j = 0; j < p; ++j)
      for (int i = 0; i < N; ++i) out[(size_t)j * N + i] = (int64_t)i * p + j;
  }
  return LMM_OK;
}

// ------------------------------------------------------------------------------------------------
// Core: per-latent exact GP logpdf / posterior over a set of independent latents
// ------------------------------------------------------------------------------------------------
namespace {

// Projection description (host): Ty = T Y  (m x p), residual |Y - Q (P Y)|², regulariser constant.
struct Projection {
  std::vector<double> T;      // m x p col-major
  std::vector<double> P, Q;   // m x p, p x m (empty: no regulariser)
  std::vector<double> noise;  // per latent diagonal noise
  double reg_c0 = 0.0;        // n * (...) part of the regulariser
  bool has_reg = false;
};

struct RunOut {
  lmm_post** post = nullptr;
  double* logpdf = nullptr;
  double* lml_terms = nullptr;
  int* info_latent = nullptr;
};

// The shared driver for OILMM (src/oilmm.jl:79-93, 116-134) and IndependentMOGP
// (src/independent_mogp.jl:74-80, 119-126).
int latents_run(lmm_ctx* ctx, int kind, const lmm_gp_desc* latents, int m, const double* x, int N, int D, int p, double sigma2,
                const double* y, const Projection& pr, const double* Hhost, const double* Uhost, const double* Shost, RunOut out,
                const double* noise_vec = nullptr /* per-point noise, m*N by outputs (IndependentMOGP with Σy = Diagonal(v)) */) {
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (double& t : ctx->timings) t = 0.0;
  int lo, hi;
  shard_range(ctx, m, lo, hi);
  const int mloc = hi - lo;
  const int nt = ntiles(N);
  const size_t npad = (size_t)nt * TILE;
  const bool keep = out.post != nullptr;

  CU(cudaEventRecord(ctx->ev[0], st));
  // ---- stage inputs
  DevBuf b_x, b_y, b_T, b_P, b_Q, b_means, b_ty, b_resid_part, b_resid, b_terms;
  CU(b_x.alloc(ctx, npad * D * sizeof(double)));
  CU(cudaMemsetAsync(b_x.p, 0, npad * D * sizeof(double), st));
  CU(copy_in(ctx, b_x.as<double>(), x, (size_t)N * D));
  const double* d_y = y;
  if (!is_device_ptr(y)) {
    CU(b_y.alloc(ctx, (size_t)p * N * sizeof(double)));
    CU(copy_in(ctx, b_y.as<double>(), y, (size_t)p * N));
    d_y = b_y.as<double>();
  }
  CU(b_T.alloc(ctx, pr.T.size() * sizeof(double)));
  CU(copy_in(ctx, b_T.as<double>(), pr.T.data(), pr.T.size()));
  if (pr.has_reg) {
    CU(b_P.alloc(ctx, pr.P.size() * sizeof(double)));
    CU(copy_in(ctx, b_P.as<double>(), pr.P.data(), pr.P.size()));
    CU(b_Q.alloc(ctx, pr.Q.size() * sizeof(double)));
    CU(copy_in(ctx, b_Q.as<double>(), pr.Q.data(), pr.Q.size()));
  }
  std::vector<double> hmeans(mloc > 0 ? mloc : 1, 0.0);
  for (int i = lo; i < hi; ++i) hmeans[i - lo] = latents[i].mean_const;
  CU(b_means.alloc(ctx, hmeans.size() * sizeof(double)));
  CU(copy_in(ctx, b_means.as<double>(), hmeans.data(), hmeans.size()));
  std::vector<LatentParams> hparams;
  fill_params(hparams, latents, pr.noise.data(), lo, hi);
  DevBuf b_params;
  CU(b_params.alloc(ctx, (hparams.size() + 1) * sizeof(LatentParams)));
  if (mloc > 0) {
    ctx->h2d += (int64_t)(hparams.size() * sizeof(LatentParams));
    CU(cudaMemcpyAsync(b_params.p, hparams.data(), hparams.size() * sizeof(LatentParams), cudaMemcpyHostToDevice, st));
  }

  // ---- projection + residual (K2/K3)
  CU(b_ty.alloc(ctx, (size_t)(mloc > 0 ? mloc : 1) * npad * sizeof(double)));
  CU(cudaMemsetAsync(b_ty.p, 0, (size_t)(mloc > 0 ? mloc : 1) * npad * sizeof(double), st));
  const int nblk = (N + 15) / 16;
  CU(b_resid_part.alloc(ctx, (size_t)nblk * sizeof(double)));
  CU(cudaMemsetAsync(b_resid_part.p, 0, (size_t)nblk * sizeof(double), st));
  CU(b_resid.alloc(ctx, sizeof(double)));
  CU(b_terms.alloc(ctx, (size_t)(m + 1) * sizeof(double)));
  CU(cudaMemsetAsync(b_terms.p, 0, (size_t)(m + 1) * sizeof(double), st));
  const bool do_reg = pr.has_reg && ctx->rank == 0;
  {
    int nb_out = 0;
    CU(launch_project(st, d_y, N, p, b_T.as<double>(), m, lo, mloc, b_means.as<double>(), b_ty.as<double>(), npad,
                      do_reg ? b_P.as<double>() : nullptr, do_reg ? b_Q.as<double>() : nullptr, b_resid_part.as<double>(), &nb_out));
    ++ctx->launches;
    if (do_reg) {
      CU(launch_sum_partials(st, b_resid_part.as<double>(), nblk, b_resid.as<double>()));
      regulariser_kernel<<<1, 1, 0, st>>>(b_terms.as<double>() + m, pr.reg_c0, b_resid.as<double>(), sigma2);
      CU(cudaGetLastError());
      ctx->launches += 2;
    }
  }
  CU(cudaEventRecord(ctx->ev[1], st));

  // ---- factor storage: all local latents when a posterior is kept, else a streamed arena
  const size_t per_lat = factor_bytes_per_latent(nt);
  int chunk = mloc;
  if (!keep && mloc > 0) {
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    size_t budget = (size_t)((double)fr * 0.80);
    size_t fit = budget / (per_lat + 6 * npad * sizeof(double));
    if (fit < 1) fit = 1;
    if ((size_t)chunk > fit) chunk = (int)fit;
  }
  DevBuf b_L, b_W, b_alpha, b_r, b_z, b_logdet, b_quad, b_info, b_nv;
  if (noise_vec) {
    const int nl = mloc > 0 ? mloc : 1;
    CU(b_nv.alloc(ctx, (size_t)nl * npad * sizeof(double)));
    CU(cudaMemsetAsync(b_nv.p, 0, (size_t)nl * npad * sizeof(double), st));
    if (mloc > 0) {
      const bool dev = is_device_ptr(noise_vec);
      if (!dev) ctx->h2d += (int64_t)((size_t)mloc * N * sizeof(double));
      CU(cudaMemcpy2DAsync(b_nv.p, npad * sizeof(double), noise_vec + (size_t)lo * N, (size_t)N * sizeof(double), (size_t)N * sizeof(double),
                           (size_t)mloc, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    }
  }
  std::vector<int> hinfo(mloc > 0 ? mloc : 1, 0);
  float ms_kmat = 0, ms_chol = 0, ms_solve = 0;
  if (mloc > 0) {
    CU(b_L.alloc(ctx, (size_t)chunk * sym_tiles(nt) * TT * sizeof(double)));
    CU(b_W.alloc(ctx, (size_t)chunk * nt * TT * sizeof(double)));
    CU(b_r.alloc(ctx, (size_t)chunk * npad * sizeof(double)));
    CU(b_z.alloc(ctx, (size_t)chunk * npad * sizeof(double)));
    if (keep) CU(b_alpha.alloc(ctx, (size_t)mloc * npad * sizeof(double)));
    CU(b_logdet.alloc(ctx, (size_t)mloc * sizeof(double)));
    CU(b_quad.alloc(ctx, (size_t)mloc * sizeof(double)));
    CU(b_info.alloc(ctx, (size_t)mloc * sizeof(int)));
    CU(cudaMemsetAsync(b_logdet.p, 0, (size_t)mloc * sizeof(double), st));
    CU(cudaMemsetAsync(b_info.p, 0, (size_t)mloc * sizeof(int), st));
    for (int c0 = 0; c0 < mloc; c0 += chunk) {
      const int nb = (c0 + chunk <= mloc) ? chunk : mloc - c0;
      TiledSym L{b_L.as<double>(), nt, sym_tiles(nt) * TT};
      double* W = b_W.as<double>();
      const size_t wstride = (size_t)nt * TT;
      const LatentParams* dp = b_params.as<LatentParams>() + c0;
      double* delta = b_ty.as<double>() + (size_t)c0 * npad;
      CU(cudaEventRecord(ctx->ev[2], st));
      CU(launch_kmat_sym(st, L, nb, b_x.as<double>(), N, D, dp, ctx->distance_form,
                         noise_vec ? b_nv.as<double>() + (size_t)c0 * npad : nullptr, npad));
      ++ctx->launches;
      CU(cudaEventRecord(ctx->ev[3], st));
      CU(chol_factor(ctx, L, W, wstride, nb, b_logdet.as<double>() + c0, b_info.as<int>() + c0));
      CU(cudaEventRecord(ctx->ev[4], st));
      CU(cudaMemcpyAsync(b_r.p, delta, (size_t)nb * npad * sizeof(double), cudaMemcpyDeviceToDevice, st));
      CU(launch_fwd_solve(st, L, W, wstride, b_r.as<double>(), b_z.as<double>(), npad, nb, &ctx->launches));
      CU(launch_sumsq(st, b_z.as<double>(), npad, (int)npad, nb, b_quad.as<double>() + c0));
      ++ctx->launches;
      if (keep) {
        CU(cudaMemcpyAsync(b_r.p, b_z.p, (size_t)nb * npad * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CU(launch_bwd_solve(st, L, W, wstride, b_r.as<double>(), b_alpha.as<double>() + (size_t)c0 * npad, npad, nb, &ctx->launches));
      }
      lml_terms_kernel<<<(nb + 127) / 128, 128, 0, st>>>(b_terms.as<double>(), lo + c0, nb, b_logdet.as<double>() + c0,
                                                         b_quad.as<double>() + c0, N, LOG2PI);
      CU(cudaGetLastError());
      ++ctx->launches;
      CU(cudaEventRecord(ctx->ev[5], st));
      {
        // accumulate stage timings per chunk: the stage events are reused by the next chunk, so they
        // are read here (one host sync per chunk; negligible next to a chunk's factorisation)
        CU(cudaEventSynchronize(ctx->ev[5]));
        float a = 0, bq = 0, c = 0;
        cudaEventElapsedTime(&a, ctx->ev[2], ctx->ev[3]);
        cudaEventElapsedTime(&bq, ctx->ev[3], ctx->ev[4]);
        cudaEventElapsedTime(&c, ctx->ev[4], ctx->ev[5]);
        ms_kmat += a; ms_chol += bq; ms_solve += c;
      }
    }
  }
  // ---- reduce the per-latent terms across ranks (one NCCL all-reduce over NVLink) and read back
  if (ctx->comm && ctx->nranks > 1) {
    int r = nccl_api().AllReduce(b_terms.p, b_terms.p, (size_t)(m + 1), NCCL_DOUBLE, NCCL_SUM, ctx->comm, st);
    if (r != 0) return ctx->fail(LMM_E_NCCL, "ncclAllReduce failed");
  }
  std::vector<double> hterms(m + 1, 0.0);
  CU(copy_out(ctx, hterms.data(), b_terms.p, (size_t)(m + 1) * sizeof(double)));
  if (mloc > 0) CU(copy_out(ctx, hinfo.data(), b_info.p, (size_t)mloc * sizeof(int)));
  CU(cudaEventRecord(ctx->ev[6], st));
  CU(cudaStreamSynchronize(st));
  {
    float tot = 0, prj = 0;
    cudaEventElapsedTime(&tot, ctx->ev[0], ctx->ev[6]);
    cudaEventElapsedTime(&prj, ctx->ev[0], ctx->ev[1]);
    ctx->timings[0] = tot; ctx->timings[1] = ms_kmat; ctx->timings[2] = ms_chol; ctx->timings[3] = ms_solve; ctx->timings[4] = prj;
  }
  for (int i = 0; i < mloc; ++i) {
    if (hinfo[i] > 0) {
      int pivot = hinfo[i] > N ? N : hinfo[i];
      if (out.info_latent) *out.info_latent = lo + i;
      char buf[128];
      snprintf(buf, sizeof buf, "PosDefException: latent %d is not positive definite (pivot %d)", lo + i, pivot);
      ctx->err = buf;
      return pivot;
    }
  }
  if (out.info_latent) *out.info_latent = -1;
  if (out.lml_terms) memcpy(out.lml_terms, hterms.data(), (size_t)(m + 1) * sizeof(double));
  if (out.logpdf) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) s += hterms[i];
    *out.logpdf = s + hterms[m];
  }
  if (keep) {
    DevBuf b_H;
    CU(b_H.alloc(ctx, (size_t)p * m * sizeof(double)));
    CU(copy_in(ctx, b_H.as<double>(), Hhost, (size_t)p * m));
    CU(cudaStreamSynchronize(st));
    lmm_post* P = new lmm_post();  // nothing below can fail: ownership of the device buffers moves to P
    P->ctx = ctx; P->kind = kind; P->m = m; P->p = p; P->N = N; P->D = D; P->nt = nt; P->lo = lo; P->hi = hi;
    P->descs.assign(latents, latents + m);
    P->noise = pr.noise;
    P->H.assign(Hhost, Hhost + (size_t)p * m);
    if (Uhost) P->U.assign(Uhost, Uhost + (size_t)p * m);
    if (Shost) P->S.assign(Shost, Shost + m);
    P->sigma2 = sigma2;
    P->bytes = (size_t)mloc * (per_lat + 2 * npad * sizeof(double)) + npad * D * sizeof(double);
    P->d_xpad = (double*)b_x.detach();
    P->d_L = (double*)b_L.detach();
    P->d_W = (double*)b_W.detach();
    P->d_alpha = (double*)b_alpha.detach();
    P->d_delta = (double*)b_ty.detach();
    P->d_params = (LatentParams*)b_params.detach();
    P->d_H = (double*)b_H.detach();
    if (noise_vec) P->d_noise_vec = (double*)b_nv.detach();
    *out.post = P;
  }
  return LMM_OK;
}

int oilmm_projection(lmm_ctx* ctx, const double* U, const double* S, int p, int m, double sigma2, int N, Projection& pr,
                     std::vector<double>& H) {
  for (int i = 0; i < m; ++i)
    if (!(S[i] > 0.0)) return ctx->fail(LMM_E_ARG, "S must have positive entries");
  if (!(sigma2 > 0.0)) return ctx->fail(LMM_E_ARG, "noise variance must be positive");
  pr.T.resize((size_t)m * p);
  pr.P.resize((size_t)m * p);
  pr.Q.assign(U, U + (size_t)p * m);
  pr.noise.resize(m);
  H.resize((size_t)p * m);
  double logdetS = 0.0;
  for (int i = 0; i < m; ++i) {
    const double rs = std::sqrt(S[i]);
    for (int j = 0; j < p; ++j) {
      const double u = U[(size_t)i * p + j];
      pr.T[(size_t)j * m + i] = u / rs;  // T = sqrt(S) \ U'        src/oilmm.jl:24
      pr.P[(size_t)j * m + i] = u;       // U'
      H[(size_t)i * p + j] = u * rs;     // U * sqrt(S)             src/oilmm.jl:69
    }
    pr.noise[i] = sigma2 * (1.0 / S[i]);  // diag(σ² * inv(S))      src/oilmm.jl:27
    logdetS += std::log(S[i]);
  }
  // -(n (logdet(S) + (p-m) log(2πσ²)) + |(I-UU')Y|²/σ²)/2          src/oilmm.jl:111-112
  pr.reg_c0 = (double)N * (logdetS + (double)(p - m) * std::log(2.0 * M_PI * sigma2));
  pr.has_reg = true;
  return LMM_OK;
}

int check_common(lmm_ctx* ctx, const lmm_gp_desc* latents, int m, const void* x, int N, int D, int p, int out_dim) {
  if (!latents || !x || m <= 0 || N <= 0 || D <= 0 || p <= 0) return ctx->fail(LMM_E_ARG, "null pointer or non-positive size");
  if (D > 64) return ctx->fail(LMM_E_UNSUPPORTED, "input dimension D > 64 is not supported");
  if (out_dim != p) return ctx->fail(LMM_E_OUT_DIM, "out dim of x != out dim of f.");
  return check_descs(ctx, latents, m);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// OILMM
// ------------------------------------------------------------------------------------------------
extern "C" int lmm_oilmm_posterior(lmm_ctx* ctx, const lmm_gp_desc* latents, int m, const double* x, int N, int D,
                                   const double* U, const double* S, int p, double sigma2, const double* y, int out_dim,
                                   lmm_post** out_post, double* out_logpdf, double* lml_terms, int* info_latent) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (out_post) *out_post = nullptr;
  int rc = check_common(ctx, latents, m, x, N, D, p, out_dim);
  if (rc) return rc;
  if (!U || !S || !y) return ctx->fail(LMM_E_ARG, "null pointer");
  if (m > p) return ctx->fail(LMM_E_ARG, "more latents than outputs");
  Projection pr;
  std::vector<double> H;
  if ((rc = oilmm_projection(ctx, U, S, p, m, sigma2, N, pr, H))) return rc;
  RunOut out{out_post, out_logpdf, lml_terms, info_latent};
  return latents_run(ctx, POST_OILMM, latents, m, x, N, D, p, sigma2, y, pr, H.data(), U, S, out);
}

extern "C" int lmm_oilmm_logpdf(lmm_ctx* ctx, const lmm_gp_desc* latents, int m, const double* x, int N, int D, const double* U,
                                const double* S, int p, double sigma2, const double* y, int out_dim, double* out_logpdf,
                                double* lml_terms, int* info_latent) {
  if (!out_logpdf && !lml_terms) return LMM_E_ARG;
  return lmm_oilmm_posterior(ctx, latents, m, x, N, D, U, S, p, sigma2, y, out_dim, nullptr, out_logpdf, lml_terms, info_latent);
}

// ------------------------------------------------------------------------------------------------
// IndependentMOGP
// ------------------------------------------------------------------------------------------------
extern "C" int lmm_imogp_posterior(lmm_ctx* ctx, const lmm_gp_desc* fs, int m, const double* x, int N, int D, double sigma2,
                                   const double* y, int out_dim, lmm_post** out_post, double* out_logpdf, int* info_latent) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (out_post) *out_post = nullptr;
  int rc = check_common(ctx, fs, m, x, N, D, m, out_dim);
  if (rc) return rc;
  if (!y) return ctx->fail(LMM_E_ARG, "null pointer");
  if (!(sigma2 > 0.0)) return ctx->fail(LMM_E_ARG, "noise variance must be positive");
  Projection pr;
  pr.T.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) pr.T[(size_t)i * m + i] = 1.0;
  pr.noise.assign(m, sigma2);
  pr.has_reg = false;
  std::vector<double> H = pr.T;
  RunOut out{out_post, out_logpdf, nullptr, info_latent};
  return latents_run(ctx, POST_IMOGP, fs, m, x, N, D, m, sigma2, y, pr, H.data(), nullptr, nullptr, out);
}

extern "C" int lmm_imogp_logpdf(lmm_ctx* ctx, const lmm_gp_desc* fs, int m, const double* x, int N, int D, double sigma2,
                                const double* y, int out_dim, double* out_logpdf, double* lml_terms, int* info_latent) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_common(ctx, fs, m, x, N, D, m, out_dim);
  if (rc) return rc;
  if (!y || (!out_logpdf && !lml_terms)) return ctx->fail(LMM_E_ARG, "null pointer");
  if (!(sigma2 > 0.0)) return ctx->fail(LMM_E_ARG, "noise variance must be positive");
  Projection pr;
  pr.T.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; ++i) pr.T[(size_t)i * m + i] = 1.0;
  pr.noise.assign(m, sigma2);
  pr.has_reg = false;
  std::vector<double> H = pr.T;
  RunOut out{nullptr, out_logpdf, lml_terms, info_latent};
  return latents_run(ctx, POST_IMOGP, fs, m, x, N, D, m, sigma2, y, pr, H.data(), nullptr, nullptr, out);
}

// ------------------------------------------------------------------------------------------------
// Posterior handle
// ------------------------------------------------------------------------------------------------
extern "C" int lmm_post_free(lmm_post* post) {
  if (!post) return LMM_E_ARG;
  lmm_ctx* ctx = post->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  cudaSetDevice(ctx->device);
  void* ptrs[] = {post->d_xpad, post->d_L, post->d_W, post->d_alpha, post->d_delta, post->d_params, post->d_H, post->d_noise_vec,
                  post->d_Ept};
  for (void* q : ptrs)
    if (q) cudaFreeAsync(q, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  delete post;
  return LMM_OK;
}

extern "C" int lmm_post_info(lmm_post* post, int* kind, int* m, int* p, int* N, int* D, int64_t* device_bytes) {
  if (!post) return LMM_E_ARG;
  if (kind) *kind = post->kind;
  if (m) *m = post->m;
  if (p) *p = post->p;
  if (N) *N = post->N;
  if (D) *D = post->D;
  if (device_bytes) *device_bytes = (int64_t)post->bytes;
  return LMM_OK;
}

extern "C" int lmm_post_export(lmm_post* post, int i, double* Lout, double* alpha, double* delta) {
  if (!post) return LMM_E_ARG;
  lmm_ctx* ctx = post->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  int b, n;
  if (post->joint()) {
    if (i != 0) return ctx->fail(LMM_E_ARG, "a joint posterior has one factor (i = 0)");
    b = 0;
    n = post->big_n;
  } else {
    if (i < post->lo || i >= post->hi) return ctx->fail(LMM_E_ARG, "latent not resident on this rank");
    b = i - post->lo;
    n = post->N;
  }
  const size_t vstride = post->joint() ? (size_t)post->big_nt * TILE : post->npad();
  if (Lout) {
    DevBuf dense;
    CU(dense.alloc(ctx, (size_t)n * n * sizeof(double)));
    CU(cudaMemsetAsync(dense.p, 0, (size_t)n * n * sizeof(double), ctx->stream));
    CU(launch_untile_lower(ctx->stream, post->Lsym(), b, dense.as<double>(), n));
    ++ctx->launches;
    CU(copy_out(ctx, Lout, dense.p, (size_t)n * n * sizeof(double)));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if (alpha) CU(copy_out(ctx, alpha, post->d_alpha + (size_t)b * vstride, (size_t)n * sizeof(double)));
  if (delta) CU(copy_out(ctx, delta, post->d_delta + (size_t)b * vstride, (size_t)n * sizeof(double)));
  CU(cudaStreamSynchronize(ctx->stream));
  return LMM_OK;
}

namespace {

// Latent posterior marginals at xs for the resident latents: ML/VL [nloc][nspad] on the device.
// mean*_i = m_i + K(x*,x) α_i ; var*_i = k(x*,x*) - colsumsq(L_i^{-1} K(x,x*))   (AbstractGPs)
int post_latent_marginals(lmm_post* post, const double* d_xspad, int Ns, int nts, double* d_ML, double* d_VL) {
  lmm_ctx* ctx = post->ctx;
  cudaStream_t st = ctx->stream;
  const int nloc = post->nloc(), nt = post->nt;
  const size_t nspad = (size_t)nts * TILE;
  if (nloc == 0) return LMM_OK;
  const size_t per_lat = (size_t)nts * nt * TT * sizeof(double);
  size_t fr = 0, tot = 0;
  CU(cudaMemGetInfo(&fr, &tot));
  size_t fit = (size_t)((double)fr * 0.8) / per_lat;
  if (fit < 1) fit = 1;
  const int chunk = (size_t)nloc < fit ? nloc : (int)fit;
  DevBuf b_V;
  CU(b_V.alloc(ctx, (size_t)chunk * per_lat));
  const TiledSym L = post->Lsym();
  for (int c0 = 0; c0 < nloc; c0 += chunk) {
    const int nb = (c0 + chunk <= nloc) ? chunk : nloc - c0;
    TiledRect V{b_V.as<double>(), nts, nt, (size_t)nts * nt * TT};
    TiledSym Lc{L.base + (size_t)c0 * L.batch_stride, nt, L.batch_stride};
    const double* Wc = post->d_W + (size_t)c0 * post->wstride();
    const LatentParams* dp = post->d_params + c0;
    CU(launch_kmat_cross(st, V, nb, d_xspad, Ns, post->d_xpad, post->N, post->D, dp, ctx->distance_form));
    CU(launch_rect_gemv(st, V, post->d_alpha + (size_t)c0 * post->npad(), post->npad(), d_ML + (size_t)c0 * nspad, nspad, dp, 1, nb));
    ctx->launches += 2;
    CU(trsm_right_lt(ctx, V, Lc, Wc, post->wstride(), nb));
    CU(launch_rect_rowsumsq(st, V, d_VL + (size_t)c0 * nspad, nspad, dp, nb));
    ++ctx->launches;
  }
  return LMM_OK;
}

}  // namespace

namespace {
int ilmm_post_mean_and_var(lmm_post* post, const double* xs, int Ns, double sigma2, double* mean, double* var);
}

extern "C" int lmm_post_mean_and_var(lmm_post* post, const double* xs, int Ns, double sigma2, double* mean, double* var) {
  if (!post || !xs || Ns <= 0 || !mean || !var) return LMM_E_ARG;
  lmm_ctx* ctx = post->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  CU(cudaSetDevice(ctx->device));
  if (post->joint()) return ilmm_post_mean_and_var(post, xs, Ns, sigma2, mean, var);
  cudaStream_t st = ctx->stream;
  for (double& t : ctx->timings) t = 0.0;
  const int nts = ntiles(Ns), nloc = post->nloc(), p = post->p, m = post->m;
  const size_t nspad = (size_t)nts * TILE;
  CU(cudaEventRecord(ctx->ev[0], st));
  DevBuf b_xs, b_ML, b_VL, b_mean, b_var;
  CU(b_xs.alloc(ctx, nspad * post->D * sizeof(double)));
  CU(cudaMemsetAsync(b_xs.p, 0, nspad * post->D * sizeof(double), st));
  CU(copy_in(ctx, b_xs.as<double>(), xs, (size_t)Ns * post->D));
  CU(b_ML.alloc(ctx, (size_t)(nloc > 0 ? nloc : 1) * nspad * sizeof(double)));
  CU(b_VL.alloc(ctx, (size_t)(nloc > 0 ? nloc : 1) * nspad * sizeof(double)));
  int rc = post_latent_marginals(post, b_xs.as<double>(), Ns, nts, b_ML.as<double>(), b_VL.as<double>());
  if (rc) return rc;
  // one buffer [mean | var] so that a single all-reduce covers both
  const size_t nout = (size_t)p * Ns;
  CU(b_mean.alloc(ctx, 2 * nout * sizeof(double)));
  double* d_mean = b_mean.as<double>();
  double* d_var = d_mean + nout;
  const bool multi = ctx->comm && ctx->nranks > 1;
  if (post->kind == POST_OILMM) {
    // M = H M_lat ; V = (H∘H)(V_lat + 1e-18) + σ²     src/oilmm.jl:61-75 (1e-18: default FiniteGP noise)
    CU(cudaMemsetAsync(d_mean, 0, 2 * nout * sizeof(double), st));
    CU(launch_backproject(st, post->d_H, p, m, post->lo, nloc, b_ML.as<double>(), b_VL.as<double>(), nspad, Ns, 1e-18, sigma2,
                          multi ? 0 : 1, d_mean, d_var));
    ++ctx->launches;
  } else {
    // IndependentMOGP: mean/var concatenated by outputs, var + σ²   src/independent_mogp.jl:50-57
    CU(cudaMemsetAsync(d_mean, 0, 2 * nout * sizeof(double), st));
    if (nloc > 0) {
      dim3 grid((unsigned)((Ns + 255) / 256), (unsigned)nloc);
      copy_add_kernel<<<grid, 256, 0, st>>>(d_mean + (size_t)post->lo * Ns, Ns, b_ML.as<double>(), nspad, Ns, 0.0);
      copy_add_kernel<<<grid, 256, 0, st>>>(d_var + (size_t)post->lo * Ns, Ns, b_VL.as<double>(), nspad, Ns, multi ? 0.0 : sigma2);
      CU(cudaGetLastError());
      ctx->launches += 2;
    }
  }
  if (multi) {
    int r = nccl_api().AllReduce(d_mean, d_mean, 2 * nout, NCCL_DOUBLE, NCCL_SUM, ctx->comm, st);
    if (r != 0) return ctx->fail(LMM_E_NCCL, "ncclAllReduce failed");
    add_scalar_kernel<<<(unsigned)((nout + 255) / 256), 256, 0, st>>>(d_var, nout, sigma2);
    CU(cudaGetLastError());
    ++ctx->launches;
  }
  CU(copy_out(ctx, mean, d_mean, nout * sizeof(double)));
  CU(copy_out(ctx, var, d_var, nout * sizeof(double)));
  CU(cudaEventRecord(ctx->ev[1], st));
  CU(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->timings[0] = ms;
  ctx->timings[5] = ms;
  return LMM_OK;
}

// ------------------------------------------------------------------------------------------------
// OILMM prior marginals: src/oilmm.jl:57-76 with GP latents (mean const, var = variance + 1e-18)
// ------------------------------------------------------------------------------------------------
extern "C" int lmm_oilmm_prior_mean_and_var(lmm_ctx* ctx, const lmm_gp_desc* latents, int m, const double* xs, int Ns, int D,
                                            const double* U, const double* S, int p, double sigma2, int out_dim, double* mean,
                                            double* var) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_common(ctx, latents, m, xs, Ns, D, p, out_dim);
  if (rc) return rc;
  if (!U || !S || !mean || !var) return ctx->fail(LMM_E_ARG, "null pointer");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  Projection pr;
  std::vector<double> H;
  if ((rc = oilmm_projection(ctx, U, S, p, m, sigma2, Ns, pr, H))) return rc;
  // latent marginals are constants along n: build ML/VL on the host side of the (tiny) m x Ns arrays
  std::vector<double> ML((size_t)m * Ns), VL((size_t)m * Ns);
  for (int i = 0; i < m; ++i)
    for (int n = 0; n < Ns; ++n) {
      ML[(size_t)i * Ns + n] = latents[i].mean_const;
      VL[(size_t)i * Ns + n] = latents[i].variance;
    }
  DevBuf b_H, b_ML, b_VL, b_out;
  CU(b_H.alloc(ctx, H.size() * sizeof(double)));
  CU(copy_in(ctx, b_H.as<double>(), H.data(), H.size()));
  CU(b_ML.alloc(ctx, ML.size() * sizeof(double)));
  CU(copy_in(ctx, b_ML.as<double>(), ML.data(), ML.size()));
  CU(b_VL.alloc(ctx, VL.size() * sizeof(double)));
  CU(copy_in(ctx, b_VL.as<double>(), VL.data(), VL.size()));
  const size_t nout = (size_t)p * Ns;
  CU(b_out.alloc(ctx, 2 * nout * sizeof(double)));
  CU(launch_backproject(st, b_H.as<double>(), p, m, 0, m, b_ML.as<double>(), b_VL.as<double>(), Ns, Ns, 1e-18, sigma2, 1,
                        b_out.as<double>(), b_out.as<double>() + nout));
  ++ctx->launches;
  CU(copy_out(ctx, mean, b_out.p, nout * sizeof(double)));
  CU(copy_out(ctx, var, b_out.as<double>() + nout, nout * sizeof(double)));
  CU(cudaStreamSynchronize(st));
  return LMM_OK;
}

// ------------------------------------------------------------------------------------------------
// Batched Cholesky primitive
// ------------------------------------------------------------------------------------------------
extern "C" int lmm_potrf_batched(lmm_ctx* ctx, const double* A, int N, int batch, double* L_out, double* logdet_out, int* info) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!A || N <= 0 || batch <= 0) return ctx->fail(LMM_E_ARG, "null pointer or non-positive size");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (double& t : ctx->timings) t = 0.0;
  const int nt = ntiles(N);
  DevBuf b_A, b_L, b_W, b_logdet, b_info;
  const double* dA = A;
  if (!is_device_ptr(A)) {
    CU(b_A.alloc(ctx, (size_t)batch * N * N * sizeof(double)));
    CU(copy_in(ctx, b_A.as<double>(), A, (size_t)batch * N * N));
    dA = b_A.as<double>();
  }
  CU(b_L.alloc(ctx, (size_t)batch * sym_tiles(nt) * TT * sizeof(double)));
  CU(b_W.alloc(ctx, (size_t)batch * nt * TT * sizeof(double)));
  CU(b_logdet.alloc(ctx, (size_t)batch * sizeof(double)));
  CU(b_info.alloc(ctx, (size_t)batch * sizeof(int)));
  CU(cudaMemsetAsync(b_logdet.p, 0, (size_t)batch * sizeof(double), st));
  CU(cudaMemsetAsync(b_info.p, 0, (size_t)batch * sizeof(int), st));
  TiledSym L{b_L.as<double>(), nt, sym_tiles(nt) * TT};
  CU(launch_tile_from_dense(st, L, batch, dA, N));
  ++ctx->launches;
  CU(cudaEventRecord(ctx->ev[0], st));
  {
    PartitionScope scope(ctx);
    CU(chol_factor(ctx, L, b_W.as<double>(), (size_t)nt * TT, batch, b_logdet.as<double>(), b_info.as<int>()));
  }
  CU(cudaEventRecord(ctx->ev[1], st));
  std::vector<int> hinfo(batch, 0);
  CU(copy_out(ctx, hinfo.data(), b_info.p, (size_t)batch * sizeof(int)));
  if (logdet_out) CU(copy_out(ctx, logdet_out, b_logdet.p, (size_t)batch * sizeof(double)));
  if (L_out) {
    DevBuf dense;
    CU(dense.alloc(ctx, (size_t)N * N * sizeof(double)));
    for (int b = 0; b < batch; ++b) {
      CU(cudaMemsetAsync(dense.p, 0, (size_t)N * N * sizeof(double), st));
      CU(launch_untile_lower(st, L, b, dense.as<double>(), N));
      ++ctx->launches;
      CU(copy_out(ctx, L_out + (size_t)b * N * N, dense.p, (size_t)N * N * sizeof(double)));
      CU(cudaStreamSynchronize(st));
    }
  }
  CU(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  ctx->timings[0] = ms;
  ctx->timings[2] = ms;
  int worst = 0;
  for (int b = 0; b < batch; ++b) {
    int v = hinfo[b] > N ? N : hinfo[b];
    if (info) info[b] = v;
    if (v > worst) worst = v;
  }
  return worst;
}

extern "C" int lmm_potrf_bench(lmm_ctx* ctx, const lmm_gp_desc* desc, const double* x, int N, int D, double noise, int batch,
                               double* logdet_out, double* out_ms_kmat, double* out_ms_chol) {
  if (!ctx) return LMM_E_ARG;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!desc || !x || N <= 0 || D <= 0 || batch <= 0) return ctx->fail(LMM_E_ARG, "null pointer or non-positive size");
  int rc = check_descs(ctx, desc, 1);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (double& t : ctx->timings) t = 0.0;
  const int nt = ntiles(N);
  const size_t npad = (size_t)nt * TILE;
  DevBuf b_x, b_L, b_W, b_logdet, b_info, b_params;
  CU(b_x.alloc(ctx, npad * D * sizeof(double)));
  CU(cudaMemsetAsync(b_x.p, 0, npad * D * sizeof(double), st));
  CU(copy_in(ctx, b_x.as<double>(), x, (size_t)N * D));
  std::vector<LatentParams> hp(batch);
  for (int b = 0; b < batch; ++b) {
    hp[b].kind = desc->kind; hp[b].pad = 0; hp[b].variance = desc->variance; hp[b].inv_ls = desc->inv_lengthscale;
    hp[b].noise = noise; hp[b].mean = desc->mean_const;
  }
  CU(b_params.alloc(ctx, hp.size() * sizeof(LatentParams)));
  CU(cudaMemcpyAsync(b_params.p, hp.data(), hp.size() * sizeof(LatentParams), cudaMemcpyHostToDevice, st));
  CU(b_L.alloc(ctx, (size_t)batch * sym_tiles(nt) * TT * sizeof(double)));
  CU(b_W.alloc(ctx, (size_t)batch * nt * TT * sizeof(double)));
  CU(b_logdet.alloc(ctx, (size_t)batch * sizeof(double)));
  CU(b_info.alloc(ctx, (size_t)batch * sizeof(int)));
  CU(cudaMemsetAsync(b_logdet.p, 0, (size_t)batch * sizeof(double), st));
  CU(cudaMemsetAsync(b_info.p, 0, (size_t)batch * sizeof(int), st));
  TiledSym L{b_L.as<double>(), nt, sym_tiles(nt) * TT};
  CU(cudaEventRecord(ctx->ev[0], st));
  CU(launch_kmat_sym(st, L, batch, b_x.as<double>(), N, D, b_params.as<LatentParams>(), ctx->distance_form));
  ++ctx->launches;
  CU(cudaEventRecord(ctx->ev[1], st));
  {
    PartitionScope scope(ctx);
    CU(chol_factor(ctx, L, b_W.as<double>(), (size_t)nt * TT, batch, b_logdet.as<double>(), b_info.as<int>()));
  }
  CU(cudaEventRecord(ctx->ev[2], st));
  std::vector<int> hinfo(batch, 0);
  CU(copy_out(ctx, hinfo.data(), b_info.p, (size_t)batch * sizeof(int)));
  if (logdet_out) CU(copy_out(ctx, logdet_out, b_logdet.p, (size_t)batch * sizeof(double)));
  CU(cudaStreamSynchronize(st));
  float a = 0, c = 0;
  cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
  cudaEventElapsedTime(&c, ctx->ev[1], ctx->ev[2]);
  if (out_ms_kmat) *out_ms_kmat = a;
  if (out_ms_chol) *out_ms_chol = c;
  ctx->timings[0] = a + c; ctx->timings[1] = a; ctx->timings[2] = c;
  int worst = 0;
  for (int b = 0; b < batch; ++b) {
    int v = hinfo[b] > N ? N : hinfo[b];
    if (v > worst) worst = v;
  }
  return worst;
}

#include "api_ext.inc"
