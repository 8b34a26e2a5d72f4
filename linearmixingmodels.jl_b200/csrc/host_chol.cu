// Batched blocked Cholesky schedules and TRSM sweeps (K6 / K10 of SURVEY.md §2.3) on top of the tile kernels, and the
// `cholesky(Symmetric(C))` primitives of the C ABI.
#include "host_internal.h"

namespace lmm_host {

// ---- batched blocked Cholesky (left-looking over blocks of `outer_block` tile columns) -------
// For every block column [s0, s1): one wide trailing update against all previous columns
// (K = s0 tiles, output written once), then per tile column: narrow update inside the block,
// diagonal-tile factor (+ inverse, logdet, info), panel TRSM as a GEMM with the inverse.
// jstart > 0 EXTENDS a factor (block-Cholesky update, AbstractGPs' sequential conditioning): the tile rows < jstart of L
// and W(J), J < jstart, are already final; only the tile rows >= jstart are computed -- for the columns J < jstart that
// is L(I,J) = (A(I,J) - sum_{k<J} L(I,k) L(J,k)') W(J)' (no diagonal-tile step), from column jstart on the ordinary
// factorisation of the Schur complement.  Cost O(N² N₂) instead of O((N + N₂)³).
cudaError_t chol_factor_stream(lmm_ctx* ctx, cudaStream_t st, TiledSym L, double* W, size_t wstride, int batch, double* logdet,
                               int* info, int jstart = 0) {
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
      const int r0 = s0 > jstart ? s0 : jstart;  // first tile row that still has to be computed
      g.i0 = r0; g.j0 = s0; g.k0 = 0; g.k1 = s0;
      if ((e = launch_gemm(st, GEMM_UPDATE, g, s1 - s0, nt - r0, batch)) != cudaSuccess) return e;
      ++ctx->launches;
      ctx->timings[6] += 1;
    }
    for (int jj = s0; jj < s1; ++jj) {
      const int r0 = jj > jstart ? jj : jstart;
      if (jj > s0) {
        g.i0 = r0; g.j0 = jj; g.k0 = s0; g.k1 = jj;
        if ((e = launch_gemm(st, GEMM_UPDATE, g, 1, nt - r0, batch)) != cudaSuccess) return e;
        ++ctx->launches;
        ctx->timings[6] += 1;
      }
      if (jj >= jstart) {
        if ((e = launch_potrf_tile(st, L, W, wstride, jj, batch, logdet, info)) != cudaSuccess) return e;
        ++ctx->launches;
      }
      const int t0 = jj + 1 > jstart ? jj + 1 : jstart;  // rows of column jj below the diagonal tile that are not final yet
      if (t0 < nt) {
        g.i0 = t0; g.j0 = jj;
        if ((e = launch_gemm(st, GEMM_TRSM, g, 1, nt - t0, batch)) != cudaSuccess) return e;
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
  // "panel_split": only the DIAGONAL tile of the next column is updated on the panel stream before its factorisation;
  // the rest of that column -- needed by the TRSM that follows the diagonal-tile kernel, not by the kernel itself -- is
  // updated on a second high-priority stream X2 meanwhile.
  const bool split = ctx->panel_split != 0;
  cudaError_t e;
  while ((int)ctx->blk_ev.size() < 2 * nblk + 2 + 2 * nt + 2) {
    cudaEvent_t ev;
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    ctx->blk_ev.push_back(ev);
  }
  cudaStream_t X = ctx->panel_stream, Y = ctx->update_stream, X2 = ctx->xchg_stream;
  cudaEvent_t* evX = ctx->blk_ev.data();          // panel(b) done on X
  cudaEvent_t* evY = ctx->blk_ev.data() + nblk;   // trailing update from block b done on Y
  cudaEvent_t* evT = ctx->blk_ev.data() + 2 * nblk + 2;  // split: TRSM of column j done on X
  cudaEvent_t* evU = evT + nt;                           // split: rest-of-column update(s) up to column j done on X2
  int last_u = -1;                                       // latest evU that X has not waited for yet
  GemmArgs g{};
  g.A = operand(L); g.B = operand(L); g.C = operand(L);
  g.W = W; g.w_batch_stride = wstride; g.sym = 1;
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(X, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if ((e = cudaStreamWaitEvent(Y, ctx->ev_fork, 0)) != cudaSuccess) return e;
  if (split && (e = cudaStreamWaitEvent(X2, ctx->ev_fork, 0)) != cudaSuccess) return e;
  for (int b = 0; b < nblk; ++b) {
    const int s0 = b * ob, s1 = (s0 + ob < nt) ? s0 + ob : nt;
    if (b >= 1) {
      // every earlier update of this block column (Y launches up to b-2) must have landed
      if (b >= 2 && (e = cudaStreamWaitEvent(X, evY[b - 2], 0)) != cudaSuccess) return e;
      g.k0 = s0 - ob; g.k1 = s0;
      if (!split) {
        g.i0 = s0; g.j0 = s0;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, s1 - s0, nt - s0, batch)) != cudaSuccess) return e;
        ++ctx->launches;
      } else {
        g.i0 = s0; g.j0 = s0;
        if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, 1, batch)) != cudaSuccess) return e;
        ++ctx->launches;
        if (s0 + 1 < nt) {  // all other tiles of the block columns (the symmetric-skip leaves row s0 to the launch above)
          if (b >= 2 && (e = cudaStreamWaitEvent(X2, evY[b - 2], 0)) != cudaSuccess) return e;
          if ((e = cudaStreamWaitEvent(X2, evT[s0 - 1], 0)) != cudaSuccess) return e;
          g.i0 = s0 + 1; g.j0 = s0;
          if ((e = launch_gemm(X2, GEMM_UPDATE, g, s1 - s0, nt - s0 - 1, batch)) != cudaSuccess) return e;
          ++ctx->launches;
          if ((e = cudaEventRecord(evU[s0], X2)) != cudaSuccess) return e;
          last_u = s0;
        }
      }
      ctx->timings[6] += 1;
    }
    for (int jj = s0; jj < s1; ++jj) {
      if (jj > s0) {
        g.k0 = s0; g.k1 = jj;
        if (!split) {
          g.i0 = jj; g.j0 = jj;
          if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, nt - jj, batch)) != cudaSuccess) return e;
          ++ctx->launches;
        } else {
          g.i0 = jj; g.j0 = jj;
          if ((e = launch_gemm(X, GEMM_UPDATE, g, 1, 1, batch)) != cudaSuccess) return e;
          ++ctx->launches;
          if (jj + 1 < nt) {
            if ((e = cudaStreamWaitEvent(X2, evT[jj - 1], 0)) != cudaSuccess) return e;
            g.i0 = jj + 1; g.j0 = jj;
            if ((e = launch_gemm(X2, GEMM_UPDATE, g, 1, nt - jj - 1, batch)) != cudaSuccess) return e;
            ++ctx->launches;
            if ((e = cudaEventRecord(evU[jj], X2)) != cudaSuccess) return e;
            last_u = jj;
          }
        }
      }
      if ((e = launch_potrf_tile(X, L, W, wstride, jj, batch, logdet, info)) != cudaSuccess) return e;
      ++ctx->launches;
      if (jj + 1 < nt) {
        if (split && last_u >= 0) {
          if ((e = cudaStreamWaitEvent(X, evU[last_u], 0)) != cudaSuccess) return e;
          last_u = -1;
        }
        g.i0 = jj + 1; g.j0 = jj;
        if ((e = launch_gemm(X, GEMM_TRSM, g, 1, nt - jj - 1, batch)) != cudaSuccess) return e;
        ++ctx->launches;
        if (split && (e = cudaEventRecord(evT[jj], X)) != cudaSuccess) return e;
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
  if (split) {  // every X2 launch is followed by a TRSM on X that waited for it; joined explicitly all the same
    if ((e = cudaEventRecord(ctx->ev_join[1], X2)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join[1], 0)) != cudaSuccess) return e;
  }
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

// Fork the batch into latent groups on separate streams (joined back into ctx->stream).  jstart > 0: extend a factor whose
// tile rows < jstart are final (see chol_factor_stream).
cudaError_t chol_factor(lmm_ctx* ctx, TiledSym L, double* W, size_t wstride, int batch, double* logdet, int* info, int jstart) {
  const int G = ctx->ngroups < batch ? ctx->ngroups : batch;
  if (jstart == 0) {  // the look-ahead / partitioned schedules factor from scratch only
    if (ctx->partition_ilmm && ctx->partition_now && batch == 1 && ctx->comm && ctx->nranks > 1 && L.nt >= 2 * ctx->nranks && nccl_api().AllGather)
      return (ctx->partition_ilmm == 2 && ctx->comm2) ? chol_factor_rowcyclic2(ctx, L, W, wstride, logdet, info)
                                                      : chol_factor_rowcyclic(ctx, L, W, wstride, logdet, info);
    if (ctx->lookahead == 2 && batch <= 2 && L.nt >= 12) return chol_factor_rightlooking(ctx, L, W, wstride, batch, logdet, info);
    if (ctx->lookahead && batch <= 2 && L.nt >= 12) return chol_factor_lookahead(ctx, L, W, wstride, batch, logdet, info);
  }
  if (G <= 1 || L.nt <= 1) return chol_factor_stream(ctx, ctx->stream, L, W, wstride, batch, logdet, info, jstart);
  cudaError_t e;
  if ((e = cudaEventRecord(ctx->ev_fork, ctx->stream)) != cudaSuccess) return e;
  for (int gi = 0; gi < G; ++gi) {
    const int b0 = (int)((int64_t)batch * gi / G), b1 = (int)((int64_t)batch * (gi + 1) / G);
    cudaStream_t st = ctx->gstream[gi];
    if ((e = cudaStreamWaitEvent(st, ctx->ev_fork, 0)) != cudaSuccess) return e;
    TiledSym Lg{L.base + (size_t)b0 * L.batch_stride, L.nt, L.batch_stride};
    if ((e = chol_factor_stream(ctx, st, Lg, W + (size_t)b0 * wstride, wstride, b1 - b0, logdet + b0, info + b0, jstart)) != cudaSuccess) return e;
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

}  // namespace lmm_host

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
  int rc = check_descs(ctx, desc, 1, D);
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
    set_params(hp[b], *desc, noise, 1.0, D);
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
