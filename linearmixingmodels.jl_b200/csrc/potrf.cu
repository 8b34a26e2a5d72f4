// Diagonal-tile factorisation: the panel step (dpotf2) of the blocked Cholesky (K6), fused with
// logdet accumulation (K7), non-PD detection and the triangular inverse W = inv(L_JJ) that turns
// every panel TRSM and every triangular vector solve into tensor-core GEMMs / plain GEMVs.
// One CTA (8 warps) per latent; the 128x128 tile lives in shared memory, column-major with
// ld = 132 (conflict-free DMMA fragment loads: column stride = 8 banks).
//
// Two kernels.  potrf_tile_kernel (first generation, "potrf_impl" = 0):
//   phase C: 16 steps of 8 columns: 8x8 diagonal block factored in registers (rsqrt pivots,
//            redundantly by every row-owner thread -> no intra-block syncs), panel rows solved
//            in registers, trailing update of the tile on the FP64 tensor pipe (m8n8k4 DMMA).
//   phase W: in-place recursive inverse of the lower triangle: 8x8 diagonal blocks by
//            substitution, then 4 doubling levels  W21 = -W22 (L21 W11)  as DMMA block products;
//            the temporary L21*W11 lives in the (free) mirrored upper block.
// potrf_tile_kernel2 (default): left-looking column updates, reciprocal pivot chain, W built beside
// the panel steps -- described above its definition.
// Latency-bound by design (it sits on the critical path of each tile column; the host runs
// several latent groups on separate streams so the other SMs keep doing trailing updates).
#include "common.cuh"
#include "kernels.h"

namespace lmm {

#ifdef LMM_POTRF_TIMING
__device__ long long g_potrf_clk[16];
__device__ long long g_potrf_acc[8];  // per-step sums: [0] column update, [1] panel (warp 0), [2] its wait at the step barrier, [3] W block row (warp 4)
#define PT(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_potrf_clk[k] = clock64(); } while (0)
#define ACC_DECL long long acc_t0 = 0
#define ACC_START(tidsel) do { if (blockIdx.x == 0 && threadIdx.x == (tidsel)) acc_t0 = clock64(); } while (0)
#define ACC_STOP(tidsel, k) do { if (blockIdx.x == 0 && threadIdx.x == (tidsel)) g_potrf_acc[k] += clock64() - acc_t0; } while (0)
#else
#define PT(k) do { } while (0)
#define ACC_DECL do { } while (0)
#define ACC_START(tidsel) do { } while (0)
#define ACC_STOP(tidsel, k) do { } while (0)
#endif

constexpr int LD = 132;
constexpr size_t POTRF_SMEM = (size_t)(TILE * LD + 2 * TILE + 32) * sizeof(double);

// 1/sqrt(x) for x > 0 in the normal range: MUFU.RSQ64H seed + two Newton-Raphson steps.  Measured on B200
// (tools/microbench/rsqrt_check.cu, max relative error over 2e7 arguments): seed 9.2e-7, one step 1.3e-12, two steps
// 2.7e-16, three steps 2.6e-16 -- the third step bought nothing and sat on the pivot chain (128 x 2 dependent FMAs).
// Non-positive / NaN pivots are reported separately by the caller: no special cases on the critical path.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

__device__ __forceinline__ void dmma884p(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// C(8x8 at R0,N0) (+)= sign * sum over 8-wide k-blocks  A(R0, K0+8kk) * B(K0+8kk, N0)
// A block element (r,k) at S[(K0+k)*LD + R0+r]; B block element (k,n) at S[(N0+n)*LD + K0+k].
__device__ __forceinline__ void block_mma(const double* S, int R0, int KA0, int KB0, int N0, int nk, double neg, double& c0, double& c1,
                                          int g, int t) {
  // two independent accumulator chains (even / odd k-blocks) hide the DMMA latency
  double e0 = 0.0, e1 = 0.0;
  int kk = 0;
  for (; kk + 1 < nk; kk += 2) {
    const int ka = KA0 + 8 * kk, kb = KB0 + 8 * kk;
    const double a0 = neg * S[(ka + t) * LD + R0 + g];
    const double a1 = neg * S[(ka + 4 + t) * LD + R0 + g];
    const double b0 = S[(N0 + g) * LD + kb + t];
    const double b1 = S[(N0 + g) * LD + kb + 4 + t];
    const double a2 = neg * S[(ka + 8 + t) * LD + R0 + g];
    const double a3 = neg * S[(ka + 12 + t) * LD + R0 + g];
    const double b2 = S[(N0 + g) * LD + kb + 8 + t];
    const double b3 = S[(N0 + g) * LD + kb + 12 + t];
    dmma884p(c0, c1, a0, b0);
    dmma884p(e0, e1, a2, b2);
    dmma884p(c0, c1, a1, b1);
    dmma884p(e0, e1, a3, b3);
  }
  if (kk < nk) {
    const int ka = KA0 + 8 * kk, kb = KB0 + 8 * kk;
    const double a0 = neg * S[(ka + t) * LD + R0 + g];
    const double a1 = neg * S[(ka + 4 + t) * LD + R0 + g];
    const double b0 = S[(N0 + g) * LD + kb + t];
    const double b1 = S[(N0 + g) * LD + kb + 4 + t];
    dmma884p(c0, c1, a0, b0);
    dmma884p(c0, c1, a1, b1);
  }
  c0 += e0;
  c1 += e1;
}

__global__ void __launch_bounds__(256, 1) potrf_tile_kernel(TiledSym L, double* __restrict__ Wbase, size_t w_batch_stride, int J,
                                                            double* __restrict__ logdet, int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];  // S[c*LD + r]
  double* dinv = S + TILE * LD;               // 1 / L[r][r]
  double* red = dinv + TILE;                  // reduction scratch
  __shared__ int fail_col;
  __shared__ unsigned char tri_bi[120], tri_bj[120];  // packed lower-triangular block order (row by row)

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* tile = L.tile(b, J, J);
  double* Wt = Wbase + (size_t)b * w_batch_stride + (size_t)J * TT;

  PT(0);
  if (tid == 0) fail_col = 0x7fffffff;
  if (tid < 120) {
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= tid) ++bi;
    tri_bi[tid] = (unsigned char)bi;
    tri_bj[tid] = (unsigned char)(tid - bi * (bi + 1) / 2);
  }
#pragma unroll 16
  for (int e = tid; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    S[c * LD + r] = tile[e];
  }
  __syncthreads();

  PT(1);
  // ---- phase C
  for (int j0 = 0; j0 < TILE; j0 += 8) {
    const int r = tid;  // row owner (threads 0..127)
    double p[8];
    double d[8][8];
    const bool active = (tid < TILE) && (r >= j0);
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k <= i; ++k) d[i][k] = S[(j0 + k) * LD + (j0 + i)];
#pragma unroll
      for (int k = 0; k < 8; ++k) p[k] = S[(j0 + k) * LD + r];
    }
    __syncthreads();
    if (active) {
      double dv[8];
      // right-looking 8x8 factor in registers (short critical path: rsqrt + 2 FMAs per column)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const double piv = d[jj][jj];
        if (!(piv > 0.0)) {
          if (r == j0) atomicMin(&fail_col, j0 + jj);
        }
        const double inv = fast_rsqrt(piv);
        dv[jj] = inv;
        d[jj][jj] = piv * inv;
#pragma unroll
        for (int i = jj + 1; i < 8; ++i) d[i][jj] *= inv;
#pragma unroll
        for (int j2 = jj + 1; j2 < 8; ++j2)
#pragma unroll
          for (int i = j2; i < 8; ++i) d[i][j2] = fma(-d[i][jj], d[j2][jj], d[i][j2]);
      }
      if (r >= j0 + 8) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          double v = p[jj];
#pragma unroll
          for (int k = 0; k < jj; ++k) v = fma(-p[k], d[jj][k], v);
          p[jj] = v * dv[jj];
        }
      } else {
        const int i = r - j0;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          double v = 0.0;
#pragma unroll
          for (int ii = 0; ii < 8; ++ii)
            if (ii == i && jj <= ii) v = d[ii][jj];
          p[jj] = v;
        }
        double di = 0.0;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
          if (ii == i) di = dv[ii];
        dinv[r] = di;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) S[(j0 + k) * LD + r] = p[k];
    }
    __syncthreads();
    if (j0 == 0) PT(2);
    // trailing update of the lower 8x8 blocks of rows/cols [j0+8, 128) on the tensor pipe
    const int nb = (TILE - j0 - 8) >> 3;
    const int nblocks = nb * (nb + 1) / 2;
    for (int base = warp; base < nblocks; base += 32) {  // 4 independent blocks per pass, branch-free
      double c0[4], c1[4], a0[4], a1[4], b0[4], b1[4];
      int R0[4], C0[4];
      bool valid[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        valid[u] = base + 8 * u < nblocks;
        const int idx = valid[u] ? base + 8 * u : nblocks - 1;  // out-of-range slots recompute a real block, never store
        R0[u] = j0 + 8 + 8 * tri_bi[idx];
        C0[u] = j0 + 8 + 8 * tri_bj[idx];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c0[u] = S[(C0[u] + 2 * t) * LD + R0[u] + g];
        c1[u] = S[(C0[u] + 2 * t + 1) * LD + R0[u] + g];
        a0[u] = -S[(j0 + t) * LD + R0[u] + g];
        a1[u] = -S[(j0 + 4 + t) * LD + R0[u] + g];
        b0[u] = S[(j0 + t) * LD + C0[u] + g];
        b1[u] = S[(j0 + 4 + t) * LD + C0[u] + g];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884p(c0[u], c1[u], a0[u], b0[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) dmma884p(c0[u], c1[u], a1[u], b1[u]);
      __syncwarp();  // all lanes have read the C blocks of this pass before anyone overwrites them
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (valid[u]) {
          S[(C0[u] + 2 * t) * LD + R0[u] + g] = c0[u];
          S[(C0[u] + 2 * t + 1) * LD + R0[u] + g] = c1[u];
        }
    }
    __syncthreads();
    if (j0 == 0) PT(3);
  }
  PT(4);

  // ---- logdet and failure report (fixed reduction order)
  if (tid < TILE) {
    double v = log(S[tid * LD + tid]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
  }
  __syncthreads();
  if (tid == 0) {
    logdet[b] += 2.0 * (((red[0] + red[1]) + red[2]) + red[3]);
    if (fail_col != 0x7fffffff && info[b] == 0) info[b] = J * TILE + fail_col + 1;
  }

  PT(5);
  // ---- write L (lower, zero upper)
#pragma unroll 8
  for (int e = tid; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    tile[e] = (r >= c) ? S[c * LD + r] : 0.0;
  }

  PT(6);
  // ---- phase W level 0: invert the 16 diagonal 8x8 blocks in place (thread = one column)
  {
    double Lb[8][8];
    const int blk = tid >> 3, col = tid & 7, o = blk * 8;
    if (tid < TILE) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < i; ++k) Lb[i][k] = S[(o + k) * LD + o + i];
    }
    __syncthreads();
    if (tid < TILE) {
      double w[8];
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) {
        double s = (rr == col) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < rr; ++k) s = fma(-Lb[rr][k], w[k], s);
        w[rr] = s * dinv[o + rr];
      }
#pragma unroll
      for (int rr = 0; rr < 8; ++rr) S[(o + col) * LD + o + rr] = w[rr];
    }
    __syncthreads();
  }
  PT(7);
  // ---- phase W doubling levels: W21 = -W22 * (L21 * W11)
  for (int s = 8, lg = 0; s < TILE; s <<= 1, ++lg) {
    const int sb = s >> 3;               // 8-blocks per side (= 1 << lg)
    const int npairs = TILE / (2 * s);
    const int nout = npairs * sb * sb;   // output 8x8 blocks per phase
    // T(i,j) = sum_{k=j}^{sb-1} L21(i,k) W11(k,j)   -> stored in the mirrored upper block (1,2)
    for (int ob = warp; ob < nout; ob += 8) {
      const int q = ob >> (2 * lg), ij = ob & (sb * sb - 1), i = ij & (sb - 1), j = ij >> lg;
      const int base = q * 2 * s;
      double c0 = 0.0, c1 = 0.0;
      // A = L21 block (rows base+s+8i, cols base+8k), B = W11 block (rows base+8k, cols base+8j)
      block_mma(S, base + s + 8 * i, base + 8 * j, base + 8 * j, base + 8 * j, sb - j, 1.0, c0, c1, g, t);
      const int TR = base + 8 * i, TC = base + s + 8 * j;
      S[(TC + 2 * t) * LD + TR + g] = c0;
      S[(TC + 2 * t + 1) * LD + TR + g] = c1;
    }
    __syncthreads();
    // W21(i,j) = -sum_{k=0}^{i} W22(i,k) T(k,j)
    for (int ob = warp; ob < nout; ob += 8) {
      const int q = ob >> (2 * lg), ij = ob & (sb * sb - 1), i = ij & (sb - 1), j = ij >> lg;
      const int base = q * 2 * s;
      double c0 = 0.0, c1 = 0.0;
      // A = W22 block (rows base+s+8i, cols base+s+8k), B = T block (rows base+8k, cols base+s+8j)
      block_mma(S, base + s + 8 * i, base + s, base, base + s + 8 * j, i + 1, -1.0, c0, c1, g, t);
      const int WR = base + s + 8 * i, WC = base + 8 * j;
      S[(WC + 2 * t) * LD + WR + g] = c0;
      S[(WC + 2 * t + 1) * LD + WR + g] = c1;
    }
    __syncthreads();
  }
  PT(8);
#pragma unroll 8
  for (int e = tid; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    Wt[e] = (r >= c) ? S[c * LD + r] : 0.0;
  }
  PT(9);
}

// ---------------------------------------------------------------------------------------------------------------------
// Second-generation diagonal-tile kernel (default).  Same results (L, W = inv(L), logdet, info), about half the cycles:
//   * LEFT-looking inside the tile.  The first kernel's right-looking trailing update re-reads and re-writes every 8x8
//     block of the tile once per panel step (6 loads + 2 stores per pair of DMMAs: shared-memory-bandwidth bound, a third
//     of its cycles); here the column block about to be factored is brought up to date in one go,
//     C(bi, bj) -= sum_{k<bj} L(bi, k) L(bj, k)^T, accumulators in registers, 3 loads per pair of DMMAs, one
//     read-modify-write per block, by all 8 warps (two blocks of the column per warp sharing the B fragments);
//   * the pivot chain of each 8x8 diagonal block runs on reciprocals (LDL^T style): next pivot = fma(-u^2, 1/piv, d),
//     one MUFU.RCP64H + 3 dependent FMAs + 1 per column; the rsqrt that scales the stored column is computed beside the
//     chain, not on it.  The row owners of the diagonal block take the common path (no divergent branch on the critical
//     warp) and the non-positive-pivot report leaves the pivot loop;
//   * the two 64x64 diagonal halves of W are built one block row at a time, by warps 4-7, WHILE warps 0-3 run the
//     (row-owner, 128-thread) panel step of the next 8 columns:  W_ii = inv(L_ii),  W_ij = -W_ii * sum_{k=j}^{i-1} L_ik W_kj
//     needs only rows of L that are already final, and restricted to its own half a block row costs less than a panel
//     step.  W lives transposed in the free upper triangle of the tile (W(r,c) at S[r*LD+c], its diagonal in dinv[]),
//     so L stays intact underneath.  The off-diagonal half W21 = -W22 (L21 W11) follows the last panel step as two fully
//     unrolled DMMA stages on all 8 warps;
//   * tile load and both stores move 16 B per thread-access.
// Measured (tools/microbench/potrf_phases.cu, B200): 47.1 us -> 36.9 us per tile (profiles/r01_panel_chain.md).
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);  // seed is good to ~2^-20 (tools/microbench/rsqrt_check.cu); cubic step -> 2^-60
  return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// element (rr, cc) of the lower-triangular 8x8 block W_oo in the mirrored storage
__device__ __forceinline__ double wdiag(const double* S, const double* dinv, int o, int rr, int cc) {
  double v = 0.0;
  if (rr > cc) v = S[(o + rr) * LD + o + cc];
  if (rr == cc) v = dinv[o + rr];
  return v;
}

// Block row i of W restricted to the block columns [jmin, i), by `nw` warps (this warp = wl) sharing the named barrier
// `bar_id`.  Deliberately light on the FP64 pipe (one block at a time per warp, two accumulator chains): the pipe is
// shared with the panel warps, whose dependent chain is the critical path -- denser DMMA issue here was measured to slow
// the whole step down.
__device__ __forceinline__ void w_block_row(double* S, const double* dinv, int i, int jmin, int wl, int nw, int lane, int bar_id) {
  const int g = lane >> 2, t = lane & 3, o = 8 * i;
  if (wl == nw - 1 && lane < 8) {  // W_ii by substitution, one thread per column
    double Lb[8][8], w[8];
    const int col = lane;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int k = 0; k < a; ++k) Lb[a][k] = S[(o + k) * LD + o + a];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      double s = (rr == col) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < rr; ++k) s = fma(-Lb[rr][k], w[k], s);
      w[rr] = s * dinv[o + rr];
    }
#pragma unroll
    for (int rr = 1; rr < 8; ++rr)
      if (rr > col) S[(o + rr) * LD + o + col] = w[rr];
  }
  __syncwarp();
  // T_j = sum_{k=j}^{i-1} L_ik W_kj, parked where W_ij will live
  for (int j = jmin + wl; j < i; j += nw) {
    const int oj = 8 * j;
    double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
    {
      const double a0 = S[(oj + t) * LD + o + g];
      const double a1 = S[(oj + 4 + t) * LD + o + g];
      const double b0 = wdiag(S, dinv, oj, t, g);
      const double b1 = wdiag(S, dinv, oj, 4 + t, g);
      dmma884p(c0, c1, a0, b0);
      dmma884p(e0, e1, a1, b1);
    }
    for (int k = j + 1; k < i; ++k) {
      const int ok = 8 * k;
      const double a0 = S[(ok + t) * LD + o + g];
      const double a1 = S[(ok + 4 + t) * LD + o + g];
      const double b0 = S[(ok + t) * LD + oj + g];
      const double b1 = S[(ok + 4 + t) * LD + oj + g];
      dmma884p(c0, c1, a0, b0);
      dmma884p(e0, e1, a1, b1);
    }
    S[(o + g) * LD + oj + 2 * t] = c0 + e0;
    S[(o + g) * LD + oj + 2 * t + 1] = c1 + e1;
  }
  named_bar(bar_id, nw * 32);  // W_ii (and every T_j) visible to the warps of this group
  const double wa0 = wdiag(S, dinv, o, g, t);
  const double wa1 = wdiag(S, dinv, o, g, 4 + t);
  for (int j = jmin + wl; j < i; j += nw) {
    const int oj = 8 * j;
    double c0 = 0.0, c1 = 0.0;
    const double b0 = S[(o + t) * LD + oj + g];
    const double b1 = S[(o + 4 + t) * LD + oj + g];
    dmma884p(c0, c1, wa0, b0);
    dmma884p(c0, c1, wa1, b1);
    __syncwarp();  // every lane has read T_j before it is overwritten
    S[(o + g) * LD + oj + 2 * t] = -c0;
    S[(o + g) * LD + oj + 2 * t + 1] = -c1;
  }
}

// The off-diagonal half of W once both 64x64 diagonal halves are inverted: W21 = -W22 (L21 W11), 8 warps, everything
// unrolled (no branches between the loads and the DMMAs: this phase has the pipe to itself and is meant to fill it).
//   stage A: warp w owns block row i = 8 + w:    T_ij = sum_{k=j}^{7} L_ik W_kj      (A fragments shared along the row)
//   stage B: warp w owns block column j = w:     W_ij = -sum_{k=8}^{i} W_ik T_kj     (B fragments shared along the column)
__device__ __forceinline__ void w_lower_left(double* S, const double* dinv, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  {
    const int o = 64 + 8 * warp;
    double c0[8], c1[8], e0[8], e1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c0[j] = c1[j] = e0[j] = e1[j] = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ok = 8 * k;
      const double a0 = S[(ok + t) * LD + o + g];
      const double a1 = S[(ok + 4 + t) * LD + o + g];
#pragma unroll
      for (int j = 0; j <= k; ++j) {
        const int oj = 8 * j;
        double b0, b1;
        if (j == k) {
          b0 = wdiag(S, dinv, oj, t, g);
          b1 = wdiag(S, dinv, oj, 4 + t, g);
        } else {
          b0 = S[(ok + t) * LD + oj + g];
          b1 = S[(ok + 4 + t) * LD + oj + g];
        }
        dmma884p(c0[j], c1[j], a0, b0);
        dmma884p(e0[j], e1[j], a1, b1);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      S[(o + g) * LD + 8 * j + 2 * t] = c0[j] + e0[j];
      S[(o + g) * LD + 8 * j + 2 * t + 1] = c1[j] + e1[j];
    }
  }
  __syncthreads();
  {
    const int oj = 8 * warp;
    double c0[8], c1[8], e0[8], e1[8];
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) c0[ii] = c1[ii] = e0[ii] = e1[ii] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int ok = 64 + 8 * kk;
      const double b0 = S[(ok + t) * LD + oj + g];
      const double b1 = S[(ok + 4 + t) * LD + oj + g];
#pragma unroll
      for (int ii = kk; ii < 8; ++ii) {
        const int o = 64 + 8 * ii;
        double a0, a1;
        if (ii == kk) {
          a0 = wdiag(S, dinv, o, g, t);
          a1 = wdiag(S, dinv, o, g, 4 + t);
        } else {
          a0 = S[(o + g) * LD + ok + t];
          a1 = S[(o + g) * LD + ok + 4 + t];
        }
        dmma884p(c0[ii], c1[ii], a0, b0);
        dmma884p(e0[ii], e1[ii], a1, b1);
      }
    }
    __syncthreads();  // every T block has been read by every warp that needs it
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int o = 64 + 8 * ii;
      S[(o + g) * LD + oj + 2 * t] = -(c0[ii] + e0[ii]);
      S[(o + g) * LD + oj + 2 * t + 1] = -(c1[ii] + e1[ii]);
    }
  }
}

__global__ void __launch_bounds__(256, 1) potrf_tile_kernel2(TiledSym L, double* __restrict__ Wbase, size_t w_batch_stride, int J,
                                                             double* __restrict__ logdet, int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];  // S[c*LD + r]: L in the lower triangle, W^T strictly above the diagonal
  double* dinv = S + TILE * LD;               // 1 / L[r][r] = W[r][r]
  double* red = dinv + TILE;
  __shared__ int fail_col;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* tile = L.tile(b, J, J);
  double* Wt = Wbase + (size_t)b * w_batch_stride + (size_t)J * TT;

  PT(0);
  if (tid == 0) fail_col = 0x7fffffff;
  pdl_wait();  // the tile was written by the previous kernel of the panel chain
#pragma unroll 16
  for (int e2 = tid; e2 < TT / 2; e2 += 256) {
    const double2 v = reinterpret_cast<const double2*>(tile)[e2];
    int r, c;
    tile_rc(2 * e2, r, c);
    S[c * LD + r] = v.x;
    S[(c + 1) * LD + r] = v.y;
  }
  __syncthreads();

  PT(1);
  ACC_DECL;
  for (int j0 = 0; j0 < TILE; j0 += 8) {
    const int bj = j0 >> 3;
    ACC_START(0);
    if (bj > 0) {
      // left-looking update of column block bj: block rows bi0 = bj + warp and bi0 + 8.  A dependent DMMA costs ~100
      // cycles, so chain length is what matters: a warp with two blocks runs 4 chains of bj DMMAs; a warp with one block
      // (always the case once bj >= 8, when the chains are longest) splits k in two halves, 4 chains of bj/2.
      const int bi0 = bj + warp;
      if (bi0 < 16) {
        const bool two = bi0 + 8 < 16;
        const int R0 = 8 * bi0, R1 = two ? R0 + 64 : R0;
        const int trips = two ? bj : (bj + 1) >> 1;
        const int koff = two ? 0 : 8 * trips;  // column offset of the second pair of chains
        const int klast = 8 * (bj - 1);
        double p0 = 0.0, p1 = 0.0, q0 = 0.0, q1 = 0.0;  // block 0 (first half of k)
        double u0 = 0.0, u1 = 0.0, v0 = 0.0, v1 = 0.0;  // block 1, or block 0's second half of k
        // fragments of trip k + 1 are in flight while the DMMAs of trip k issue
        int c2 = koff < klast ? koff : klast;
        double bb0 = S[t * LD + j0 + g], bb1 = S[(4 + t) * LD + j0 + g];
        double a00 = S[t * LD + R0 + g], a01 = S[(4 + t) * LD + R0 + g];
        double bb2 = S[(c2 + t) * LD + j0 + g], bb3 = S[(c2 + 4 + t) * LD + j0 + g];
        double a10 = S[(c2 + t) * LD + R1 + g], a11 = S[(c2 + 4 + t) * LD + R1 + g];
        for (int k = 0; k < trips; ++k) {
          const int cn = 8 * (k + 1 < trips ? k + 1 : k);
          const int c2n = cn + koff < klast ? cn + koff : klast;
          const double nb0 = S[(cn + t) * LD + j0 + g];
          const double nb1 = S[(cn + 4 + t) * LD + j0 + g];
          const double n00 = S[(cn + t) * LD + R0 + g];
          const double n01 = S[(cn + 4 + t) * LD + R0 + g];
          const double nb2 = S[(c2n + t) * LD + j0 + g];
          const double nb3 = S[(c2n + 4 + t) * LD + j0 + g];
          const double n10 = S[(c2n + t) * LD + R1 + g];
          const double n11 = S[(c2n + 4 + t) * LD + R1 + g];
          dmma884p(p0, p1, a00, bb0);
          dmma884p(q0, q1, a01, bb1);
          if (two || 8 * k + koff <= klast) {  // warp-uniform; false only for the odd trip out of a split k range
            dmma884p(u0, u1, a10, bb2);
            dmma884p(v0, v1, a11, bb3);
          }
          bb0 = nb0; bb1 = nb1; a00 = n00; a01 = n01;
          bb2 = nb2; bb3 = nb3; a10 = n10; a11 = n11;
        }
        double* c0p = &S[(j0 + 2 * t) * LD + R0 + g];
        if (two) {
          c0p[0] -= p0 + q0;
          c0p[LD] -= p1 + q1;
          double* c1p = &S[(j0 + 2 * t) * LD + R1 + g];
          c1p[0] -= u0 + v0;
          c1p[LD] -= u1 + v1;
        } else {
          c0p[0] -= (p0 + q0) + (u0 + v0);
          c0p[LD] -= (p1 + q1) + (u1 + v1);
        }
      }
      __syncthreads();  // the column block is up to date for its row owners
    }
    ACC_STOP(0, 0);
    if (j0 == 8) PT(2);
    ACC_START(0);
    ACC_START(128);
    if (warp < 4) {
      const int r = tid;  // row owner
      double p[8];
      double d[8][8];
      const bool active = r >= j0;
      if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int k = 0; k <= i; ++k) d[i][k] = S[(j0 + k) * LD + (j0 + i)];
#pragma unroll
        for (int k = 0; k < 8; ++k) p[k] = S[(j0 + k) * LD + r];
      }
      named_bar(1, 128);  // every row owner has read the diagonal block before its rows are overwritten
      if (active) {
        double dv[8], tm[8][8];
        int failj = 8;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const double piv = d[jj][jj];
          if (!(piv > 0.0) && failj == 8) failj = jj;
          const double rc = fast_rcp(piv);
          dv[jj] = fast_rsqrt(piv);
#pragma unroll
          for (int j2 = jj + 1; j2 < 8; ++j2) {
            d[j2][j2] = fma(-(d[j2][jj] * d[j2][jj]), rc, d[j2][j2]);  // the pivot chain: one FMA behind the reciprocal
            tm[j2][jj] = d[j2][jj] * rc;
#pragma unroll
            for (int i = j2 + 1; i < 8; ++i) d[i][j2] = fma(-d[i][jj], tm[j2][jj], d[i][j2]);
          }
        }
        // every row, the 8 rows of the diagonal block included (their entries right of the diagonal are never stored):
        // u_j = a_j - sum_{k<j} u_k (u_jk / piv_k),  L_rj = u_j / sqrt(piv_j)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          double v = p[jj];
#pragma unroll
          for (int k = 0; k < jj; ++k) v = fma(-p[k], tm[jj][k], v);
          p[jj] = v;
        }
        const int within = r - j0;  // >= 8 below the diagonal block
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k <= within) S[(j0 + k) * LD + r] = p[k] * dv[k];
        if (r == j0) {
#pragma unroll
          for (int k = 0; k < 8; ++k) dinv[j0 + k] = dv[k];
          if (failj < 8) atomicMin(&fail_col, j0 + failj);
        }
      }
    } else if (j0 > 0) {
      w_block_row(S, dinv, bj - 1, bj - 1 < 8 ? 0 : 8, warp - 4, 4, lane, 2);
    }
    ACC_STOP(0, 1);
    ACC_STOP(128, 3);
    ACC_START(0);
    __syncthreads();
    ACC_STOP(0, 2);
    if (j0 == 8) PT(3);
  }
  PT(4);
  w_block_row(S, dinv, 15, 8, warp, 8, lane, 0);  // ends the inversion of the lower-right 64x64 half
  w_lower_left(S, dinv, warp, lane);

  if (tid < TILE) {
    double v = log(S[tid * LD + tid]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
  }
  __syncthreads();
  PT(5);
  if (tid == 0) {
    logdet[b] += 2.0 * (((red[0] + red[1]) + red[2]) + red[3]);
    if (fail_col != 0x7fffffff && info[b] == 0) info[b] = J * TILE + fail_col + 1;
  }
  PT(6);
  pdl_trigger();  // the column's TRSM may be scheduled while L and W are stored
  PT(7);
#pragma unroll 8
  for (int e2 = tid; e2 < TT / 2; e2 += 256) {
    int r, c;
    tile_rc(2 * e2, r, c);
    double2 l, w;
    l.x = (r >= c) ? S[c * LD + r] : 0.0;
    l.y = (r >= c + 1) ? S[(c + 1) * LD + r] : 0.0;
    const double2 wt = *reinterpret_cast<const double2*>(&S[r * LD + c]);  // 16-byte aligned: LD and c are even
    w.x = (r > c) ? wt.x : (r == c ? dinv[r] : 0.0);
    w.y = (r > c + 1) ? wt.y : (r == c + 1 ? dinv[r] : 0.0);
    reinterpret_cast<double2*>(tile)[e2] = l;
    reinterpret_cast<double2*>(Wt)[e2] = w;
  }
  PT(8);
  PT(9);
}

static int g_potrf_impl = 1;  // 0: first-generation kernel (right-looking, W after L); 1: left-looking, W rows overlapped
void set_potrf_impl(int v) { g_potrf_impl = v; }

cudaError_t launch_potrf_tile(cudaStream_t st, TiledSym L, double* W, size_t w_batch_stride, int J, int batch, double* logdet,
                              int* info) {
  static bool configured_dev[64] = {false};  // function attributes are per device
  int dev = 0;
  cudaGetDevice(&dev);
  bool& configured = configured_dev[dev & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(potrf_tile_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (g_potrf_impl == 1)
    return launch_pdl(pdl_enabled(), potrf_tile_kernel2, dim3(batch), dim3(256), POTRF_SMEM, st, L, W, w_batch_stride, J, logdet, info);
  else
    potrf_tile_kernel<<<batch, 256, POTRF_SMEM, st>>>(L, W, w_batch_stride, J, logdet, info);
  return cudaGetLastError();
}

}  // namespace lmm
