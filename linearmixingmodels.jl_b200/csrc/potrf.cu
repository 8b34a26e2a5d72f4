// Diagonal-tile factorisation: the panel step (dpotf2) of the blocked Cholesky (K6), fused with
// logdet accumulation (K7), non-PD detection and the triangular inverse W = inv(L_JJ) that turns
// every panel TRSM and every triangular vector solve into tensor-core GEMMs / plain GEMVs.
// One CTA per latent; the 128x128 tile lives in shared memory (column-major, ld = 129).
// Latency-bound by design (it sits on the critical path of each tile column; batching over
// latents keeps the SMs busy meanwhile).
#include "common.cuh"
#include "kernels.h"

namespace lmm {

constexpr int LD = TILE + 1;
constexpr size_t POTRF_SMEM = (size_t)(TILE * LD + 2 * TILE + 32) * sizeof(double);

__global__ void __launch_bounds__(256, 1) potrf_tile_kernel(TiledSym L, double* __restrict__ Wbase, size_t w_batch_stride, int J,
                                                            double* __restrict__ logdet, int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];  // S[c*LD + r]
  double* dinv = S + TILE * LD;               // 1 / L[r][r]
  double* red = dinv + TILE;                  // reduction scratch (TILE doubles)
  __shared__ int fail_col;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  double* tile = L.tile(b, J, J);
  double* Wt = Wbase + (size_t)b * w_batch_stride + (size_t)J * TT;

  if (tid == 0) fail_col = 0x7fffffff;
  for (int e = tid; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    S[c * LD + r] = tile[e];
  }
  __syncthreads();

  // ---- blocked right-looking Cholesky, 8 columns per step
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
      // factor the 8x8 diagonal block in registers (redundantly per thread)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        double piv = d[jj][jj];
#pragma unroll
        for (int k = 0; k < jj; ++k) piv = fma(-d[jj][k], d[jj][k], piv);
        if (!(piv > 0.0)) {
          if (r == j0) atomicMin(&fail_col, j0 + jj);
        }
        const double l = sqrt(piv);
        d[jj][jj] = l;
        const double inv = 1.0 / l;
        dv[jj] = inv;
#pragma unroll
        for (int i = jj + 1; i < 8; ++i) {
          double v = d[i][jj];
#pragma unroll
          for (int k = 0; k < jj; ++k) v = fma(-d[i][k], d[jj][k], v);
          d[i][jj] = v * inv;
        }
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
        dinv[r] = 0.0;
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
          if (ii == i) dinv[r] = dv[ii];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) S[(j0 + k) * LD + r] = p[k];
    }
    __syncthreads();
    // trailing update of rows/cols [j0+8, 128): 4x4 strided register blocks, lower part only
    const int n = TILE - j0 - 8;
    if (n > 0) {
      const int nq = n >> 2;  // n is a multiple of 8
      const int base = j0 + 8;
      for (int task = tid; task < nq * nq; task += 256) {
        const int tr = task % nq, tc = task / nq;
        double pr[4][8], pc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            pr[i][k] = S[(j0 + k) * LD + base + tr + i * nq];
            pc[i][k] = S[(j0 + k) * LD + base + tc + i * nq];
          }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jx = 0; jx <= i; ++jx) {
            if (i == jx && tr < tc) continue;
            const int rr = base + tr + i * nq, cc = base + tc + jx * nq;
            double v = S[cc * LD + rr];
#pragma unroll
            for (int k = 0; k < 8; ++k) v = fma(-pr[i][k], pc[jx][k], v);
            S[cc * LD + rr] = v;
          }
      }
    }
    __syncthreads();
  }

  // ---- logdet and failure report
  if (tid < TILE) red[tid] = log(S[tid * LD + tid]);
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int i = 0; i < TILE; ++i) s += red[i];
    logdet[b] += 2.0 * s;
    if (fail_col != 0x7fffffff && info[b] == 0) info[b] = J * TILE + fail_col + 1;
  }

  // ---- write L (lower, zero upper)
  for (int e = tid; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    tile[e] = (r >= c) ? S[c * LD + r] : 0.0;
  }
  __syncthreads();

  // ---- W = inv(L): thread c owns column c; W[r][c] (r > c) is kept at S[r*LD + c] (the unused
  // upper triangle, transposed); diagonal in dinv.  Rows are processed 4 at a time: the sums over
  // k below the row block are 4 independent FMA chains, then a 4x4 triangular finish.
  if (tid < TILE) {
    const int c = tid;
    const int c0 = c & ~31;  // warp-uniform start so that reads of L broadcast
    for (int rb = (c0 & ~3); rb < TILE; rb += 4) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int k = c0; k < rb; ++k) {
        const double w = (k > c) ? S[k * LD + c] : ((k == c) ? dinv[c] : 0.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = fma(S[k * LD + rb + i], w, s[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = rb + i;
        // contributions from rows inside the block (k in [rb, rr))
        double w_rr;
#pragma unroll
        for (int kk = 0; kk < i; ++kk) {
          const int k = rb + kk;
          const double w = (k > c) ? S[k * LD + c] : ((k == c) ? dinv[c] : 0.0);
          s[i] = fma(S[k * LD + rr], w, s[i]);
        }
        w_rr = -s[i] * dinv[rr];
        if (rr > c) S[rr * LD + c] = w_rr;
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    Wt[e] = (r > c) ? S[r * LD + c] : ((r == c) ? dinv[r] : 0.0);
  }
}

cudaError_t launch_potrf_tile(cudaStream_t st, TiledSym L, double* W, size_t w_batch_stride, int J, int batch, double* logdet,
                              int* info) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  potrf_tile_kernel<<<batch, 256, POTRF_SMEM, st>>>(L, W, w_batch_stride, J, logdet, info);
  return cudaGetLastError();
}

}  // namespace lmm
