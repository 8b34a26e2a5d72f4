// Triangular vector solves, quadratic form and prediction reductions on the tiled factor
// (K8/K9/K10 of SURVEY.md §2.3; reference: AbstractGPs `C.U' \ δ`, `C \ δ`, posterior mean / var
// reached from src/oilmm.jl:90,128,61).  All of these read each 128 KB tile exactly once:
// HBM-bound GEMV-shaped work.  With W(J) = inv(L_JJ) from the panel kernel a solve is a chain of
// nt column steps:  z_J = W_J r_J ;  r_I -= L(I,J) z_J  (I > J)   -- one launch per step, all
// tiles of the column (and all latents) in parallel.
#include "common.cuh"
#include "kernels.h"

namespace lmm {

// Per-thread partial products of y = T x over one tile (interleaved layout), 256 threads.
// Thread t, iteration it touches element e = it*256 + t:  row = ((it&1)*8 + warp)*8 + (t>>2)&7,
// col = (it>>1)*4 + (t&3).  acc0 collects even it (row rlo), acc1 odd it (row rlo + 64).
__device__ __forceinline__ void tile_gemv_n_acc(const double* __restrict__ tile, const double* xs, double& acc0, double& acc1) {
  const int t = threadIdx.x;
#pragma unroll 8
  for (int it = 0; it < 64; it += 2) {
    const double x = xs[(it >> 1) * 4 + (t & 3)];
    acc0 = fma(tile[it * 256 + t], x, acc0);
    acc1 = fma(tile[(it + 1) * 256 + t], x, acc1);
  }
}
// Finish: quad-reduce and scatter to ys[128] (shared).  Caller syncs afterwards.
__device__ __forceinline__ void tile_gemv_n_finish(double acc0, double acc1, double* ys) {
  const int t = threadIdx.x;
  acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1);
  acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
  acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
  acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
  if ((t & 3) == 0) {
    const int rlo = (t >> 5) * 8 + ((t >> 2) & 7);
    ys[rlo] = acc0;
    ys[rlo + 64] = acc1;
  }
}
// y = T^T x over one tile: ys[c] = sum_r T(r,c) xs[r].  part: 8*128 doubles of shared scratch.
// Ends with a __syncthreads(); ys valid for all threads afterwards.
__device__ __forceinline__ void tile_gemv_t(const double* __restrict__ tile, const double* xs, double* ys, double* part) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int rlo = warp * 8 + ((t >> 2) & 7);
  const double x0 = xs[rlo], x1 = xs[rlo + 64];
#pragma unroll 8
  for (int it = 0; it < 64; it += 2) {
    double v = tile[it * 256 + t] * x0;
    v = fma(tile[(it + 1) * 256 + t], x1, v);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    if (lane < 4) part[warp * TILE + (it >> 1) * 4 + lane] = v;
  }
  __syncthreads();
  if (t < TILE) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w * TILE + t];
    ys[t] = s;
  }
  __syncthreads();
}

// grid (nt - J, batch).  rvec: running right-hand side (destroyed), zvec: solution.
__global__ void __launch_bounds__(256) fwd_step_kernel(TiledSym L, const double* __restrict__ W, size_t w_batch_stride,
                                                       double* __restrict__ rvec, double* __restrict__ zvec, size_t vec_stride, int J) {
  __shared__ double xs[TILE], zs[TILE], ys[TILE];
  const int b = blockIdx.y, i = blockIdx.x, t = threadIdx.x;
  double* r = rvec + (size_t)b * vec_stride;
  if (t < TILE) xs[t] = r[J * TILE + t];
  __syncthreads();
  double a0 = 0.0, a1 = 0.0;
  tile_gemv_n_acc(W + (size_t)b * w_batch_stride + (size_t)J * TT, xs, a0, a1);
  tile_gemv_n_finish(a0, a1, zs);
  __syncthreads();
  if (i == 0) {
    if (t < TILE) zvec[(size_t)b * vec_stride + J * TILE + t] = zs[t];
    return;
  }
  const int I = J + i;
  a0 = a1 = 0.0;
  tile_gemv_n_acc(L.tile(b, I, J), zs, a0, a1);
  tile_gemv_n_finish(a0, a1, ys);
  __syncthreads();
  if (t < TILE) r[I * TILE + t] -= ys[t];
}

// grid (J + 1, batch).  rvec: running rhs (destroyed), avec: solution of L^T a = r.
__global__ void __launch_bounds__(256) bwd_step_kernel(TiledSym L, const double* __restrict__ W, size_t w_batch_stride,
                                                       double* __restrict__ rvec, double* __restrict__ avec, size_t vec_stride, int J) {
  __shared__ double xs[TILE], as[TILE], ys[TILE];
  __shared__ double part[8 * TILE];
  const int b = blockIdx.y, k = blockIdx.x, t = threadIdx.x;
  double* r = rvec + (size_t)b * vec_stride;
  if (t < TILE) xs[t] = r[J * TILE + t];
  __syncthreads();
  tile_gemv_t(W + (size_t)b * w_batch_stride + (size_t)J * TT, xs, as, part);
  if (k == J) {
    if (t < TILE) avec[(size_t)b * vec_stride + J * TILE + t] = as[t];
    return;
  }
  tile_gemv_t(L.tile(b, J, k), as, ys, part);
  if (t < TILE) r[k * TILE + t] -= ys[t];
}

cudaError_t launch_fwd_solve(cudaStream_t st, TiledSym L, const double* W, size_t w_batch_stride, double* rvec, double* zvec,
                             size_t vec_stride, int batch, int64_t* launches) {
  for (int J = 0; J < L.nt; ++J) {
    dim3 grid((unsigned)(L.nt - J), (unsigned)batch);
    fwd_step_kernel<<<grid, 256, 0, st>>>(L, W, w_batch_stride, rvec, zvec, vec_stride, J);
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

cudaError_t launch_bwd_solve(cudaStream_t st, TiledSym L, const double* W, size_t w_batch_stride, double* rvec, double* avec,
                             size_t vec_stride, int batch, int64_t* launches) {
  for (int J = L.nt - 1; J >= 0; --J) {
    dim3 grid((unsigned)(J + 1), (unsigned)batch);
    bwd_step_kernel<<<grid, 256, 0, st>>>(L, W, w_batch_stride, rvec, avec, vec_stride, J);
    if (launches) ++*launches;
  }
  return cudaGetLastError();
}

// out[b] = sum v^2, one CTA per latent, fixed summation order.
__global__ void __launch_bounds__(256) sumsq_kernel(const double* __restrict__ v, size_t stride, int n, double* __restrict__ out) {
  __shared__ double red[256];
  const double* p = v + (size_t)blockIdx.x * stride;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s = fma(p[i], p[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}
cudaError_t launch_sumsq(cudaStream_t st, const double* v, size_t stride, int n, int batch, double* out) {
  sumsq_kernel<<<batch, 256, 0, st>>>(v, stride, n, out);
  return cudaGetLastError();
}

// grid (ntr, batch): y[b][R*128 + r] = (add_mean ? mean_b : 0) + sum_J T(R,J) x[b][J*128 + :]
__global__ void __launch_bounds__(256) rect_gemv_kernel(TiledRect A, const double* __restrict__ x, size_t x_stride,
                                                        double* __restrict__ y, size_t y_stride,
                                                        const LatentParams* __restrict__ params, int add_mean) {
  __shared__ double xs[2][TILE], ys[TILE];
  const int b = blockIdx.y, R = blockIdx.x, t = threadIdx.x;
  const double* xb = x + (size_t)b * x_stride;
  double a0 = 0.0, a1 = 0.0;
  for (int J = 0; J < A.ntc; ++J) {
    if (t < TILE) xs[J & 1][t] = xb[J * TILE + t];
    __syncthreads();
    tile_gemv_n_acc(A.tile(b, R, J), xs[J & 1], a0, a1);
  }
  tile_gemv_n_finish(a0, a1, ys);
  __syncthreads();
  if (t < TILE) y[(size_t)b * y_stride + R * TILE + t] = ys[t] + (add_mean ? params[b].mean : 0.0);
}
cudaError_t launch_rect_gemv(cudaStream_t st, TiledRect A, const double* x, size_t x_stride, double* y, size_t y_stride,
                             const LatentParams* params, int add_mean, int batch) {
  dim3 grid((unsigned)A.ntr, (unsigned)batch);
  rect_gemv_kernel<<<grid, 256, 0, st>>>(A, x, x_stride, y, y_stride, params, add_mean);
  return cudaGetLastError();
}

// grid (ntr, batch): y[b][R*128 + r] = variance_b - sum_J sum_c T(R,J)(r,c)^2
__global__ void __launch_bounds__(256) rect_rowsumsq_kernel(TiledRect A, double* __restrict__ y, size_t y_stride,
                                                            const LatentParams* __restrict__ params) {
  __shared__ double ys[TILE];
  const int b = blockIdx.y, R = blockIdx.x, t = threadIdx.x;
  double a0 = 0.0, a1 = 0.0;
  for (int J = 0; J < A.ntc; ++J) {
    const double* tile = A.tile(b, R, J);
#pragma unroll 8
    for (int it = 0; it < 64; it += 2) {
      const double v0 = tile[it * 256 + t], v1 = tile[(it + 1) * 256 + t];
      a0 = fma(v0, v0, a0);
      a1 = fma(v1, v1, a1);
    }
  }
  tile_gemv_n_finish(a0, a1, ys);
  __syncthreads();
  if (t < TILE) y[(size_t)b * y_stride + R * TILE + t] = params[b].variance - ys[t];
}
cudaError_t launch_rect_rowsumsq(cudaStream_t st, TiledRect A, double* y, size_t y_stride, const LatentParams* params, int batch) {
  dim3 grid((unsigned)A.ntr, (unsigned)batch);
  rect_rowsumsq_kernel<<<grid, 256, 0, st>>>(A, y, y_stride, params);
  return cudaGetLastError();
}

// grid (nt, batch): y_I = sum_{J <= I} L(I,J) z_J   (diagonal tiles have a zero upper part)
__global__ void __launch_bounds__(256) lower_gemv_kernel(TiledSym L, const double* __restrict__ z, size_t z_stride,
                                                         double* __restrict__ y, size_t y_stride) {
  __shared__ double xs[2][TILE], ys[TILE];
  const int b = blockIdx.y, I = blockIdx.x, t = threadIdx.x;
  const double* zb = z + (size_t)b * z_stride;
  double a0 = 0.0, a1 = 0.0;
  for (int J = 0; J <= I; ++J) {
    if (t < TILE) xs[J & 1][t] = zb[J * TILE + t];
    __syncthreads();
    tile_gemv_n_acc(L.tile(b, I, J), xs[J & 1], a0, a1);
  }
  tile_gemv_n_finish(a0, a1, ys);
  __syncthreads();
  if (t < TILE) y[(size_t)b * y_stride + I * TILE + t] = ys[t];
}
cudaError_t launch_lower_gemv(cudaStream_t st, TiledSym L, const double* z, size_t z_stride, double* y, size_t y_stride, int batch) {
  dim3 grid((unsigned)L.nt, (unsigned)batch);
  lower_gemv_kernel<<<grid, 256, 0, st>>>(L, z, z_stride, y, y_stride);
  return cudaGetLastError();
}

// grid (lower tiles): dense[c*N + r] = L(r, c) for r >= c (dense pre-zeroed by the caller)
__global__ void __launch_bounds__(256) untile_lower_kernel(TiledSym L, int b, double* __restrict__ dense, int N) {
  const int tl = blockIdx.x;
  int I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
  while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
  while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
  const int J = tl - (int)((size_t)I * (I + 1) / 2);
  const double* tile = L.tile(b, I, J);
  for (int idx = threadIdx.x; idx < TT; idx += 256) {
    const int r = idx & 127, c = idx >> 7;  // coalesced dense writes
    const int gr = I * TILE + r, gc = J * TILE + c;
    if (gr < N && gc < N && gr >= gc) dense[(size_t)gc * N + gr] = tile[tile_elem(r, c)];
  }
}
cudaError_t launch_untile_lower(cudaStream_t st, TiledSym L, int b, double* dense, int N) {
  untile_lower_kernel<<<(unsigned)sym_tiles(L.nt), 256, 0, st>>>(L, b, dense, N);
  return cudaGetLastError();
}

// grid (lower tiles, batch): tiles <- dense symmetric (lower triangle read), identity padding
__global__ void __launch_bounds__(256) tile_from_dense_kernel(TiledSym L, const double* __restrict__ dense, int N) {
  const int tl = blockIdx.x, b = blockIdx.y;
  int I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
  while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
  while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
  const int J = tl - (int)((size_t)I * (I + 1) / 2);
  double* tile = L.tile(b, I, J);
  const double* A = dense + (size_t)b * N * N;
  for (int idx = threadIdx.x; idx < TT; idx += 256) {
    const int r = idx & 127, c = idx >> 7;
    const int gr = I * TILE + r, gc = J * TILE + c;
    double v;
    if (gr >= N || gc >= N) v = (gr == gc) ? 1.0 : 0.0;
    else v = (gr >= gc) ? A[(size_t)gc * N + gr] : A[(size_t)gr * N + gc];
    tile[tile_elem(r, c)] = v;
  }
}
cudaError_t launch_tile_from_dense(cudaStream_t st, TiledSym L, int batch, const double* dense, int N) {
  dim3 grid((unsigned)sym_tiles(L.nt), (unsigned)batch);
  tile_from_dense_kernel<<<grid, 256, 0, st>>>(L, dense, N);
  return cudaGetLastError();
}

// grid (ntr*ntc, batch): dense block-diagonal scatter of rectangular tiled matrices:
// dense[(b*Na + r) + (b*Nb + c) * ld] = A_b(r, c)   for r < Na, c < Nb   (dense pre-zeroed)
__global__ void __launch_bounds__(256) untile_rect_blockdiag_kernel(TiledRect A, int Na, int Nb, double* __restrict__ dense, size_t ld) {
  const int b = blockIdx.y, R = blockIdx.x / A.ntc, J = blockIdx.x % A.ntc;
  const double* tile = A.tile(b, R, J);
  for (int idx = threadIdx.x; idx < TT; idx += 256) {
    const int r = idx & 127, c = idx >> 7;
    const int gr = R * TILE + r, gc = J * TILE + c;
    if (gr < Na && gc < Nb) dense[((size_t)b * Nb + gc) * ld + (size_t)b * Na + gr] = tile[tile_elem(r, c)];
  }
}
cudaError_t launch_untile_rect_blockdiag(cudaStream_t st, TiledRect A, int batch, int Na, int Nb, double* dense, size_t ld) {
  dim3 grid((unsigned)(A.ntr * A.ntc), (unsigned)batch);
  untile_rect_blockdiag_kernel<<<grid, 256, 0, st>>>(A, Na, Nb, dense, ld);
  return cudaGetLastError();
}


// ---- row-cyclic block-column exchange (multi-GPU Cholesky of one large factor) ---------------------
// Tile rows [ra, rb) of block column [s0, s1); rank r owns the rows I = r (mod G).  Slot q of rank r is its q-th
// own row >= ra.  grid (slots, ob): one CTA copies one tile (16384 doubles, 128-bit accesses).
__global__ void __launch_bounds__(256) rowcyclic_copy_kernel(TiledSym L, int s0, int s1, int ra, int rb, int G, int rank_lo, int rank_hi,
                                                             int skip_rank, int slots, double* __restrict__ buf, int to_buf) {
  const int ob = s1 - s0, q = blockIdx.x, c = blockIdx.y;
  for (int r = rank_lo; r < rank_hi; ++r) {
    if (r == skip_rank) continue;
    const int first = ra + ((r - ra % G) % G + G) % G;  // first tile row >= ra owned by rank r
    const int I = first + q * G;
    if (I >= rb) continue;
    const int J = s0 + c;
    if (J > I) continue;  // above the diagonal inside the block
    double2* t = reinterpret_cast<double2*>(L.tile(0, I, J));
    double2* bslot = reinterpret_cast<double2*>(buf + (((size_t)(to_buf ? 0 : r) * slots + q) * ob + c) * TT);
    for (int e = threadIdx.x; e < TT / 2; e += 256) {
      if (to_buf) bslot[e] = t[e];
      else t[e] = bslot[e];
    }
  }
}
cudaError_t launch_rowcyclic_pack(cudaStream_t st, TiledSym L, int s0, int s1, int ra, int rb, int G, int rank, int slots, double* buf) {
  dim3 grid((unsigned)slots, (unsigned)(s1 - s0));
  rowcyclic_copy_kernel<<<grid, 256, 0, st>>>(L, s0, s1, ra, rb, G, rank, rank + 1, -1, slots, buf, 1);
  return cudaGetLastError();
}
cudaError_t launch_rowcyclic_unpack(cudaStream_t st, TiledSym L, int s0, int s1, int ra, int rb, int G, int rank, int slots,
                                    const double* all) {
  dim3 grid((unsigned)slots, (unsigned)(s1 - s0));
  rowcyclic_copy_kernel<<<grid, 256, 0, st>>>(L, s0, s1, ra, rb, G, 0, G, rank, slots, const_cast<double*>(all), 0);
  return cudaGetLastError();
}

}  // namespace lmm
