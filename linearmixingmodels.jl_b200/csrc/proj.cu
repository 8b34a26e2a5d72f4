// Projection of the observations onto the latents + regulariser residual (K2/K3), and the
// back-projection of latent marginals to the outputs (K11).
// Reference: src/oilmm.jl:85 `Ty = T*Y`, :111-112 `sum(abs2, (I - U*U') * Y)`, :69-75
// `M = U*sqrt(S)*M_latent; V = abs2.(U*sqrt(S))*V_latent .+ σ²; vec(M')`; general ILMM:
// src/ilmm.jl:157-158, :179-180 (`Y .- H*T*Y`).
// Y is never transposed or copied: the by-outputs vector y IS the N x p column-major matrix, so
// row j of Y is the contiguous run y[j*N ...] and every global access below is coalesced along n.
// The projection is a skinny GEMM (m x p times p x N); at p = m = 64 its arithmetic intensity is
// 4pmN / (8(p+m)N) = 16 flop/B, above the B200 FP64 balance (37 TFLOP/s / 6.5 TB/s = 5.7 flop/B), so
// it is FP64-pipe bound, not HBM bound: the kernel is a register-tiled (4 x NT per thread)
// shared-memory GEMM -- Y block staged once in smem and reused by the three products T*Y, P*Y and
// Q*(P*Y); the residual never forms (I - UU') (O(pmN), not O(p²N)).
#include "common.cuh"
#include "kernels.h"

namespace lmm {

// acc[4][NT] += A[row0 + 0..3, 0:K] * Bs[0:K, n0 + 0..NT-1];  A is column-major with leading
// dimension lda (rows beyond `rows` read as zero); Bs is shared, [K][NB].
template <int NT, int NB>
__device__ __forceinline__ void tile_gemm(const double* __restrict__ A, int lda, int rows, int row0, const double* Bs, int K, int n0,
                                          double (&acc)[4][NT]) {
  const bool full = row0 + 3 < rows && (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  for (int k = 0; k < K; ++k) {
    double a[4];
    if (full) {
      const double2 v0 = __ldg(reinterpret_cast<const double2*>(A + (size_t)k * lda + row0));
      const double2 v1 = __ldg(reinterpret_cast<const double2*>(A + (size_t)k * lda + row0 + 2));
      a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (row0 + i < rows) ? __ldg(A + (size_t)k * lda + row0 + i) : 0.0;
    }
    double b[NT];
#pragma unroll
    for (int q = 0; q < NT; ++q) b[q] = Bs[k * NB + n0 + q];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[i][q] = fma(a[i], b[q], acc[i][q]);
  }
}

// One CTA = NB = 16*NT columns of Y.  256 threads = 16 (row groups of 4) x 16 (column groups of NT).
template <int NT>
__global__ void __launch_bounds__(256) project_kernel(const double* __restrict__ y, int N, int p, const double* __restrict__ T, int m,
                                                      int lat0, int mloc, const double* __restrict__ means, double* __restrict__ ty,
                                                      size_t ty_stride, const double* __restrict__ P, const double* __restrict__ Q,
                                                      double* __restrict__ resid_partial, double* __restrict__ resid_out, double* __restrict__ z_out) {
  constexpr int NB = 16 * NT;
  extern __shared__ __align__(16) double sm[];
  double* Ys = sm;            // [p][NB]
  double* Zs = sm + p * NB;   // [m][NB]
  __shared__ double red[256];
  const int t = threadIdx.x, tr = t >> 4, tc = t & 15;
  const int nb0 = blockIdx.x * NB, n0 = tc * NT;

  for (int idx = t; idx < p * NB; idx += 256) {
    const int j = idx / NB, nn = idx % NB;
    Ys[idx] = (nb0 + nn < N) ? y[(size_t)j * N + nb0 + nn] : 0.0;
  }
  __syncthreads();
  // Ty rows owned by this rank: rows lat0 .. lat0+mloc of T (m x p, column-major)
  for (int r0 = tr * 4; r0 < mloc; r0 += 64) {
    double acc[4][NT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[i][q] = 0.0;
    tile_gemm<NT, NB>(T + lat0, m, mloc, r0, Ys, p, n0, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + i < mloc) {
        const double mu = means[r0 + i];
#pragma unroll
        for (int q = 0; q < NT; ++q)
          if (nb0 + n0 + q < N) ty[(size_t)(r0 + i) * ty_stride + nb0 + n0 + q] = acc[i][q] - mu;
      }
  }
  if (P == nullptr) return;
  // Z = P Y (m x NB) into shared memory
  for (int r0 = tr * 4; r0 < m; r0 += 64) {
    double acc[4][NT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[i][q] = 0.0;
    tile_gemm<NT, NB>(P, m, m, r0, Ys, p, n0, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + i < m)
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          Zs[(r0 + i) * NB + n0 + q] = acc[i][q];
          if (z_out && nb0 + n0 + q < N) z_out[(size_t)(r0 + i) * N + nb0 + n0 + q] = acc[i][q];
        }
  }
  __syncthreads();
  // R = Y - Q Z (p x NB), summed squares
  double ss = 0.0;
  for (int r0 = tr * 4; r0 < p; r0 += 64) {
    double acc[4][NT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[i][q] = 0.0;
    tile_gemm<NT, NB>(Q, p, p, r0, Zs, m, n0, acc);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (r0 + i < p)
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          const double r = Ys[(r0 + i) * NB + n0 + q] - acc[i][q];
          ss = fma(r, r, ss);
          if (resid_out && nb0 + n0 + q < N) resid_out[(size_t)(r0 + i) * N + nb0 + n0 + q] = r;
        }
  }
  red[t] = ss;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) red[t] += red[t + w];
    __syncthreads();
  }
  if (t == 0) resid_partial[blockIdx.x] = red[0];
}

int project_block_cols(int p, int m) {
  // largest column block whose Y and Z tiles fit in shared memory (<= 200 KB)
  for (int nt : {4, 2, 1})
    if ((size_t)(p + m) * 16 * nt * sizeof(double) <= 200 * 1024) return 16 * nt;
  return 0;
}

cudaError_t launch_project(cudaStream_t st, const double* y, int N, int p, const double* T, int m, int lat0, int mloc,
                           const double* means, double* ty, size_t ty_stride, const double* P, const double* Q,
                           double* resid_partial, int* nblocks_out, double* resid_out, double* z_out) {
  const int nbcols = project_block_cols(p, m);
  if (nbcols == 0) return cudaErrorInvalidValue;
  const int nblocks = (N + nbcols - 1) / nbcols;
  if (nblocks_out) *nblocks_out = nblocks;
  const size_t smem = (size_t)(p + m) * nbcols * sizeof(double);
  cudaError_t e = cudaSuccess;
#define LMM_LAUNCH_PROJECT(NT_)                                                                                          \
  do {                                                                                                                    \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(project_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                       \
    project_kernel<NT_><<<nblocks, 256, smem, st>>>(y, N, p, T, m, lat0, mloc, means, ty, ty_stride, P, Q, resid_partial, resid_out, z_out); \
  } while (0)
  if (nbcols == 64) LMM_LAUNCH_PROJECT(4);
  else if (nbcols == 32) LMM_LAUNCH_PROJECT(2);
  else LMM_LAUNCH_PROJECT(1);
#undef LMM_LAUNCH_PROJECT
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}
cudaError_t launch_sum_partials(cudaStream_t st, const double* partial, int n, double* out) {
  sum_partials_kernel<<<1, 256, 0, st>>>(partial, n, out);
  return cudaGetLastError();
}

// grid (ceil(Ns/256), p): one output row j per blockIdx.y, coalesced along n.
__global__ void __launch_bounds__(256) backproject_kernel(const double* __restrict__ H, int p, int m, int lat0, int mloc,
                                                          const double* __restrict__ ML, const double* __restrict__ VL,
                                                          size_t lat_stride, int Ns, double jitter, double sigma2, int add_noise,
                                                          double* __restrict__ mean, double* __restrict__ var) {
  const int n = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if (n >= Ns) return;
  double sm_ = 0.0, sv = 0.0;
  for (int i = 0; i < mloc; ++i) {
    const double h = H[(size_t)(lat0 + i) * p + j];
    sm_ = fma(h, ML[(size_t)i * lat_stride + n], sm_);
    sv = fma(h * h, VL[(size_t)i * lat_stride + n] + jitter, sv);
  }
  mean[(size_t)j * Ns + n] = sm_;
  var[(size_t)j * Ns + n] = add_noise ? sv + sigma2 : sv;
}
cudaError_t launch_backproject(cudaStream_t st, const double* H, int p, int m, int lat0, int mloc, const double* ML,
                               const double* VL, size_t lat_stride, int Ns, double jitter, double sigma2, int add_noise,
                               double* mean, double* var) {
  dim3 grid((unsigned)((Ns + 255) / 256), (unsigned)p);
  backproject_kernel<<<grid, 256, 0, st>>>(H, p, m, lat0, mloc, ML, VL, lat_stride, Ns, jitter, sigma2, add_noise, mean, var);
  return cudaGetLastError();
}

// out[i + j*ra] (+)= scale * sum_n A[i*lda + n] * B[j*ldb + n]   (A: ra x N, B: rb x N, rows contiguous)
// grid (ra, rb): one CTA per output entry, fixed-order block reduction.
__global__ void __launch_bounds__(256) abt_kernel(const double* __restrict__ A, size_t lda, const double* __restrict__ B, size_t ldb, int N,
                                                  double scale, double* __restrict__ out, int ra) {
  __shared__ double red[256];
  const int i = blockIdx.x, j = blockIdx.y, t = threadIdx.x;
  const double* a = A + (size_t)i * lda;
  const double* b = B + (size_t)j * ldb;
  double s = 0.0;
  for (int n = t; n < N; n += 256) s = fma(a[n], b[n], s);
  red[t] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) red[t] += red[t + w];
    __syncthreads();
  }
  if (t == 0) out[(size_t)j * ra + i] += scale * red[0];
}
cudaError_t launch_abt(cudaStream_t st, const double* A, size_t lda, int ra, const double* B, size_t ldb, int rb, int N, double scale,
                       double* out) {
  if (ra <= 0 || rb <= 0) return cudaSuccess;
  dim3 grid((unsigned)ra, (unsigned)rb);
  abt_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, N, scale, out, ra);
  return cudaGetLastError();
}

}  // namespace lmm
