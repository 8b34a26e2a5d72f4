// Projection of the observations onto the latents + regulariser residual (K2/K3), and the
// back-projection of latent marginals to the outputs (K11).
// Reference: src/oilmm.jl:85 `Ty = T*Y`, :111-112 `sum(abs2, (I - U*U') * Y)`, :69-75
// `M = U*sqrt(S)*M_latent; V = abs2.(U*sqrt(S))*V_latent .+ σ²; vec(M')`; general ILMM:
// src/ilmm.jl:157-158, :179-180 (`Y .- H*T*Y`).
// Y is never transposed or copied: the by-outputs vector y IS the N x p column-major matrix, so
// row j of Y is the contiguous run y[j*N ...] and every load below is coalesced along n.
// HBM-bound: 8pN bytes read + 8mN written; the residual never forms (I - UU') (O(pmN), not O(p²N)).
#include "common.cuh"
#include "kernels.h"

namespace lmm {

constexpr int PCOLS = 32;  // columns of Y per CTA

__global__ void __launch_bounds__(256) project_kernel(const double* __restrict__ y, int N, int p, const double* __restrict__ T,
                                                      int m, int lat0, int mloc, const double* __restrict__ means,
                                                      double* __restrict__ ty, size_t ty_stride, const double* __restrict__ P,
                                                      const double* __restrict__ Q, double* __restrict__ resid_partial) {
  extern __shared__ __align__(16) double sm[];
  double* Ys = sm;               // [p][PCOLS]
  double* Zs = sm + p * PCOLS;   // [m][PCOLS]
  __shared__ double red[256];
  const int n0 = blockIdx.x * PCOLS, t = threadIdx.x;

  for (int idx = t; idx < p * PCOLS; idx += 256) {
    const int j = idx / PCOLS, nn = idx % PCOLS;
    Ys[idx] = (n0 + nn < N) ? y[(size_t)j * N + n0 + nn] : 0.0;
  }
  __syncthreads();
  // projection rows owned by this rank
  for (int idx = t; idx < mloc * PCOLS; idx += 256) {
    const int i = idx / PCOLS, nn = idx % PCOLS;
    if (n0 + nn < N) {
      double s = 0.0;
      for (int j = 0; j < p; ++j) s = fma(T[(size_t)j * m + lat0 + i], Ys[j * PCOLS + nn], s);
      ty[(size_t)i * ty_stride + n0 + nn] = s - means[i];
    }
  }
  if (P == nullptr) return;
  // residual |Y - Q (P Y)|^2 over this block's columns
  for (int idx = t; idx < m * PCOLS; idx += 256) {
    const int i = idx / PCOLS, nn = idx % PCOLS;
    double s = 0.0;
    for (int j = 0; j < p; ++j) s = fma(P[(size_t)j * m + i], Ys[j * PCOLS + nn], s);
    Zs[idx] = s;
  }
  __syncthreads();
  double acc = 0.0;
  for (int idx = t; idx < p * PCOLS; idx += 256) {
    const int j = idx / PCOLS, nn = idx % PCOLS;
    double s = 0.0;
    for (int i = 0; i < m; ++i) s = fma(Q[(size_t)i * p + j], Zs[i * PCOLS + nn], s);
    const double r = Ys[idx] - s;
    acc = fma(r, r, acc);
  }
  red[t] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) red[t] += red[t + w];
    __syncthreads();
  }
  if (t == 0) resid_partial[blockIdx.x] = red[0];
}

cudaError_t launch_project(cudaStream_t st, const double* y, int N, int p, const double* T, int m, int lat0, int mloc,
                           const double* means, double* ty, size_t ty_stride, const double* P, const double* Q,
                           double* resid_partial, int* nblocks_out) {
  const int nblocks = (N + PCOLS - 1) / PCOLS;
  if (nblocks_out) *nblocks_out = nblocks;
  const size_t smem = (size_t)(p + m) * PCOLS * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  project_kernel<<<nblocks, 256, smem, st>>>(y, N, p, T, m, lat0, mloc, means, ty, ty_stride, P, Q, resid_partial);
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

}  // namespace lmm
