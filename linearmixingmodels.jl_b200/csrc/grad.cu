// Fused kernel-gradient reduction for the logpdf rrule (reference call sites
// test/oilmm.jl:31-32, test/ilmm.jl:31-32, test/independent_mogp.jl:65-66 `gradient(logpdf, fx, y)`).
// With G = d lml / dC = (αα' - C^{-1})/2 the hyper-parameter gradients of one latent are
//   d/d variance = <G, κ>,  d/d inv_lengthscale = <G, dK/ds>,  d/d noise = tr(G),
// contracted tile by tile: the kernel values and their lengthscale derivatives are recomputed in
// registers from the staged inputs (never stored), NegCinv = -C^{-1} comes from the batched
// `potri` (triangular TRSM sweep + SYRK on the DMMA kernel).  HBM-bound: one read of C^{-1}.
#include "common.cuh"
#include "kernels.h"

namespace lmm {

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// κ and dK/ds (both already multiplied by the variance where appropriate): returns κ (unscaled by variance)
__device__ __forceinline__ void kappa_and_ds(int kind, double d2, double rinv_ls, double& kap, double& dkds, double param = 1.0) {
  if (kind == 0) {
    kap = exp_nonpos(-0.5 * d2);
    dkds = -kap * d2 * rinv_ls;
  } else if (kind == 4) {  // (1 + d²/2α)^(-α);  d/ds = -(d²/s) (1 + d²/2α)^(-α-1)
    const double base = 1.0 + d2 / (2.0 * param);
    kap = pow(base, -param);
    dkds = -d2 * rinv_ls * kap / base;
  } else {
    const double d = sqrt(d2);
    if (kind == 1) {
      const double s = 1.7320508075688772 * d;
      const double e = exp_nonpos(-s);
      kap = (1.0 + s) * e;
      dkds = -3.0 * d2 * e * rinv_ls;
    } else if (kind == 2) {
      const double s = 2.23606797749979 * d;
      const double e = exp_nonpos(-s);
      kap = (1.0 + s + (d * d) * 1.6666666666666667) * e;
      dkds = -(5.0 / 3.0) * d2 * (1.0 + s) * e * rinv_ls;
    } else {  // exp(-d);  d/ds = -(d/s) exp(-d)
      kap = exp_nonpos(-d);
      dkds = -d * kap * rinv_ls;
    }
  }
}

// grid (lower tiles, batch).  partial[(b*ntiles + tile)*3 + {0,1,2}] = {<G,κ>, <G,dK/ds>/variance.., tr G}
__global__ void __launch_bounds__(256) kgrad_kernel(TiledSym negCinv, const double* __restrict__ xpad, int N, int D,
                                                    const LatentParams* __restrict__ params, const double* __restrict__ alpha,
                                                    size_t alpha_stride, int form, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* xa = sm;                 // [128*D] scaled
  double* xb = xa + TILE * D;
  double* sa = xb + TILE * D;
  double* sb = sa + TILE;
  double* aa = sb + TILE;          // alpha rows
  double* ab = aa + TILE;          // alpha cols
  __shared__ double red[3][256];

  const int b = blockIdx.y, tl = blockIdx.x, t = threadIdx.x;
  int I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
  while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
  while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
  const int J = tl - (int)((size_t)I * (I + 1) / 2);
  const LatentParams lp = params[b];
  const double rinv_ls = 1.0 / lp.inv_ls;
  for (int i = t; i < TILE * D; i += 256) {
    const double s = input_scale(params + b, i % D);
    xa[i] = xpad[(size_t)I * TILE * D + i] * s;
    xb[i] = xpad[(size_t)J * TILE * D + i] * s;
  }
  if (t < TILE) {
    aa[t] = alpha[(size_t)b * alpha_stride + I * TILE + t];
    ab[t] = alpha[(size_t)b * alpha_stride + J * TILE + t];
  }
  __syncthreads();
  for (int i = t; i < 2 * TILE; i += 256) {
    const double* v = (i < TILE) ? xa + (size_t)i * D : xb + (size_t)(i - TILE) * D;
    double s = 0.0;
    for (int k = 0; k < D; ++k) s = fma(v[k], v[k], s);
    if (i < TILE) sa[i] = s; else sb[i - TILE] = s;
  }
  __syncthreads();

  const double* tile = negCinv.tile(b, I, J);
  const int r = (((2 * t) >> 5) & 15) * 8 + (((2 * t) >> 2) & 7);
  const int gr = I * TILE + r;
  double gv = 0.0, gs = 0.0, gn = 0.0;
  for (int it = 0; it < 32; ++it) {
    const int c = 4 * it + ((2 * t) & 3);
    const double2 nc = *reinterpret_cast<const double2*>(tile + it * 512 + 2 * t);
    const double ncv[2] = {nc.x, nc.y};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int cc = c + q, gc = J * TILE + cc;
      if (gr >= N || gc >= N) continue;
      if (I == J && cc > r) continue;  // lower triangle only; off-diagonal entries count twice
      const double G = 0.5 * (aa[r] * ab[cc] + ncv[q]);
      if (gr == gc) {
        gn += G;
        gv = fma(G, 1.0, gv);  // κ(0) = 1, dK/ds = 0 on the diagonal
      } else {
        const double d2 = sqdist(xa + (size_t)r * D, xb + (size_t)cc * D, D, sa[r], sb[cc], form);
        double kap, dk;
        kappa_and_ds(lp.kind, d2, rinv_ls, kap, dk, lp.param);
        gv = fma(2.0 * G, kap, gv);
        gs = fma(2.0 * G, dk, gs);
      }
    }
  }
  red[0][t] = gv; red[1][t] = gs; red[2][t] = gn;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) {
      red[0][t] += red[0][t + w];
      red[1][t] += red[1][t + w];
      red[2][t] += red[2][t + w];
    }
    __syncthreads();
  }
  if (t == 0) {
    double* out = partial + ((size_t)b * gridDim.x + tl) * 3;
    out[0] = red[0][0];
    out[1] = red[1][0] * lp.variance;  // dK/ds carries the variance
    out[2] = red[2][0];
  }
}

cudaError_t launch_kgrad(cudaStream_t st, TiledSym negCinv, int batch, const double* xpad, int N, int D, const LatentParams* params,
                         const double* alpha, size_t alpha_stride, int form, double* partial) {
  const size_t smem = (size_t)(2 * TILE * D + 4 * TILE) * sizeof(double);
  if (smem + 1024 > 48 * 1024) {  // static shared memory counts against the default limit too
    cudaError_t e = cudaFuncSetAttribute(kgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)sym_tiles(negCinv.nt), (unsigned)batch);
  kgrad_kernel<<<grid, 256, smem, st>>>(negCinv, xpad, N, D, params, alpha, alpha_stride, form, partial);
  return cudaGetLastError();
}

// out[b*4 + {0,1,2}] = fixed-order sums of the tile partials; out[b*4+3] = sum(alpha_b) (d/d mean)
__global__ void __launch_bounds__(256) kgrad_finish_kernel(const double* __restrict__ partial, int ntiles_, const double* __restrict__ alpha,
                                                           size_t alpha_stride, int N, double* __restrict__ out) {
  __shared__ double red[4][256];
  const int b = blockIdx.x, t = threadIdx.x;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int i = t; i < ntiles_; i += 256) {
    const double* p = partial + ((size_t)b * ntiles_ + i) * 3;
    s0 += p[0]; s1 += p[1]; s2 += p[2];
  }
  for (int i = t; i < N; i += 256) s3 += alpha[(size_t)b * alpha_stride + i];
  red[0][t] = s0; red[1][t] = s1; red[2][t] = s2; red[3][t] = s3;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w)
      for (int k = 0; k < 4; ++k) red[k][t] += red[k][t + w];
    __syncthreads();
  }
  if (t < 4) out[(size_t)b * 4 + t] = red[t][0];
}
cudaError_t launch_kgrad_finish(cudaStream_t st, const double* partial, int ntiles_, int batch, const double* alpha, size_t alpha_stride,
                                int N, double* out) {
  kgrad_finish_kernel<<<batch, 256, 0, st>>>(partial, ntiles_, alpha, alpha_stride, N, out);
  return cudaGetLastError();
}

// X (rectangular tiles, nt x nt per latent) <- identity
__global__ void __launch_bounds__(256) rect_identity_kernel(TiledRect X) {
  const int b = blockIdx.y, R = blockIdx.x / X.ntc, J = blockIdx.x % X.ntc;
  double* tile = X.tile(b, R, J);
  for (int e = threadIdx.x; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    tile[e] = (R == J && r == c) ? 1.0 : 0.0;
  }
}
cudaError_t launch_rect_identity(cudaStream_t st, TiledRect X, int batch) {
  dim3 grid((unsigned)(X.ntr * X.ntc), (unsigned)batch);
  rect_identity_kernel<<<grid, 256, 0, st>>>(X);
  return cudaGetLastError();
}


// ---- general-ILMM gradient: the joint (mN x mN) matrix C = blockdiag(K_a) + ΣT ⊗ I -----------------
// Latent blocks start at multiples of N, not of the tile size, so elements are addressed one by one.
__device__ __forceinline__ double sym_get(const double* __restrict__ base, int r, int c) {  // r >= c
  return base[sym_tile_index(r / TILE, c / TILE) * TT + tile_elem(r % TILE, c % TILE)];
}

// grid (chunks of the lower triangle of one N x N latent block, m).  x is [N][D] (unpadded), alpha the
// joint vector (latent a at a*N).  partial[(a*nchunks + chunk)*2 + {0,1}] = {<G_aa, κ>, <G_aa, dK_a/ds>}.
__global__ void __launch_bounds__(256) kgrad_block_kernel(TiledSym negCinv, const double* __restrict__ x, int N, int D,
                                                          const LatentParams* __restrict__ params, const double* __restrict__ alpha,
                                                          int form, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* xa = sm;
  double* xb = xa + TILE * D;
  double* sa = xb + TILE * D;
  double* sb = sa + TILE;
  double* aa = sb + TILE;
  double* ab = aa + TILE;
  __shared__ double red[2][256];
  const int a = blockIdx.y, tl = blockIdx.x, t = threadIdx.x;
  int I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
  while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
  while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
  const int J = tl - (int)((size_t)I * (I + 1) / 2);
  const LatentParams lp = params[a];
  const double rinv_ls = 1.0 / lp.inv_ls;
  for (int i = t; i < TILE * D; i += 256) {
    const int ra = I * TILE + i / D, rb = J * TILE + i / D;
    const double s = input_scale(params + a, i % D);
    xa[i] = ra < N ? x[(size_t)ra * D + i % D] * s : 0.0;
    xb[i] = rb < N ? x[(size_t)rb * D + i % D] * s : 0.0;
  }
  if (t < TILE) {
    aa[t] = (I * TILE + t < N) ? alpha[(size_t)a * N + I * TILE + t] : 0.0;
    ab[t] = (J * TILE + t < N) ? alpha[(size_t)a * N + J * TILE + t] : 0.0;
  }
  __syncthreads();
  for (int i = t; i < 2 * TILE; i += 256) {
    const double* v = (i < TILE) ? xa + (size_t)i * D : xb + (size_t)(i - TILE) * D;
    double s = 0.0;
    for (int k = 0; k < D; ++k) s = fma(v[k], v[k], s);
    if (i < TILE) sa[i] = s; else sb[i - TILE] = s;
  }
  __syncthreads();
  const double* base = negCinv.base;
  double gv = 0.0, gs = 0.0;
  for (int e = t; e < TT; e += 256) {
    const int c = ((e >> 9) << 2) + (e & 3), r = (e >> 2) & 127;
    const int gr = I * TILE + r, gc = J * TILE + c;
    if (gr >= N || gc >= N || gc > gr) continue;
    const double G = 0.5 * (aa[r] * ab[c] + sym_get(base, a * N + gr, a * N + gc));
    if (gr == gc) {
      gv += G;
    } else {
      const double d2 = sqdist(xa + (size_t)r * D, xb + (size_t)c * D, D, sa[r], sb[c], form);
      double kap, dk;
      kappa_and_ds(lp.kind, d2, rinv_ls, kap, dk, lp.param);
      gv = fma(2.0 * G, kap, gv);
      gs = fma(2.0 * G, dk, gs);
    }
  }
  red[0][t] = gv; red[1][t] = gs;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) {
      red[0][t] += red[0][t + w];
      red[1][t] += red[1][t + w];
    }
    __syncthreads();
  }
  if (t == 0) {
    double* out = partial + ((size_t)a * gridDim.x + tl) * 2;
    out[0] = red[0][0];
    out[1] = red[1][0] * lp.variance;
  }
}

// out[a*3 + {0,1,2}] = {Σ chunks <G,κ>, Σ chunks <G,dK/ds>, Σ_n α[a*N+n]} in a fixed order.
__global__ void __launch_bounds__(256) kgrad_block_finish_kernel(const double* __restrict__ partial, int nchunks,
                                                                 const double* __restrict__ alpha, int N, double* __restrict__ out) {
  __shared__ double red[3][256];
  const int a = blockIdx.x, t = threadIdx.x;
  double s0 = 0, s1 = 0, s2 = 0;
  for (int i = t; i < nchunks; i += 256) {
    s0 += partial[((size_t)a * nchunks + i) * 2];
    s1 += partial[((size_t)a * nchunks + i) * 2 + 1];
  }
  for (int i = t; i < N; i += 256) s2 += alpha[(size_t)a * N + i];
  red[0][t] = s0; red[1][t] = s1; red[2][t] = s2;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w)
      for (int k = 0; k < 3; ++k) red[k][t] += red[k][t + w];
    __syncthreads();
  }
  if (t < 3) out[(size_t)a * 3 + t] = red[t][0];
}

// B[a,b] = Σ_n G[(a,n),(b,n)] = Σ_n (α_a[n] α_b[n] - C^{-1}[(a,n),(b,n)])/2  (m x m, symmetric): d lml / dΣT.
__global__ void __launch_bounds__(256) block_trace_kernel(TiledSym negCinv, const double* __restrict__ alpha, int N, int m,
                                                          double* __restrict__ B) {
  __shared__ double red[256];
  const int a = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  if (b > a) return;
  double s = 0.0;
  for (int n = t; n < N; n += 256) s += 0.5 * (alpha[(size_t)a * N + n] * alpha[(size_t)b * N + n] + sym_get(negCinv.base, a * N + n, b * N + n));
  red[t] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) red[t] += red[t + w];
    __syncthreads();
  }
  if (t == 0) {
    B[(size_t)b * m + a] = red[0];
    B[(size_t)a * m + b] = red[0];
  }
}

cudaError_t launch_kgrad_joint(cudaStream_t st, TiledSym negCinv, const double* x, int N, int D, const LatentParams* params, int m,
                               const double* alpha, int form, double* partial, double* out3, double* B) {
  const size_t smem = (size_t)(2 * TILE * D + 4 * TILE) * sizeof(double);
  if (smem + 1024 > 48 * 1024) {  // static shared memory counts against the default limit too
    cudaError_t e = cudaFuncSetAttribute(kgrad_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int ntn = (N + TILE - 1) / TILE;
  const int nchunks = (int)sym_tiles(ntn);
  dim3 grid((unsigned)nchunks, (unsigned)m);
  kgrad_block_kernel<<<grid, 256, smem, st>>>(negCinv, x, N, D, params, alpha, form, partial);
  kgrad_block_finish_kernel<<<m, 256, 0, st>>>(partial, nchunks, alpha, N, out3);
  dim3 g2((unsigned)m, (unsigned)m);
  block_trace_kernel<<<g2, 256, 0, st>>>(negCinv, alpha, N, m, B);
  return cudaGetLastError();
}

cudaError_t launch_block_trace(cudaStream_t st, TiledSym negCinv, const double* alpha, int N, int m, double* B) {
  dim3 g2((unsigned)m, (unsigned)m);
  block_trace_kernel<<<g2, 256, 0, st>>>(negCinv, alpha, N, m, B);
  return cudaGetLastError();
}

// ---- gradient w.r.t. the ARD multipliers (only launched when a latent has an ARDTransform and the caller asks) ------
// dK/d a_k = variance κ'(d²) 2 u_k² / a_k with u = scaled coordinate difference; κ'(d²) 2 = (dK/ds per unit variance) s / d².
// One generic kernel for both storages: per-latent factors (mat_batch_stride = tile storage of one latent, off_per_lat = 0)
// and the joint ILMM matrix (mat_batch_stride = 0, off_per_lat = N).  grid (chunks of the lower triangle, latents).
__global__ void __launch_bounds__(256) kgrad_ard_kernel(const double* __restrict__ mat_base, size_t mat_batch_stride, int off_per_lat,
                                                        const double* __restrict__ x, int N, int D,
                                                        const LatentParams* __restrict__ params, const double* __restrict__ alpha,
                                                        size_t alpha_stride, int form, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* xa = sm;
  double* xb = xa + TILE * D;
  double* sa = xb + TILE * D;
  double* sb = sa + TILE;
  double* aa = sb + TILE;
  double* ab = aa + TILE;
  __shared__ double red[MAX_ARD][256];
  const int lat = blockIdx.y, tl = blockIdx.x, t = threadIdx.x;
  double ga[MAX_ARD];
#pragma unroll
  for (int k = 0; k < MAX_ARD; ++k) ga[k] = 0.0;
  const LatentParams lp = params[lat];
  if (lp.ard_dim > 0) {  // block-uniform
    int I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
    while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
    while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
    const int J = tl - (int)((size_t)I * (I + 1) / 2);
    const double* base = mat_base + (size_t)lat * mat_batch_stride;
    const int off = lat * off_per_lat;
    const double* al = alpha + (size_t)lat * alpha_stride;
    const double rinv_ls = 1.0 / lp.inv_ls;
    for (int i = t; i < TILE * D; i += 256) {
      const int ra = I * TILE + i / D, rb = J * TILE + i / D;
      const double s = input_scale(params + lat, i % D);
      xa[i] = ra < N ? x[(size_t)ra * D + i % D] * s : 0.0;
      xb[i] = rb < N ? x[(size_t)rb * D + i % D] * s : 0.0;
    }
    if (t < TILE) {
      aa[t] = (I * TILE + t < N) ? al[I * TILE + t] : 0.0;
      ab[t] = (J * TILE + t < N) ? al[J * TILE + t] : 0.0;
    }
    __syncthreads();
    for (int i = t; i < 2 * TILE; i += 256) {
      const double* v = (i < TILE) ? xa + (size_t)i * D : xb + (size_t)(i - TILE) * D;
      double s = 0.0;
      for (int k = 0; k < D; ++k) s = fma(v[k], v[k], s);
      if (i < TILE) sa[i] = s; else sb[i - TILE] = s;
    }
    __syncthreads();
    for (int e = t; e < TT; e += 256) {
      const int c = ((e >> 9) << 2) + (e & 3), r = (e >> 2) & 127;
      const int gr = I * TILE + r, gc = J * TILE + c;
      if (gr >= N || gc >= N || gc >= gr) continue;  // strictly lower: the diagonal does not depend on the inputs
      const double d2 = sqdist(xa + (size_t)r * D, xb + (size_t)c * D, D, sa[r], sb[c], form);
      if (!(d2 > 0.0)) continue;
      const double G = 0.5 * (aa[r] * ab[c] + sym_get(base, off + gr, off + gc));
      double kap, dk;
      kappa_and_ds(lp.kind, d2, rinv_ls, kap, dk, lp.param);
      const double w = 2.0 * G * dk * lp.inv_ls / d2;
#pragma unroll
      for (int k = 0; k < MAX_ARD; ++k)
        if (k < D) {
          const double u = xa[(size_t)r * D + k] - xb[(size_t)c * D + k];
          ga[k] = fma(w, u * u, ga[k]);
        }
    }
  }
#pragma unroll
  for (int k = 0; k < MAX_ARD; ++k) red[k][t] = ga[k];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w)
#pragma unroll
      for (int k = 0; k < MAX_ARD; ++k) red[k][t] += red[k][t + w];
    __syncthreads();
  }
  if (t < MAX_ARD) partial[((size_t)lat * gridDim.x + tl) * MAX_ARD + t] = red[t][0] * lp.variance / (lp.ard_dim > 0 ? params[lat].ard[t] : 1.0);
}
// out[lat*MAX_ARD + k] = Σ chunks (fixed order)
__global__ void __launch_bounds__(256) kgrad_ard_finish_kernel(const double* __restrict__ partial, int nchunks, double* __restrict__ out) {
  __shared__ double red[256];
  const int lat = blockIdx.x, k = blockIdx.y, t = threadIdx.x;
  double s = 0.0;
  for (int i = t; i < nchunks; i += 256) s += partial[((size_t)lat * nchunks + i) * MAX_ARD + k];
  red[t] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (t < w) red[t] += red[t + w];
    __syncthreads();
  }
  if (t == 0) out[(size_t)lat * MAX_ARD + k] = red[0];
}
cudaError_t launch_kgrad_ard(cudaStream_t st, const double* mat_base, size_t mat_batch_stride, int off_per_lat, const double* x, int N, int D,
                             const LatentParams* params, int nlat, const double* alpha, size_t alpha_stride, int form, double* partial,
                             double* out) {
  const size_t smem = (size_t)(2 * TILE * D + 4 * TILE) * sizeof(double);
  const int ntn = (N + TILE - 1) / TILE;
  const int nchunks = (int)sym_tiles(ntn);
  dim3 grid((unsigned)nchunks, (unsigned)nlat);
  kgrad_ard_kernel<<<grid, 256, smem, st>>>(mat_base, mat_batch_stride, off_per_lat, x, N, D, params, alpha, alpha_stride, form, partial);
  dim3 g2((unsigned)nlat, (unsigned)MAX_ARD);
  kgrad_ard_finish_kernel<<<g2, 256, 0, st>>>(partial, nchunks, out);
  return cudaGetLastError();
}

// v[i] = a[i] * sa - b[i]
__global__ void scale_sub_kernel(double* __restrict__ v, const double* __restrict__ a, double sa, const double* __restrict__ b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = a[i] * sa - b[i];
}
cudaError_t launch_scale_sub(cudaStream_t st, double* v, const double* a, double sa, const double* b, size_t n) {
  scale_sub_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, a, sa, b, n);
  return cudaGetLastError();
}

}  // namespace lmm
