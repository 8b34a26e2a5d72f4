// Small elementwise / bookkeeping kernels of the host drivers and their launchers.
#include "common.cuh"
#include "kernels.h"

namespace lmm {

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

__global__ void fill_noise_kernel(double* nv, size_t stride, int n_old, int n_new, const double* old_vec, size_t old_stride,
                                  const LatentParams* old_params, const double* new_noise) {
  const int i = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_old + n_new) return;
  double v;
  if (k < n_old) v = old_vec ? old_vec[(size_t)i * old_stride + k] : old_params[i].noise;
  else v = new_noise[i];
  nv[(size_t)i * stride + k] = v;
}

// dst[i][:] = E (m*m doubles) for i in [0, n)
__global__ void repeat_block_kernel(double* __restrict__ dst, const double* __restrict__ E, int mm, size_t total) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < total) dst[k] = E[k % mm];
}

cudaError_t launch_lml_terms(cudaStream_t st, double* terms, int slot0, int nb, const double* logdet, const double* quad, int n, double log2pi) {
  if (nb <= 0) return cudaSuccess;
  lml_terms_kernel<<<(nb + 127) / 128, 128, 0, st>>>(terms, slot0, nb, logdet, quad, n, log2pi);
  return cudaGetLastError();
}
cudaError_t launch_regulariser(cudaStream_t st, double* slot, double c0, const double* resid, double sigma2) {
  regulariser_kernel<<<1, 1, 0, st>>>(slot, c0, resid, sigma2);
  return cudaGetLastError();
}
cudaError_t launch_add_scalar(cudaStream_t st, double* v, size_t n, double s) {
  if (n == 0) return cudaSuccess;
  add_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, s);
  return cudaGetLastError();
}
cudaError_t launch_copy_add(cudaStream_t st, int nlat, double* out, size_t stride_out, const double* in, size_t stride_in, int n, double s) {
  if (nlat <= 0 || n <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)nlat);
  copy_add_kernel<<<grid, 256, 0, st>>>(out, stride_out, in, stride_in, n, s);
  return cudaGetLastError();
}
cudaError_t launch_add_mean(cudaStream_t st, int nlat, double* v, size_t stride, int n, const LatentParams* params) {
  if (nlat <= 0 || n <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)nlat);
  add_mean_kernel<<<grid, 256, 0, st>>>(v, stride, n, params);
  return cudaGetLastError();
}
cudaError_t launch_axpy(cudaStream_t st, double* y, const double* x, size_t n, double a) {
  if (n == 0) return cudaSuccess;
  axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, x, n, a);
  return cudaGetLastError();
}
cudaError_t launch_fill_noise(cudaStream_t st, int nlat, double* nv, size_t stride, int n_old, int n_new, const double* old_vec, size_t old_stride,
                              const LatentParams* old_params, const double* new_noise) {
  if (nlat <= 0) return cudaSuccess;
  dim3 grid((unsigned)((n_old + n_new + 255) / 256), (unsigned)nlat);
  fill_noise_kernel<<<grid, 256, 0, st>>>(nv, stride, n_old, n_new, old_vec, old_stride, old_params, new_noise);
  return cudaGetLastError();
}
cudaError_t launch_repeat_block(cudaStream_t st, double* dst, const double* E, int mm, size_t total) {
  if (total == 0) return cudaSuccess;
  repeat_block_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dst, E, mm, total);
  return cudaGetLastError();
}

}  // namespace lmm
