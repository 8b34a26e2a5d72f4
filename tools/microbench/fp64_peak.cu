// FP64 peak probes for B200 (sm_100a): register-resident DMMA / DFMA loops, cuBLAS DGEMM/DSYRK,
// cuSOLVER DPOTRF. Output: one JSON object on stdout. This is measurement tooling, not product code:
// the cuBLAS/cuSOLVER numbers are the *denominator* (roofline) and the vendor bar, nothing links
// them into liblmm.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s @%d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double seed) {
  double acc[NACC][2];
  double a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = seed * (threadIdx.x + i + 1); b[i] = seed * (threadIdx.x * 3 + i + 2); }
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = 0; acc[i][1] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a[i & 3], b[(i >> 2) & 3]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double seed) {
  double acc[NACC][4];
  double a[2][4], b[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[j][i] = seed * (threadIdx.x + i + 1 + j);
    b[j][0] = seed * (threadIdx.x + 7 + j); b[j][1] = seed * (threadIdx.x + 9 + j);
  }
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma1688(acc[i], a[i & 1], b[(i >> 1) & 1]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double seed) {
  double acc[NACC][4];
  double a[8], b[2][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + i + 1);
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) b[j][i] = seed * (threadIdx.x + 3 * i + j);
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma16816(acc[i], a, b[i & 1]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double seed) {
  double acc[NACC];
  double a = seed * threadIdx.x, b = seed + 1.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <class F>
static float time_ms(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int nsm = prop.multiProcessorCount;
  double* dout; CK(cudaMalloc(&dout, 64));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"mem_gb\": %.1f", prop.name, nsm, prop.major, prop.minor, prop.totalGlobalMem / 1e9);

  const int iters = 20000;
  // ---- DMMA m8n8k4: 256 FMA = 512 flop / instr / warp
  printf(", \"dmma_m8n8k4\": {");
  bool first = true;
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { k_dmma884<16><<<nsm, warps * 32>>>(dout, iters, 1e-9); });
    double tf = (double)nsm * warps * iters * 16 * 512.0 / (ms * 1e-3) / 1e12;
    printf("%s\"w%d\": %.2f", first ? "" : ", ", warps, tf); first = false;
  }
  {
    float ms = time_ms([&] { k_dmma884<16><<<nsm * 2, 8 * 32>>>(dout, iters, 1e-9); });
    printf(", \"2cta_w8\": %.2f", (double)nsm * 2 * 8 * iters * 16 * 512.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma884<4><<<nsm, 4 * 32>>>(dout, iters, 1e-9); });
    printf(", \"w4_acc4\": %.2f", (double)nsm * 4 * iters * 4 * 512.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_dmma884<32><<<nsm, 4 * 32>>>(dout, iters, 1e-9); });
    printf(", \"w4_acc32\": %.2f", (double)nsm * 4 * iters * 32 * 512.0 / (ms * 1e-3) / 1e12);
  }
  printf("}");
  printf(", \"dmma_m16n8k8\": {"); first = true;
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { k_dmma1688<8><<<nsm, warps * 32>>>(dout, iters, 1e-9); });
    double tf = (double)nsm * warps * iters * 8 * 2048.0 / (ms * 1e-3) / 1e12;
    printf("%s\"w%d\": %.2f", first ? "" : ", ", warps, tf); first = false;
  }
  printf("}");
  printf(", \"dmma_m16n8k16\": {"); first = true;
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { k_dmma16816<8><<<nsm, warps * 32>>>(dout, iters, 1e-9); });
    double tf = (double)nsm * warps * iters * 8 * 4096.0 / (ms * 1e-3) / 1e12;
    printf("%s\"w%d\": %.2f", first ? "" : ", ", warps, tf); first = false;
  }
  printf("}");
  printf(", \"dfma\": {"); first = true;
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { k_dfma<16><<<nsm, warps * 32>>>(dout, iters * 4, 1e-9); });
    double tf = (double)nsm * warps * 32 * (iters * 4.0) * 16 * 2.0 / (ms * 1e-3) / 1e12;
    printf("%s\"w%d\": %.2f", first ? "" : ", ", warps, tf); first = false;
  }
  printf("}");
  fflush(stdout);

  // ---- HBM copy
  {
    size_t n = (size_t)1 << 30;  // 1 GiB each way
    char *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n));
    CK(cudaMemset(a, 1, n));
    float ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n, cudaMemcpyDeviceToDevice)); }, 5);
    printf(", \"hbm_copy_gbs\": %.1f", 2.0 * n / (ms * 1e-3) / 1e9);
    CK(cudaFree(a)); CK(cudaFree(b));
  }

  // ---- cuBLAS DGEMM / DSYRK, cuSOLVER DPOTRF
  cublasHandle_t hb; cublasCreate(&hb);
  cusolverDnHandle_t hs; cusolverDnCreate(&hs);
  printf(", \"cublas\": {"); first = true;
  for (int n : {4096, 8192, 16384}) {
    double *A, *B, *C; size_t bytes = (size_t)n * n * 8;
    CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
    std::vector<double> h((size_t)n * n);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    CK(cudaMemcpy(A, h.data(), bytes, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, h.data(), bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(C, 0, bytes));
    double one = 1.0, zero = 0.0, mone = -1.0;
    float ms = time_ms([&] { cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n); });
    printf("%s\"dgemm_nt_%d\": %.2f", first ? "" : ", ", n, 2.0 * n * (double)n * n / (ms * 1e-3) / 1e12); first = false;
    ms = time_ms([&] { cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); });
    printf(", \"dgemm_nn_%d\": %.2f", n, 2.0 * n * (double)n * n / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { cublasDsyrk(hb, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &mone, A, n, &one, C, n); });
    printf(", \"dsyrk_%d\": %.2f", n, 1.0 * n * (double)n * n / (ms * 1e-3) / 1e12);
    if (n == 8192) {  // sustained: back-to-back for ~3 s
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      int reps = 60; cudaEventRecord(e0);
      for (int r = 0; r < reps; ++r) cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1);
      printf(", \"dgemm_nt_8192_sustained\": %.2f", reps * 2.0 * n * (double)n * n / (t * 1e-3) / 1e12);
    }
    // potrf: build SPD matrix C = A*A^T/n + I*n
    {
      // diag-dominant SPD: reuse C := 0.001*A*A^T then add n on the diagonal via a tiny host loop on a strided memcpy
      double sc = 1e-3;
      cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &sc, A, n, A, n, &zero, C, n);
      std::vector<double> dg(n, 0.0);
      CK(cudaMemcpy2D(dg.data(), 8, C, (size_t)(n + 1) * 8, 8, n, cudaMemcpyDeviceToHost));
      for (int i = 0; i < n; ++i) dg[i] += 10.0;
      CK(cudaMemcpy2D(C, (size_t)(n + 1) * 8, dg.data(), 8, 8, n, cudaMemcpyHostToDevice));
      int lwork = 0; cusolverDnDpotrf_bufferSize(hs, CUBLAS_FILL_MODE_LOWER, n, C, n, &lwork);
      double* work; CK(cudaMalloc(&work, (size_t)lwork * 8)); int* info; CK(cudaMalloc(&info, 4));
      CK(cudaMemcpy(B, C, bytes, cudaMemcpyDeviceToDevice));
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        CK(cudaMemcpy(C, B, bytes, cudaMemcpyDeviceToDevice));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); cusolverDnDpotrf(hs, CUBLAS_FILL_MODE_LOWER, n, C, n, work, lwork, info);
        cudaEventRecord(e1); cudaEventSynchronize(e1); float t; cudaEventElapsedTime(&t, e0, e1); if (t < best) best = t;
      }
      int hinfo = -1; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
      printf(", \"dpotrf_%d_tflops\": %.2f, \"dpotrf_%d_ms\": %.2f, \"dpotrf_%d_info\": %d", n, (double)n * n * n / 3.0 / (best * 1e-3) / 1e12, n, best, n, hinfo);
      CK(cudaFree(work)); CK(cudaFree(info));
    }
    CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
    fflush(stdout);
  }
  printf("}}\n");
  return 0;
}
