// Does the FP64 tensor sub-pipe (DMMA) share its datapath with the FP64 FMA pipe (DFMA) on B200?
// ncu books DMMA under smsp__pipe_tensor_subpipe_dmma and DFMA under sm__pipe_fp64, both peak at
// 64 FMA/clk/SM when run alone (fp64_peak.cu).  This probe runs both instruction kinds at once --
// interleaved in one warp, and split over specialised warps -- and prints the combined TFLOP/s: a sum
// above the single-pipe 37 TFLOP/s would mean the trailing update could use both.
// Measurement tooling only (not linked into liblmm).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s @%d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// NM DMMA (256 FMA per warp instr) and NF DFMA (32 FMA per warp instr) per iteration, all independent chains
template <int NM, int NF>
__global__ void k_mix(double* out, int iters, double seed) {
  double am[NM > 0 ? NM : 1][2], af[NF > 0 ? NF : 1];
  const double a = seed * (threadIdx.x + 1), b = seed * (threadIdx.x * 3 + 2), c = seed + 1.0;
#pragma unroll
  for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { am[i][0] = 0; am[i][1] = 0; }
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) af[i] = i;
  for (int it = 0; it < iters; ++it) {
    constexpr int R = NM > 0 ? (NF / NM) : 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      dmma884(am[i][0], am[i][1], a, b);
#pragma unroll
      for (int j = 0; j < R; ++j) af[i * R + j] = fma(af[i * R + j], a, c);
    }
    if (NM == 0) {
#pragma unroll
      for (int j = 0; j < NF; ++j) af[j] = fma(af[j], a, c);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += am[i][0] + am[i][1];
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += af[i];
  if (s == 123.456) out[0] = s;
}

// warp-specialised: warps with (warp % period) < nd run DMMA only, the others DFMA only
__global__ void k_split(double* out, int iters, double seed, int period, int nd, unsigned long long* flops) {
  const int warp = threadIdx.x >> 5;
  const double a = seed * (threadIdx.x + 1), b = seed * (threadIdx.x * 3 + 2), c = seed + 1.0;
  double s = 0;
  if ((warp % period) < nd) {
    double am[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { am[i][0] = 0; am[i][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(am[i][0], am[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += am[i][0] + am[i][1];
  } else {
    double af[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) af[i] = i;
    for (int it = 0; it < iters * 8; ++it) {  // 8 x 16 DFMA = the FMA count of 16 DMMA
#pragma unroll
      for (int i = 0; i < 16; ++i) af[i] = fma(af[i], a, c);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += af[i];
  }
  if (s == 123.456) out[0] = s;
}

template <class F>
static float time_ms(F f) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount, iters = 20000, warps = 8;
  double* dout; CK(cudaMalloc(&dout, 64));
  printf("{\"gpu\": \"%s\"", prop.name);
#define RUN(NM, NF)                                                                                              \
  {                                                                                                              \
    float ms = time_ms([&] { k_mix<NM, NF><<<nsm * 2, warps * 32>>>(dout, iters, 1e-9); });                      \
    double fl = (double)nsm * 2 * warps * iters * (NM * 512.0 + NF * 64.0);                                      \
    printf(", \"mix_dmma%d_dfma%d\": %.2f", NM, NF, fl / (ms * 1e-3) / 1e12);                                    \
  }
  RUN(8, 0) RUN(0, 64) RUN(8, 8) RUN(8, 16) RUN(8, 32) RUN(8, 64) RUN(4, 64)
  for (int nd : {8, 6, 4, 2, 0}) {
    float ms = time_ms([&] { k_split<<<nsm * 2, 8 * 32>>>(dout, iters, 1e-9, 8, nd, nullptr); });
    double fl = (double)nsm * 2 * 8 * iters * 16 * 512.0;  // every warp does the same FMA count
    printf(", \"split_%ddmma_%ddfma_warps\": %.2f", nd, 8 - nd, fl / (ms * 1e-3) / 1e12);
  }
  printf("}\n");
  return 0;
}
