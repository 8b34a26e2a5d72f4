// Phase-level cycle counts of potrf_tile_kernel (clock64 at phase boundaries): tuning aid.
#define LMM_POTRF_TIMING 1
#include "../../linearmixingmodels.jl_b200/csrc/potrf.cu"
#include <cstdio>
#include <vector>
using namespace lmm;
int main() {
  const int batch = 8;
  std::vector<double> h((size_t)batch * TT);
  for (int b = 0; b < batch; ++b)
    for (int r = 0; r < TILE; ++r)
      for (int c = 0; c < TILE; ++c) h[(size_t)b * TT + tile_elem(r, c)] = exp(-0.5 * 0.01 * (r - c) * (r - c)) + (r == c ? 0.1 : 0.0);
  double *dL, *dW, *dld; int* dinfo;
  cudaMalloc(&dL, h.size() * 8); cudaMalloc(&dW, h.size() * 8); cudaMalloc(&dld, batch * 8); cudaMalloc(&dinfo, batch * 4);
  cudaMemset(dld, 0, batch * 8); cudaMemset(dinfo, 0, batch * 4);
  TiledSym L{dL, 1, (size_t)TT};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(dL, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    launch_potrf_tile(0, L, dW, (size_t)TT, 0, batch, dld, dinfo);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk[16]; cudaMemcpyFromSymbol(clk, g_potrf_clk, sizeof clk);
    const char* names[] = {"load", "C:first panel(a)", "C:first trailing(b)", "C:remaining 15 steps", "logdet", "store L", "W level0", "W levels", "store W"};
    printf("rep %d: %.1f us total;", rep, ms * 1e3);
    for (int k = 0; k < 9; ++k) printf(" %s=%lld", names[k], clk[k + 1] - clk[k]);
    printf(" cycles (sum %lld)\n", clk[9] - clk[0]);
  }
  int info[8]; cudaMemcpy(info, dinfo, sizeof info, cudaMemcpyDeviceToHost); printf("info0=%d err=%s\n", info[0], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
