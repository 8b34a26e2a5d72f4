// Phase-level cycle counts of the two diagonal-tile kernels (clock64 at phase boundaries) and a check that they agree
// on L, W = inv(L) and logdet: tuning aid.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a potrf_phases.cu -o potrf_phases
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
  TiledSym L{dL, 1, (size_t)TT};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<double> outL[2], outW[2]; double ld[2] = {0, 0};
  for (int impl = 0; impl < 2; ++impl) {
    set_potrf_impl(impl);
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemcpy(dL, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
      cudaMemset(dld, 0, batch * 8); cudaMemset(dinfo, 0, batch * 4); cudaMemset(dW, 0xff, h.size() * 8);
      cudaEventRecord(e0);
      launch_potrf_tile(0, L, dW, (size_t)TT, 0, batch, dld, dinfo);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long clk[16]; cudaMemcpyFromSymbol(clk, g_potrf_clk, sizeof clk);
      // impl 0: load | first panel | first trailing | other 15 steps | logdet | store L | W level 0 | W levels | store W
      // impl 1: load | first panel + second update | second panel | other 14 steps | W row 15 + logdet | - | - | stores | -
      printf("impl %d rep %d: %.1f us total; phase cycles:", impl, rep, ms * 1e3);
      for (int k = 0; k < 9; ++k) printf(" %lld", clk[k + 1] - clk[k]);
      printf(" (sum %lld)\n", clk[9] - clk[0]);
      if (impl == 1) {
        long long acc[8]; cudaMemcpyFromSymbol(acc, g_potrf_acc, sizeof acc);
        printf("    step sums: column update %lld | panel (warp 0) %lld | warp 0 waiting at the step barrier %lld | W block rows (warp 4) %lld\n", acc[0], acc[1], acc[2], acc[3]);
        long long z[8] = {0}; cudaMemcpyToSymbol(g_potrf_acc, z, sizeof z);
      }
    }
    outL[impl].resize(h.size()); outW[impl].resize(h.size());
    cudaMemcpy(outL[impl].data(), dL, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(outW[impl].data(), dW, h.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&ld[impl], dld, 8, cudaMemcpyDeviceToHost);
    int info[8]; cudaMemcpy(info, dinfo, sizeof info, cudaMemcpyDeviceToHost);
    printf("impl %d: info0=%d logdet0=%.17g err=%s\n", impl, info[0], ld[impl], cudaGetErrorString(cudaGetLastError()));
  }
  for (int impl = 1; impl < 2; ++impl) {
    double dl = 0, dw = 0, wmax = 0;
    for (size_t i = 0; i < h.size(); ++i) {
      dl = fmax(dl, fabs(outL[0][i] - outL[impl][i])); dw = fmax(dw, fabs(outW[0][i] - outW[impl][i])); wmax = fmax(wmax, fabs(outW[0][i]));
    }
    printf("impl %d vs 0: max |dL| = %.3e   max |dW| = %.3e (max |W| = %.3e)   logdet diff = %.3e\n", impl, dl, dw, wmax, fabs(ld[0] - ld[impl]));
  }
  return 0;
}
