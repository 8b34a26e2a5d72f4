// Phase-level cycle counts of the diagonal-tile kernel (clock64 at phase boundaries) and a check of L, W = inv(L) and logdet
// against a plain host Cholesky of the same tile: tuning aid.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a potrf_phases.cu -o potrf_phases
#define LMM_POTRF_TIMING 1
#include "../../linearmixingmodels.jl_b200/csrc/potrf.cu"
namespace lmm { bool pdl_enabled() { return false; } }  // defined in gemm.cu in the library
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
  for (int impl = 1; impl < 2; ++impl) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemcpy(dL, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
      cudaMemset(dld, 0, batch * 8); cudaMemset(dinfo, 0, batch * 4); cudaMemset(dW, 0xff, h.size() * 8);
      cudaEventRecord(e0);
      launch_potrf_tile(0, L, dW, (size_t)TT, 0, batch, dld, dinfo);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long clk[16]; cudaMemcpyFromSymbol(clk, g_potrf_clk, sizeof clk);
      // phases: load | first panel + second update | second panel | other 14 steps | W row 15 + logdet | - | - | stores | -
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
  {  // host reference: unblocked Cholesky of tile 0 and the product W L = I
    std::vector<double> A(TILE * TILE), Lh(TILE * TILE, 0.0);
    for (int r = 0; r < TILE; ++r)
      for (int c = 0; c < TILE; ++c) A[r * TILE + c] = h[tile_elem(r, c)];
    double ldh = 0.0;
    for (int j = 0; j < TILE; ++j) {
      double d = A[j * TILE + j];
      for (int k = 0; k < j; ++k) d -= Lh[j * TILE + k] * Lh[j * TILE + k];
      Lh[j * TILE + j] = sqrt(d);
      ldh += 2.0 * log(Lh[j * TILE + j]);
      for (int i = j + 1; i < TILE; ++i) {
        double v = A[i * TILE + j];
        for (int k = 0; k < j; ++k) v -= Lh[i * TILE + k] * Lh[j * TILE + k];
        Lh[i * TILE + j] = v / Lh[j * TILE + j];
      }
    }
    double dl = 0, dwl = 0;
    for (int r = 0; r < TILE; ++r)
      for (int c = 0; c < TILE; ++c) {
        dl = fmax(dl, fabs(outL[1][tile_elem(r, c)] - Lh[r * TILE + c]));
        double s = 0.0;
        for (int k = 0; k < TILE; ++k) s += outW[1][tile_elem(r, k)] * outL[1][tile_elem(k, c)];
        dwl = fmax(dwl, fabs(s - (r == c ? 1.0 : 0.0)));
      }
    printf("vs host: max |dL| = %.3e   max |W L - I| = %.3e   logdet diff = %.3e\n", dl, dwl, fabs(ld[1] - ldh));
  }
  return 0;
}
