// Projection of the observations onto the latents + regulariser residual (K2/K3), and the
// back-projection of latent marginals to the outputs (K11).
// Reference: src/oilmm.jl:85 `Ty = T*Y`, :111-112 `sum(abs2, (I - U*U') * Y)`, :69-75
// `M = U*sqrt(S)*M_latent; V = abs2.(U*sqrt(S))*V_latent .+ σ²; vec(M')`; general ILMM:
// src/ilmm.jl:157-158, :179-180 (`Y .- H*T*Y`).
// Y is never transposed or copied: the by-outputs vector y IS the N x p column-major matrix, so
// row j of Y is the contiguous run y[j*N ...] and every global access below is coalesced along n.
//
// The projection is three skinny GEMMs over one block of columns of Y -- T*Y, Z = P*Y and Y - Q*Z (the residual
// never forms I - UU': O(pmN), not O(p²N)) -- i.e. 6pmN flop against 8(p+m)N bytes: 3pm/(4(p+m)) flop/B = 24 at
// p = m = 64, 14.5 at the reference's notebook shape (p = 600, m = 20), above the B200 FP64 balance of
// 37 TFLOP/s / 6.5 TB/s = 5.7 flop/B.  It is therefore FP64-pipe bound, and like every other true contraction in this
// library it runs on the FP64 tensor cores: project_dmma_kernel stages the Y block once in shared memory (cp.async),
// keeps Z in shared memory, and issues mma.m8n8k4.f64 (DMMA) for all three products.  Floor at C4: 4.0e8 flop /
// 37 TFLOP/s = 10.8 us, i.e. at most 16.8 MB / 10.8 us = 1.55 TB/s = 24 % of the HBM peak whatever the kernel does.
// The grid is sized for the SM count, not for the tile: the column block shrinks from 64 to 8 columns until there are
// >= 148 CTAs, and at small N (the notebook's N = 552: 69 blocks of 8 columns) the p rows of the residual are split over
// up to 4 CTAs per column block (each recomputes the small Z), where the round-1 kernel ran 18 CTAs on 148 SMs.
// project_kernel (round 1: register-tiled scalar-FMA GEMM) is kept selectable ("project_impl" = 0) as the cross-check.
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

// ------------------------------------------------------------------------------------------------
// DMMA projection kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884q(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

struct ProjArgs {
  const double* y; int N, p;
  const double* T; int m, lat0, mloc;
  const double* means; double* ty; size_t ty_stride;
  const double* P; const double* Q;
  double* resid_partial; double* resid_out; double* z_out;
  int psplit;  // CTAs per column block along the p rows of the residual (grid.y)
};

// One warp: C[8 x 8*NCB] += A[8 x k-range] * Bs[k-range x 8*NCB] for row block `rb` of A (column-major, leading dimension
// lda, `rows` valid rows starting at global row a_row0) and the column blocks cb0 .. cb0+NCB-1 of the shared operand Bs
// ([K4][LD], K4 = K rounded up to 4, rows beyond K are zero), over the k4-steps [ks0, ks1).  mma.m8n8k4: lane = (r, k)
// holds A[r][k] with r = lane/4, k = lane%4; B[k][n] with k = lane%4, n = lane/4; C[r][2*(lane%4) + {0,1}].  The A
// fragments of the next FOUR k-steps are in flight (through L1 / L2) while the current four feed the tensor pipe: with a
// one-step prefetch every k-step paid an L2 round trip at the notebook shape (p = 600: 150 dependent steps).
template <int NCB>
__device__ __forceinline__ void warp_gemm(const double* __restrict__ A, int lda, int a_row0, int rows, int rb, const double* Bs, int LD,
                                          int K, int cb0, int ks0, int ks1, double (&acc)[NCB][2]) {
  const int lane = threadIdx.x & 31, r = lane >> 2, k = lane & 3;
  const int row = rb * 8 + r;
  const bool rok = row < rows;
  const double* ap = A + (size_t)(a_row0 + row);
  const double* bp = Bs + (size_t)k * LD + cb0 * 8 + r;
  auto lda_at = [&](int ks) -> double {
    const int kk = ks * 4 + k;
    return (rok && ks < ks1 && kk < K) ? __ldg(ap + (size_t)kk * lda) : 0.0;
  };
  double a[4], an[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) a[u] = lda_at(ks0 + u);
  for (int ks = ks0; ks < ks1; ks += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) an[u] = lda_at(ks + 4 + u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (ks + u < ks1) {  // warp-uniform
        const double* b = bp + (size_t)(ks + u) * 4 * LD;
#pragma unroll
        for (int c = 0; c < NCB; ++c) dmma884q(acc[c][0], acc[c][1], a[u], b[c * 8]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = an[u];
  }
}

// All 8 warps: the row blocks [rb_lo, rb_hi) x NCG column groups of C = A * Bs, `epi(rb, cb0, acc)` called by the warp
// that holds the finished accumulators.  With fewer than 8 (row block, column group) tasks the K range is split over the
// idle warps (2, 4 or 8 slices of at least 8 k4-steps each) and the partial accumulators are summed through `scratch`
// (8 x NCB x 32 x 2 doubles) in a fixed order -- the notebook shape has m = 20 (3 row blocks) against K = p = 600.
template <int NCB, int NCG, class Epi>
__device__ __forceinline__ void gemm_rows(const double* __restrict__ A, int lda, int a_row0, int rows, int rb_lo, int rb_hi,
                                          const double* Bs, int LD, int K, double* scratch, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntask = (rb_hi - rb_lo) * NCG;
  const int ksteps = (K + 3) >> 2;
  int KS = 1;
  while (KS * 2 * ntask <= 8 && KS * 2 * 8 <= ksteps) KS *= 2;
  if (KS == 1) {
    for (int task = warp; task < ntask; task += 8) {
      const int rb = rb_lo + task / NCG, cb0 = (task % NCG) * NCB;
      double acc[NCB][2];
#pragma unroll
      for (int c = 0; c < NCB; ++c) acc[c][0] = acc[c][1] = 0.0;
      warp_gemm<NCB>(A, lda, a_row0, rows, rb, Bs, LD, K, cb0, 0, ksteps, acc);
      epi(rb, cb0, acc);
    }
    return;
  }
  const int task = warp / KS, ks = warp % KS;  // ntask * KS <= 8: at most one task per warp
  const bool active = task < ntask;
  const int rb = rb_lo + task / NCG, cb0 = (task % NCG) * NCB;
  double acc[NCB][2];
#pragma unroll
  for (int c = 0; c < NCB; ++c) acc[c][0] = acc[c][1] = 0.0;
  if (active) {
    const int k0 = (int)((long long)ksteps * ks / KS), k1 = (int)((long long)ksteps * (ks + 1) / KS);
    warp_gemm<NCB>(A, lda, a_row0, rows, rb, Bs, LD, K, cb0, k0, k1, acc);
    if (ks > 0) {
#pragma unroll
      for (int c = 0; c < NCB; ++c) {
        scratch[((size_t)(warp * NCB + c) * 32 + lane) * 2] = acc[c][0];
        scratch[((size_t)(warp * NCB + c) * 32 + lane) * 2 + 1] = acc[c][1];
      }
    }
  }
  __syncthreads();
  if (active && ks == 0) {
    for (int s2 = 1; s2 < KS; ++s2) {
#pragma unroll
      for (int c = 0; c < NCB; ++c) {
        acc[c][0] += scratch[((size_t)((warp + s2) * NCB + c) * 32 + lane) * 2];
        acc[c][1] += scratch[((size_t)((warp + s2) * NCB + c) * 32 + lane) * 2 + 1];
      }
    }
    epi(rb, cb0, acc);
  }
  __syncthreads();  // scratch may be reused by the next product
}

constexpr int PROJ_SCRATCH = 8 * 4 * 32 * 2;  // doubles (16 KB)

// grid (ceil(N / NB), psplit), 256 threads.  NB columns of Y per CTA; LD = NB + 8 for NB >= 16 (row stride = 8 mod 16
// doubles: the 4 x 8 B-fragment of a warp falls into two conflict-free 128-byte wavefronts), 8 for NB = 8.
template <int NB>
__global__ void __launch_bounds__(256) project_dmma_kernel(ProjArgs a) {
  constexpr int LD = NB >= 16 ? NB + 8 : NB;
  constexpr int NCBT = NB / 8;              // 8-column blocks per CTA
  constexpr int NCB = NCBT >= 4 ? 4 : NCBT;  // column blocks per warp task
  constexpr int NCG = NCBT / NCB;           // column groups
  extern __shared__ __align__(16) double sm[];
  const int p = a.p, m = a.m, N = a.N;
  const int p4 = (p + 3) & ~3, m4 = (m + 3) & ~3;
  double* Ys = sm;                            // [p4][LD]
  double* Zs = sm + (size_t)p4 * LD;          // [m4][LD]
  double* scratch = Zs + (size_t)m4 * LD;     // [PROJ_SCRATCH]
  __shared__ double red[8];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, r = lane >> 2, q = lane & 3;
  const int nb0 = blockIdx.x * NB;

  // ---- stage the Y block (zero beyond p rows / N columns); Z rows beyond m are zeroed once
  for (int idx = t; idx < p4 * NB; idx += 256) {
    const int j = idx / NB, nn = idx % NB;
    double* dst = Ys + (size_t)j * LD + nn;
    if (j < p && nb0 + nn < N) cp_async8(dst, a.y + (size_t)j * N + nb0 + nn);
    else *dst = 0.0;
  }
  for (int idx = t; idx < (m4 - m) * NB; idx += 256) Zs[(size_t)(m + idx / NB) * LD + idx % NB] = 0.0;
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const bool same_tp = a.P != nullptr && a.P == a.T && a.lat0 == 0 && a.mloc == m;  // general ILMM: Z = T Y is the projection itself
  // ---- Ty = T[lat0 : lat0 + mloc, :] Y - mean     (slice 0 only)
  if (blockIdx.y == 0 && !same_tp) {
    gemm_rows<NCB, NCG>(a.T, m, a.lat0, a.mloc, 0, (a.mloc + 7) >> 3, Ys, LD, p, scratch, [&](int rb, int cb0, double (&acc)[NCB][2]) {
      const int row = rb * 8 + r;
      if (row < a.mloc) {
        const double mu = a.means[row];
#pragma unroll
        for (int c = 0; c < NCB; ++c) {
          const int col = nb0 + (cb0 + c) * 8 + 2 * q;
          double* o = a.ty + (size_t)row * a.ty_stride + col;
          if (col < N) o[0] = acc[c][0] - mu;
          if (col + 1 < N) o[1] = acc[c][1] - mu;
        }
      }
    });
  }
  if (a.P == nullptr) return;
  // ---- Z = P Y  (m x NB) into shared memory (every slice needs it)
  gemm_rows<NCB, NCG>(a.P, m, 0, m, 0, (m + 7) >> 3, Ys, LD, p, scratch, [&](int rb, int cb0, double (&acc)[NCB][2]) {
    const int row = rb * 8 + r;
    if (row < m) {
      const double mu = same_tp ? a.means[row] : 0.0;
#pragma unroll
      for (int c = 0; c < NCB; ++c) {
        const int lc = (cb0 + c) * 8 + 2 * q, col = nb0 + lc;
        Zs[(size_t)row * LD + lc] = acc[c][0];
        Zs[(size_t)row * LD + lc + 1] = acc[c][1];
        if (blockIdx.y == 0) {
          if (a.z_out) {
            if (col < N) a.z_out[(size_t)row * N + col] = acc[c][0];
            if (col + 1 < N) a.z_out[(size_t)row * N + col + 1] = acc[c][1];
          }
          if (same_tp) {
            double* o = a.ty + (size_t)row * a.ty_stride + col;
            if (col < N) o[0] = acc[c][0] - mu;
            if (col + 1 < N) o[1] = acc[c][1] - mu;
          }
        }
      }
    }
  });
  __syncthreads();
  // ---- R = Y - Q Z over this slice's rows of p; summed squares in a fixed order
  double ss = 0.0;
  {
    const int nrb = (p + 7) >> 3;
    const int rb_lo = (int)((long long)nrb * blockIdx.y / a.psplit), rb_hi = (int)((long long)nrb * (blockIdx.y + 1) / a.psplit);
    gemm_rows<NCB, NCG>(a.Q, p, 0, p, rb_lo, rb_hi, Zs, LD, m, scratch, [&](int rb, int cb0, double (&acc)[NCB][2]) {
      const int row = rb * 8 + r;
      if (row < p) {
#pragma unroll
        for (int c = 0; c < NCB; ++c) {
          const int lc = (cb0 + c) * 8 + 2 * q, col = nb0 + lc;
          const double r0 = Ys[(size_t)row * LD + lc] - acc[c][0];      // columns beyond N hold Y = 0 and Z = 0: r = 0
          const double r1 = Ys[(size_t)row * LD + lc + 1] - acc[c][1];
          ss = fma(r0, r0, ss);
          ss = fma(r1, r1, ss);
          if (a.resid_out) {
            if (col < N) a.resid_out[(size_t)row * N + col] = r0;
            if (col + 1 < N) a.resid_out[(size_t)row * N + col + 1] = r1;
          }
        }
      }
    });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    a.resid_partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
  }
}

static int g_project_impl = 1;  // 1: DMMA kernel (default); 0: round-1 register-tiled scalar-FMA kernel
void set_project_impl(int v) { g_project_impl = v; }

// Upper bound of the number of residual partial sums launch_project writes for N columns (callers size and zero the
// buffer with it, then sum that many entries).
int project_max_partials(int N) { return ((N + 7) / 8) * 4; }

template <int NB>
static cudaError_t launch_project_dmma_nb(cudaStream_t st, const ProjArgs& a, int nblocks, size_t smem) {
  if (smem + 1024 > 48 * 1024) {  // static shared memory counts against the 48 KB default limit too
    cudaError_t e = cudaFuncSetAttribute(project_dmma_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)nblocks, (unsigned)a.psplit);
  project_dmma_kernel<NB><<<grid, 256, smem, st>>>(a);
  return cudaGetLastError();
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
  if (g_project_impl == 1) {
    const int p4 = (p + 3) & ~3, m4 = (m + 3) & ~3;
    int num_sms = 148;
    {
      static int cached_sms[64] = {0};
      int dev = 0;
      cudaGetDevice(&dev);
      if (cached_sms[dev & 63] == 0) cudaDeviceGetAttribute(&cached_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
      if (cached_sms[dev & 63] > 0) num_sms = cached_sms[dev & 63];
    }
    // the widest column block that fits in shared memory and still gives every SM a CTA; else the narrowest that fits
    int nb = 0;
    for (int c : {64, 32, 16, 8}) {
      const size_t sm_c = ((size_t)(p4 + m4) * (c >= 16 ? c + 8 : c) + PROJ_SCRATCH) * sizeof(double);
      if (sm_c > 200 * 1024) continue;
      nb = c;
      if ((N + c - 1) / c >= num_sms) break;
    }
    if (nb != 0) {
      const int nblocks = (N + nb - 1) / nb;
      ProjArgs a{y, N, p, T, m, lat0, mloc, means, ty, ty_stride, P, Q, resid_partial, resid_out, z_out, 1};
      if (P != nullptr && nblocks < num_sms) {  // small N: split the p rows of the residual over up to 4 CTAs per column block
        int s = num_sms / nblocks;
        const int nrb = (p + 7) / 8;
        if (s > 4) s = 4;
        if (s > nrb) s = nrb;
        if (s >= 1) a.psplit = s;
      }
      if (nblocks_out) *nblocks_out = nblocks * a.psplit;
      const size_t smem = ((size_t)(p4 + m4) * (nb >= 16 ? nb + 8 : nb) + PROJ_SCRATCH) * sizeof(double);
      switch (nb) {
        case 64: return launch_project_dmma_nb<64>(st, a, nblocks, smem);
        case 32: return launch_project_dmma_nb<32>(st, a, nblocks, smem);
        case 16: return launch_project_dmma_nb<16>(st, a, nblocks, smem);
        default: return launch_project_dmma_nb<8>(st, a, nblocks, smem);
      }
    }
    // p + m too large for one column block of 8 in shared memory: fall through to the register-tiled kernel's own check
  }
  const int nbcols = project_block_cols(p, m);
  if (nbcols == 0) return cudaErrorInvalidValue;
  const int nblocks = (N + nbcols - 1) / nbcols;
  if (nblocks_out) *nblocks_out = nblocks;
  const size_t smem = (size_t)(p + m) * nbcols * sizeof(double);
  cudaError_t e = cudaSuccess;
#define LMM_LAUNCH_PROJECT(NT_)                                                                                          \
  do {                                                                                                                    \
    if (smem + 4096 > 48 * 1024) e = cudaFuncSetAttribute(project_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
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
