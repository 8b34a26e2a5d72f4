// Dense ILMM assembly (K14/K15 of SURVEY.md §2.3): the (mN x mN) projected covariance
// blockdiag(K_a) + ΣT ⊗ I_N of src/ilmm.jl:160-162 and the (pN x pN) form H K H' + σ²I
// (test/ilmm.jl:5, src/ilmm.jl:136) are written tile by tile straight into the Cholesky
// workspace: no kron(), no BlockDiagonal -> Matrix copy, K_a entries are recomputed in registers.
// Also the block-diagonal cross-covariance K(x*, x) of the joint ILMM posterior and the per-point
// Gram reduction its predictive variance needs.
#include "common.cuh"
#include "kernels.h"

namespace lmm {

__device__ __forceinline__ double kernel_pair(const LatentParams& lp, const double* __restrict__ xa, const double* __restrict__ xb,
                                              int D, int form, bool same_point) {
  if (needs_raw_points(&lp)) return kernel_value_raw(&lp, xa, xb, D, form, same_point);  // composite / periodic latent
  if (same_point) return kappa_eval(lp.kind, lp.variance, 0.0, lp.param);
  double a[64], b[64];
  double sa = 0.0, sb = 0.0;
  for (int k = 0; k < D; ++k) {
    const double s = input_scale(&lp, k);
    a[k] = xa[k] * s;
    b[k] = xb[k] * s;
    sa = fma(a[k], a[k], sa);
    sb = fma(b[k], b[k], sb);
  }
  return kappa_eval(lp.kind, lp.variance, sqdist(a, b, D, sa, sb, form), lp.param);
}

// mode 0 (projected): dim = m*N, val = [a==b] k_a(i,j) + E[a,b] [i==j]      (E = ΣT, m x m col-major)
// mode 1 (dense):     dim = q*N, val = sum_l Hm[a,l] Hm[b,l] k_l(i,j) + E0 [a==b][i==j]   (Hm: q x m col-major)
// mode 2 (dense Σy):  dim = m*N, val = [a==b] k_a(i,j) + E[gr, gc]          (E = Σy, dim x dim col-major)
// mode 3 (per point): dim = m*N, val = [a==b] k_a(i,j) + E[i][a,b] [i==j]   (E: N blocks of m x m col-major)
__global__ void __launch_bounds__(256) assemble_ilmm_kernel(TiledSym out, const double* __restrict__ x, int N, int D,
                                                            const LatentParams* __restrict__ params, int m, int q,
                                                            const double* __restrict__ E, const double* __restrict__ Hm, int mode,
                                                            int form) {
  int I, J;
  if (out.cyc_G) {  // distributed storage: grid (tile columns, own rows); this rank builds the tile rows it owns
    I = out.cyc_r + (int)blockIdx.y * out.cyc_G;
    J = blockIdx.x;
    if (J > I) return;
  } else {
    const int tl = blockIdx.x;
    I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
    while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
    while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
    J = tl - (int)((size_t)I * (I + 1) / 2);
  }
  const int dim = q * N;
  double* tile = out.tile(0, I, J);
  for (int e = threadIdx.x; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    const int gr = I * TILE + r, gc = J * TILE + c;
    double val;
    if (gr >= dim || gc >= dim) {
      val = (gr == gc) ? 1.0 : 0.0;
    } else {
      const int a = gr / N, i = gr % N, b = gc / N, j = gc % N;
      const double* xi = x + (size_t)i * D;
      const double* xj = x + (size_t)j * D;
      if (mode == 0 || mode == 3) {
        val = (a == b) ? kernel_pair(params[a], xi, xj, D, form, i == j) : 0.0;
        if (i == j) val += E[(mode == 3 ? (size_t)i * m * m : 0) + (size_t)b * m + a];
      } else if (mode == 2) {
        val = ((a == b) ? kernel_pair(params[a], xi, xj, D, form, i == j) : 0.0) + E[(size_t)gc * dim + gr];
      } else {
        val = 0.0;
        for (int l = 0; l < m; ++l)
          val = fma(Hm[(size_t)l * q + a] * Hm[(size_t)l * q + b], kernel_pair(params[l], xi, xj, D, form, i == j), val);
        if (gr == gc) val += E[0];
      }
    }
    tile[e] = val;
  }
}

cudaError_t launch_assemble_ilmm(cudaStream_t st, TiledSym out, const double* x, int N, int D, const LatentParams* params, int m,
                                 int q, const double* E, const double* Hm, int mode, int form) {
  if (out.cyc_G) {  // out.nt = tile rows of the matrix proper (a right-hand-side row below it is written by launch_rhs_row)
    if (out.nt <= out.cyc_r) return cudaSuccess;
    dim3 grid((unsigned)out.nt, (unsigned)((out.nt - 1 - out.cyc_r) / out.cyc_G + 1));
    assemble_ilmm_kernel<<<grid, 256, 0, st>>>(out, x, N, D, params, m, q, E, Hm, mode, form);
    return cudaGetLastError();
  }
  assemble_ilmm_kernel<<<(unsigned)sym_tiles(out.nt), 256, 0, st>>>(out, x, N, D, params, m, q, E, Hm, mode, form);
  return cudaGetLastError();
}

// Block-diagonal cross-covariance: rows (a, n) over x* (m*Ns), cols (b, i) over x (m*N).
__global__ void __launch_bounds__(256) assemble_cross_blockdiag_kernel(TiledRect out, const double* __restrict__ xs, int Ns,
                                                                       const double* __restrict__ x, int N, int D,
                                                                       const LatentParams* __restrict__ params, int m, int form) {
  const int R = blockIdx.x / out.ntc, J = blockIdx.x % out.ntc;
  double* tile = out.tile(0, R, J);
  for (int e = threadIdx.x; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    const int gr = R * TILE + r, gc = J * TILE + c;
    double val = 0.0;
    if (gr < m * Ns && gc < m * N) {
      const int a = gr / Ns, n = gr % Ns, b = gc / N, i = gc % N;
      if (a == b) val = kernel_pair(params[a], xs + (size_t)n * D, x + (size_t)i * D, D, form, false);
    }
    tile[e] = val;
  }
}
cudaError_t launch_assemble_cross_blockdiag(cudaStream_t st, TiledRect out, const double* xs, int Ns, const double* x, int N, int D,
                                            const LatentParams* params, int m, int form) {
  assemble_cross_blockdiag_kernel<<<(unsigned)(out.ntr * out.ntc), 256, 0, st>>>(out, xs, Ns, x, N, D, params, m, form);
  return cudaGetLastError();
}

// ILMM predictive marginals from the joint latent posterior.  V = Kc L^{-T} (rows (a,n)).
// grid (Ns).  mean[j*Ns+n] = sum_a H[j,a] (mean_a + mlat[a*Ns+n]);
// var[j*Ns+n] = sum_ab H[j,a] H[j,b] ( [a==b](variance_a + 1e-18) - sum_r V[(a,n),r] V[(b,n),r] ) + sigma2
__global__ void __launch_bounds__(256) ilmm_predict_kernel(TiledRect V, int Ns, int m, int p, const double* __restrict__ H,
                                                           const LatentParams* __restrict__ params,
                                                           const double* __restrict__ mlat, double sigma2,
                                                           double* __restrict__ mean, double* __restrict__ var) {
  extern __shared__ double G[];  // m*m
  __shared__ double red[256];
  const int n = blockIdx.x, t = threadIdx.x;
  const int ncols = V.ntc * TILE;
  for (int a = 0; a < m; ++a) {
    for (int b = 0; b <= a; ++b) {
      const int ra = a * Ns + n, rb = b * Ns + n;
      const double* ta = V.base + (size_t)(ra / TILE) * V.ntc * TT;
      const double* tb = V.base + (size_t)(rb / TILE) * V.ntc * TT;
      double s = 0.0;
      for (int c = t; c < ncols; c += 256) {
        const size_t off = (size_t)(c / TILE) * TT;
        s = fma(ta[off + tile_elem(ra % TILE, c % TILE)], tb[off + tile_elem(rb % TILE, c % TILE)], s);
      }
      red[t] = s;
      __syncthreads();
      for (int w = 128; w > 0; w >>= 1) {
        if (t < w) red[t] += red[t + w];
        __syncthreads();
      }
      if (t == 0) {
        G[a * m + b] = red[0];
        G[b * m + a] = red[0];
      }
      __syncthreads();
    }
  }
  for (int j = t; j < p; j += 256) {
    double sm_ = 0.0, sv = 0.0;
    for (int a = 0; a < m; ++a) {
      const double ha = H[(size_t)a * p + j];
      sm_ = fma(ha, params[a].mean + mlat[(size_t)a * Ns + n], sm_);
      for (int b = 0; b < m; ++b) {
        const double c = ((a == b) ? (params[a].kdiag + 1e-18) : 0.0) - G[a * m + b];
        sv = fma(ha * H[(size_t)b * p + j], c, sv);
      }
    }
    mean[(size_t)j * Ns + n] = sm_;
    var[(size_t)j * Ns + n] = sv + sigma2;
  }
}
cudaError_t launch_ilmm_predict(cudaStream_t st, TiledRect V, int Ns, int m, int p, const double* H, const LatentParams* params,
                                const double* mlat, double sigma2, double* mean, double* var) {
  ilmm_predict_kernel<<<Ns, 256, (size_t)m * m * sizeof(double), st>>>(V, Ns, m, p, H, params, mlat, sigma2, mean, var);
  return cudaGetLastError();
}

// Heterotopic / missing-data dense model: rows and columns run over the OBSERVED entries obs[k] = j*N + i (output j at
// input i, by outputs) of y ~ N((H ⊗ I) m, Σ_l (h_l h_l') ⊗ K_l + σ² I).   val = Σ_l H[a,l] H[b,l] k_l(x_i, x_j) + σ² [r == c].
__global__ void __launch_bounds__(256) assemble_masked_kernel(TiledSym out, const int* __restrict__ obs, int nobs,
                                                              const double* __restrict__ x, int N, int D,
                                                              const LatentParams* __restrict__ params, int m, int p,
                                                              const double* __restrict__ Hm, double sigma2, int form) {
  const int tl = blockIdx.x;
  int I = (int)((sqrt(8.0 * (double)tl + 1.0) - 1.0) * 0.5);
  while ((size_t)(I + 1) * (I + 2) / 2 <= (size_t)tl) ++I;
  while ((size_t)I * (I + 1) / 2 > (size_t)tl) --I;
  const int J = tl - (int)((size_t)I * (I + 1) / 2);
  double* tile = out.tile(0, I, J);
  for (int e = threadIdx.x; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    const int gr = I * TILE + r, gc = J * TILE + c;
    double val;
    if (gr >= nobs || gc >= nobs) {
      val = (gr == gc) ? 1.0 : 0.0;
    } else {
      const int oa = obs[gr], ob = obs[gc];
      const int a = oa / N, i = oa % N, b = ob / N, j = ob % N;
      val = 0.0;
      for (int l = 0; l < m; ++l)
        val = fma(Hm[(size_t)l * p + a] * Hm[(size_t)l * p + b], kernel_pair(params[l], x + (size_t)i * D, x + (size_t)j * D, D, form, i == j), val);
      if (gr == gc) val += sigma2;
    }
    tile[e] = val;
  }
}
cudaError_t launch_assemble_masked(cudaStream_t st, TiledSym out, const int* obs, int nobs, const double* x, int N, int D,
                                   const LatentParams* params, int m, int p, const double* Hm, double sigma2, int form) {
  assemble_masked_kernel<<<(unsigned)sym_tiles(out.nt), 256, 0, st>>>(out, obs, nobs, x, N, D, params, m, p, Hm, sigma2, form);
  return cudaGetLastError();
}

// Cross covariance of the dense model between ALL outputs at x* (rows (j, n), by outputs, p*Ns) and the observed entries
// (columns k over obs): Σ_l H[j,l] H[b_k,l] k_l(x*_n, x_{i_k}).
__global__ void __launch_bounds__(256) assemble_masked_cross_kernel(TiledRect out, const double* __restrict__ xs, int Ns,
                                                                    const int* __restrict__ obs, int nobs, const double* __restrict__ x,
                                                                    int N, int D, const LatentParams* __restrict__ params, int m, int p,
                                                                    const double* __restrict__ Hm, int form) {
  const int Rt = blockIdx.x / out.ntc, Jt = blockIdx.x % out.ntc;
  double* tile = out.tile(0, Rt, Jt);
  for (int e = threadIdx.x; e < TT; e += 256) {
    int r, c;
    tile_rc(e, r, c);
    const int gr = Rt * TILE + r, gc = Jt * TILE + c;
    double val = 0.0;
    if (gr < p * Ns && gc < nobs) {
      const int j = gr / Ns, n = gr % Ns, ob = obs[gc], b = ob / N, i = ob % N;
      for (int l = 0; l < m; ++l)
        val = fma(Hm[(size_t)l * p + j] * Hm[(size_t)l * p + b], kernel_pair(params[l], xs + (size_t)n * D, x + (size_t)i * D, D, form, false), val);
    }
    tile[e] = val;
  }
}
cudaError_t launch_assemble_masked_cross(cudaStream_t st, TiledRect out, const double* xs, int Ns, const int* obs, int nobs, const double* x,
                                         int N, int D, const LatentParams* params, int m, int p, const double* Hm, int form) {
  assemble_masked_cross_kernel<<<(unsigned)(out.ntr * out.ntc), 256, 0, st>>>(out, xs, Ns, obs, nobs, x, N, D, params, m, p, Hm, form);
  return cudaGetLastError();
}

// Gather rows: dst[u][:] = src[idx[u]][:]  (delta replication for the hyper-parameter sweep)
__global__ void gather_rows_kernel(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ idx, size_t stride) {
  const int u = blockIdx.y;
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < stride) dst[(size_t)u * stride + k] = src[(size_t)idx[u] * stride + k];
}
cudaError_t launch_gather_rows(cudaStream_t st, double* dst, const double* src, const int* idx, size_t stride, int n) {
  dim3 grid((unsigned)((stride + 255) / 256), (unsigned)n);
  gather_rows_kernel<<<grid, 256, 0, st>>>(dst, src, idx, stride);
  return cudaGetLastError();
}

}  // namespace lmm

namespace lmm {

// Dense output covariance from independent latent covariances (TiledSym C, one per resident latent):
//   out[(j,n) + (j2,n2)*pNs] = sum_a H[j,a] H[j2,a] C_a(n,n2)      (the caller adds σ² on the diagonal
//   after the cross-rank reduction).  src/ilmm.jl:132-139 without kron(H, I): C = (H⊗I) C_lat (H⊗I)'.
// grid (ceil(pNs/16), ceil(pNs/16)), block 16x16.
__global__ void __launch_bounds__(256) mix_cov_kernel(TiledSym C, int nloc, int lat0, const double* __restrict__ H, int p, int Ns,
                                                      double* __restrict__ out) {
  const int dim = p * Ns;
  const int row = blockIdx.x * 16 + threadIdx.x, col = blockIdx.y * 16 + threadIdx.y;
  if (row >= dim || col >= dim) return;
  const int j = row / Ns, n = row % Ns, j2 = col / Ns, n2 = col % Ns;
  const int hi = n >= n2 ? n : n2, lo = n >= n2 ? n2 : n;
  const size_t off = sym_tile_index(hi / TILE, lo / TILE) * TT + tile_elem(hi % TILE, lo % TILE);
  double s = 0.0;
  for (int a = 0; a < nloc; ++a)
    s = fma(H[(size_t)(lat0 + a) * p + j] * H[(size_t)(lat0 + a) * p + j2], C.base[(size_t)a * C.batch_stride + off], s);
  out[(size_t)col * dim + row] = s;
}
cudaError_t launch_mix_cov(cudaStream_t st, TiledSym C, int nloc, int lat0, const double* H, int p, int Ns, double* out) {
  const int dim = p * Ns;
  dim3 grid((unsigned)((dim + 15) / 16), (unsigned)((dim + 15) / 16)), block(16, 16);
  mix_cov_kernel<<<grid, block, 0, st>>>(C, nloc, lat0, H, p, Ns, out);
  return cudaGetLastError();
}

// Same from a joint latent covariance Cl (TiledSym over m*Ns, batch 1):
//   out[(j,n),(j2,n2)] = sum_{a,b} H[j,a] H[j2,b] Cl[(a,n),(b,n2)]
__global__ void __launch_bounds__(256) mix_cov_joint_kernel(TiledSym Cl, int m, const double* __restrict__ H, int p, int Ns,
                                                            double* __restrict__ out) {
  const int dim = p * Ns;
  const int row = blockIdx.x * 16 + threadIdx.x, col = blockIdx.y * 16 + threadIdx.y;
  if (row >= dim || col >= dim) return;
  const int j = row / Ns, n = row % Ns, j2 = col / Ns, n2 = col % Ns;
  double s = 0.0;
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      const int r = a * Ns + n, c = b * Ns + n2;
      const int hi = r >= c ? r : c, lo = r >= c ? c : r;
      const double v = Cl.base[sym_tile_index(hi / TILE, lo / TILE) * TT + tile_elem(hi % TILE, lo % TILE)];
      s = fma(H[(size_t)a * p + j] * H[(size_t)b * p + j2], v, s);
    }
  out[(size_t)col * dim + row] = s;
}
cudaError_t launch_mix_cov_joint(cudaStream_t st, TiledSym Cl, int m, const double* H, int p, int Ns, double* out) {
  const int dim = p * Ns;
  dim3 grid((unsigned)((dim + 15) / 16), (unsigned)((dim + 15) / 16)), block(16, 16);
  mix_cov_joint_kernel<<<grid, block, 0, st>>>(Cl, m, H, p, Ns, out);
  return cudaGetLastError();
}

// out[i + i*dim] += s
__global__ void add_diag_kernel(double* out, int dim, double s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) out[(size_t)i * dim + i] += s;
}
cudaError_t launch_add_diag(cudaStream_t st, double* out, int dim, double s) {
  add_diag_kernel<<<(dim + 255) / 256, 256, 0, st>>>(out, dim, s);
  return cudaGetLastError();
}

}  // namespace lmm
