#!/usr/bin/env python
"""Lane-level NumPy model of `project_dmma_kernel` (csrc/proj.cu): the same task split over warps, the same per-lane
addresses for the A / B fragments of mma.m8n8k4.f64 and the same epilogue index arithmetic, executed on the CPU -- checks a
change of the kernel's index maps without a GPU (tests/test_project_dmma_model.py).

mma.m8n8k4 (row.col) fragment ownership: lane = 4 r + k holds A[r][k]; lane = 4 n + k holds B[k][n]; lane = 4 r + q holds
C[r][2q], C[r][2q + 1]."""
import numpy as np


def dmma(acc, a_lane, b_lane):
    """acc: (32, 2) per-lane accumulators of one 8 x 8 block; a_lane, b_lane: (32,) per-lane operand values."""
    A = a_lane.reshape(8, 4)          # [r][k]
    B = b_lane.reshape(8, 4).T        # lane = 4 n + k -> B[k][n]
    C = A @ B                         # 8 x 8
    for lane in range(32):
        r, q = lane >> 2, lane & 3
        acc[lane, 0] += C[r, 2 * q]
        acc[lane, 1] += C[r, 2 * q + 1]


def warp_gemm(A, lda, a_row0, rows, rb, Bs, LD, K, cb0, NCB):
    """A: flat column-major array; Bs: flat shared array.  Returns acc[NCB][32][2]."""
    acc = np.zeros((NCB, 32, 2))
    lanes = np.arange(32)
    r, k = lanes >> 2, lanes & 3
    row = rb * 8 + r
    rok = row < rows
    ksteps = (K + 3) >> 2
    for ks in range(ksteps):
        kk = ks * 4 + k
        ok = rok & (kk < K)
        a = np.where(ok, A[np.where(ok, a_row0 + row + kk * lda, 0)], 0.0)
        for c in range(NCB):
            b = Bs[(ks * 4 + k) * LD + (cb0 + c) * 8 + r]
            dmma(acc[c], a, b)
    return acc


def project_block(y, N, p, T, m, lat0, mloc, means, P, Q, NB, bx, by, psplit, ty, z_out=None, resid_out=None):
    """One CTA (bx, by).  y: flat p*N; T, P: flat m*p column-major; Q: flat p*m column-major.  Writes ty [mloc][N]; returns ss."""
    LD = NB + 8 if NB >= 16 else NB
    NCBT = NB // 8
    NCB = 4 if NCBT >= 4 else NCBT
    NCG = NCBT // NCB
    p4, m4 = (p + 3) & ~3, (m + 3) & ~3
    Ys = np.full(p4 * LD, np.nan)
    Zs = np.full(m4 * LD, np.nan)
    nb0 = bx * NB
    for idx in range(p4 * NB):
        j, nn = idx // NB, idx % NB
        Ys[j * LD + nn] = y[j * N + nb0 + nn] if (j < p and nb0 + nn < N) else 0.0
    for idx in range((m4 - m) * NB):
        Zs[(m + idx // NB) * LD + idx % NB] = 0.0
    lanes = np.arange(32)
    r, q = lanes >> 2, lanes & 3
    same_tp = P is not None and P is T and lat0 == 0 and mloc == m
    if by == 0 and not same_tp:
        nrb = (mloc + 7) >> 3
        for task in range(nrb * NCG):
            rb, cb0 = task // NCG, (task % NCG) * NCB
            acc = warp_gemm(T, m, lat0, mloc, rb, Ys, LD, p, cb0, NCB)
            for lane in range(32):
                row = rb * 8 + r[lane]
                if row < mloc:
                    for c in range(NCB):
                        col = nb0 + (cb0 + c) * 8 + 2 * q[lane]
                        if col < N:
                            ty[row, col] = acc[c, lane, 0] - means[row]
                        if col + 1 < N:
                            ty[row, col + 1] = acc[c, lane, 1] - means[row]
    if P is None:
        return None
    nrb = (m + 7) >> 3
    for task in range(nrb * NCG):
        rb, cb0 = task // NCG, (task % NCG) * NCB
        acc = warp_gemm(P, m, 0, m, rb, Ys, LD, p, cb0, NCB)
        for lane in range(32):
            row = rb * 8 + r[lane]
            if row < m:
                for c in range(NCB):
                    lc = (cb0 + c) * 8 + 2 * q[lane]
                    col = nb0 + lc
                    Zs[row * LD + lc] = acc[c, lane, 0]
                    Zs[row * LD + lc + 1] = acc[c, lane, 1]
                    if by == 0:
                        if z_out is not None:
                            if col < N:
                                z_out[row, col] = acc[c, lane, 0]
                            if col + 1 < N:
                                z_out[row, col + 1] = acc[c, lane, 1]
                        if same_tp:
                            if col < N:
                                ty[row, col] = acc[c, lane, 0] - means[row]
                            if col + 1 < N:
                                ty[row, col + 1] = acc[c, lane, 1] - means[row]
    ss = 0.0
    nrb = (p + 7) >> 3
    rb_lo, rb_hi = nrb * by // psplit, nrb * (by + 1) // psplit
    for task in range((rb_hi - rb_lo) * NCG):
        rb, cb0 = rb_lo + task // NCG, (task % NCG) * NCB
        acc = warp_gemm(Q, p, 0, p, rb, Zs, LD, m, cb0, NCB)
        for lane in range(32):
            row = rb * 8 + r[lane]
            if row < p:
                for c in range(NCB):
                    lc = (cb0 + c) * 8 + 2 * q[lane]
                    col = nb0 + lc
                    r0 = Ys[row * LD + lc] - acc[c, lane, 0]
                    r1 = Ys[row * LD + lc + 1] - acc[c, lane, 1]
                    ss += r0 * r0 + r1 * r1
                    if resid_out is not None:
                        if col < N:
                            resid_out[row, col] = r0
                        if col + 1 < N:
                            resid_out[row, col + 1] = r1
    return ss


def project(Y, T, lat0, mloc, means, P, Q, NB, psplit):
    """Whole grid.  Y: p x N; T, P: m x p; Q: p x m (2-D arrays).  Returns (Ty [mloc x N], resid ss, Z, R)."""
    p, N = Y.shape
    m = T.shape[0]
    y = Y.reshape(-1)
    Tf = np.asfortranarray(T).reshape(-1, order="F")
    Pf = Tf if P is T else (None if P is None else np.asfortranarray(P).reshape(-1, order="F"))
    Qf = None if Q is None else np.asfortranarray(Q).reshape(-1, order="F")
    ty = np.full((mloc, N), np.nan)
    z = np.full((m, N), np.nan)
    R = np.full((p, N), np.nan)
    total = 0.0
    nblocks = (N + NB - 1) // NB
    for by in range(psplit):
        for bx in range(nblocks):
            ss = project_block(y, N, p, Tf, m, lat0, mloc, means, Pf, Qf, NB, bx, by, psplit, ty, z, R)
            if ss is not None:
                total += ss
    return ty, total, z, R
