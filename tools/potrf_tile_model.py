#!/usr/bin/env python
"""Lane-level NumPy model of `potrf_tile_kernel2` (linearmixingmodels.jl_b200/csrc/potrf.cu): the same shared-memory indices,
the same m8n8k4 fragment maps (lane = 4 g + t: A(g, t), B(t, g), C(g, 2t), C(g, 2t+1)), the same work split over warps and the
same order of phases -- left-looking column updates with the k range split for single-block warps, reciprocal (LDL^T-style)
pivot chain of the 8x8 diagonal block, W = inv(L) built block row by block row in the transposed upper triangle, the
off-diagonal half W21 = -W22 (L21 W11) at the end.  It lets a change of the kernel's index arithmetic be checked on a machine
without a GPU (tests/test_potrf_tile_model.py); the upper triangle of the tile is poisoned with NaN to prove that nothing the
kernel reads lives there before it is written.  Not product code: the CUDA kernel is the implementation, this is its model."""
import numpy as np

LD, T = 132, 128


class TileModel:
    def __init__(self, A):
        self.S = np.zeros(T * LD)
        for c in range(T):
            self.S[c * LD:c * LD + T] = A[:, c]
        for r in range(T):  # poison: the kernel must never read the upper triangle of its input
            for c in range(r + 1, T):
                self.S[c * LD + r] = np.nan
        self.dinv = np.zeros(T)

    @staticmethod
    def dmma(c0, c1, a, b):
        """One warp-wide mma.m8n8k4.f64: lanes supply A(g, t) and B(t, g) and own C(g, 2t), C(g, 2t + 1)."""
        Am, Bm = np.zeros((8, 4)), np.zeros((4, 8))
        for l in range(32):
            g, t = l >> 2, l & 3
            Am[g, t], Bm[t, g] = a[l], b[l]
        C = Am @ Bm
        for l in range(32):
            g, t = l >> 2, l & 3
            c0[l] += C[g, 2 * t]
            c1[l] += C[g, 2 * t + 1]

    def wdiag(self, o, rr, cc):
        """Element (rr, cc) of the lower-triangular 8x8 block W_oo in the mirrored storage."""
        if rr > cc:
            return self.S[(o + rr) * LD + o + cc]
        return self.dinv[o + rr] if rr == cc else 0.0

    def frag(self, fn):
        return np.array([fn(l >> 2, l & 3) for l in range(32)])

    # ---- left-looking update of column block bj (all 8 warps) ----
    def column_update(self, bj):
        S, j0 = self.S, 8 * bj
        for warp in range(8):
            bi0 = bj + warp
            if bi0 >= 16:
                continue
            two = bi0 + 8 < 16
            R0 = 8 * bi0
            R1 = R0 + 64 if two else R0
            trips = bj if two else (bj + 1) >> 1
            koff = 0 if two else 8 * trips
            klast = 8 * (bj - 1)
            p0, p1, q0, q1, u0, u1, v0, v1 = (np.zeros(32) for _ in range(8))
            for k in range(trips):
                ck = 8 * k
                c2 = min(ck + koff, klast)
                self.dmma(p0, p1, self.frag(lambda g, t: S[(ck + t) * LD + R0 + g]), self.frag(lambda g, t: S[(ck + t) * LD + j0 + g]))
                self.dmma(q0, q1, self.frag(lambda g, t: S[(ck + 4 + t) * LD + R0 + g]), self.frag(lambda g, t: S[(ck + 4 + t) * LD + j0 + g]))
                if two or 8 * k + koff <= klast:
                    self.dmma(u0, u1, self.frag(lambda g, t: S[(c2 + t) * LD + R1 + g]), self.frag(lambda g, t: S[(c2 + t) * LD + j0 + g]))
                    self.dmma(v0, v1, self.frag(lambda g, t: S[(c2 + 4 + t) * LD + R1 + g]),
                              self.frag(lambda g, t: S[(c2 + 4 + t) * LD + j0 + g]))
            for l in range(32):
                g, t = l >> 2, l & 3
                if two:
                    S[(j0 + 2 * t) * LD + R0 + g] -= p0[l] + q0[l]
                    S[(j0 + 2 * t + 1) * LD + R0 + g] -= p1[l] + q1[l]
                    S[(j0 + 2 * t) * LD + R1 + g] -= u0[l] + v0[l]
                    S[(j0 + 2 * t + 1) * LD + R1 + g] -= u1[l] + v1[l]
                else:
                    S[(j0 + 2 * t) * LD + R0 + g] -= (p0[l] + q0[l]) + (u0[l] + v0[l])
                    S[(j0 + 2 * t + 1) * LD + R0 + g] -= (p1[l] + q1[l]) + (u1[l] + v1[l])

    # ---- panel step of 8 columns: every row owner r >= j0 (threads 0..127) ----
    def panel_step(self, j0):
        S = self.S
        d0 = np.zeros((8, 8))
        for i in range(8):
            for k in range(i + 1):
                d0[i, k] = S[(j0 + k) * LD + j0 + i]
        out = {}
        for r in range(j0, T):
            d = d0.copy()
            p = np.array([S[(j0 + k) * LD + r] for k in range(8)])
            dv, tm = np.zeros(8), np.zeros((8, 8))
            for jj in range(8):
                piv = d[jj, jj]
                assert piv > 0.0
                rc = 1.0 / piv
                dv[jj] = 1.0 / np.sqrt(piv)
                for j2 in range(jj + 1, 8):
                    d[j2, j2] = d[j2, j2] - (d[j2, jj] * d[j2, jj]) * rc
                    tm[j2, jj] = d[j2, jj] * rc
                    for i in range(j2 + 1, 8):
                        d[i, j2] -= d[i, jj] * tm[j2, jj]
            for jj in range(8):
                v = p[jj]
                for k in range(jj):
                    v -= p[k] * tm[jj, k]
                p[jj] = v
            out[r] = (p * dv, dv)
        return out  # stored after the concurrent W block row has run: the two touch disjoint parts of the tile

    def panel_store(self, j0, out):
        for r in range(j0, T):
            for k in range(8):
                if k <= r - j0:
                    self.S[(j0 + k) * LD + r] = out[r][0][k]
            if r == j0:
                self.dinv[j0:j0 + 8] = out[r][1]

    # ---- block row i of W restricted to block columns [jmin, i), nw warps ----
    def w_block_row(self, i, jmin, nw):
        S, o = self.S, 8 * i
        for col in range(8):  # W_ii by substitution, one thread per column
            w = np.zeros(8)
            for rr in range(8):
                s = 1.0 if rr == col else 0.0
                for k in range(rr):
                    s -= S[(o + k) * LD + o + rr] * w[k]
                w[rr] = s * self.dinv[o + rr]
            for rr in range(1, 8):
                if rr > col:
                    S[(o + rr) * LD + o + col] = w[rr]
        for wl in range(nw):  # T_j = sum_{k=j}^{i-1} L_ik W_kj
            for j in range(jmin + wl, i, nw):
                oj = 8 * j
                c0, c1, e0, e1 = (np.zeros(32) for _ in range(4))
                self.dmma(c0, c1, self.frag(lambda g, t: S[(oj + t) * LD + o + g]), self.frag(lambda g, t: self.wdiag(oj, t, g)))
                self.dmma(e0, e1, self.frag(lambda g, t: S[(oj + 4 + t) * LD + o + g]), self.frag(lambda g, t: self.wdiag(oj, 4 + t, g)))
                for k in range(j + 1, i):
                    ok = 8 * k
                    self.dmma(c0, c1, self.frag(lambda g, t: S[(ok + t) * LD + o + g]), self.frag(lambda g, t: S[(ok + t) * LD + oj + g]))
                    self.dmma(e0, e1, self.frag(lambda g, t: S[(ok + 4 + t) * LD + o + g]),
                              self.frag(lambda g, t: S[(ok + 4 + t) * LD + oj + g]))
                for l in range(32):
                    g, t = l >> 2, l & 3
                    S[(o + g) * LD + oj + 2 * t] = c0[l] + e0[l]
                    S[(o + g) * LD + oj + 2 * t + 1] = c1[l] + e1[l]
        for wl in range(nw):  # W_ij = -W_ii T_j (after the group barrier)
            for j in range(jmin + wl, i, nw):
                oj = 8 * j
                c0, c1 = np.zeros(32), np.zeros(32)
                b0 = self.frag(lambda g, t: S[(o + t) * LD + oj + g])
                b1 = self.frag(lambda g, t: S[(o + 4 + t) * LD + oj + g])
                self.dmma(c0, c1, self.frag(lambda g, t: self.wdiag(o, g, t)), b0)
                self.dmma(c0, c1, self.frag(lambda g, t: self.wdiag(o, g, 4 + t)), b1)
                for l in range(32):
                    g, t = l >> 2, l & 3
                    S[(o + g) * LD + oj + 2 * t] = -c0[l]
                    S[(o + g) * LD + oj + 2 * t + 1] = -c1[l]

    # ---- W21 = -W22 (L21 W11): stage A by block rows, stage B by block columns ----
    def w_lower_left(self):
        S = self.S
        for warp in range(8):
            o = 64 + 8 * warp
            acc = [[np.zeros(32) for _ in range(4)] for _ in range(8)]
            for k in range(8):
                ok = 8 * k
                a0 = self.frag(lambda g, t: S[(ok + t) * LD + o + g])
                a1 = self.frag(lambda g, t: S[(ok + 4 + t) * LD + o + g])
                for j in range(k + 1):
                    oj = 8 * j
                    if j == k:
                        b0, b1 = self.frag(lambda g, t: self.wdiag(oj, t, g)), self.frag(lambda g, t: self.wdiag(oj, 4 + t, g))
                    else:
                        b0 = self.frag(lambda g, t: S[(ok + t) * LD + oj + g])
                        b1 = self.frag(lambda g, t: S[(ok + 4 + t) * LD + oj + g])
                    self.dmma(acc[j][0], acc[j][1], a0, b0)
                    self.dmma(acc[j][2], acc[j][3], a1, b1)
            for j in range(8):
                for l in range(32):
                    g, t = l >> 2, l & 3
                    S[(o + g) * LD + 8 * j + 2 * t] = acc[j][0][l] + acc[j][2][l]
                    S[(o + g) * LD + 8 * j + 2 * t + 1] = acc[j][1][l] + acc[j][3][l]
        results = []
        for warp in range(8):  # every warp computes before anyone stores (CTA barrier in the kernel)
            oj = 8 * warp
            acc = [[np.zeros(32) for _ in range(4)] for _ in range(8)]
            for kk in range(8):
                ok = 64 + 8 * kk
                b0 = self.frag(lambda g, t: S[(ok + t) * LD + oj + g])
                b1 = self.frag(lambda g, t: S[(ok + 4 + t) * LD + oj + g])
                for ii in range(kk, 8):
                    o = 64 + 8 * ii
                    if ii == kk:
                        a0, a1 = self.frag(lambda g, t: self.wdiag(o, g, t)), self.frag(lambda g, t: self.wdiag(o, g, 4 + t))
                    else:
                        a0 = self.frag(lambda g, t: S[(o + g) * LD + ok + t])
                        a1 = self.frag(lambda g, t: S[(o + g) * LD + ok + 4 + t])
                    self.dmma(acc[ii][0], acc[ii][1], a0, b0)
                    self.dmma(acc[ii][2], acc[ii][3], a1, b1)
            results.append(acc)
        for warp in range(8):
            oj = 8 * warp
            for ii in range(8):
                o = 64 + 8 * ii
                for l in range(32):
                    g, t = l >> 2, l & 3
                    S[(o + g) * LD + oj + 2 * t] = -(results[warp][ii][0][l] + results[warp][ii][2][l])
                    S[(o + g) * LD + oj + 2 * t + 1] = -(results[warp][ii][1][l] + results[warp][ii][3][l])

    def run(self):
        for j0 in range(0, T, 8):
            bj = j0 >> 3
            if bj > 0:
                self.column_update(bj)
            out = self.panel_step(j0)          # warps 0-3 ...
            if j0 > 0:                         # ... while warps 4-7 build W block row bj - 1 inside its 64x64 half
                self.w_block_row(bj - 1, 0 if bj - 1 < 8 else 8, 4)
            self.panel_store(j0, out)
        self.w_block_row(15, 8, 8)
        self.w_lower_left()
        L, W = np.zeros((T, T)), np.zeros((T, T))
        for r in range(T):
            for c in range(r + 1):
                L[r, c] = self.S[c * LD + r]
                W[r, c] = self.dinv[r] if r == c else self.S[r * LD + c]
        return L, W


def demo_matrix(seed=0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0, 3, T))
    return np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2) + 0.1 * np.eye(T)


if __name__ == "__main__":
    A = demo_matrix()
    L, W = TileModel(A).run()
    Lr = np.linalg.cholesky(A)
    print("max |L - chol(A)| =", np.abs(L - Lr).max(), " max |W L - I| =", np.abs(W @ Lr - np.eye(T)).max(),
          " NaN reached the outputs:", bool(np.isnan(L).any() or np.isnan(W).any()))
