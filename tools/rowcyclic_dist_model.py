#!/usr/bin/env python
"""NumPy model of the distributed-storage row-cyclic factorisation (csrc/host_chol.cu: chol_factor_rowcyclic_dist; index map
common.cuh: cyc_tile_index): every rank holds only the tile rows I = r (mod G), cyclic-packed, plus the window of the current block
column; per block column: look-ahead update of the own rows from the previous window, exchange (all-gather) into the window, panel on
the window by every rank, own rows of the window back home, trailing update of the own rows from the window.  The right-hand side rides
along as one extra tile row, so z = L^{-1} rhs falls out.  Same schedule and the same index arithmetic as the CUDA host code, tiles of
any size T; the exchange is a callable (torch.distributed all_gather under gloo in tests/test_distributed_cpu.py, a loop over simulated
ranks in the single-process test)."""
import numpy as np


def cyc_tile_index(I, J, G, r):
    l = (I - r) // G
    return l * (r + 1) + G * (l * (l - 1) // 2) + J


def cyc_tiles(nrows, G, r):
    if nrows <= r:
        return 0
    cnt = (nrows - 1 - r) // G + 1
    return cyc_tile_index(r + cnt * G, 0, G, r)


class RankState:
    """What ONE rank stores: its own tile rows of the (nc + 1)-row matrix (last row = right-hand side), packed."""

    def __init__(self, A, rhs, T, G, r):
        n = A.shape[0]
        assert n % T == 0
        self.T, self.G, self.r, self.nc, self.nrows = T, G, r, n // T, n // T + 1
        self.store = np.full((cyc_tiles(self.nrows, G, r), T, T), np.nan)
        for I in range(r, self.nrows, G):
            for J in range(min(I, self.nc - 1) + 1):
                if I < self.nc:
                    self.store[cyc_tile_index(I, J, G, r)] = A[I * T:(I + 1) * T, J * T:(J + 1) * T]
                else:  # the right-hand side in row 0 of the extra tile row
                    t = np.zeros((T, T))
                    t[0] = rhs[J * T:(J + 1) * T]
                    self.store[cyc_tile_index(I, J, G, r)] = t
        self.logdet = 0.0
        self.z = np.zeros(n)

    def tile(self, I, J):
        assert I % self.G == self.r
        return self.store[cyc_tile_index(I, J, self.G, self.r)]

    def first_own(self, s):
        return s + ((self.r - s % self.G) % self.G + self.G) % self.G

    def pack(self, s0, ob):
        """own rows >= s0 of the block column [s0, s0 + ob): [slot][ob] tiles (tiles above the diagonal stay NaN: never read)."""
        slots = (self.nrows - s0 + self.G - 1) // self.G
        buf = np.full((slots, ob, self.T, self.T), np.nan)
        for q in range(slots):
            I = self.first_own(s0) + q * self.G
            if I >= self.nrows:
                continue
            for c in range(ob):
                if s0 + c <= I and s0 + c < self.nc:
                    buf[q, c] = self.tile(I, s0 + c)
        return buf


def unpack(gathered, s0, ob, nrows, G, T):
    """all ranks' packed rows -> the window {(I, c)}: rows >= s0, columns s0 .. s0 + ob - 1."""
    win = {}
    for r in range(G):
        first = s0 + ((r - s0 % G) % G + G) % G
        for q in range(gathered[r].shape[0]):
            I = first + q * G
            if I >= nrows:
                continue
            for c in range(ob):
                if s0 + c <= I:
                    win[(I, s0 + c)] = gathered[r][q, c].copy()
    return win


def factor(states, ob, exchange):
    """states: the RankState objects this process simulates (all ranks in the single-process test, one under gloo);
    exchange(list of packed buffers of those states) -> for each of them the list of all G ranks' buffers."""
    st0 = states[0]
    T, G, nc, nrows = st0.T, st0.G, st0.nc, st0.nrows
    prev = [None] * len(states)
    for s0 in range(0, nc, ob):
        s1 = min(s0 + ob, nc)
        for si, st in enumerate(states):  # look-ahead update of the own rows of this block column from the previous window
            if prev[si] is not None:
                pw, p0, p1 = prev[si]
                for I in range(st.first_own(s0), nrows, G):
                    for J in range(s0, min(s1, I + 1)):
                        t = st.tile(I, J)
                        for k in range(p0, p1):
                            t -= pw[(I, k)] @ pw[(J, k)].T
        gathered = exchange([st.pack(s0, ob) for st in states])
        for si, st in enumerate(states):
            win = unpack(gathered[si], s0, ob, nrows, G, T)
            for jj in range(s0, s1):  # the panel, on the window: every rank, all rows
                for I in range(jj, nrows):
                    for k in range(s0, jj):
                        win[(I, jj)] -= win[(I, k)] @ win[(jj, k)].T
                Ljj = np.linalg.cholesky(win[(jj, jj)])
                win[(jj, jj)] = Ljj
                st.logdet += 2.0 * np.sum(np.log(np.diag(Ljj)))
                Winv = np.linalg.inv(Ljj)
                for I in range(jj + 1, nrows):
                    win[(I, jj)] = win[(I, jj)] @ Winv.T
            for I in range(st.first_own(s0), nrows, G):  # the finished rows go home
                for J in range(s0, min(s1, I + 1)):
                    st.tile(I, J)[...] = win[(I, J)]
            for J in range(s0, s1):  # z from the extra tile row
                st.z[J * T:(J + 1) * T] = win[(nc, J)][0]
            s2 = s1 + ob  # trailing update of the own rows right of the NEXT block column, operands from the window
            for I in range(st.first_own(s2), nrows, G) if s2 < nc else []:
                for J in range(s2, min(nc, I + 1)):
                    t = st.tile(I, J)
                    for k in range(s0, s1):
                        t -= win[(I, k)] @ win[(J, k)].T
            prev[si] = (win, s0, s1)
    return states
