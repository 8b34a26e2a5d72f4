"""CPU tests of the integer-slice (Ozaki) arithmetic through its NumPy model (tools/ozaki_model.py: the formulas of csrc/ozaki.cu): digit
ranges, exact extraction, int32 bound, truncation error of the recombined product.  The GPU kernels are tested in test_ozaki_gpu.py."""
import numpy as np
import pytest

from tools.ozaki_model import int32_bound_ok, product, reconstruct, row_scale, slice_planes


def cholesky_rows(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    d = np.exp(rng.uniform(-3, 3, n))
    A = (M @ M.T / n + np.eye(n)) * d[:, None] * d[None, :]
    return A, np.linalg.cholesky(A)


@pytest.mark.parametrize("S,bits", [(8, 7), (7, 7), (6, 7), (7, 8), (6, 8)])
def test_digit_planes_reconstruct_to_the_last_bit_kept(S, bits):
    A, L = cholesky_rows(192, seed=S + bits)
    sc = row_scale(np.diag(A))
    assert np.all(np.abs(L) <= sc[:, None] * 64.0 * (1 + 1e-12))  # |L(i,k)| <= sqrt(A_ii) < 2^E
    planes = slice_planes(L, sc, S, bits)
    assert np.abs(planes[0]).max() <= 65
    if bits == 7:
        assert np.abs(planes).max() <= 64
    err = np.abs(reconstruct(planes, sc, bits) - L) / sc[:, None]
    ulp = 0.5 * 2.0 ** (-bits * (S - 1))  # half a unit of the last plane, in units of the scale
    assert err.max() <= ulp * (1 + 1e-9), (err.max(), ulp)


@pytest.mark.parametrize("S,bits,tol", [(8, 7, 2e-14), (7, 8, 8e-14), (6, 7, 5e-10)])
def test_sliced_product_matches_fp64(S, bits, tol):
    """C = L1 L2' over K = 512 columns: error relative to |L1||L2|' -- the normwise bound an FP64 GEMM has (with a larger constant for
    fewer planes); every per-d sum stays below 2^31."""
    A, L = cholesky_rows(512, seed=3)
    sc = row_scale(np.diag(A))
    L1, L2 = L[256:384, :256], L[384:512, :256]
    p1, p2 = slice_planes(L1, sc[256:384], S, bits), slice_planes(L2, sc[384:512], S, bits)
    C = product(p1, sc[256:384], p2, sc[384:512], bits)
    ref = L1.astype(np.longdouble) @ L2.astype(np.longdouble).T
    bound = np.abs(L1) @ np.abs(L2).T
    rowprod = (sc[256:384, None] * 64) * (sc[None, 384:512] * 64)  # the fixed-point grid is relative to the row scales
    assert float(np.max(np.abs(C - ref) / rowprod)) < tol
    assert float(np.max(np.abs(C - ref) / bound)) < tol * 1e4  # and still tiny against |L1||L2|'


def test_int32_bound():
    assert int32_bound_ok(16256, 7, 8) and int32_bound_ok(16256, 8, 7)
    assert not int32_bound_ok(18725, 7, 8) and int32_bound_ok(18724, 7, 8)
    assert int32_bound_ok(65535, 8, 7) and not int32_bound_ok(65536, 8, 7)
    # worst case really fits: all digits at their extreme value
    for S, bits, K in ((7, 8, 18724), (8, 7, 65535)):
        q = 128 if bits == 8 else 64
        assert S * K * q * q < 2 ** 31


@pytest.mark.parametrize("S", [6, 7, 8])
def test_issue_order_covers_every_pair_once(S):
    """The plane-major issue loops of the kernel enumerate exactly the pairs t + u < S, every accumulator of a pass is first written by
    plane t = 0 (the kernel initialises accumulators only there), MMAs that share the A plane are consecutive and form one
    fill -> use* -> lastuse group, and a pass never needs more than the four accumulators tensor memory holds."""
    from tools.ozaki_model import issue_order

    seen = set()
    for pas in (0, 1):
        order = issue_order(S, pas)
        accs = {}
        for i, (t, u, acc, mode) in enumerate(order):
            assert 0 <= acc < 4
            assert (t, u) not in seen
            seen.add((t, u))
            accs.setdefault(acc, t)
        assert all(first_t == 0 for first_t in accs.values())
        # collector groups
        i = 0
        while i < len(order):
            t = order[i][0]
            grp = [o for o in order if o[0] == t]
            assert order[i:i + len(grp)] == grp  # consecutive
            modes = [o[3] for o in grp]
            assert modes == ([0] if len(grp) == 1 else [1] + [2] * (len(grp) - 2) + [3])
            i += len(grp)
    assert seen == {(t, u) for t in range(S) for u in range(S) if t + u < S}
    assert len(seen) == S * (S + 1) // 2


def test_stage_counts_and_issuer_split():
    from tools.ozaki_model import stages

    assert stages(4) == 6 and stages(7) == 4 and stages(6) == 4  # even: two issuing threads alternate
    assert stages(8) == 3                                        # odd: one issuer (a shared stage would let the parity wait pass a phase early)
    for planes in (4, 6, 7, 8):
        assert stages(planes) * 2 * planes * 4096 <= 224 * 1024
